// pipe_bench.cu -- issue-rate microbenchmark for the integer/FP instructions the codec kernels are built from.
// Reports warp-instructions per clock per SM sub-partition (SMSP) for each op alone and for ALU+FMA mixes, so the
// per-value instruction budgets in DESIGN.md rest on measured pipe rates.  Build: nvcc -arch=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 512;
constexpr int ILP = 8;

#define BENCH_KERNEL(NAME, BODY)                                                            \
  __global__ void __launch_bounds__(256) NAME(uint32_t* out, uint32_t seed) {               \
    uint32_t r[ILP];                                                                        \
    float f[ILP];                                                                           \
    _Pragma("unroll") for (int i = 0; i < ILP; ++i) { r[i] = seed * (threadIdx.x + i + 1); f[i] = (float)r[i] * 1e-9f; } \
    const uint32_t c0 = seed | 0x01010101u, c1 = seed ^ 0x5410u;                             \
    const float fc = (float)seed * 1e-3f;                                                    \
    _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) {                                 \
      _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                        \
        _Pragma("unroll") for (int i = 0; i < ILP; ++i) { BODY }                             \
      }                                                                                     \
    }                                                                                       \
    uint32_t acc = 0;                                                                       \
    _Pragma("unroll") for (int i = 0; i < ILP; ++i) acc += r[i] + __float_as_uint(f[i]);     \
    if (acc == 0x12345678u) out[threadIdx.x] = acc;                                          \
  }

BENCH_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_shf, asm volatile("shr.u32 %0, %0, 1;" : "+r"(r[i]));  r[i] |= 0x80000000u;)
BENCH_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_imadshl, asm volatile("mad.lo.u32 %0, %0, 16, %1;" : "+r"(r[i]) : "r"(c1));)
BENCH_KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(c0));)
BENCH_KERNEL(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc), "f"(1.0f));)
BENCH_KERNEL(k_ffma_imm, asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3F000000;" : "+f"(f[i]));)
BENCH_KERNEL(k_fmul, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fc));)
BENCH_KERNEL(k_fadd, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fc));)
BENCH_KERNEL(k_fmnmx, asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fc));)
BENCH_KERNEL(k_dp4a, asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_vmin2, asm volatile("min.u16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(c0));)
BENCH_KERNEL(k_fhadd, { unsigned short h = (unsigned short)r[i]; asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(h)); })
BENCH_KERNEL(k_hadd2f32, { float t; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(t) : "r"(r[i])); f[i] += t; })
BENCH_KERNEL(k_f2fp, { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(fc)); f[i] = __uint_as_float(r[i]); })
BENCH_KERNEL(k_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_popc, asm volatile("popc.b32 %0, %0;" : "+r"(r[i])); r[i] += c0;)
BENCH_KERNEL(k_vimnmx, asm volatile("max.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(c0));)
// mixes: one ALU op + one FMA-pipe op per slot
BENCH_KERNEL(k_mix_lop_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[(i + 4) % ILP]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_mix_lop_ffma, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc), "f"(1.0f));)
BENCH_KERNEL(k_mix_lop_prmt, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[(i + 4) % ILP]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_mix_imad_ffma, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc), "f"(1.0f));)
BENCH_KERNEL(k_mix_lop_dp4a, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[(i + 4) % ILP]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_mix_lop_fhadd, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); { unsigned short h = (unsigned short)c1; asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(h)); })
BENCH_KERNEL(k_mix_lop_f2fp, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); { uint32_t t; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(f[i]), "f"(fc)); f[i] = __uint_as_float(t); })
BENCH_KERNEL(k_mix_imad_f2fp, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1)); { uint32_t t; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(f[i]), "f"(fc)); f[i] = __uint_as_float(t); })
BENCH_KERNEL(k_mix_lop_vmin2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("min.u16x2 %0, %0, %1;" : "+r"(r[(i + 4) % ILP]) : "r"(c0));)
BENCH_KERNEL(k_mix_imad_dp4a, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[(i + 4) % ILP]) : "r"(c0), "r"(c1));)
BENCH_KERNEL(k_mix_imad_fhadd, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1)); { unsigned short h = (unsigned short)c1; asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(h)); })
BENCH_KERNEL(k_mix_ffma_fmul, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc), "f"(1.0f)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[(i + 4) % ILP]) : "f"(fc));)

template <class K>
static void run(const char* name, K kern, int ops_per_slot, int sms, double clk_hz) {
  uint32_t* out;
  cudaMalloc(&out, 4096);
  const int blocks = sms * 8;
  kern<<<blocks, 256>>>(out, 3);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    kern<<<blocks, 256>>>(out, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double warp_instr = (double)blocks * 8 * ITERS * 8 * ILP * ops_per_slot;
  const double per_clk_smsp = warp_instr / (best * 1e-3 * clk_hz) / (sms * 4.0);
  printf("%-18s %8.3f ms  %6.3f warp-instr/clk/SMSP (%d op/slot)\n", name, best, per_clk_smsp, ops_per_slot);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  printf("%s: %d SMs, clock %.0f MHz (rates assume this clock)\n", p.name, p.multiProcessorCount, clk / 1e6);
  const int sms = p.multiProcessorCount;
#define R(k, n) run(#k, k, n, sms, clk)
  R(k_lop3, 1); R(k_prmt, 1); R(k_shf, 1); R(k_iadd3, 1); R(k_vimnmx, 1); R(k_vmin2, 1); R(k_fmnmx, 1); R(k_popc, 2);
  R(k_imad, 1); R(k_imadshl, 1); R(k_dp4a, 1); R(k_ffma, 1); R(k_ffma_imm, 1); R(k_fmul, 1); R(k_fadd, 1); R(k_hfma2, 1);
  R(k_fhadd, 1); R(k_hadd2f32, 2); R(k_f2fp, 1);
  R(k_mix_lop_imad, 2); R(k_mix_lop_ffma, 2); R(k_mix_lop_prmt, 2); R(k_mix_imad_ffma, 2); R(k_mix_lop_dp4a, 2); R(k_mix_imad_dp4a, 2);
  R(k_mix_lop_fhadd, 2); R(k_mix_imad_fhadd, 2); R(k_mix_lop_f2fp, 2); R(k_mix_imad_f2fp, 2); R(k_mix_lop_vmin2, 2); R(k_mix_ffma_fmul, 2);
  return 0;
}

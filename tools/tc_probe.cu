// tc_probe.cu -- hardware probe of the tcgen05 (kind::i8) operand conventions the decode-attention kernel relies on.
// One CTA, one 128 x 16 x 128 integer MMA (4 K-steps of 32) with run-time selected shared-memory layouts / descriptors;
// the result is compared with a CPU product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu
// Run: tools/tc_probe <case>   (see main(); every wait is bounded, a hang is reported as a timeout instead)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Cfg {
  uint32_t a_lbo, a_sbo, a_layout, a_major, a_kstep;   // bytes; layout = descriptor layout_type (0 none, 2 SW128)
  uint32_t b_lbo, b_sbo, b_layout, b_major, b_kstep;
  uint32_t a_signed, b_signed;
  uint32_t a_bytes, b_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < 2000000; ++it) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

__global__ void __launch_bounds__(128) probe_kernel(Cfg c, const uint8_t* a_img, const uint8_t* b_img, int32_t* out, int* status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (base & 1023u)) & 1023u;
  uint8_t* A = smem_raw + pad;
  uint8_t* B = A + 65536;
  for (uint32_t i = threadIdx.x * 16; i < c.a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(A + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = threadIdx.x * 16; i < c.b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(B + i) = *reinterpret_cast<const uint4*>(b_img + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    // instruction descriptor: c_format S32 (2) @4, a_format @7, b_format @10, a_major @15, b_major @16, N>>3 @17, M>>4 @24
    const uint32_t idesc = (2u << 4) | (c.a_signed << 7) | (c.b_signed << 10) | (c.a_major << 15) | (c.b_major << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    for (uint32_t k = 0; k < 4; ++k) {
      const uint64_t da = make_desc(smem_u32(A) + k * c.a_kstep, c.a_lbo, c.a_sbo, c.a_layout);
      const uint64_t db = make_desc(smem_u32(B) + k * c.b_kstep, c.b_lbo, c.b_sbo, c.b_layout);
      const uint32_t acc = k > 0;
      asm volatile(
          "{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p; }"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  __syncwarp();
  const bool ok = mbar_wait_bounded(smem_u32(&s_bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (threadIdx.x == 0) *status = 1; }
  else {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// physical byte offset of element (token-like row, byte column) in a [rows][128 B] tile with the 128-byte swizzle
static uint32_t sw128(uint32_t row, uint32_t col) { return row * 128 + (((col >> 4) ^ (row & 7)) << 4) + (col & 15); }

static float frand_seed = 1;
static uint32_t lcg() { static uint32_t s = 12345; s = s * 1664525u + 1013904223u; return s >> 8; }

int main(int argc, char** argv) {
  (void)frand_seed;
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  // logical operands: A[m][k] (m < 128, k < 128) unsigned 0..15 ; B[n][k] (n < 16)
  std::vector<int> a(128 * 128), b(16 * 128);
  for (auto& x : a) x = lcg() % 16;
  Cfg c;
  memset(&c, 0, sizeof(c));
  std::vector<uint8_t> ai(65536, 0), bi(8192, 0);
  c.a_bytes = 16384; c.b_bytes = 2048;
  const char* name = "";
  // cases 0/1: S = K . Q^T  -- A = K tile, K-major SW128 (row = token m, col = dim k); B = Q limbs, K-major SW128, signed
  // cases 2/3: O = V^T . W  -- A = V tile, MN-major SW128 (row = token k, col = dim m); B = W bytes [token k][16 n], MN-major, no swizzle
  if (which == 0 || which == 1) {
    name = which == 0 ? "A K-major SW128 (u8), B K-major SW128 (s8)" : "same, LBO=0";
    for (auto& x : b) x = (int)(lcg() % 256) - 128;
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 128; ++k) ai[sw128(m, k)] = (uint8_t)a[m * 128 + k];
    for (int n = 0; n < 16; ++n) for (int k = 0; k < 128; ++k) bi[sw128(n, k)] = (uint8_t)(int8_t)b[n * 128 + k];
    c.a_lbo = which == 0 ? 16 : 0; c.a_sbo = 1024; c.a_layout = 2; c.a_major = 0; c.a_kstep = 32;
    c.b_lbo = which == 0 ? 16 : 0; c.b_sbo = 1024; c.b_layout = 2; c.b_major = 0; c.b_kstep = 32;
    c.a_signed = 0; c.b_signed = 1;
  } else {
    for (auto& x : b) x = lcg() % 256;
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 128; ++k) ai[sw128(k, m)] = (uint8_t)a[m * 128 + k];
    for (int n = 0; n < 16; ++n) for (int k = 0; k < 128; ++k) bi[k * 16 + n] = (uint8_t)b[n * 128 + k];
    c.a_layout = 2; c.a_major = 1; c.a_kstep = 4096; c.a_sbo = 1024;
    c.b_layout = 0; c.b_major = 1; c.b_kstep = 512;
    c.a_signed = 0; c.b_signed = 0;
    if (which == 2) { name = "A MN-major SW128, B MN-major none (LBO=128 k-stride, SBO=256)"; c.a_lbo = 16384; c.b_lbo = 128; c.b_sbo = 256; }
    else if (which == 3) { name = "A MN-major SW128, B MN-major none (SBO=128 k-stride, LBO=256)"; c.a_lbo = 16384; c.b_lbo = 256; c.b_sbo = 128; }
    else if (which == 4) { name = "A MN-major SW128 (LBO=1 unit), B MN-major none (LBO=128)"; c.a_lbo = 16; c.b_lbo = 128; c.b_sbo = 256; }
    else { name = "A MN-major SW128 (LBO/SBO swapped), B MN-major none (LBO=128)"; c.a_lbo = 1024; c.a_sbo = 16384; c.b_lbo = 128; c.b_sbo = 256; }
  }
  uint8_t *da, *db; int32_t* dout; int* dstat;
  cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dout, 128 * 16 * 4); cudaMalloc(&dstat, 4);
  cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xFF, 128 * 16 * 4); cudaMemset(dstat, 0, 4);
  const int smem = 65536 + 8192 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(c, da, db, dout, dstat);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case %d [%s]: CUDA error %s\n", which, name, cudaGetErrorString(e)); return 2; }
  std::vector<int32_t> out(128 * 16); int stat = 0;
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
  if (stat) { printf("case %d [%s]: TIMEOUT waiting for the MMA commit\n", which, name); return 3; }
  int bad = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    long ref = 0;
    for (int k = 0; k < 128; ++k) ref += (long)a[m * 128 + k] * b[n * 128 + k];
    if (out[m * 16 + n] != (int32_t)ref) { if (bad < 4) printf("  mismatch m=%d n=%d got %d want %ld\n", m, n, out[m * 16 + n], ref); ++bad; }
  }
  printf("case %d [%s]: %s (%d / 2048 mismatches)\n", which, name, bad ? "FAIL" : "PASS", bad);
  return bad ? 1 : 0;
}

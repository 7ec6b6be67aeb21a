// tma_probe.cu -- hardware probe of cp.async.bulk.tensor.2d tile::gather4 (sm_100a): four rows of a 2-D tensor, chosen by four
// row indices, land back to back in shared memory with ONE instruction.  The decode-attention producer uses it to fetch four
// KV pages (and four scale rows) per instruction straight from the block table.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu      Run: tools/tma_probe <box_rows>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, uint32_t row_bytes, uint8_t* out, int* status) {
  extern __shared__ __align__(128) uint8_t buf[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < 4 * row_bytes; i += blockDim.x) buf[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4 * row_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  bool ok = false;
  for (int it = 0; it < 4000000 && !ok; ++it) {
    uint32_t p;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0, 1, 0, p; }" : "=r"(p) : "r"(smem_u32(&bar)) : "memory");
    ok = p != 0;
  }
  if (!ok) { if (threadIdx.x == 0) *status = 1; return; }
  for (uint32_t i = threadIdx.x; i < 4 * row_bytes; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int which = argc > 2 ? atoi(argv[2]) : 0;            // 0: 2048-byte rows of u64, 1: 64-byte rows of u32
  const uint32_t row_bytes = which == 0 ? 2048 : 64;
  const int R = 64;
  std::vector<uint8_t> h((size_t)R * row_bytes);
  for (int r = 0; r < R; ++r) for (uint32_t c = 0; c < row_bytes; ++c) h[(size_t)r * row_bytes + c] = (uint8_t)(r * 7 + c * 13 + (c >> 8));
  uint8_t *d, *dout; int* dstat;
  cudaMalloc(&d, h.size()); cudaMalloc(&dout, 4 * row_bytes); cudaMalloc(&dstat, 4);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMemset(dstat, 0, 4); cudaMemset(dout, 0, 4 * row_bytes);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 4; }
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {which == 0 ? 256u : 16u, (cuuint64_t)R};
  const cuuint64_t gstr[1] = {row_bytes};
  const cuuint32_t box[2] = {which == 0 ? 256u : 16u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = ((EncodeTiled)fn)(&tm, which == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("box_rows %d kind %d: cuTensorMapEncodeTiled failed (%d)\n", box_rows, which, (int)cr); return 5; }
  const int rows[4] = {5, 60, 17, 2};
  probe<<<1, 128, 4 * row_bytes>>>(tm, rows[0], rows[1], rows[2], rows[3], row_bytes, dout, dstat);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("box_rows %d kind %d: CUDA error %s\n", box_rows, which, cudaGetErrorString(e)); return 2; }
  int stat = 0; cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
  if (stat) { printf("box_rows %d kind %d: TIMEOUT\n", box_rows, which); return 3; }
  std::vector<uint8_t> out(4 * row_bytes);
  cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 4; ++i) for (uint32_t c = 0; c < row_bytes; ++c) bad += out[i * row_bytes + c] != h[(size_t)rows[i] * row_bytes + c];
  printf("box_rows %d kind %d (row %u B): %s (%d mismatching bytes; first bytes %02x %02x, row1 %02x)\n", box_rows, which, row_bytes, bad ? "FAIL" : "PASS", bad,
         out[0], out[1], out[row_bytes]);
  return bad ? 1 : 0;
}

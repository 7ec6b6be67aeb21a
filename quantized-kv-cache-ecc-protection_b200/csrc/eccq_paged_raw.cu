// eccq_paged_raw.cu -- paged scatter / gather of UNENCODED rows: the unprotected comparison modes of the shim (codec "fp16":
// the K/V rows as they are; codec "fp8": their float8_e4m3fn bytes, with the reference's per-(token, head) fault injection
// on the raw bytes), kv_cache/ecc_shim.py:511-570 (write) and :802-872 (read).  Same paged layout as the protected modes
// (SimpleBlockManager, ecc_shim.py:271-316) without the scale arrays: cache [num_blocks, L, Hkv, block_size * row_bytes].
//
// Pure data movement (16-byte loads / stores, one warp per (token, head) row), HBM-bound by construction.
#include "eccq_cache_geom.cuh"

namespace eccq {

struct RawArgs {
  uint8_t* k_cache;
  uint8_t* v_cache;
  const uint8_t* k;          // write: [nseq, S, H, row_bytes]
  const uint8_t* v;
  uint8_t* k_out;            // read: [nseq, T, H, row_bytes]
  uint8_t* v_out;
  const int32_t* block_table;
  const int32_t* start_pos;
  const int32_t* context_lens;
  int row_bytes, L, H, bs, max_blocks, nseq, S, T, layer;
  int inject, n_bits;
  float ber;
  uint32_t seed_lo;
};

// one warp per row (i, s, h); 16-byte chunks strided over the lanes
template <bool INJECT>
__global__ void __launch_bounds__(256) paged_rows_write_kernel(RawArgs a, uint64_t* counters) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)a.nseq * a.S * a.H;
  uint32_t flips = 0;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int h = (int)(r % a.H);
    const int64_t tok = r / a.H;
    const int s = (int)(tok % a.S), i = (int)(tok / a.S);
    const int pos = (a.start_pos ? a.start_pos[i] : 0) + s;
    const int lb = pos / a.bs, slot = pos - lb * a.bs;
    if (lb >= a.max_blocks) continue;
    const int blk = a.block_table[(int64_t)i * a.max_blocks + lb];
    if (blk < 0) continue;
    const int64_t dst = ((((int64_t)blk * a.L + a.layer) * a.H + h) * a.bs + slot) * a.row_bytes;
    for (int c = lane * 16; c < a.row_bytes; c += 512) {
      uint4 kx = ld_stream(a.k + r * a.row_bytes + c), vx = ld_stream(a.v + r * a.row_bytes + c);
      if (INJECT) {
        // the reference injects per (token, head) row: one launch of N = row_bytes one-byte elements, seed_K = seed + count,
        // seed_V = seed_K + 1, count += 1 (ecc_shim.py:547-556) -- the bit-identical Philox stream of the flat injector
        uint32_t kw[4] = {kx.x, kx.y, kx.z, kx.w}, vw[4] = {vx.x, vx.y, vx.z, vx.w};
        const uint32_t seed_k = a.seed_lo + (uint32_t)r;
#pragma unroll 4
        for (int e = 0; e < 16; ++e) {
          const uint32_t fk = ref_flip_mask(seed_k, (uint32_t)a.row_bytes, (uint32_t)(c + e), a.n_bits, a.ber);
          const uint32_t fv = ref_flip_mask(seed_k + 1u, (uint32_t)a.row_bytes, (uint32_t)(c + e), a.n_bits, a.ber);
          kw[e >> 2] ^= fk << (8 * (e & 3));
          vw[e >> 2] ^= fv << (8 * (e & 3));
          flips += __popc(fk) + __popc(fv);
        }
        kx = make_uint4(kw[0], kw[1], kw[2], kw[3]);
        vx = make_uint4(vw[0], vw[1], vw[2], vw[3]);
      }
      st_stream(a.k_cache + dst + c, kx);
      st_stream(a.v_cache + dst + c, vx);
    }
  }
  if (INJECT) {
    const int idx[1] = {ECCQ_CTR_FLIPPED};
    const uint32_t val[1] = {flips};
    block_accumulate<1>(counters, idx, val);
  }
}

__global__ void __launch_bounds__(256) paged_rows_read_kernel(RawArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)a.nseq * a.T * a.H;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int h = (int)(r % a.H);
    const int64_t tok = r / a.H;
    const int t = (int)(tok % a.T), i = (int)(tok / a.T);
    if (t >= a.context_lens[i]) continue;
    const int lb = t / a.bs, slot = t - lb * a.bs;
    if (lb >= a.max_blocks) continue;
    const int blk = a.block_table[(int64_t)i * a.max_blocks + lb];
    if (blk < 0) continue;
    const int64_t src = ((((int64_t)blk * a.L + a.layer) * a.H + h) * a.bs + slot) * a.row_bytes;
    for (int c = lane * 16; c < a.row_bytes; c += 512) {
      st_stream(a.k_out + r * a.row_bytes + c, ld_stream(a.k_cache + src + c));
      st_stream(a.v_out + r * a.row_bytes + c, ld_stream(a.v_cache + src + c));
    }
  }
}

}  // namespace eccq

using namespace eccq;

static int raw_common(RawArgs& a, const char* what, void* k_cache, void* v_cache, int row_bytes, int L, int H, int bs, const int32_t* block_table,
                      int max_blocks, int nseq, int layer) {
  if (!k_cache || !v_cache || !block_table || row_bytes <= 0 || row_bytes % 16 != 0 || L <= 0 || H <= 0 || bs <= 0 || max_blocks <= 0 || nseq < 0 ||
      layer < 0 || layer >= L || !aligned16(k_cache) || !aligned16(v_cache))
    return fail(ECCQ_EINVAL, what);
  a.k_cache = (uint8_t*)k_cache; a.v_cache = (uint8_t*)v_cache; a.block_table = block_table;
  a.row_bytes = row_bytes; a.L = L; a.H = H; a.bs = bs; a.max_blocks = max_blocks; a.nseq = nseq; a.layer = layer;
  return ECCQ_OK;
}

extern "C" int eccq_paged_rows_write(void* k_cache, void* v_cache, const void* k, const void* v, int row_bytes, int num_layers, int num_kv_heads,
                                     int block_size, const int32_t* block_table, int max_blocks, const int32_t* start_pos, int nseq, int S,
                                     int layer, int inject, int n_bits, float ber, int64_t seed, int64_t inj_base, uint64_t* counters,
                                     void* stream) {
  RawArgs a{};
  if (int rc = raw_common(a, "paged_rows_write: bad arguments", k_cache, v_cache, row_bytes, num_layers, num_kv_heads, block_size, block_table,
                          max_blocks, nseq, layer))
    return rc;
  if (S < 0 || (nseq * (int64_t)S > 0 && (!k || !v || !aligned16(k) || !aligned16(v)))) return fail(ECCQ_EINVAL, "paged_rows_write: bad arguments");
  if (inject != ECCQ_INJECT_NONE && inject != ECCQ_INJECT_PHILOX_REF) return fail(ECCQ_EUNSUPPORTED, "paged_rows_write: injector must be none or philox_ref");
  if (inject == ECCQ_INJECT_PHILOX_REF && (n_bits < 1 || n_bits > 8)) return fail(ECCQ_EINVAL, "paged_rows_write: n_bits must be 1..8 (one-byte elements)");
  const int64_t rows = (int64_t)nseq * S * num_kv_heads;
  if (rows == 0) return ECCQ_OK;
  a.k = (const uint8_t*)k; a.v = (const uint8_t*)v; a.start_pos = start_pos; a.S = S;
  a.inject = inject; a.n_bits = n_bits; a.ber = ber; a.seed_lo = (uint32_t)((uint64_t)seed + (uint64_t)inj_base);
  const int grid = warps_grid(rows);
  if (inject == ECCQ_INJECT_PHILOX_REF && ber > 0.0f)
    paged_rows_write_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a, counters);
  else
    paged_rows_write_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, counters);
  return check_launch("paged_rows_write");
}

extern "C" int eccq_paged_rows_read(const void* k_cache, const void* v_cache, void* k_out, void* v_out, int row_bytes, int num_layers,
                                    int num_kv_heads, int block_size, const int32_t* block_table, int max_blocks, const int32_t* context_lens,
                                    int nseq, int out_stride_tokens, int layer, void* stream) {
  RawArgs a{};
  if (int rc = raw_common(a, "paged_rows_read: bad arguments", const_cast<void*>(k_cache), const_cast<void*>(v_cache), row_bytes, num_layers,
                          num_kv_heads, block_size, block_table, max_blocks, nseq, layer))
    return rc;
  if (out_stride_tokens < 0 || !context_lens) return fail(ECCQ_EINVAL, "paged_rows_read: bad arguments");
  const int64_t rows = (int64_t)nseq * out_stride_tokens * num_kv_heads;
  if (rows == 0) return ECCQ_OK;
  if (!k_out || !v_out || !aligned16(k_out) || !aligned16(v_out)) return fail(ECCQ_EINVAL, "paged_rows_read: bad output pointers");
  a.k_out = (uint8_t*)k_out; a.v_out = (uint8_t*)v_out; a.context_lens = context_lens; a.T = out_stride_tokens;
  paged_rows_read_kernel<<<warps_grid(rows), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("paged_rows_read");
}

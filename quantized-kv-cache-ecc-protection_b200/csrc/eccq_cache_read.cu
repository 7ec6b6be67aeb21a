// eccq_cache_read.cu -- the fused paged-cache read.
//
//   cache read   = read half of ECCBackend.attend (ecc_shim.py:931-1071): gather pages + syndrome decode + correct +
//                  SECDED flag + temporal interpolation of double-detected values + dequantise, K and V, ONE launch;
//   correct_pages = the same decode into corrected INT4 pages (decode attention with interpolation / Golay);
//   dequantize_rows = the arithmetic without the paging (fused_decode_dequantize_hamming84, fused_kernels.py:372-437).
#include "eccq_cache_geom.cuh"

namespace eccq {

// ---------------------------------------------------------------------------------------------- fused read
struct ReadArgs {
  const int32_t* block_table;
  const int32_t* context_lens;
  void* k_out;
  void* v_out;
  uint8_t* nib_k;
  uint8_t* nib_v;
  int max_blocks, nseq, T, layer;  // T = out_stride_tokens
  // eccq_cache_correct_pages: nib_k / nib_v are INT4 pages of a one-layer cache with the SAME paged layout (page = block*H + h)
  // and the row scales are copied beside them
  int paged_nib = 0;
  float* sck_out = nullptr;
  float* scv_out = nullptr;
};

template <typename TO>
__device__ __forceinline__ void store16(TO* dst, const uint32_t (&nib)[4], float scale);
template <>
__device__ __forceinline__ void store16<float>(float* dst, const uint32_t (&nib)[4], float scale) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 o;
    o.x = __float_as_uint(byte_to_float_m8(nib[k], 0) * scale);
    o.y = __float_as_uint(byte_to_float_m8(nib[k], 1) * scale);
    o.z = __float_as_uint(byte_to_float_m8(nib[k], 2) * scale);
    o.w = __float_as_uint(byte_to_float_m8(nib[k], 3) * scale);
    st_stream(dst + 4 * k, o);
  }
}
template <>
__device__ __forceinline__ void store16<__half>(__half* dst, const uint32_t (&nib)[4], float scale) {
  uint32_t o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 a = __floats2half2_rn(byte_to_float_m8(nib[k], 0) * scale, byte_to_float_m8(nib[k], 1) * scale);
    const __half2 b = __floats2half2_rn(byte_to_float_m8(nib[k], 2) * scale, byte_to_float_m8(nib[k], 3) * scale);
    o[2 * k] = *reinterpret_cast<const uint32_t*>(&a);
    o[2 * k + 1] = *reinterpret_cast<const uint32_t*>(&b);
  }
  st_stream(dst, make_uint4(o[0], o[1], o[2], o[3]));
  st_stream(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
}
template <>
__device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* dst, const uint32_t (&nib)[4], float scale) {
  uint32_t o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(byte_to_float_m8(nib[k], 0) * scale, byte_to_float_m8(nib[k], 1) * scale);
    const __nv_bfloat162 b = __floats2bfloat162_rn(byte_to_float_m8(nib[k], 2) * scale, byte_to_float_m8(nib[k], 3) * scale);
    o[2 * k] = *reinterpret_cast<const uint32_t*>(&a);
    o[2 * k + 1] = *reinterpret_cast<const uint32_t*>(&b);
  }
  st_stream(dst, make_uint4(o[0], o[1], o[2], o[3]));
  st_stream(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
}

// decode one 16-byte chunk: nibbles, per-byte double flags; counters accumulate as four byte-wide sums per word
// (flushed by the caller before a byte can overflow)
template <int CODEC>
__device__ __forceinline__ void decode16(const uint4& x, uint32_t (&nib)[4], uint32_t (&dbl)[4], uint32_t& acc_corr, uint32_t& acc_det) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (CODEC == ECCQ_CODEC_H84) {
      const Dec4 d = h84_decode4(w[k]);
      nib[k] = d.nib;
      dbl[k] = h84_double4(d);
      acc_corr += h84_single4(d);
      acc_det += dbl[k];
    } else if (CODEC == ECCQ_CODEC_H74) {
      const Dec4 d = h74_decode4(w[k]);
      nib[k] = d.nib; dbl[k] = 0;
      acc_corr += d.nz;
    } else {
      nib[k] = w[k]; dbl[k] = 0;  // int4: raw nibble bytes (ecc_shim.py:913-914 uses the byte as is)
    }
  }
}

// Fast path: D % 16 == 0, C = D/16 lanes per token row (power of two <= 32), groups = 32/C, block_size % groups == 0,
// SPL = block_size/groups consecutive token slots per lane (<= MAXSPL). One warp per (tensor, sequence, block, head) page:
// every lane keeps SPL consecutive tokens of its 16-byte column in registers, so the temporal neighbours of a
// double-detected value are a register, a lane shuffle, or (page edge, rare) one extra 16-byte load.
template <typename TO, int CODEC, bool INTERP, int MAXSPL>
__global__ void __launch_bounds__(256, ((MAXSPL <= 4 && !INTERP) ? 3 : 2)) cache_read_kernel(CacheGeom g, ReadArgs a, int nblk_logical, uint64_t* counters) {
  const int C = g.D >> 4;
  const int groups = 32 / C;
  const int SPL = g.bs / groups;
  const int lane = threadIdx.x & 31;
  const int c = lane & (C - 1);
  const int grp = lane / C;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = (int64_t)2 * a.nseq * nblk_logical * g.H;
  const int64_t page_bytes = (int64_t)g.bs * g.D;
  uint32_t ncorr = 0, ndet = 0, nint = 0;

  for (int64_t u = warp; u < units; u += nwarps) {
    // u = ((i*nblk + j)*H + h)*2 + kv : K and V pages of a block are handled by adjacent warps
    const int kv = (int)(u & 1);
    const int h = (int)((u >> 1) % g.H);
    const int64_t ij = (u >> 1) / g.H;
    const int j = (int)(ij % nblk_logical);
    const int i = (int)(ij / nblk_logical);
    const int ctx = min(a.context_lens[i], a.T);
    const int t0 = j * g.bs;
    if (t0 >= ctx || j >= a.max_blocks) continue;
    const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;
    const int blk = bt[j];
    if (blk < 0) continue;
    const uint8_t* cache = kv ? g.v_cache : g.k_cache;
    const float* scales = kv ? g.v_scales : g.k_scales;
    const int64_t page = page_index(g, blk, a.layer, h);
    const uint8_t* pbase = cache + page * page_bytes + c * 16;
    TO* out = reinterpret_cast<TO*>(kv ? a.v_out : a.k_out);
    uint8_t* nout = kv ? a.nib_v : a.nib_k;
    const int slot0 = grp * SPL;

    uint4 raw[MAXSPL];
    float sc[MAXSPL];
#pragma unroll
    for (int s = 0; s < MAXSPL; ++s) {
      const bool valid = s < SPL && t0 + slot0 + s < ctx;
      raw[s] = make_uint4(0, 0, 0, 0);
      sc[s] = 0.0f;
      if (valid) {
        raw[s] = ld_stream(pbase + (int64_t)(slot0 + s) * g.D);
        sc[s] = __ldg(scales + page * g.bs + slot0 + s);
      }
    }
    uint32_t nib[MAXSPL][4], dbl[MAXSPL][4];
    uint32_t acc_c = 0, acc_d = 0;  // <= 4*MAXSPL <= 32 per byte
#pragma unroll
    for (int s = 0; s < MAXSPL; ++s) {
      const bool valid = s < SPL && t0 + slot0 + s < ctx;
      if (valid) decode16<CODEC>(raw[s], nib[s], dbl[s], acc_c, acc_d);
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) nib[s][k] = dbl[s][k] = 0;
      }
    }
    ncorr += bytesum(acc_c);
    ndet += bytesum(acc_d);

    uint32_t up[4], dn[4];
    if (INTERP && CODEC == ECCQ_CODEC_H84) {
      // neighbours held by the adjacent lane groups (pre-interpolation nibbles)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t last = nib[0][k];
#pragma unroll
        for (int s = 1; s < MAXSPL; ++s) if (s < SPL) last = nib[s][k];
        up[k] = __shfl_up_sync(0xFFFFFFFFu, last, C);
        dn[k] = __shfl_down_sync(0xFFFFFFFFu, nib[0][k], C);
      }
    }
#pragma unroll
    for (int s = 0; s < MAXSPL; ++s) {
      const int slot = slot0 + s;
      const int tok = t0 + slot;
      if (!(s < SPL && tok < ctx)) continue;
      uint32_t o4[4] = {nib[s][0], nib[s][1], nib[s][2], nib[s][3]};
      if (INTERP && CODEC == ECCQ_CODEC_H84) {
        const uint32_t any = dbl[s][0] | dbl[s][1] | dbl[s][2] | dbl[s][3];
        if (any != 0) {
          uint32_t left[4], right[4];
          // ---- left neighbour: token max(tok-1, 0)
          if (tok == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) left[k] = nib[s][k];
          } else if (s > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) left[k] = nib[s > 0 ? s - 1 : 0][k];
          } else if (grp > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) left[k] = up[k];
          } else {  // last slot of the previous logical block
            const int pb = bt[j - 1];
            uint4 x = make_uint4(0, 0, 0, 0);
            if (pb >= 0) x = ld_stream(cache + page_index(g, pb, a.layer, h) * page_bytes + (int64_t)(g.bs - 1) * g.D + c * 16);
            uint32_t d2[4], c0 = 0, c1 = 0;
            decode16<CODEC>(x, left, d2, c0, c1);
          }
          // ---- right neighbour: token min(tok+1, ctx-1)
          if (tok + 1 >= ctx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) right[k] = nib[s][k];
          } else if (s + 1 < SPL) {
#pragma unroll
            for (int k = 0; k < 4; ++k) right[k] = nib[s + 1 < MAXSPL ? s + 1 : s][k];
          } else if (grp + 1 < groups) {
#pragma unroll
            for (int k = 0; k < 4; ++k) right[k] = dn[k];
          } else {  // first slot of the next logical block
            const int nb = (j + 1 < a.max_blocks) ? bt[j + 1] : -1;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (nb >= 0) x = ld_stream(cache + page_index(g, nb, a.layer, h) * page_bytes + c * 16);
            uint32_t d2[4], c0 = 0, c1 = 0;
            decode16<CODEC>(x, right, d2, c0, c1);
          }
          uint32_t acc_i = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            o4[k] = interp4(nib[s][k], left[k], right[k], dbl[s][k]);
            acc_i += dbl[s][k];
          }
          nint += bytesum(acc_i);
        }
      }
      const int64_t o = (((int64_t)i * a.T + tok) * g.H + h) * g.D + c * 16;
      if (out) store16<TO>(out + o, o4, sc[s]);
      if (nout) st_stream(nout + o, make_uint4(o4[0], o4[1], o4[2], o4[3]));
    }
  }
  const int idx[3] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED, ECCQ_CTR_INTERPOLATED};
  const uint32_t val[3] = {ncorr, ndet, nint};
  block_accumulate<3>(counters, idx, val);
}

// Generic read: any D / block size / codec (Golay included). One warp per (tensor, sequence, token, head) row.
__device__ __forceinline__ uint32_t decode_byte_nib(int codec, uint32_t b, uint32_t& single, uint32_t& dbl) {
  single = dbl = 0;
  if (codec == ECCQ_CODEC_H84) { const Dec4 d = h84_decode4(b & 0xFFu); single = h84_single4(d) & 1u; dbl = h84_double4(d) & 1u; return d.nib & 0xFu; }
  if (codec == ECCQ_CODEC_H74) { const Dec4 d = h74_decode4(b & 0xFFu); single = d.nz & 1u; return d.nib & 0xFu; }
  return b & 0xFFu;
}

template <typename TO>
__global__ void __launch_bounds__(256) cache_read_generic_kernel(CacheGeom g, ReadArgs a, int interp, uint64_t* counters) {
  __shared__ uint16_t s_lo[64], s_hi[64];
  __shared__ __align__(16) int32_t s_syn[4096];
  if (g.codec == ECCQ_CODEC_GOLAY) { golay_load_parity(s_lo, s_hi); golay_load_syndrome(s_syn); }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = (int64_t)2 * a.nseq * a.T * g.H;
  uint32_t ncorr = 0, ndet = 0, nint = 0;
  for (int64_t u = warp; u < units; u += nwarps) {
    const int kv = (int)(u & 1);
    const int h = (int)((u >> 1) % g.H);
    const int64_t it = (u >> 1) / g.H;
    const int tok = (int)(it % a.T);
    const int i = (int)(it / a.T);
    const int ctx = min(a.context_lens[i], a.T);
    if (tok >= ctx) continue;
    const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;
    const int lb = tok / g.bs, slot = tok - lb * g.bs;
    if (lb >= a.max_blocks) continue;
    const int blk = bt[lb];
    if (blk < 0) continue;
    const uint8_t* cache = kv ? g.v_cache : g.k_cache;
    const int64_t page = page_index(g, blk, a.layer, h);
    const float scale = (kv ? g.v_scales : g.k_scales)[page * g.bs + slot];
    TO* out = reinterpret_cast<TO*>(kv ? a.v_out : a.k_out);
    uint8_t* nout = kv ? a.nib_v : a.nib_k;
    const int64_t o = (((int64_t)i * a.T + tok) * g.H + h) * g.D;
    auto row_ptr = [&](int t) -> const uint8_t* {  // codeword row of token t (same head), nullptr if unmapped
      const int b = t / g.bs;
      const int pb = bt[b];
      if (pb < 0) return nullptr;
      return cache + (page_index(g, pb, a.layer, h) * (int64_t)g.bs + (t - b * g.bs)) * g.cw * (g.codec == ECCQ_CODEC_GOLAY ? 4 : 1);
    };
    const uint8_t* row = row_ptr(tok);
    if (g.codec != ECCQ_CODEC_GOLAY) {
      for (int e = lane; e < g.D; e += 32) {
        uint32_t sg, db;
        uint32_t n = decode_byte_nib(g.codec, row[e], sg, db);
        ncorr += sg; ndet += db;
        if (interp && db) {
          uint32_t s2, d2;
          const uint8_t* lr = tok > 0 ? row_ptr(tok - 1) : row;
          const uint8_t* rr = tok + 1 < ctx ? row_ptr(tok + 1) : row;
          const uint32_t l = lr ? decode_byte_nib(g.codec, lr[e], s2, d2) : 0u;
          const uint32_t rgt = rr ? decode_byte_nib(g.codec, rr[e], s2, d2) : 0u;
          n = min((l + rgt + 1u) >> 1, 15u);
          ++nint;
        }
        if (out) out[o + e] = Cvt<TO>::to(((float)(int)n - 8.0f) * scale);
        if (nout) nout[o + e] = (uint8_t)n;
      }
    } else {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
      for (int j = lane; j < g.cw; j += 32) {
        uint32_t cnt;
        const uint32_t d = golay_decode1(s_lo, s_hi, s_syn, rw[j], cnt);
        if (cnt == 4u) ++ndet; else ncorr += cnt;
        for (int t = 0; t < 3; ++t) {
          const int e = 3 * j + t;
          if (e < g.D) {
            const uint32_t n = (d >> (4 * t)) & 0xFu;
            if (out) out[o + e] = Cvt<TO>::to(((float)(int)n - 8.0f) * scale);
            if (nout) nout[o + e] = (uint8_t)n;
          }
        }
      }
    }
  }
  const int idx[3] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED, ECCQ_CTR_INTERPOLATED};
  const uint32_t val[3] = {ncorr, ndet, nint};
  block_accumulate<3>(counters, idx, val);
}

template <typename TO>
__global__ void __launch_bounds__(256) dequantize_rows_kernel(const uint8_t* __restrict__ q, const float* __restrict__ scales,
                                                               TO* __restrict__ out, int64_t rows, int row_len, int codec,
                                                               int zero_double, uint64_t* counters) {
  const int64_t n = rows * row_len;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  uint32_t ncorr = 0, ndet = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    uint32_t sg = 0, db = 0;
    uint32_t nb = codec < 0 ? (uint32_t)q[i] : decode_byte_nib(codec, q[i], sg, db);
    if (zero_double && db) nb = 0;
    ncorr += sg; ndet += db;
    out[i] = Cvt<TO>::to(((float)(int)nb - 8.0f) * scales[i / row_len]);
  }
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {ncorr, ndet};
  block_accumulate<2>(counters, idx, val);
}


}  // namespace eccq
#include "eccq_cache_read2.cuh"
#include "eccq_cache_golay.cuh"
namespace eccq {

template <typename TO, int CODEC, int C, int SPL, int HP = 1>
static void launch_read2(const CacheGeom& g, const ReadArgs& a, int nblk, bool interp, uint64_t* counters, cudaStream_t st) {
  static const int chunk_env = env_int("ECCQ_READ_CHUNK", 0), hf = env_int("ECCQ_READ_HEAD_FASTEST", 0);
  auto launch = [&](auto k) {
    const int64_t pages = (int64_t)2 * a.nseq * nblk * g.H;
    const int grid = persistent_grid(k, 256, (pages + 7) / 8);
    // default: split every (sequence, tensor, head) run into the fewest equal chunks that give each resident warp >= 1 unit
    int chunk = chunk_env;
    if (chunk <= 0) {
      const int64_t runs = (int64_t)2 * a.nseq * (g.H / HP), warps = (int64_t)grid * 8;
      const int64_t per_run = std::max<int64_t>(1, (4 * warps + runs - 1) / runs);   // ~4 units per warp
      chunk = (int)std::max<int64_t>(4, (nblk + per_run - 1) / per_run);
    }
    k<<<grid, 256, 0, st>>>(g, a, nblk, chunk, hf, counters);
  };
  if (interp) launch(cache_read2_kernel<TO, CODEC, true, C, SPL, HP>);
  else launch(cache_read2_kernel<TO, CODEC, false, C, SPL, HP>);
}

template <typename TO, int CODEC>
static void launch_read_fast(const CacheGeom& g, const ReadArgs& a, int nblk, bool interp, uint64_t* counters, cudaStream_t st) {
  const int64_t units = (int64_t)2 * a.nseq * nblk * g.H;
  if (units < 0x7FFFFFFFll && aligned16(g.k_scales) && aligned16(g.v_scales)) {   // production geometries: head_dim 128 / 64 with 16- or 32-token blocks (vector loads of the row scales)
    if (g.D == 128 && g.bs == 16) { launch_read2<TO, CODEC, 8, 4>(g, a, nblk, interp, counters, st); return; }
    // two heads per warp: +30 % with interpolation (GPT-2 geometry, 3.8 -> 4.9 TB/s); without it the one-head mapping is 5 % faster
    if (g.D == 64 && g.bs == 16 && g.H % 2 == 0 && interp) { launch_read2<TO, CODEC, 4, 4, 2>(g, a, nblk, interp, counters, st); return; }
    if (g.D == 64 && g.bs == 16) { launch_read2<TO, CODEC, 4, 2>(g, a, nblk, interp, counters, st); return; }
    if (g.D == 64 && g.bs == 32) { launch_read2<TO, CODEC, 4, 4>(g, a, nblk, interp, counters, st); return; }
  }
  const int grid = warps_grid(units);
  const int SPL = g.bs / (32 / (g.D >> 4));
  if (SPL <= 4) {
    if (interp) cache_read_kernel<TO, CODEC, true, 4><<<grid, 256, 0, st>>>(g, a, nblk, counters);
    else cache_read_kernel<TO, CODEC, false, 4><<<grid, 256, 0, st>>>(g, a, nblk, counters);
  } else {
    if (interp) cache_read_kernel<TO, CODEC, true, 8><<<grid, 256, 0, st>>>(g, a, nblk, counters);
    else cache_read_kernel<TO, CODEC, false, 8><<<grid, 256, 0, st>>>(g, a, nblk, counters);
  }
}
template <typename TO>
static void launch_read(const CacheGeom& g, const ReadArgs& a, bool interp, uint64_t* counters, cudaStream_t st) {
  const int C = g.D >> 4;
  bool fast = g.codec != ECCQ_CODEC_GOLAY && (g.D % 16 == 0) && is_pow2(C) && C <= 32;
  if (fast) {
    const int groups = 32 / C;
    fast = (g.bs % groups == 0) && (g.bs / groups) <= 8 && aligned16(g.k_cache) && aligned16(g.v_cache) &&
           (!a.k_out || aligned16(a.k_out)) && (!a.v_out || aligned16(a.v_out)) && (!a.nib_k || aligned16(a.nib_k)) &&
           (!a.nib_v || aligned16(a.nib_v));
  }
  if (!fast && g.codec == ECCQ_CODEC_GOLAY && (a.paged_nib || (!a.nib_k && !a.nib_v))) {
    const int chunks = g.bs * g.cw / 4;
    const int nblk = std::min((a.T + g.bs - 1) / g.bs, a.max_blocks);
    const int64_t units = (int64_t)2 * a.nseq * nblk * g.H;
    if ((g.bs * g.cw) % 4 == 0 && chunks <= 256 && g.bs <= 32 && (g.D * sizeof(TO)) % 16 == 0 && units < 0x7FFFFFFFll &&
        aligned16(g.k_cache) && aligned16(g.v_cache) && (!a.k_out || aligned16(a.k_out)) && (!a.v_out || aligned16(a.v_out))) {
      if ((g.D == 128 || g.D == 64) && g.bs == 16 && aligned16(g.k_scales) && aligned16(g.v_scales)) {
        auto launch2 = [&](auto k, size_t smem) {
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          int per_sm = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem) != cudaSuccess || per_sm <= 0) per_sm = 1;
          const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + 7) / 8, (int64_t)sm_count() * per_sm));
          k<<<grid, 256, smem, st>>>(g, a, nblk, counters);
        };
        const size_t words = (size_t)g.bs * g.cw + g.bs;
        const size_t smem2 = kGolay2TableBytes + 8 * 2 * words * 4;
        if (g.D == 128) launch2(cache_read_golay2_kernel<TO, 128, 16>, smem2);
        else launch2(cache_read_golay2_kernel<TO, 64, 16>, smem2);
        return;
      }
      GolayReadArgs ga;
      ga.div_cw = FastDiv::make((uint32_t)g.cw);
      ga.chunks = chunks;
      ga.out_chunks_row = (int)(g.D * sizeof(TO) / 16);
      const size_t smem = 16384 + 256 + (size_t)8 * (g.bs * g.D * sizeof(TO) + g.bs * 4);
      auto k = cache_read_golay_kernel<TO>;
      static size_t smem_set = 0;
      if (smem > smem_set) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); smem_set = smem; }
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem) != cudaSuccess || per_sm <= 0) per_sm = 1;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + 7) / 8, (int64_t)sm_count() * per_sm));
      k<<<grid, 256, smem, st>>>(g, a, ga, nblk, counters);
      return;
    }
  }
  if (!fast) {
    const int64_t units = (int64_t)2 * a.nseq * a.T * g.H;
    cache_read_generic_kernel<TO><<<warps_grid(units), 256, 0, st>>>(g, a, interp ? 1 : 0, counters);
    return;
  }
  const int nblk = std::min((a.T + g.bs - 1) / g.bs, a.max_blocks);
  switch (g.codec) {
    case ECCQ_CODEC_H74: launch_read_fast<TO, ECCQ_CODEC_H74>(g, a, nblk, false, counters, st); break;
    case ECCQ_CODEC_H84: launch_read_fast<TO, ECCQ_CODEC_H84>(g, a, nblk, interp, counters, st); break;
    default: launch_read_fast<TO, ECCQ_CODEC_INT4>(g, a, nblk, false, counters, st); break;
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) dequantize_rows16_kernel(const uint8_t* __restrict__ q, const float* __restrict__ scales,
                                                                 TO* __restrict__ out, int64_t nchunks, int G, int codec,
                                                                 int zero_double, uint64_t* counters) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  FlagCount fc;
  for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += nthreads) {
    const uint4 x = ld_stream(q + (ch << 4));
    const float sc = __ldg(scales + ch / G);
    uint32_t nib[4], flag[4];
    if (codec == ECCQ_CODEC_H84) decode_words4<ECCQ_CODEC_H84>(x, nib, flag);
    else if (codec == ECCQ_CODEC_H74) decode_words4<ECCQ_CODEC_H74>(x, nib, flag);
    else decode_words4<ECCQ_CODEC_INT4>(x, nib, flag);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      fc.add(flag[k]);
      if (zero_double) nib[k] &= ~((flag[k] & M01) * 15u);   // fused_kernels.py:343-344 (Q4)
    }
    dequant_store16<TO>(out + (ch << 4), nib, sc);
  }
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {fc.singles(), fc.doubles()};
  block_accumulate<2>(counters, idx, val);
}


}  // namespace eccq

using namespace eccq;

ECCQ_DEFINE_TU_TABLES(cache_read)

extern "C" int eccq_cache_read(const eccq_cache_desc* c, const int32_t* block_table, int max_blocks, const int32_t* context_lens,
                               int nseq, int out_stride_tokens, int layer, int use_interpolation, int out_dtype, void* k_out,
                               void* v_out, uint8_t* nib_out_k, uint8_t* nib_out_v, uint64_t* counters, void* stream) {
  CacheGeom g;
  if (!make_geom(c, g)) return fail(ECCQ_EINVAL, "cache_read: bad cache descriptor");
  if (nseq < 0 || out_stride_tokens < 0 || layer < 0 || layer >= g.L || max_blocks <= 0 || !block_table || !context_lens)
    return fail(ECCQ_EINVAL, "cache_read: bad arguments");
  if (nseq == 0 || out_stride_tokens == 0) return ECCQ_OK;
  if (int rc = ensure_tables()) return rc;
  ReadArgs a;
  a.block_table = block_table; a.context_lens = context_lens; a.k_out = k_out; a.v_out = v_out;
  a.nib_k = nib_out_k; a.nib_v = nib_out_v; a.max_blocks = max_blocks; a.nseq = nseq; a.T = out_stride_tokens; a.layer = layer;
  const bool interp = use_interpolation && g.codec == ECCQ_CODEC_H84;  // the reference interpolates hamming84 only
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_dtype) {
    case ECCQ_F32: launch_read<float>(g, a, interp, counters, st); break;
    case ECCQ_F16: launch_read<__half>(g, a, interp, counters, st); break;
    case ECCQ_BF16: launch_read<__nv_bfloat16>(g, a, interp, counters, st); break;
    default: return fail(ECCQ_EINVAL, "cache_read: bad dtype");
  }
  return check_launch("cache_read");
}

extern "C" int eccq_cache_correct_pages(const eccq_cache_desc* c, const int32_t* block_table, int max_blocks,
                                        const int32_t* context_lens, int nseq, int layer, int use_interpolation, uint8_t* k_pages,
                                        uint8_t* v_pages, float* k_page_scales, float* v_page_scales, uint64_t* counters,
                                        void* stream) {
  CacheGeom g;
  if (!make_geom(c, g)) return fail(ECCQ_EINVAL, "cache_correct_pages: bad cache descriptor");
  if (nseq < 0 || layer < 0 || layer >= g.L || max_blocks <= 0 || !block_table || !context_lens || !k_pages || !v_pages ||
      !k_page_scales || !v_page_scales)
    return fail(ECCQ_EINVAL, "cache_correct_pages: bad arguments");
  if (nseq == 0) return ECCQ_OK;
  const bool geom_ok = g.codec == ECCQ_CODEC_GOLAY ? ((g.D == 128 || g.D == 64) && g.bs == 16)
                                                    : ((g.D == 128 && g.bs == 16) || (g.D == 64 && (g.bs == 16 || g.bs == 32)));
  if (!geom_ok || !aligned16(g.k_cache) || !aligned16(g.v_cache) || !aligned16(g.k_scales) ||
      !aligned16(g.v_scales) || !aligned16(k_pages) || !aligned16(v_pages) ||
      (int64_t)2 * nseq * max_blocks * g.H >= 0x7FFFFFFFll)
    return fail(ECCQ_EUNSUPPORTED, "cache_correct_pages: geometry / codec outside the streaming read kernel");
  if (int rc = ensure_tables()) return rc;   // the Golay branch reads g_golay; this may be the first library call on the device
  ReadArgs a;
  a.block_table = block_table; a.context_lens = context_lens; a.k_out = nullptr; a.v_out = nullptr;
  a.nib_k = k_pages; a.nib_v = v_pages; a.max_blocks = max_blocks; a.nseq = nseq; a.T = max_blocks * g.bs; a.layer = layer;
  a.paged_nib = 1; a.sck_out = k_page_scales; a.scv_out = v_page_scales;
  const bool interp = use_interpolation && g.codec == ECCQ_CODEC_H84;
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.codec) {
    case ECCQ_CODEC_H74: launch_read_fast<__half, ECCQ_CODEC_H74>(g, a, max_blocks, false, counters, st); break;
    case ECCQ_CODEC_H84: launch_read_fast<__half, ECCQ_CODEC_H84>(g, a, max_blocks, interp, counters, st); break;
    case ECCQ_CODEC_GOLAY: launch_read<__half>(g, a, false, counters, st); break;   // takes the cache_read_golay2_kernel branch
    default: launch_read_fast<__half, ECCQ_CODEC_INT4>(g, a, max_blocks, false, counters, st); break;
  }
  return check_launch("cache_correct_pages");
}

extern "C" int eccq_dequantize_rows(const uint8_t* q, const float* scales, void* out, int out_dtype, int64_t rows, int row_len,
                                    int codec, int zero_double, uint64_t* counters, void* stream) {
  if (rows < 0 || row_len <= 0 || (rows && (!q || !scales || !out))) return fail(ECCQ_EINVAL, "dequantize_rows: bad arguments");
  if (codec > ECCQ_CODEC_H84 && codec != ECCQ_CODEC_INT4) return fail(ECCQ_EUNSUPPORTED, "dequantize_rows: codec must be h74/h84/int4/<0");
  if (rows == 0) return ECCQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = rows * row_len;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8));
  const int cdc = codec == ECCQ_CODEC_INT4 ? -1 : codec;
  if (row_len % 16 == 0 && aligned16(q) && aligned16(out)) {
    const int64_t nchunks = n >> 4;
    const int G = row_len / 16;
    const int grid16 = (int)std::max<int64_t>(1, std::min<int64_t>((nchunks + 255) / 256, (int64_t)sm_count() * 8));
    const int cd = cdc < 0 ? ECCQ_CODEC_INT4 : cdc;
    switch (out_dtype) {
      case ECCQ_F32: dequantize_rows16_kernel<float><<<grid16, 256, 0, st>>>(q, scales, (float*)out, nchunks, G, cd, zero_double, counters); break;
      case ECCQ_F16: dequantize_rows16_kernel<__half><<<grid16, 256, 0, st>>>(q, scales, (__half*)out, nchunks, G, cd, zero_double, counters); break;
      case ECCQ_BF16: dequantize_rows16_kernel<__nv_bfloat16><<<grid16, 256, 0, st>>>(q, scales, (__nv_bfloat16*)out, nchunks, G, cd, zero_double, counters); break;
      default: return fail(ECCQ_EINVAL, "dequantize_rows: bad dtype");
    }
    return check_launch("dequantize_rows");
  }
  switch (out_dtype) {
    case ECCQ_F32: dequantize_rows_kernel<float><<<grid, 256, 0, st>>>(q, scales, (float*)out, rows, row_len, cdc, zero_double, counters); break;
    case ECCQ_F16: dequantize_rows_kernel<__half><<<grid, 256, 0, st>>>(q, scales, (__half*)out, rows, row_len, cdc, zero_double, counters); break;
    case ECCQ_BF16: dequantize_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(q, scales, (__nv_bfloat16*)out, rows, row_len, cdc, zero_double, counters); break;
    default: return fail(ECCQ_EINVAL, "dequantize_rows: bad dtype");
  }
  return check_launch("dequantize_rows");
}


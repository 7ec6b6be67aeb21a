// eccq_attention.cu -- ECC-fused paged decode attention (q_len == 1), split over the context (FlashDecoding).
//
// Replaces paged_attention_ecc / paged_attention_ecc_kernel (kv_cache/attention_ecc.py:620-780, :264-426): the reference
// walks the whole context serially with one program per (batch, q-head), re-decoding every K/V row once per q-head
// and mis-indexing GQA caches.  Here one CTA owns a (sequence, kv-head, context-chunk): it decodes each
// codeword row once for the whole GQA group, keeps an online-softmax state per q-head and writes a partial
// (m, l, acc); a second small kernel merges the chunks.  Codeword bytes are consumed directly: the Hamming syndrome
// decode (or the Golay table decode) is the K/V prologue, per-token scales are applied to the score / to P.
//
// v0 (this file): CUDA-core dot products, warp per token.  The tcgen05/TMEM tile version builds on the same
// partial/merge protocol.
#include "eccq_attention.h"

namespace eccq {

constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;

template <typename T>
__device__ __forceinline__ float ldq(const void* q, int64_t i) { return Cvt<T>::from(reinterpret_cast<const T*>(q)[i]); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// EPL = values per lane. Hamming/int4: EPL = 4*NCH (one 32-bit word of codeword bytes per 128-value chunk);
// Golay: EPL = 3*NIT (one codeword per 32-codeword sweep).
// PREFILL = true turns the same kernel into the ECC-fused causal prefill attention (replaces the gather + decode + fp32
// materialisation + repeat_interleave + SDPA of ECCBackend.attend, kv_cache/ecc_shim.py:1067-1089, :1138-1164): the GP
// "heads" of a CTA become GP consecutive query positions of ONE q-head, every token row is decoded once per CTA in the K/V
// prologue exactly as for decode attention (interpolation included), a causal mask is applied per query, and the CTA writes
// the normalised output itself (no context split, no merge kernel).  grid = (query blocks, Hq, batch rows).
template <typename TQ, int CODEC, int GP, int NCH, bool PREFILL = false>
__global__ void __launch_bounds__(kAttnThreads) attn_partial_kernel(CacheGeom g, AttnArgs a, uint64_t* counters) {
  constexpr bool GOLAY = CODEC == ECCQ_CODEC_GOLAY;
  constexpr int VPC = GOLAY ? 3 : 4;   // values per lane per chunk
  constexpr int EPL = VPC * NCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [golay tables] q[GP][EPL*32] floats, then per-warp merge area
  uint16_t* s_lo = nullptr; uint16_t* s_hi = nullptr; int32_t* s_syn = nullptr;
  unsigned char* p = smem_raw;
  if (GOLAY) {
    s_syn = reinterpret_cast<int32_t*>(p); p += 4096 * 4;
    s_lo = reinterpret_cast<uint16_t*>(p); p += 128;
    s_hi = reinterpret_cast<uint16_t*>(p); p += 128;
    golay_load_parity(s_lo, s_hi);
    golay_load_syndrome(s_syn);
  }
  float* s_red = reinterpret_cast<float*>(p);  // [kAttnWarps][GP][EPL*32 + 2]

  const int split = PREFILL ? 0 : blockIdx.x;
  const int G = a.Hq / g.H;
  const int npass = (G + GP - 1) / GP;
  const int h = PREFILL ? (int)blockIdx.y / G : (int)blockIdx.y / npass;
  const int pass = PREFILL ? 0 : blockIdx.y - h * npass;
  const int i = PREFILL ? (a.nseq == 1 ? 0 : (int)blockIdx.z) : (int)blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ctx = min(a.context_lens[i], a.max_blocks * g.bs);   // never walk past the block table
  // prefill: queries t0 .. t0+ng-1 of q-head blockIdx.y, batch row blockIdx.z; query t sits at position qstart + t
  const int t0 = PREFILL ? (int)blockIdx.x * GP : 0;
  const int qstart = ctx - a.Sq;
  const int c0 = PREFILL ? 0 : split * a.chunk;
  const int g0 = PREFILL ? 0 : pass * GP;   // first q-head (within the group) of this pass
  const int ng = PREFILL ? min(GP, a.Sq - t0) : min(GP, G - g0);
  const int c1 = PREFILL ? max(0, min(ctx, qstart + t0 + ng)) : min(ctx, c0 + a.chunk);   // prefill: tokens the block's last query sees
  const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;
  const int rowbytes = GOLAY ? g.cw * 4 : g.D;
  const int64_t page_bytes = (int64_t)g.bs * rowbytes;

  // element index of value v (0..VPC-1) of chunk ch for this lane, or -1
  auto elem = [&](int ch, int v) -> int {
    const int e = GOLAY ? 3 * (ch * 32 + lane) + v : ch * 128 + lane * 4 + v;
    return e < g.D ? e : -1;
  };

  // q for this lane: qreg[gi][ch*VPC+v], pre-multiplied by sm_scale
  float qreg[GP][EPL];
#pragma unroll
  for (int gi = 0; gi < GP; ++gi)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int v = 0; v < VPC; ++v) {
        const int e = elem(ch, v);
        const int64_t qrow = PREFILL ? ((int64_t)blockIdx.z * a.Hq + blockIdx.y) * a.Sq + t0 + gi : (int64_t)i * a.Hq + (int64_t)h * G + g0 + gi;
        qreg[gi][ch * VPC + v] = (gi < ng && e >= 0) ? ldq<TQ>(a.q, qrow * g.D + e) * a.sm_scale : 0.0f;
      }
  if (GOLAY) __syncthreads();

  float m[GP], l[GP], acc[GP][EPL];
#pragma unroll
  for (int gi = 0; gi < GP; ++gi) {
    m[gi] = -INFINITY; l[gi] = 0.0f;
#pragma unroll
    for (int x = 0; x < EPL; ++x) acc[gi][x] = 0.0f;
  }
  uint32_t ncorr = 0, ndet = 0;
  // prefill: the last query block of the first q-head of a group (first batch row of a shared context) sees every token once
  const bool do_count = PREFILL ? a.count_errors && (int)blockIdx.x == (int)gridDim.x - 1 && (int)blockIdx.y % G == 0 && (a.nseq > 1 || blockIdx.z == 0)
                                : a.count_errors && pass == 0;

  // The codeword words this lane owns of a token's K and V rows + its scales, loaded one loop iteration ahead of their use: the
  // block-table entry, the row and the scales are three dependent global loads, and a warp walks its tokens serially, so
  // without the prefetch every token pays that round trip (short prefill launches are bound by exactly this chain).
  struct Raw { uint32_t k[NCH], v[NCH]; float ksc, vsc; int blk; };
  auto load_raw = [&](int tok) -> Raw {
    Raw r;
    const int lb = tok / g.bs, slot = tok - lb * g.bs;
    r.blk = bt[lb];
    r.ksc = r.vsc = 0.0f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) r.k[ch] = r.v[ch] = 0u;
    if (r.blk < 0) return r;
    const int64_t pg = page_index(g, r.blk, a.layer, h);
    const int64_t off = pg * page_bytes + (int64_t)slot * rowbytes;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int w = GOLAY ? ch * 32 + lane : ch * 32 + lane;           // word index inside the row
      const bool in = GOLAY ? w < g.cw : w * 4 < g.D;
      if (in) {
        r.k[ch] = __ldg(reinterpret_cast<const uint32_t*>(g.k_cache + off) + w);
        r.v[ch] = __ldg(reinterpret_cast<const uint32_t*>(g.v_cache + off) + w);
      }
    }
    r.ksc = __ldg(g.k_scales + pg * g.bs + slot);
    r.vsc = a.v_use_k_scales ? r.ksc : __ldg(g.v_scales + pg * g.bs + slot);
    return r;
  };
  // decode the values this lane owns of a row from its loaded words
  auto decode_row = [&](const uint32_t (&raw)[NCH], float (&val)[EPL], uint32_t (&dbl)[NCH], bool count) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      dbl[ch] = 0;
      if (GOLAY) {
        const int j = ch * 32 + lane;
        uint32_t d = 0;
        if (j < g.cw) {
          uint32_t cnt;
          d = golay_decode1(s_lo, s_hi, s_syn, raw[ch], cnt);
          if (count) { if (cnt == 4u) ++ndet; else ncorr += cnt; }
        }
        val[ch * 3] = (float)(int)(d & 0xF) - 8.0f;
        val[ch * 3 + 1] = (float)(int)((d >> 4) & 0xF) - 8.0f;
        val[ch * 3 + 2] = (float)(int)((d >> 8) & 0xF) - 8.0f;
      } else {
        const int e0 = ch * 128 + lane * 4;
        uint32_t nib = 0x08080808u;
        if (e0 < g.D) {
          const uint32_t w = raw[ch];
          if (CODEC == ECCQ_CODEC_H84) {
            const Dec4 d = h84_decode4(w);
            nib = d.nib; dbl[ch] = h84_double4(d);
            if (count) { ncorr += __popc(h84_single4(d)); ndet += __popc(dbl[ch]); }
          } else if (CODEC == ECCQ_CODEC_H74) {
            const Dec4 d = h74_decode4(w);
            nib = d.nib;
            if (count) ncorr += __popc(d.nz);
          } else {
            nib = w;
          }
        }
        // stash the nibble word in val[] bit patterns until interpolation is resolved
        val[ch * 4] = __uint_as_float(nib);
      }
    }
  };
  // Hamming: nibble word of a neighbour token for chunk ch (pre-interpolation decode)
  auto neighbour_word = [&](const uint8_t* cache, int tok, int ch) -> uint32_t {
    const int lb = tok / g.bs, slot = tok - lb * g.bs;
    const int blk = bt[lb];
    if (blk < 0) return 0u;
    const uint8_t* row = cache + page_index(g, blk, a.layer, h) * page_bytes + (int64_t)slot * rowbytes;
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row + ch * 128 + lane * 4));
    return h84_decode4(w).nib;
  };
  auto finish_row = [&](const uint8_t* cache, int tok, float (&val)[EPL], const uint32_t (&dbl)[NCH]) {
    if constexpr (!GOLAY) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t nib = __float_as_uint(val[ch * 4]);
      if (CODEC == ECCQ_CODEC_H84 && a.interp && dbl[ch]) {
        const uint32_t left = tok > 0 ? neighbour_word(cache, tok - 1, ch) : nib;
        const uint32_t right = tok + 1 < ctx ? neighbour_word(cache, tok + 1, ch) : nib;
        nib = interp4(nib, left, right, dbl[ch]);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) val[ch * 4 + v] = byte_to_float_m8(nib, v);
    }
    }
  };

  Raw nxt;
  if (c0 + warp < c1) nxt = load_raw(c0 + warp);
  for (int tok = c0 + warp; tok < c1; tok += kAttnWarps) {
    const Raw cur = nxt;
    if (tok + kAttnWarps < c1) nxt = load_raw(tok + kAttnWarps);
    if (cur.blk < 0) continue;          // unmapped block: its tokens take no part (the reference's gather skips it)
    float kv[EPL];
    uint32_t dbl[NCH];
    decode_row(cur.k, kv, dbl, do_count);
    finish_row(g.k_cache, tok, kv, dbl);
    const float ksc = cur.ksc, vsc = cur.vsc;
    float s[GP];
#pragma unroll
    for (int gi = 0; gi < GP; ++gi) {
      float d = 0.0f;
#pragma unroll
      for (int x = 0; x < EPL; ++x) d = fmaf(qreg[gi][x], kv[x], d);
      s[gi] = warp_sum(d) * ksc;
    }
    float vv[EPL];
    decode_row(cur.v, vv, dbl, do_count);
    finish_row(g.v_cache, tok, vv, dbl);
#pragma unroll
    for (int gi = 0; gi < GP; ++gi) {
      if (PREFILL && tok > qstart + t0 + gi) continue;   // causal mask (warp-uniform)
      const float mn = fmaxf(m[gi], s[gi]);
      const float alpha = __expf(m[gi] - mn);   // exp(-inf) = 0 on the first token
      const float pw = __expf(s[gi] - mn);
      l[gi] = l[gi] * alpha + pw;
      const float pv = pw * vsc;
#pragma unroll
      for (int x = 0; x < EPL; ++x) acc[gi][x] = fmaf(acc[gi][x], alpha, pv * vv[x]);
      m[gi] = mn;
    }
  }

  // ---- merge the warps of this CTA through shared memory, write the chunk partial
  constexpr int RS = EPL * 32 + 2;
#pragma unroll
  for (int gi = 0; gi < GP; ++gi) {
    float* dst = s_red + ((size_t)warp * GP + gi) * RS;
#pragma unroll
    for (int x = 0; x < EPL; ++x) dst[x * 32 + lane] = acc[gi][x];
    if (lane == 0) { dst[EPL * 32] = m[gi]; dst[EPL * 32 + 1] = l[gi]; }
  }
  __syncthreads();
  for (int gi = warp; gi < ng; gi += kAttnWarps) {
    float M = -INFINITY;
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, s_red[((size_t)w * GP + gi) * RS + EPL * 32]);
    float Lsum = 0.0f, o[EPL];
#pragma unroll
    for (int x = 0; x < EPL; ++x) o[x] = 0.0f;
    for (int w = 0; w < kAttnWarps; ++w) {
      const float* src = s_red + ((size_t)w * GP + gi) * RS;
      const float mw = src[EPL * 32];
      const float f = (mw == -INFINITY) ? 0.0f : __expf(mw - M);
      Lsum += src[EPL * 32 + 1] * f;
#pragma unroll
      for (int x = 0; x < EPL; ++x) o[x] += src[x * 32 + lane] * f;
    }
    if constexpr (PREFILL) {
      TQ* orow = reinterpret_cast<TQ*>(a.out) + (((int64_t)blockIdx.z * a.Hq + blockIdx.y) * a.Sq + t0 + gi) * g.D;
      const float inv = Lsum > 0.0f ? 1.0f / Lsum : 0.0f;      // no visible token (empty context): zeros
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int v = 0; v < VPC; ++v) {
          const int e = elem(ch, v);
          if (e >= 0) orow[e] = Cvt<TQ>::to(o[ch * VPC + v] * inv);
        }
      continue;
    }
    const int qh = h * G + g0 + gi;
    float* wsrow = a.ws + (((int64_t)i * a.Hq + qh) * a.nsplit + split) * (g.D + 2);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int v = 0; v < VPC; ++v) {
        const int e = elem(ch, v);
        if (e >= 0) wsrow[e] = o[ch * VPC + v];
      }
    if (lane == 0) { wsrow[g.D] = M; wsrow[g.D + 1] = Lsum; }
  }
  if (a.count_errors) {
    const int idx[3] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED, ECCQ_CTR_INTERPOLATED};
    const uint32_t val[3] = {ncorr, ndet, a.interp ? ndet : 0u};   // every double-detected value is interpolated (ecc_shim.py:1038-1059)
    block_accumulate<3>(counters, idx, val);
  }
}

// out[i, qh, :] = sum_s acc_s * exp(m_s - M) / sum_s l_s * exp(m_s - M); empty context -> zeros
template <typename TQ>
__global__ void attn_merge_kernel(const float* __restrict__ ws, TQ* __restrict__ out, int Hq, int D, int nsplit) {
  const int qh = blockIdx.x, i = blockIdx.y;
  const float* base = ws + ((int64_t)i * Hq + qh) * nsplit * (D + 2);
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, base[(int64_t)s * (D + 2) + D]);
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    float num = 0.0f, den = 0.0f;
    for (int s = 0; s < nsplit; ++s) {
      const float* r = base + (int64_t)s * (D + 2);
      const float f = (r[D] == -INFINITY) ? 0.0f : __expf(r[D] - M);
      num += r[e] * f;
      den += r[D + 1] * f;
    }
    out[((int64_t)i * Hq + qh) * D + e] = Cvt<TQ>::to(den > 0.0f ? num / den : 0.0f);
  }
}

static inline void attn_plan(int max_context, int& nsplit, int& chunk) {
  chunk = 512;
  nsplit = (max_context + chunk - 1) / chunk;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 128) { nsplit = 128; chunk = ((max_context + nsplit - 1) / nsplit + 15) / 16 * 16; }
}

template <typename TQ, int CODEC, int GP, int NCH>
static int launch_attn(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  constexpr int VPC = CODEC == ECCQ_CODEC_GOLAY ? 3 : 4;
  const size_t smem = (CODEC == ECCQ_CODEC_GOLAY ? 4096 * 4 + 256 : 0) + sizeof(float) * kAttnWarps * GP * (VPC * NCH * 32 + 2);
  auto kern = attn_partial_kernel<TQ, CODEC, GP, NCH>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
  }
  const int G = a.Hq / g.H;
  const int npass = (G + GP - 1) / GP;
  dim3 grid(a.nsplit, g.H * npass, a.nseq);
  kern<<<grid, kAttnThreads, smem, st>>>(g, a, counters);
  return ECCQ_OK;
}

template <typename TQ, int CODEC>
static int dispatch_attn(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  const int G = a.Hq / g.H;
  const int per_chunk = CODEC == ECCQ_CODEC_GOLAY ? 96 : 128;
  const int nch = (g.D + per_chunk - 1) / per_chunk;
  if (nch > (CODEC == ECCQ_CODEC_GOLAY ? 3 : 2)) return fail(ECCQ_EUNSUPPORTED, "paged_decode_attention: head_dim too large (max 256)");
#define ECCQ_ATTN_CASE(GPV)                                                                                        \
  if (CODEC == ECCQ_CODEC_GOLAY && nch == 3) return launch_attn<TQ, CODEC, GPV, (CODEC == ECCQ_CODEC_GOLAY ? 3 : 2)>(g, a, counters, st); \
  return nch == 1 ? launch_attn<TQ, CODEC, GPV, 1>(g, a, counters, st) : launch_attn<TQ, CODEC, GPV, 2>(g, a, counters, st)
  if (G == 1) { ECCQ_ATTN_CASE(1); }
  if (G == 2) { ECCQ_ATTN_CASE(2); }
  if (G <= 4) { ECCQ_ATTN_CASE(4); }
  ECCQ_ATTN_CASE(8);
#undef ECCQ_ATTN_CASE
}

template <typename TQ>
static int try_attn_mma(const CacheGeom& g, AttnArgs& a, uint64_t* counters, cudaStream_t st);  // eccq_attention_mma.cuh (may coarsen a.chunk / a.nsplit)

template <typename TQ> struct DtypeCode;
template <> struct DtypeCode<float> { static constexpr int value = ECCQ_F32; };
template <> struct DtypeCode<__half> { static constexpr int value = ECCQ_F16; };
template <> struct DtypeCode<__nv_bfloat16> { static constexpr int value = ECCQ_BF16; };

template <typename TQ>
static int run_attn(const CacheGeom& g, const AttnArgs& a_in, uint64_t* counters, cudaStream_t st) {
  AttnArgs a = a_in;
  int rc;
  if (!a.force_general) {
    // Blackwell-native tile kernel first (tcgen05 / TMEM / bulk TMA, eccq_attention_tc.cu); < 0 = launch error
    const int took = try_attn_tc(g, a, DtypeCode<TQ>::value, counters, st);
    if (took < 0) return took;
    if (took) {
      dim3 mg1(a.Hq, a.nseq);
      attn_merge_kernel<TQ><<<mg1, 128, 0, st>>>(a.ws, reinterpret_cast<TQ*>(a.out), a.Hq, g.D, a.nsplit);
      return ECCQ_OK;
    }
  }
  const int took_mma = a.force_general ? 0 : try_attn_mma<TQ>(g, a, counters, st);
  if (took_mma < 0) return took_mma;
  if (took_mma) {
    dim3 mg0(a.Hq, a.nseq);
    attn_merge_kernel<TQ><<<mg0, 128, 0, st>>>(a.ws, reinterpret_cast<TQ*>(a.out), a.Hq, g.D, a.nsplit);
    return ECCQ_OK;
  }
  switch (g.codec) {
    case ECCQ_CODEC_H74: rc = dispatch_attn<TQ, ECCQ_CODEC_H74>(g, a, counters, st); break;
    case ECCQ_CODEC_H84: rc = dispatch_attn<TQ, ECCQ_CODEC_H84>(g, a, counters, st); break;
    case ECCQ_CODEC_GOLAY: rc = dispatch_attn<TQ, ECCQ_CODEC_GOLAY>(g, a, counters, st); break;
    default: rc = dispatch_attn<TQ, ECCQ_CODEC_INT4>(g, a, counters, st); break;
  }
  if (rc) return rc;
  dim3 mg(a.Hq, a.nseq);
  attn_merge_kernel<TQ><<<mg, 128, 0, st>>>(a.ws, reinterpret_cast<TQ*>(a.out), a.Hq, g.D, a.nsplit);
  return ECCQ_OK;
}

// ---- ECC-fused causal prefill, tiled (head_dim 32 / 64 / 128): the kernel ECCBackend.attend uses for q_len > 1 -----------------------
// Replaces gather + decode + interpolation + fp32 K/V materialisation + repeat_interleave + SDPA (kv_cache/ecc_shim.py:979-1089,
// :1138-1164).  CTA = 32 query rows of ONE kv-head (row = (query position, q-head of the GQA group): the group shares every decoded
// K/V row), 8 warps x 4 rows.  Per 32-token tile: (1) the warps decode 4 token rows each straight from the codeword pages -- the
// row prologue of attn_partial_kernel (all codecs, temporal interpolation, counters) -- into shared memory as nibbles, one byte per
// value, rows padded to D + 4 bytes; (2) QK^T with lane = token: the lane walks ITS row's nibbles (conflict-free through the
// padding), the query values are 16-byte shared-memory broadcasts, fp32 FMAs, no shuffles; (3) online softmax per query row in base
// 2, one max / sum reduction over the lanes per row and tile; (4) PV with lane = dims: the weights go through shared memory, every
// lane accumulates its D/32 dims.  ~13 warp-instructions per (query, token) pair against ~77 for the warp-per-token kernel above.
constexpr int kPfRows = 32;                        // query rows per CTA
constexpr int kPfRPW = kPfRows / kAttnWarps;       // query rows per warp
constexpr int kPfTile = 32;                        // tokens per tile (= lanes)
#ifndef ECCQ_PF_MINBLOCKS
#define ECCQ_PF_MINBLOCKS 2   // resident CTAs per SM the prefill tile kernel is compiled for (measured: 1 -> 0.28, 2 -> 0.20, 3 -> 0.21, 4 -> 0.24 ms on the GPT-2 1 024-token layer)
#endif
template <typename TQ, int CODEC, int DPL>
__global__ void __launch_bounds__(kAttnThreads, ECCQ_PF_MINBLOCKS) attn_prefill_tile_kernel(CacheGeom g, AttnArgs a, uint64_t* counters) {
  constexpr bool GOLAY = CODEC == ECCQ_CODEC_GOLAY;
  constexpr int D = 32 * DPL;
  constexpr int KS = D + 4;                        // padded row stride (bytes) of the nibble tiles
  constexpr int CW = (D + 2) / 3;                  // Golay codewords per row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  uint16_t* s_lo = nullptr; uint16_t* s_hi = nullptr; int32_t* s_syn = nullptr;
  if (GOLAY) {
    s_syn = reinterpret_cast<int32_t*>(p); p += 4096 * 4;
    s_lo = reinterpret_cast<uint16_t*>(p); p += 128;
    s_hi = reinterpret_cast<uint16_t*>(p); p += 128;
    golay_load_parity(s_lo, s_hi);
    golay_load_syndrome(s_syn);
  }
  float* Qs = reinterpret_cast<float*>(p); p += kPfRows * D * 4;                 // [row][D], pre-multiplied by sm_scale * log2(e)
  float* Ps = reinterpret_cast<float*>(p); p += kAttnWarps * kPfRPW * kPfTile * 4;   // [warp][row][token]: weights x v scale
  float* s_ks = reinterpret_cast<float*>(p); p += kPfTile * 4;
  float* s_vs = reinterpret_cast<float*>(p); p += kPfTile * 4;
  uint8_t* Kn = p; p += kPfTile * KS;                                            // nibbles of the tile's K rows
  uint8_t* Vn = p;

  const int G = a.Hq / g.H;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = a.nseq == 1 ? 0 : b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ctx = min(a.context_lens[i], a.max_blocks * g.bs);
  const int qstart = ctx - a.Sq;                   // query t sits at position qstart + t
  const int nrows = a.Sq * G;
  const int nqblk = (nrows + kPfRows - 1) / kPfRows;
  const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;
  const int rowbytes = GOLAY ? g.cw * 4 : D;
  const int64_t page_bytes = (int64_t)g.bs * rowbytes;
  uint32_t ncorr = 0, ndet = 0;
  // Causal load balance: a CTA takes query block n-1-x and then block x -- together always n+1 tiles of work -- instead of one block
  // whose cost grows with its position (the last block's CTA used to run 32 x as long as the first one's and set the kernel time).
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
  const int qblk = pass == 0 ? nqblk - 1 - (int)blockIdx.x : (int)blockIdx.x;
  if (pass == 1 && qblk >= nqblk - 1 - (int)blockIdx.x) break;   // odd count: the middle block was pass 0
  if (pass == 1) __syncthreads();                  // every warp is done with the first block's queries in shared memory
  const int row0 = qblk * kPfRows;                 // first (position, head) row of this block inside the kv-head: row = pos * G + head
  const int pos_last = min(a.Sq - 1, (row0 + kPfRows - 1) / G);
  const int tmax = max(0, min(ctx, qstart + pos_last + 1));   // tokens [0, tmax) are visible to some query of the CTA
  const bool do_count = a.count_errors && qblk == nqblk - 1 && (a.nseq > 1 || b == 0);   // the last block sees every token

  // ---- queries of the block -> shared memory
  for (int e = threadIdx.x; e < kPfRows * D; e += kAttnThreads) {
    const int r = e / D, d = e - r * D;
    const int rg = row0 + r;
    float v = 0.0f;
    if (rg < nrows) {
      const int pos = rg / G, hq = h * G + (rg - pos * G);
      v = ldq<TQ>(a.q, (((int64_t)b * a.Hq + hq) * a.Sq + pos) * D + d) * (a.sm_scale * 1.4426950408889634f);
    }
    Qs[e] = v;
  }
  // this warp's query rows: position of each (for the causal mask), running state; lane = dims lane*DPL .. +DPL-1 for the output
  int qpos[kPfRPW];
  float m[kPfRPW], l[kPfRPW], acc[kPfRPW][DPL];
#pragma unroll
  for (int q = 0; q < kPfRPW; ++q) {
    const int rg = row0 + warp * kPfRPW + q;
    qpos[q] = rg < nrows ? qstart + rg / G : -1;   // -1: no such row, sees nothing
    m[q] = -INFINITY; l[q] = 0.0f;
#pragma unroll
    for (int x = 0; x < DPL; ++x) acc[q][x] = 0.0f;
  }

  // nibble word (4 values) of a neighbour token's K or V row for the interpolation of a double-detected value (rare)
  auto neighbour_word = [&](const uint8_t* cache, int tok, int w) -> uint32_t {
    const int lb = tok / g.bs, slot = tok - lb * g.bs;
    const int blk = bt[lb];
    if (blk < 0) return 0u;
    const uint8_t* row = cache + page_index(g, blk, a.layer, h) * page_bytes + (int64_t)slot * rowbytes;
    return h84_decode4(__ldg(reinterpret_cast<const uint32_t*>(row) + w)).nib;
  };

  // The codeword words of the tile's rows this warp decodes (4 rows x K, V) and their scales are loaded one tile AHEAD, while the
  // current tile is being multiplied: block table -> row -> words is a chain of dependent global loads, and a CTA walks its tiles
  // serially, so without the prefetch every tile paid that round trip (0.73 -> 0.26 ms came from the tiling, the rest from this).
  constexpr int RR = kPfTile / kAttnWarps;         // token rows per warp and tile
  constexpr int NW = GOLAY ? (CW + 31) / 32 : 1;   // words per lane, row and tensor
  struct TileRaw { uint32_t k[RR][NW], v[RR][NW]; float ks[RR], vs[RR]; int blk[RR]; };
  auto load_tile = [&](int tile0, TileRaw& r) {
#pragma unroll
    for (int rr = 0; rr < RR; ++rr) {
      const int tok = tile0 + warp * RR + rr;
      r.blk[rr] = -1; r.ks[rr] = 0.0f; r.vs[rr] = 0.0f;
#pragma unroll
      for (int x = 0; x < NW; ++x) { r.k[rr][x] = 0u; r.v[rr][x] = 0u; }
      if (tok < tmax) {
        const int lb = tok / g.bs, slot = tok - lb * g.bs;
        const int blk = bt[lb];
        r.blk[rr] = blk;
        if (blk >= 0) {
          const int64_t pg = page_index(g, blk, a.layer, h);
          const int64_t off = pg * page_bytes + (int64_t)slot * rowbytes;
          r.ks[rr] = __ldg(g.k_scales + pg * g.bs + slot);
          r.vs[rr] = a.v_use_k_scales ? r.ks[rr] : __ldg(g.v_scales + pg * g.bs + slot);
#pragma unroll
          for (int x = 0; x < NW; ++x) {
            const int w = x * 32 + lane;
            if (GOLAY ? w < CW : w < D / 4) {
              r.k[rr][x] = __ldg(reinterpret_cast<const uint32_t*>(g.k_cache + off) + w);
              r.v[rr][x] = __ldg(reinterpret_cast<const uint32_t*>(g.v_cache + off) + w);
            }
          }
        }
      }
    }
  };
  TileRaw cur;
  if (tmax > 0) load_tile(0, cur);
  for (int tile0 = 0; tile0 < tmax; tile0 += kPfTile) {
    __syncthreads();                               // the previous tile has been consumed (and Qs is written, first time round)
    // ---- (1) decode: warp w owns token rows 4w .. 4w+3 of the tile
#pragma unroll
    for (int rr = 0; rr < RR; ++rr) {
      const int trow = warp * RR + rr;
      const int tok = tile0 + trow;
      uint8_t* kdst = Kn + trow * KS;
      uint8_t* vdst = Vn + trow * KS;
      if (cur.blk[rr] < 0) {                       // past the visible range or unmapped: the token takes no part
        if (lane == 0) { s_ks[trow] = 0.0f; s_vs[trow] = -1.0f; }   // v scale < 0 marks the row invalid
        for (int w = lane; w < D / 4; w += 32) { reinterpret_cast<uint32_t*>(kdst)[w] = 0x08080808u; reinterpret_cast<uint32_t*>(vdst)[w] = 0x08080808u; }
        continue;
      }
      if (lane == 0) { s_ks[trow] = cur.ks[rr]; s_vs[trow] = cur.vs[rr]; }
      if (GOLAY) {
#pragma unroll
        for (int x = 0; x < NW; ++x) {
          const int j = x * 32 + lane;
          if (j < CW) {
#pragma unroll
            for (int kv = 0; kv < 2; ++kv) {
              uint32_t cnt;
              const uint32_t d = golay_decode1(s_lo, s_hi, s_syn, kv ? cur.v[rr][x] : cur.k[rr][x], cnt);
              if (do_count) { if (cnt == 4u) ++ndet; else ncorr += cnt; }
              uint8_t* dst = kv ? vdst : kdst;
#pragma unroll
              for (int v = 0; v < 3; ++v)
                if (3 * j + v < D) dst[3 * j + v] = (uint8_t)((d >> (4 * v)) & 0xFu);
            }
          }
        }
      } else {
        const int w = lane;
        if (w < D / 4) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            const uint8_t* cache = kv ? g.v_cache : g.k_cache;
            const uint32_t cwv = kv ? cur.v[rr][0] : cur.k[rr][0];
            uint32_t nib = cwv, dbl = 0;
            if (CODEC == ECCQ_CODEC_H84) {
              const Dec4 d = h84_decode4(cwv);
              nib = d.nib; dbl = h84_double4(d);
              if (do_count) { ncorr += __popc(h84_single4(d)); ndet += __popc(dbl); }
              if (a.interp && dbl) {
                const uint32_t left = tok > 0 ? neighbour_word(cache, tok - 1, w) : nib;
                const uint32_t right = tok + 1 < ctx ? neighbour_word(cache, tok + 1, w) : nib;
                nib = interp4(nib, left, right, dbl);
              }
            } else if (CODEC == ECCQ_CODEC_H74) {
              const Dec4 d = h74_decode4(cwv);
              nib = d.nib;
              if (do_count) ncorr += __popc(d.nz);
            }
            reinterpret_cast<uint32_t*>(kv ? vdst : kdst)[w] = nib & 0x0F0F0F0Fu;
          }
        }
      }
    }
    if (tile0 + kPfTile < tmax) load_tile(tile0 + kPfTile, cur);   // in flight during the products below
    __syncthreads();
    // ---- (2) QK^T, lane = token of the tile
    const int tok = tile0 + lane;
    const float ksc = s_ks[lane], vsc = s_vs[lane];
    const bool tvalid = vsc >= 0.0f;
    float sc[kPfRPW];
#pragma unroll
    for (int q = 0; q < kPfRPW; ++q) sc[q] = 0.0f;
    const uint32_t* krow = reinterpret_cast<const uint32_t*>(Kn + lane * KS);
#pragma unroll 4
    for (int d4 = 0; d4 < D / 4; ++d4) {
      const uint32_t kw = krow[d4];
      const float k0 = byte_to_float_m8(kw, 0), k1 = byte_to_float_m8(kw, 1), k2 = byte_to_float_m8(kw, 2), k3 = byte_to_float_m8(kw, 3);
#pragma unroll
      for (int q = 0; q < kPfRPW; ++q) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + (warp * kPfRPW + q) * D + d4 * 4);
        sc[q] = fmaf(qv.x, k0, fmaf(qv.y, k1, fmaf(qv.z, k2, fmaf(qv.w, k3, sc[q]))));
      }
    }
    // ---- (3) online softmax per query row (base 2), weights x v scale -> shared memory
#pragma unroll
    for (int q = 0; q < kPfRPW; ++q) {
      const bool vis = tvalid && tok <= qpos[q];
      const float sv = vis ? sc[q] * ksc : -INFINITY;
      float mt = sv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xFFFFFFFFu, mt, o));
      const float mn = fmaxf(m[q], mt);
      float pw = 0.0f, alpha = 1.0f;
      if (mn > -INFINITY) {                        // warp-uniform
        alpha = exp2f(m[q] - mn);                  // exp2(-inf) = 0 on the row's first visible tile
        pw = vis ? exp2f(sv - mn) : 0.0f;
        m[q] = mn;
      }
      l[q] = l[q] * alpha + warp_sum(pw);
#pragma unroll
      for (int x = 0; x < DPL; ++x) acc[q][x] *= alpha;
      Ps[(warp * kPfRPW + q) * kPfTile + lane] = tvalid ? pw * vsc : 0.0f;
    }
    __syncwarp();
    // ---- (4) PV, lane = dims lane*DPL .. lane*DPL+DPL-1
#pragma unroll 2
    for (int t4 = 0; t4 < kPfTile / 4; ++t4) {
      float4 pv[kPfRPW];
#pragma unroll
      for (int q = 0; q < kPfRPW; ++q) pv[q] = *reinterpret_cast<const float4*>(Ps + (warp * kPfRPW + q) * kPfTile + t4 * 4);
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const uint8_t* vrow = Vn + (t4 * 4 + tt) * KS + lane * DPL;
        uint32_t vw;
        if (DPL == 4) vw = *reinterpret_cast<const uint32_t*>(vrow);
        else if (DPL == 2) vw = *reinterpret_cast<const uint16_t*>(vrow);
        else vw = *vrow;
        float vf[DPL];
#pragma unroll
        for (int x = 0; x < DPL; ++x) vf[x] = byte_to_float_m8(vw, x);
#pragma unroll
        for (int q = 0; q < kPfRPW; ++q) {
          const float pq = tt == 0 ? pv[q].x : tt == 1 ? pv[q].y : tt == 2 ? pv[q].z : pv[q].w;
#pragma unroll
          for (int x = 0; x < DPL; ++x) acc[q][x] = fmaf(pq, vf[x], acc[q][x]);
        }
      }
    }
    __syncwarp();                                  // Ps of this warp is rewritten by the next tile
  }
  // ---- output rows of this warp: out[b, hq, pos, lane*DPL ..]
#pragma unroll
  for (int q = 0; q < kPfRPW; ++q) {
    const int rg = row0 + warp * kPfRPW + q;
    if (rg >= nrows) continue;
    const int pos = rg / G, hq = h * G + (rg - pos * G);
    TQ* orow = reinterpret_cast<TQ*>(a.out) + (((int64_t)b * a.Hq + hq) * a.Sq + pos) * D + lane * DPL;
    const float inv = l[q] > 0.0f ? 1.0f / l[q] : 0.0f;   // no visible token: zeros
#pragma unroll
    for (int x = 0; x < DPL; ++x) orow[x] = Cvt<TQ>::to(acc[q][x] * inv);
  }
  }   // pass
  if (a.count_errors) {
    const int idx[3] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED, ECCQ_CTR_INTERPOLATED};
    const uint32_t val[3] = {ncorr, ndet, a.interp ? ndet : 0u};
    block_accumulate<3>(counters, idx, val);
  }
}

template <typename TQ, int CODEC, int DPL>
static int launch_prefill_tile(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  constexpr int D = 32 * DPL;
  const size_t smem = (CODEC == ECCQ_CODEC_GOLAY ? 4096 * 4 + 256 : 0) + (size_t)kPfRows * D * 4 + kAttnWarps * kPfRPW * kPfTile * 4 + 2 * kPfTile * 4 +
                      2 * kPfTile * (D + 4);
  auto kern = attn_prefill_tile_kernel<TQ, CODEC, DPL>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
  }
  const int G = a.Hq / g.H;
  const int nqblk = (a.Sq * G + kPfRows - 1) / kPfRows;
  dim3 grid((nqblk + 1) / 2, g.H, a.qbatch);      // a CTA takes query blocks n-1-x and x
  kern<<<grid, kAttnThreads, smem, st>>>(g, a, counters);
  return ECCQ_OK;
}
// 1 = launched, 0 = geometry outside the tiled kernel (the warp-per-token kernel below serves it)
template <typename TQ, int CODEC>
static int try_prefill_tile(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  static const int env = [] { const char* v = std::getenv("ECCQ_PREFILL_TILE"); return v && *v ? std::atoi(v) : 1; }();
  if (!env || g.H > 65535) return 0;
  int rc;
  if (g.D == 128) rc = launch_prefill_tile<TQ, CODEC, 4>(g, a, counters, st);
  else if (g.D == 64) rc = launch_prefill_tile<TQ, CODEC, 2>(g, a, counters, st);
  else if (g.D == 32) rc = launch_prefill_tile<TQ, CODEC, 1>(g, a, counters, st);
  else return 0;
  return rc ? rc : 1;
}

// ---- prefill (q_len > 1): the PREFILL instantiation of attn_partial_kernel, GP = queries per CTA
template <typename TQ, int CODEC, int GP, int NCH>
static int launch_prefill(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  constexpr int VPC = CODEC == ECCQ_CODEC_GOLAY ? 3 : 4;
  const size_t smem = (CODEC == ECCQ_CODEC_GOLAY ? 4096 * 4 + 256 : 0) + sizeof(float) * kAttnWarps * GP * (VPC * NCH * 32 + 2);
  auto kern = attn_partial_kernel<TQ, CODEC, GP, NCH, true>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
  }
  dim3 grid((a.Sq + GP - 1) / GP, a.Hq, a.qbatch);
  kern<<<grid, kAttnThreads, smem, st>>>(g, a, counters);
  return ECCQ_OK;
}
template <typename TQ, int CODEC>
static int dispatch_prefill(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  const int took = try_prefill_tile<TQ, CODEC>(g, a, counters, st);
  if (took < 0) return took;
  if (took) return ECCQ_OK;
  const int per_chunk = CODEC == ECCQ_CODEC_GOLAY ? 96 : 128;
  const int nch = (g.D + per_chunk - 1) / per_chunk;
  if (nch > (CODEC == ECCQ_CODEC_GOLAY ? 3 : 2)) return fail(ECCQ_EUNSUPPORTED, "paged_prefill_attention: head_dim too large (max 256)");
  if (nch == 1) return launch_prefill<TQ, CODEC, 8, 1>(g, a, counters, st);
  if (nch == 2) return launch_prefill<TQ, CODEC, 4, 2>(g, a, counters, st);
  return launch_prefill<TQ, CODEC, 4, (CODEC == ECCQ_CODEC_GOLAY ? 3 : 2)>(g, a, counters, st);
}
template <typename TQ>
static int run_prefill(const CacheGeom& g, const AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  switch (g.codec) {
    case ECCQ_CODEC_H74: return dispatch_prefill<TQ, ECCQ_CODEC_H74>(g, a, counters, st);
    case ECCQ_CODEC_H84: return dispatch_prefill<TQ, ECCQ_CODEC_H84>(g, a, counters, st);
    case ECCQ_CODEC_GOLAY: return dispatch_prefill<TQ, ECCQ_CODEC_GOLAY>(g, a, counters, st);
    default: return dispatch_prefill<TQ, ECCQ_CODEC_INT4>(g, a, counters, st);
  }
}

}  // namespace eccq

#include "eccq_attention_mma.cuh"

using namespace eccq;

ECCQ_DEFINE_TU_TABLES(attention)

extern "C" int64_t eccq_attention_workspace_bytes(int nseq, int Hq, int D, int max_context) {
  if (nseq <= 0 || Hq <= 0 || D <= 0) return 0;
  int nsplit, chunk;
  attn_plan(max_context, nsplit, chunk);
  return (int64_t)nseq * Hq * nsplit * (D + 2) * (int64_t)sizeof(float);
}

extern "C" int eccq_paged_decode_attention(const eccq_cache_desc* c, const void* q, void* out, int q_dtype,
                                           const int32_t* block_table, int max_blocks, const int32_t* context_lens, int nseq,
                                           int Hq, int layer, float sm_scale, int use_interpolation, int v_use_k_scales,
                                           int count_errors, void* workspace, int64_t workspace_bytes, uint64_t* counters,
                                           void* stream) {
  CacheGeom g;
  if (!make_geom(c, g)) return fail(ECCQ_EINVAL, "paged_decode_attention: bad cache descriptor");
  if (nseq < 0 || Hq <= 0 || Hq % g.H != 0 || layer < 0 || layer >= g.L || max_blocks <= 0 || !block_table || !context_lens)
    return fail(ECCQ_EINVAL, "paged_decode_attention: bad arguments");
  if (nseq == 0) return ECCQ_OK;
  if (!q || !out || !workspace) return fail(ECCQ_EINVAL, "paged_decode_attention: null pointer");
  if (g.codec != ECCQ_CODEC_GOLAY && (g.D % 4 != 0 || (reinterpret_cast<uintptr_t>(g.k_cache) & 3) || (reinterpret_cast<uintptr_t>(g.v_cache) & 3)))
    return fail(ECCQ_EUNSUPPORTED, "paged_decode_attention: head_dim must be a multiple of 4 for byte codecs");
  if (int rc = ensure_tables()) return rc;
  AttnArgs a;
  a.q = q; a.out = out; a.block_table = block_table; a.context_lens = context_lens; a.ws = (float*)workspace;
  a.max_blocks = max_blocks; a.nseq = nseq; a.Hq = Hq; a.layer = layer; a.sm_scale = sm_scale;
  a.force_general = getenv("ECCQ_ATTN_GENERAL") != nullptr;   // debugging aid: route everything through the general kernel
  a.interp = use_interpolation && g.codec == ECCQ_CODEC_H84; a.v_use_k_scales = v_use_k_scales; a.count_errors = count_errors && counters;
  attn_plan(max_blocks * g.bs, a.nsplit, a.chunk);
  if (workspace_bytes < (int64_t)nseq * Hq * a.nsplit * (g.D + 2) * (int64_t)sizeof(float))
    return fail(ECCQ_EINVAL, "paged_decode_attention: workspace too small (size it with max_context = max_blocks*block_size)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch (q_dtype) {
    case ECCQ_F32: rc = run_attn<float>(g, a, counters, st); break;
    case ECCQ_F16: rc = run_attn<__half>(g, a, counters, st); break;
    case ECCQ_BF16: rc = run_attn<__nv_bfloat16>(g, a, counters, st); break;
    default: return fail(ECCQ_EINVAL, "paged_decode_attention: bad dtype");
  }
  if (rc) return rc;
  return check_launch("paged_decode_attention");
}

extern "C" int eccq_paged_prefill_attention(const eccq_cache_desc* c, const void* q, void* out, int q_dtype,
                                            const int32_t* block_table, int max_blocks, const int32_t* context_lens, int nseq,
                                            int batch, int Hq, int Sq, int layer, float sm_scale, int use_interpolation,
                                            int count_errors, uint64_t* counters, void* stream) {
  CacheGeom g;
  if (!make_geom(c, g)) return fail(ECCQ_EINVAL, "paged_prefill_attention: bad cache descriptor");
  if (nseq < 0 || batch < 0 || Hq <= 0 || Hq % g.H != 0 || Sq <= 0 || layer < 0 || layer >= g.L || max_blocks <= 0 || !block_table || !context_lens)
    return fail(ECCQ_EINVAL, "paged_prefill_attention: bad arguments");
  if (nseq != 1 && nseq != batch) return fail(ECCQ_EINVAL, "paged_prefill_attention: nseq must be 1 (shared context) or equal to batch");
  if (nseq == 0 || batch == 0) return ECCQ_OK;
  if (!q || !out) return fail(ECCQ_EINVAL, "paged_prefill_attention: null pointer");
  if (batch > 65535 || Hq > 65535) return fail(ECCQ_EUNSUPPORTED, "paged_prefill_attention: batch / head count above the grid limits");
  if (g.codec != ECCQ_CODEC_GOLAY && (g.D % 4 != 0 || (reinterpret_cast<uintptr_t>(g.k_cache) & 3) || (reinterpret_cast<uintptr_t>(g.v_cache) & 3)))
    return fail(ECCQ_EUNSUPPORTED, "paged_prefill_attention: head_dim must be a multiple of 4 for byte codecs");
  if (int rc = ensure_tables()) return rc;
  AttnArgs a;
  a.q = q; a.out = out; a.block_table = block_table; a.context_lens = context_lens; a.ws = nullptr;
  a.max_blocks = max_blocks; a.nseq = nseq; a.Hq = Hq; a.layer = layer; a.sm_scale = sm_scale;
  a.nsplit = 1; a.chunk = max_blocks * g.bs; a.force_general = 1;
  a.interp = use_interpolation && g.codec == ECCQ_CODEC_H84; a.v_use_k_scales = 0; a.count_errors = count_errors && counters;
  a.Sq = Sq; a.qbatch = batch;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch (q_dtype) {
    case ECCQ_F32: rc = run_prefill<float>(g, a, counters, st); break;
    case ECCQ_F16: rc = run_prefill<__half>(g, a, counters, st); break;
    case ECCQ_BF16: rc = run_prefill<__nv_bfloat16>(g, a, counters, st); break;
    default: return fail(ECCQ_EINVAL, "paged_prefill_attention: bad dtype");
  }
  if (rc) return rc;
  return check_launch("paged_prefill_attention");
}

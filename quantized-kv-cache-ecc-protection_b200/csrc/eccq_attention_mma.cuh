// eccq_attention_mma.cuh -- ECC-fused paged decode attention, tensor-core tile version.
//
// One CTA = (sequence, kv-head, context chunk), four warps; each warp walks 16-token tiles (one tile never straddles a
// page: block_size % 16 == 0).  Per tile:
//   cp.async   the K and V codeword rows of the tile (2 x 16 x D bytes) and their 2 x 16 fp32 scales are staged in shared
//              memory one tile ahead (double buffer per warp), straight from the paged cache -- no gather pass;
//   prologue   the codeword bytes are SWAR-decoded in registers (PRMT-LUT syndrome decode, eccq_common.cuh) and expanded
//              to exact fp16 integers with the 0x6400|n magic (1024+n) minus 1032 -- the operands never exist in memory;
//   QK^T       S^T[16 tokens x 8 heads] = K[16 x D] . Q^T on mma.sync m16n8k16 (A = K fragments built in registers with a
//              dim permutation that makes every lane's bytes contiguous, B = Q fragments resident in registers);
//   softmax    online, base 2 (log2e and sm_scale folded into Q), per-token K scale applied to S, V scale folded into P;
//   PV         O^T[D x 8 heads] += V^T[D x 16 tokens] . P on mma.sync (P moved from accumulator to operand layout with
//              movmatrix.trans; V fragments gathered across token rows with PRMT).
// The GQA group (<= 8 q-heads) rides on the N dimension, so each codeword is decoded once for the whole group.
// With fp32 queries Q and P are split hi+lo (two MMAs) so the 1e-4 bar of the reference's fp32 tests holds.
//
// The contraction issues legacy HMMA (under 2 % of the instruction stream); the kernel is bound by the decode prologue and the
// tile loop around it, not by the tensor pipe.  Two prologues share this tile loop: SWAR (hamming_sel4, used when decode
// errors are counted) and a shared-memory table (default, see attn_mma_kernel<..., LUT>).  Interpolation and Golay reach this
// kernel through eccq_cache_correct_pages (corrected INT4 pages).  DESIGN.md 4.6 has the measurements and the tcgen05 note.
#pragma once
#include "eccq_attention.h"

namespace eccq {

constexpr int kMmaWarps = 4;
constexpr int kMmaThreads = kMmaWarps * 32;
constexpr int kTile = 16;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// two nibble bytes -> half2 of exact integers (n - 8): bytes {b, 0x64} = 1024 + n, then -1032
__device__ __forceinline__ uint32_t nib2_to_half2(uint32_t w, uint32_t sel) {
  const uint32_t h = __byte_perm(w, 0x64646464u, sel);
  const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&h), __halves2half2(__ushort_as_half(0xE408), __ushort_as_half(0xE408)));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// corrected nibbles of four codewords; flag bytes (0x80 single, 0xFF double) only when the caller counts errors
template <int CODEC, bool COUNT>
__device__ __forceinline__ uint32_t decode_word(uint32_t w, uint32_t& flag) {
  if (CODEC == ECCQ_CODEC_H84 || CODEC == ECCQ_CODEC_H74) {
    constexpr int HC = CODEC == ECCQ_CODEC_H74 ? ECCQ_CODEC_H74 : ECCQ_CODEC_H84;
    const uint32_t sel = hamming_sel4<HC>(w);
    if (COUNT) flag = prmt(kFlagLutLo, kFlagLutHi, sel);
    return lop3<kLopXorAnd>(w, prmt(kFlipLutLo, kFlipLutHi, sel), M0F);
  }
  flag = 0;
  return w;
}

// LUT = true: the "table" variant.  The decode prologue of the SWAR variant keeps the ALU pipe 71 % busy (10 ALU operations
// per four codewords + 2 for the fp16 expansion) while HBM idles at 40 %.  Here a 256-entry table in shared memory maps a
// codeword byte straight to its corrected value as an fp16 operand: the entry is the half 0xHB00 = fp16(n - 8) (the low byte
// of fp16(n - 8) is zero for every integer in [-8, 7]).  The table is replicated per bank -- entry c of lane L lives at byte
// c*256 + L*4 -- so every look-up is conflict-free, and the address is ONE PRMT ({0, 0, codeword byte, lane*4}).  Two
// look-ups are merged into a half2 operand with one IMAD (FMA-heavy pipe).  Per value: 1 ALU operation + 1 LDS instead of
// 3.6 ALU operations.  (A 65 536-entry pair table was tried first: 5.4 bank-conflict wavefronts per look-up made it
// shared-memory bound and slower than SWAR.)  The table takes 64 KB, so the kernel is persistent: one 8-warp CTA per SM
// builds it once and loops over (sequence, kv-head, chunk) items.  Used when errors are not counted (the counters need the
// flag bytes of the SWAR decode).
constexpr int kLutWarps = 16;
constexpr int kLutLowWarps = 6;        // warps whose staging lies below the table (6 x 9.25 KB < 63 KB)
constexpr int kLutTableBytes = 65536;

// dynamic shared memory of attn_mma_kernel; the table variant keeps its look-up table at the start of it
extern __shared__ __align__(16) unsigned char eccq_attn_smem[];

// The table sits at the ABSOLUTE shared-window address 0x10000 (64 KB aligned), so the complete address of a look-up --
// 0x10000 + byte K of x * 256 + lane * 4 -- is composed by ONE PRMT from x and the register `lane4t` = 0x00010000 | lane*4;
// no add.  The staging areas of the warps are placed around it (see the kernel prologue).
constexpr uint32_t kLutTableAddr = 0x10000u;
template <int K>
__device__ __forceinline__ uint32_t lut_byte(uint32_t, uint32_t x, uint32_t lane4t) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(prmt(x, lane4t, 0x7604u | (K << 4))));
  return v;
}
__device__ __forceinline__ uint32_t half_pair(uint32_t lo, uint32_t hi) { return hi * 65536u + lo; }

template <typename TQ, int CODEC, int D, bool PRECISE, bool COUNT, int LUT>
__global__ void __launch_bounds__(LUT ? kLutWarps * 32 : kMmaThreads, LUT ? 1 : 4) attn_mma_kernel(CacheGeom g, AttnArgs a, uint64_t* counters,
                                                                                                  const uint16_t* pair_tab) {
  constexpr int NW = LUT ? kLutWarps : kMmaWarps;   // warps per CTA
  constexpr bool LUT_K = LUT == 1 || LUT == 3, LUT_V = LUT == 1 || LUT == 2;   // which operand uses the table
  constexpr int KS = D / 16;          // k-steps of QK^T == m-tiles of PV
  constexpr int KW = D / 16;          // 32-bit words of a K row owned by a lane (D/4 bytes)
  constexpr int VW = D / 32;          // words of a V row owned by a lane (D/8 bytes)
  constexpr int ROW = D + 16;         // padded smem row stride: conflict-free 16-byte reads
  constexpr int STAGE = 2 * kTile * ROW + 2 * kTile * 4;   // K rows, V rows, k scales, v scales
  static_assert(!(LUT && COUNT), "the table variant does not produce error flags");
  // Shared-memory layout of the table variant (absolute shared-window addresses; `base` = start of the dynamic window):
  //   [base, 0x10000)            staging of warps 0 .. kLutLowWarps-1
  //   [0x10000, 0x20000)         the look-up table
  //   [0x20000, ...)             staging of the remaining warps; reused as the merge area at the end of an item
  unsigned char* const smem_all = eccq_attn_smem;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(eccq_attn_smem);
  unsigned char* const upper = smem_all + (0x20000u - base);
  unsigned char* smem = LUT ? upper : smem_all;   // merge area (and staging of the non-table variant)
  const uint32_t tab = kLutTableAddr;
  if (LUT) {
    // the absolute-address layout assumes the dynamic window starts low enough for kLutLowWarps staging pairs below the table
    if (base + kLutLowWarps * 2 * STAGE > kLutTableAddr) __trap();
    const uint8_t* one = reinterpret_cast<const uint8_t*>(pair_tab) + (CODEC == ECCQ_CODEC_H84 ? 256 : 0);   // hb(decode(c)), c = 0..255
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_all + (kLutTableAddr - base));
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) dst[(i >> 5) * 64 + (i & 31)] = (uint32_t)__ldg(one + (i >> 5)) << 8;
    __syncthreads();
  }
  const int G = a.Hq / g.H;
  const int npass = (G + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grow = lane >> 2, quad = lane & 3;
  const uint32_t lane4 = (uint32_t)lane * 4u | (LUT ? kLutTableAddr : 0u);   // see lut_byte
  const int64_t page_bytes = (int64_t)g.bs * D;
  unsigned char* wsm = !LUT ? smem + (size_t)warp * 2 * STAGE
                            : (warp < kLutLowWarps ? smem_all + (size_t)warp * 2 * STAGE : upper + (size_t)(warp - kLutLowWarps) * 2 * STAGE);
  const float* vsc_base = a.v_use_k_scales ? g.k_scales : g.v_scales;
  FlagCount fc;
  const int ny = g.H * npass;
  const int nitems = LUT ? a.nsplit * ny * a.nseq : 1;
#pragma unroll 1
  for (int item = LUT ? (int)blockIdx.x : 0; item < nitems; item += LUT ? (int)gridDim.x : 1) {
  // item = (i * ny + hy) * nsplit + split: consecutive CTAs walk the chunks of one (sequence, head)
  const int split = LUT ? item % a.nsplit : (int)blockIdx.x;
  const int hy = LUT ? (item / a.nsplit) % ny : (int)blockIdx.y;
  const int i = LUT ? item / (a.nsplit * ny) : (int)blockIdx.z;
  const int h = hy / npass;
  const int pass = hy - h * npass;
  const int ctx = min(a.context_lens[i], a.max_blocks * g.bs);   // never walk past the block table
  const int c0 = split * a.chunk;
  const int c1 = min(ctx, c0 + a.chunk);
  const int g0 = pass * 8;
  const int ng = min(8, G - g0);
  const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;

  // ---- Q fragments (B operand of QK^T): head n = lane/4, dims quad*(D/4) + 4s + {0,1 | 2,3}; sm_scale*log2e folded in
  uint32_t qf[KS][2], qf_lo[PRECISE ? KS : 1][2];
  {
    const bool valid = grow < ng;
    const int64_t qbase = ((int64_t)i * a.Hq + (int64_t)h * G + g0 + (valid ? grow : 0)) * D;
    const float sc = a.sm_scale * 1.4426950408889634f;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = valid ? ldq<TQ>(a.q, qbase + quad * (D / 4) + 4 * s + e) * sc : 0.0f;
      qf[s][0] = pack_half2(v[0], v[1]);
      qf[s][1] = pack_half2(v[2], v[3]);
      if (PRECISE) {
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&qf[s][0]));
        const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&qf[s][1]));
        qf_lo[s][0] = pack_half2(v[0] - h0.x, v[1] - h0.y);
        qf_lo[s][1] = pack_half2(v[2] - h1.x, v[3] - h1.y);
      }
    }
  }

  float o[KS][4];
#pragma unroll
  for (int mt = 0; mt < KS; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[mt][e] = 0.0f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};   // heads 2*quad, 2*quad+1
  const bool do_count = COUNT && pass == 0;

  // tiles of this warp: first tile index (in units of 16 tokens) is c0/16 + warp, stride kMmaWarps
  const int tile_end = (c1 + kTile - 1) / kTile;
  auto stage_tile = [&](int tile, int st) {
    unsigned char* dst = wsm + (size_t)st * STAGE;
    const int tok0 = tile * kTile;
    const int lb = tok0 / g.bs;
    const int blk = bt[lb];
    if (blk >= 0) {
      const int64_t page = page_index(g, blk, a.layer, h);
      const int slot0 = tok0 - lb * g.bs;
      const uint8_t* ksrc = g.k_cache + page * page_bytes + (int64_t)slot0 * D;
      const uint8_t* vsrc = g.v_cache + page * page_bytes + (int64_t)slot0 * D;
#pragma unroll
      for (int j = 0; j < kTile * D / 16 / 32; ++j) {
        const int c = lane + 32 * j;
        const int row = c / (D / 16), col = c % (D / 16);
        cp_async16(dst + row * ROW + col * 16, ksrc + row * D + col * 16);
        cp_async16(dst + kTile * ROW + row * ROW + col * 16, vsrc + row * D + col * 16);
      }
      if (lane < 4) cp_async16(dst + 2 * kTile * ROW + lane * 16, g.k_scales + page * g.bs + slot0 + lane * 4);
      else if (lane < 8) cp_async16(dst + 2 * kTile * ROW + kTile * 4 + (lane - 4) * 16, vsc_base + page * g.bs + slot0 + (lane - 4) * 4);
    }
    return blk >= 0;
  };

  int tile = c0 / kTile + warp;
  bool cur_ok = false;
  if (tile < tile_end) cur_ok = stage_tile(tile, 0);
  cp_async_commit();
  int st = 0;
  for (; tile < tile_end; tile += NW, st ^= 1) {
    const int ntile = tile + NW;
    bool next_ok = false;
    if (ntile < tile_end) next_ok = stage_tile(ntile, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    if (cur_ok) {
      const unsigned char* sK = wsm + (size_t)st * STAGE;
      const unsigned char* sV = sK + kTile * ROW;
      const float* sKs = reinterpret_cast<const float*>(sK + 2 * kTile * ROW);
      const float* sVs = sKs + kTile;
      const int tok0 = tile * kTile;

      // ---------------- QK^T: rows grow and grow+8, this lane's D/4 bytes of each
      float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      {
        uint32_t kr0[KW], kr1[KW];
#pragma unroll
        for (int w4 = 0; w4 < KW / 4; ++w4) {
          const uint4 x0 = *reinterpret_cast<const uint4*>(sK + grow * ROW + quad * (D / 4) + 16 * w4);
          const uint4 x1 = *reinterpret_cast<const uint4*>(sK + (grow + 8) * ROW + quad * (D / 4) + 16 * w4);
          kr0[4 * w4] = x0.x; kr0[4 * w4 + 1] = x0.y; kr0[4 * w4 + 2] = x0.z; kr0[4 * w4 + 3] = x0.w;
          kr1[4 * w4] = x1.x; kr1[4 * w4 + 1] = x1.y; kr1[4 * w4 + 2] = x1.z; kr1[4 * w4 + 3] = x1.w;
        }
        const bool v0 = do_count && tok0 + grow < c1, v1 = do_count && tok0 + grow + 8 < c1;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          uint32_t af[4];
          if (LUT_K) {
            af[0] = half_pair(lut_byte<0>(tab, kr0[s], lane4), lut_byte<1>(tab, kr0[s], lane4));
            af[1] = half_pair(lut_byte<0>(tab, kr1[s], lane4), lut_byte<1>(tab, kr1[s], lane4));
            af[2] = half_pair(lut_byte<2>(tab, kr0[s], lane4), lut_byte<3>(tab, kr0[s], lane4));
            af[3] = half_pair(lut_byte<2>(tab, kr1[s], lane4), lut_byte<3>(tab, kr1[s], lane4));
          } else {
            uint32_t f0 = 0, f1 = 0;
            const uint32_t n0 = decode_word<CODEC, COUNT>(kr0[s], f0);
            const uint32_t n1 = decode_word<CODEC, COUNT>(kr1[s], f1);
            if (COUNT) { fc.add(v0 ? f0 : 0u); fc.add(v1 ? f1 : 0u); }
            af[0] = nib2_to_half2(n0, 0x4140); af[1] = nib2_to_half2(n1, 0x4140); af[2] = nib2_to_half2(n0, 0x4342); af[3] = nib2_to_half2(n1, 0x4342);
          }
          mma16816(s4, af, qf[s][0], qf[s][1]);
          if (PRECISE) mma16816(s4, af, qf_lo[s][0], qf_lo[s][1]);
        }
      }
      // ---------------- scale, mask, online softmax (heads 2*quad, 2*quad+1; tokens grow, grow+8)
      const float ks0 = sKs[grow], ks1 = sKs[grow + 8];
      const bool t0ok = tok0 + grow < c1, t1ok = tok0 + grow + 8 < c1;
      s4[0] = t0ok ? s4[0] * ks0 : -INFINITY; s4[1] = t0ok ? s4[1] * ks0 : -INFINITY;
      s4[2] = t1ok ? s4[2] * ks1 : -INFINITY; s4[3] = t1ok ? s4[3] * ks1 : -INFINITY;
      float mx[2] = {fmaxf(s4[0], s4[2]), fmaxf(s4[1], s4[3])};
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xFFFFFFFFu, mx[0], off));
        mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xFFFFFFFFu, mx[1], off));
      }
      float alpha[2], p[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float mn = fmaxf(m_run[e], mx[e]);
        alpha[e] = (mn == -INFINITY) ? 1.0f : exp2f(m_run[e] - mn);
        m_run[e] = mn;
      }
      p[0] = (m_run[0] == -INFINITY) ? 0.0f : exp2f(s4[0] - m_run[0]);
      p[1] = (m_run[1] == -INFINITY) ? 0.0f : exp2f(s4[1] - m_run[1]);
      p[2] = (m_run[0] == -INFINITY) ? 0.0f : exp2f(s4[2] - m_run[0]);
      p[3] = (m_run[1] == -INFINITY) ? 0.0f : exp2f(s4[3] - m_run[1]);
      l_run[0] = l_run[0] * alpha[0] + p[0] + p[2];   // per-lane partial sums; reduced over the 8 row-lanes at the end
      l_run[1] = l_run[1] * alpha[1] + p[1] + p[3];
      if (__any_sync(0xFFFFFFFFu, alpha[0] != 1.0f || alpha[1] != 1.0f)) {   // the running maxima rarely move after the first tiles
#pragma unroll
        for (int mt = 0; mt < KS; ++mt) { o[mt][0] *= alpha[0]; o[mt][1] *= alpha[1]; o[mt][2] *= alpha[0]; o[mt][3] *= alpha[1]; }
      }
      // ---------------- P (x V scale) -> B fragments of PV
      const float vs0 = sVs[grow], vs1 = sVs[grow + 8];
      // masked rows (past the context / unmapped page) contribute exactly 0 even when the caller's scale slot holds Inf / NaN
      const float pv[4] = {t0ok ? p[0] * vs0 : 0.0f, t0ok ? p[1] * vs0 : 0.0f, t1ok ? p[2] * vs1 : 0.0f, t1ok ? p[3] * vs1 : 0.0f};
      const uint32_t ph0 = pack_half2(pv[0], pv[1]), ph1 = pack_half2(pv[2], pv[3]);
      const uint32_t pb0 = movmatrix_trans(ph0), pb1 = movmatrix_trans(ph1);
      uint32_t pl0 = 0, pl1 = 0;
      if (PRECISE) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&ph0));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&ph1));
        pl0 = movmatrix_trans(pack_half2(pv[0] - f0.x, pv[1] - f0.y));
        pl1 = movmatrix_trans(pack_half2(pv[2] - f1.x, pv[3] - f1.y));
      }
      // ---------------- PV: tokens 2q, 2q+1, 2q+8, 2q+9; this lane's D/8 bytes [grow*D/8, +D/8) of each
      {
        uint32_t vr[4][VW];
        const int toks[4] = {2 * quad, 2 * quad + 1, 2 * quad + 8, 2 * quad + 9};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const unsigned char* src = sV + toks[t] * ROW + grow * (D / 8);
          if (VW == 4) {
            const uint4 x = *reinterpret_cast<const uint4*>(src);
            vr[t][0] = x.x; vr[t][1] = x.y; vr[t][2 % VW] = x.z; vr[t][3 % VW] = x.w;
          } else {
            const uint2 x = *reinterpret_cast<const uint2*>(src);
            vr[t][0] = x.x; vr[t][1 % VW] = x.y;
          }
          const bool vok = do_count && tok0 + toks[t] < c1;
          if (!LUT_V) {
#pragma unroll
            for (int w = 0; w < VW; ++w) {
              uint32_t f = 0;
              vr[t][w] = decode_word<CODEC, COUNT>(vr[t][w], f);
              if (COUNT) fc.add(vok ? f : 0u);
            }
          }
        }
#pragma unroll
        for (int mt = 0; mt < KS; ++mt) {
          const int w = mt >> 1;
          uint32_t af[4];
          if (LUT_V) {   // operand pairs (T0.b, T1.b): two byte look-ups each, no gather
            if (mt & 1) {
              af[0] = half_pair(lut_byte<2>(tab, vr[0][w], lane4), lut_byte<2>(tab, vr[1][w], lane4));
              af[1] = half_pair(lut_byte<3>(tab, vr[0][w], lane4), lut_byte<3>(tab, vr[1][w], lane4));
              af[2] = half_pair(lut_byte<2>(tab, vr[2][w], lane4), lut_byte<2>(tab, vr[3][w], lane4));
              af[3] = half_pair(lut_byte<3>(tab, vr[2][w], lane4), lut_byte<3>(tab, vr[3][w], lane4));
            } else {
              af[0] = half_pair(lut_byte<0>(tab, vr[0][w], lane4), lut_byte<0>(tab, vr[1][w], lane4));
              af[1] = half_pair(lut_byte<1>(tab, vr[0][w], lane4), lut_byte<1>(tab, vr[1][w], lane4));
              af[2] = half_pair(lut_byte<0>(tab, vr[2][w], lane4), lut_byte<0>(tab, vr[3][w], lane4));
              af[3] = half_pair(lut_byte<1>(tab, vr[2][w], lane4), lut_byte<1>(tab, vr[3][w], lane4));
            }
          } else {
            const uint32_t sel = (mt & 1) ? 0x7362u : 0x5140u;
            const uint32_t x01 = __byte_perm(vr[0][w], vr[1][w], sel);   // {T0.b, T1.b, T0.b+1, T1.b+1}
            const uint32_t x23 = __byte_perm(vr[2][w], vr[3][w], sel);
            af[0] = nib2_to_half2(x01, 0x4140); af[1] = nib2_to_half2(x01, 0x4342); af[2] = nib2_to_half2(x23, 0x4140); af[3] = nib2_to_half2(x23, 0x4342);
          }
          mma16816(o[mt], af, pb0, pb1);
          if (PRECISE) mma16816(o[mt], af, pl0, pl1);
        }
      }
    }
    __syncwarp();
    cur_ok = next_ok;
  }
  cp_async_wait<0>();

  // ---- reduce the row sums over the 8 row-lanes (m_run is already uniform across them)
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    l_run[0] += __shfl_xor_sync(0xFFFFFFFFu, l_run[0], off);
    l_run[1] += __shfl_xor_sync(0xFFFFFFFFu, l_run[1], off);
  }
  // ---- merge the four warps through shared memory (reuse the staging area), write the chunk partial
  __syncthreads();
  constexpr int RS = D + 2;
  float* red = reinterpret_cast<float*>(smem);            // [warp][8 heads][D + 2]
  {
    float* dst = red + (size_t)warp * 8 * RS;
#pragma unroll
    for (int mt = 0; mt < KS; ++mt) {
      const int d0 = grow * (D / 8) + 2 * mt;              // rows grow / grow+8 of m-tile mt -> dims d0 / d0+1
      dst[(2 * quad) * RS + d0] = o[mt][0];
      dst[(2 * quad + 1) * RS + d0] = o[mt][1];
      dst[(2 * quad) * RS + d0 + 1] = o[mt][2];
      dst[(2 * quad + 1) * RS + d0 + 1] = o[mt][3];
    }
    if (grow == 0) {
      dst[(2 * quad) * RS + D] = m_run[0]; dst[(2 * quad) * RS + D + 1] = l_run[0];
      dst[(2 * quad + 1) * RS + D] = m_run[1]; dst[(2 * quad + 1) * RS + D + 1] = l_run[1];
    }
  }
  __syncthreads();
  for (int gi = warp; gi < ng; gi += NW) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, red[((size_t)w * 8 + gi) * RS + D]);
    float f[NW], L = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = red[((size_t)w * 8 + gi) * RS + D];
      f[w] = (mw == -INFINITY) ? 0.0f : exp2f(mw - M);
      L += red[((size_t)w * 8 + gi) * RS + D + 1] * f[w];
    }
    const int qh = h * G + g0 + gi;
    float* wsrow = a.ws + (((int64_t)i * a.Hq + qh) * a.nsplit + split) * (D + 2);
    for (int d = lane; d < D; d += 32) {
      float acc = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) acc += red[((size_t)w * 8 + gi) * RS + d] * f[w];
      wsrow[d] = acc;
    }
    // the merge kernel works in natural-log units: m_ln = m_log2 * ln2
    if (lane == 0) { wsrow[D] = (M == -INFINITY) ? -INFINITY : M * 0.6931471805599453f; wsrow[D + 1] = L; }
  }
  if (LUT) __syncthreads();   // the staging area is reused by the next item
  }   // items
  if (COUNT) {
    const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
    const uint32_t val[2] = {fc.singles(), fc.doubles()};
    block_accumulate<2>(counters, idx, val);
  }
}

template <typename TQ, int CODEC, int D, bool PRECISE>
static int launch_attn_mma(const CacheGeom& g, AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  const size_t stage = 2 * kTile * (D + 16) + 2 * kTile * 4;
  const int G = a.Hq / g.H;
  const int npass = (G + 7) / 8;
  constexpr bool kHamming = CODEC == ECCQ_CODEC_H84 || CODEC == ECCQ_CODEC_H74;
  static const int lut_env = [] { const char* v = std::getenv("ECCQ_ATTN_LUT"); return v && *v ? std::atoi(v) : 1; }();
  if (kHamming && !a.count_errors && lut_env) {
    // table variant: persistent, one 8-warp CTA per SM
    const uint16_t* tab = ensure_pair_tables();
    if (tab != nullptr) {
      // a persistent CTA of kLutWarps warps needs items long enough to amortise the per-item pipeline fill and merge:
      // coarsen the split-KV plan (the partial / merge protocol is unchanged, only fewer partials are written)
      static const int coarsen = [] { const char* v = std::getenv("ECCQ_ATTN_LUT_COARSEN"); return v && *v ? std::atoi(v) : 4; }();
      if (coarsen > 1 && a.nsplit > 1) {
        a.chunk *= coarsen;
        a.nsplit = (a.nsplit + coarsen - 1) / coarsen;
      }
      // dynamic window: everything up to the table's end at 0x20000 (the window starts at >= 0 and, with the 1 KB the
      // system reserves, normally at 1024) + the upper staging / merge area
      const size_t upper = std::max<size_t>((kLutWarps - kLutLowWarps) * 2 * stage, sizeof(float) * kLutWarps * 8 * (D + 2));
      const size_t smem = 0x20000 + upper;
      const int items = a.nsplit * g.H * npass * a.nseq;
      auto go = [&](auto kern) -> int {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
        kern<<<std::min(items, sm_count()), kLutWarps * 32, smem, st>>>(g, a, counters, tab);
        return check_launch("paged_decode_attention (mma.sync, table prologue)");
      };
      constexpr int HC = kHamming ? CODEC : ECCQ_CODEC_H84;
      if (lut_env == 2) return go(attn_mma_kernel<TQ, HC, D, PRECISE, false, 2>);
      if (lut_env == 3) return go(attn_mma_kernel<TQ, HC, D, PRECISE, false, 3>);
      return go(attn_mma_kernel<TQ, HC, D, PRECISE, false, 1>);
    }
  }
  const size_t smem = std::max<size_t>(kMmaWarps * 2 * stage, sizeof(float) * kMmaWarps * 8 * (D + 2));
  dim3 grid(a.nsplit, g.H * npass, a.nseq);
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
    }
    kern<<<grid, kMmaThreads, smem, st>>>(g, a, counters, nullptr);
    return check_launch("paged_decode_attention (mma.sync)");
  };
  if (a.count_errors) return go(attn_mma_kernel<TQ, CODEC, D, PRECISE, true, 0>);
  return go(attn_mma_kernel<TQ, CODEC, D, PRECISE, false, 0>);
}

// returns 1 when the tile kernel took the launch, 0 when the shape/codec needs the general kernel
template <typename TQ>
static int try_attn_mma(const CacheGeom& g, AttnArgs& a, uint64_t* counters, cudaStream_t st) {
  if (g.codec == ECCQ_CODEC_GOLAY || a.interp) return 0;
  if ((g.D != 64 && g.D != 128) || g.bs % kTile != 0 || a.chunk % kTile != 0) return 0;
  if (!aligned16(g.k_cache) || !aligned16(g.v_cache) || !aligned16(g.k_scales) || !aligned16(g.v_scales)) return 0;
  constexpr bool PRECISE = sizeof(TQ) == 4;
#define ECCQ_MMA_CASE(C)                                                                       \
  {                                                                                            \
    const int rc = g.D == 128 ? launch_attn_mma<TQ, C, 128, PRECISE>(g, a, counters, st)       \
                              : launch_attn_mma<TQ, C, 64, PRECISE>(g, a, counters, st);       \
    return rc ? rc : 1;   /* 1 = took the launch, < 0 = error */                               \
  }
  switch (g.codec) {
    case ECCQ_CODEC_H84: ECCQ_MMA_CASE(ECCQ_CODEC_H84);
    case ECCQ_CODEC_H74: ECCQ_MMA_CASE(ECCQ_CODEC_H74);
    default: ECCQ_MMA_CASE(ECCQ_CODEC_INT4);
  }
#undef ECCQ_MMA_CASE
}

}  // namespace eccq

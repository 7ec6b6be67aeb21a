// eccq_cache_read2.cuh -- fused cache read, second generation (included by eccq_cache.cu; uses CacheGeom / ReadArgs).
//
// One warp per (tensor, sequence, head, block) page as in cache_read_kernel -- every lane keeps SPL consecutive token slots
// of its 16-byte column in registers -- but each warp STREAMS through a chunk of consecutive pages of one
// (sequence, tensor, head) run, so that
//   * the work is split evenly over the resident warps (no tail wave, no per-page index divisions);
//   * the next page's codewords and scales are loaded while the current page is decoded (2 pages in flight per warp);
//   * the temporal neighbours across page edges are already on chip: the previous page's last row is carried in
//     registers, the next page's first row is the prefetched one (shuffled from lane group 0 and decoded under a vote).
// Built for the ALU-pipe budget measured by tools/pipe_bench.cu:
//   * hamming_decode4: 11 ALU + 7 IMAD per four codewords, flags counted with two IDP.4A (FMA-heavy pipe);
//   * temporal interpolation is branch-free SWAR on the nibble words (the flag byte 0xFF of a double error IS the select
//     mask; 0x80 of a single error selects nothing because nibbles and averages stay below 16), entered under one warp vote;
//   * dequantisation: PRMT builds fp16 pairs 0x6400|n (= 1024+n), the mixed-precision add FHADD turns each half into the
//     fp32 value n-8 in one FP32-pipe instruction, FMUL by the row scale, F2FP packs -- identical roundings to
//     ((n - 8) * scale).to(dtype) of the reference (ecc_shim.py:1040-1052).
// Compile-time geometry: C = head_dim/16 lanes per token row, SPL slots per lane, block_size = (32/C)*SPL.
#pragma once

namespace eccq {

// fp16 pair (1024+a, 1024+b) -> (a-8, b-8) as fp32, exact
__device__ __forceinline__ void halfpair_m8(uint32_t hp, float& lo, float& hi) {
  asm("{.reg .f16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "add.rn.f32.f16 %0, l, %3;\n\t"
      "add.rn.f32.f16 %1, h, %3;}"
      : "=f"(lo), "=f"(hi) : "r"(hp), "f"(-1032.0f));
}

template <typename TO> struct Pack2;
template <> struct Pack2<__half> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Pack2<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Pack2<float> {
  static __device__ __forceinline__ uint32_t pack(float a, float) { return __float_as_uint(a); }
};

// 16 nibble bytes (four words) * scale -> 16 outputs at dst
template <typename TO>
__device__ __forceinline__ void dequant_store16(TO* dst, const uint32_t (&nib)[4], float scale) {
  float f[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t p01 = prmt(nib[k], 0x64646464u, 0x4140u);   // {0x64, b1, 0x64, b0}
    const uint32_t p23 = prmt(nib[k], 0x64646464u, 0x4342u);   // {0x64, b3, 0x64, b2}
    halfpair_m8(p01, f[4 * k], f[4 * k + 1]);
    halfpair_m8(p23, f[4 * k + 2], f[4 * k + 3]);
  }
  if constexpr (sizeof(TO) == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      st_stream(dst + 4 * k, make_uint4(__float_as_uint(f[4 * k] * scale), __float_as_uint(f[4 * k + 1] * scale),
                                        __float_as_uint(f[4 * k + 2] * scale), __float_as_uint(f[4 * k + 3] * scale)));
  } else {
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = Pack2<TO>::pack(f[2 * q] * scale, f[2 * q + 1] * scale);
    st_stream(dst, make_uint4(o[0], o[1], o[2], o[3]));
    st_stream(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
  }
}

template <int CODEC>
__device__ __forceinline__ void decode_words4(const uint4& x, uint32_t (&nib)[4], uint32_t (&flag)[4]) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (CODEC == ECCQ_CODEC_H84 || CODEC == ECCQ_CODEC_H74) {
      hamming_decode4<CODEC == ECCQ_CODEC_H74 ? ECCQ_CODEC_H74 : ECCQ_CODEC_H84>(w[k], nib[k], flag[k]);
    } else {
      nib[k] = w[k]; flag[k] = 0;   // int4: the byte as is (ecc_shim.py:913-914)
    }
  }
}

template <int SPL>
__device__ __forceinline__ void load_scales(const float* p, float (&sc)[SPL]) {
  if constexpr (SPL == 4) {
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(p));
    sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
  } else if constexpr (SPL == 2) {
    const float2 s2 = __ldg(reinterpret_cast<const float2*>(p));
    sc[0] = s2.x; sc[1] = s2.y;
  } else {
#pragma unroll
    for (int s = 0; s < SPL; ++s) sc[s] = __ldg(p + s);
  }
}


// INTERP = true: the next page is prefetched into registers (it is also the right neighbour of this page's last row);
// INTERP = false: pages loaded at the top of their own iteration.  2 CTAs per SM (128 registers) everywhere except the
// 1 KB-page mapping without interpolation (C=4, SPL=2), which fits 80 registers and is 20 % faster at 3 CTAs per SM; the
// other non-interpolating variants spill ~100 B at 80 registers and run 12-15 % faster at 2 (same-box A/B).
// Prefetching the next page in the non-interpolating variants as well (the registers are there at 2 CTAs per SM, and ncu
// shows them waiting on their loads: long-scoreboard 3.5 cycles per issue) measured 9-17 % SLOWER -- not adopted.
// HP = heads per warp: with head_dim 64 and 16-token blocks a page is only 1 KB, so a warp takes the pages of TWO adjacent
// heads of a block (contiguous in the cache): each half-warp runs the 4-lanes-per-row x 4-slots mapping on its own head,
// which halves the per-page overhead.  Lane shuffles that cross the half-warp only reach values that are overwritten
// (first / last lane group of a page).
template <typename TO, int CODEC, bool INTERP, int C, int SPL, int HP = 1>
__global__ void __launch_bounds__(256, (C == 4 && SPL == 2 && !(INTERP && CODEC == ECCQ_CODEC_H84)) ? 3 : 2) cache_read2_kernel(CacheGeom g, ReadArgs a, int nblk_logical, int chunk, int head_fastest,
                                                              uint64_t* counters) {
  constexpr int D = 16 * C, HL = 32 / HP, GROUPS = HL / C, BS = GROUPS * SPL;   // HL = lanes per head
  constexpr bool kHamming = CODEC == ECCQ_CODEC_H84 || CODEC == ECCQ_CODEC_H74;
  constexpr bool kInterp = INTERP && CODEC == ECCQ_CODEC_H84;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr int64_t page_bytes = (int64_t)BS * D;
  const int lane = threadIdx.x & 31;
  const int hsel = lane / HL, hbase = hsel * HL;   // which of the warp's heads this lane works on; its first lane
  const int c = lane & (C - 1);
  const int grp = (lane - hbase) / C;
  const int slot0 = grp * SPL;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  // Work unit = (sequence i, tensor kv, chunk of kReadChunk consecutive blocks, head h) with h fastest: the warps of a CTA
  // work on adjacent heads of the same blocks, i.e. on one contiguous region of the cache (H adjacent pages per block) and
  // one contiguous region of the output ([token][head][D]) at a time, which keeps the DRAM accesses of a CTA page-local.
  const uint32_t nblk = (uint32_t)nblk_logical;
  const uint32_t nchunk = (nblk + chunk - 1) / chunk;
  const uint32_t HU = (uint32_t)g.H / HP;   // head units per (sequence, tensor)
  const uint32_t units = 2u * (uint32_t)a.nseq * HU * nchunk;
  FlagCount fc;

  for (uint32_t u = warp; u < units; u += nwarps) {
    // ---- set up a segment: pages [j, jend) of one (sequence, head, tensor) run
    uint32_t ikv;
    int h, j;
    if (head_fastest) {          // u = ((i*2 + kv)*nchunk + chunk_index)*H + h
      const uint32_t rest = u / HU;
      h = (int)(u - rest * HU);
      ikv = rest / nchunk;
      j = (int)(rest - ikv * nchunk) * chunk;
    } else {                     // u = ((i*2 + kv)*H + h)*nchunk + chunk_index
      const uint32_t rest = u / nchunk;
      j = (int)(u - rest * nchunk) * chunk;
      ikv = rest / HU;
      h = (int)(rest - ikv * HU);
    }
    h = h * HP + hsel;
    const int kv = (int)(ikv & 1u);
    const int i = (int)(ikv >> 1);
    const int ctx = min(a.context_lens[i], a.T);
    const int nb_ctx = min(min((ctx + BS - 1) / BS, (int)nblk), a.max_blocks);   // pages of this run inside the context
    if (j >= nb_ctx) continue;
    const int jend = min(nb_ctx, j + chunk);
    const int32_t* bt = a.block_table + (int64_t)i * a.max_blocks;
    const uint8_t* cache = (kv ? g.v_cache : g.k_cache) + c * 16;
    const float* scales = kv ? g.v_scales : g.k_scales;
    const int64_t lh = (int64_t)a.layer * g.H + h;        // page index = blk * (L*H) + lh
    const int64_t LH = (int64_t)g.L * g.H;
    TO* out = reinterpret_cast<TO*>(kv ? a.v_out : a.k_out);
    uint8_t* nout = kv ? a.nib_v : a.nib_k;
    const int64_t ostride = (int64_t)g.H * D;
    int64_t obase = (((int64_t)i * a.T + (int64_t)j * BS + slot0) * g.H + h) * D + c * 16;

    uint32_t carry[4] = {0, 0, 0, 0};   // pre-interpolation nibbles of the previous page's last row (column c)
    if (kInterp && j > 0) {
      const int pb = bt[j - 1];
      uint4 x = make_uint4(0, 0, 0, 0);
      if (pb >= 0) x = ld_stream(cache + ((int64_t)pb * LH + lh) * page_bytes + (BS - 1) * D);
      uint32_t nf[4];
      decode_words4<CODEC>(x, carry, nf);
    }
    int blk = bt[j];
    int blk_n = j + 1 < nb_ctx ? bt[j + 1] : -1;
    uint4 raw[SPL];
    float sc[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) raw[s] = make_uint4(0, 0, 0, 0);
    if (kInterp && blk >= 0) {
      const int64_t page = (int64_t)blk * LH + lh;
#pragma unroll
      for (int s = 0; s < SPL; ++s) raw[s] = ld_stream(cache + page * page_bytes + (slot0 + s) * D);
      load_scales<SPL>(scales + page * BS + slot0, sc);
    }

#pragma unroll 1
    for (; j < jend; ++j) {
      // ---- prefetch the next page of the run (also past this warp's range: its first row is the right neighbour)
      uint4 rawn[SPL];
      float scn[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) { rawn[s] = make_uint4(0, 0, 0, 0); scn[s] = 0.0f; }
      if (kInterp) {
        if (blk_n >= 0) {
          const int64_t page = (int64_t)blk_n * LH + lh;
#pragma unroll
          for (int s = 0; s < SPL; ++s) rawn[s] = ld_stream(cache + page * page_bytes + (slot0 + s) * D);
          load_scales<SPL>(scales + page * BS + slot0, scn);
        }
      } else if (blk >= 0) {   // no prefetch: this page's loads
        const int64_t page = (int64_t)blk * LH + lh;
#pragma unroll
        for (int s = 0; s < SPL; ++s) raw[s] = ld_stream(cache + page * page_bytes + (slot0 + s) * D);
        load_scales<SPL>(scales + page * BS + slot0, sc);
      }
      const int blk_nn = j + 2 < nb_ctx ? bt[j + 2] : -1;
      const int t0 = j * BS;
      const int nvalid = min(ctx - t0, BS);   // tokens of this page inside the context

      if (blk >= 0) {
        if (nvalid < BS) {   // ragged last page: slots past the context must decode clean
#pragma unroll
          for (int s = 0; s < SPL; ++s)
            if (slot0 + s >= nvalid) raw[s] = make_uint4(0, 0, 0, 0);
        }
        uint32_t nib[SPL][4], flag[SPL][4];
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
          decode_words4<CODEC>(raw[s], nib[s], flag[s]);
          if (kHamming) {
#pragma unroll
            for (int k = 0; k < 4; ++k) fc.add(flag[s][k]);
          }
        }
        if constexpr (kInterp) {
          uint32_t last[4];   // this page's last row before interpolation = the next page's left neighbour
#pragma unroll
          for (int k = 0; k < 4; ++k) last[k] = __shfl_sync(FULL, nib[SPL - 1][k], hbase + (GROUPS - 1) * C + c);
          uint32_t anyd = 0;
#pragma unroll
          for (int s = 0; s < SPL; ++s) anyd |= flag[s][0] | flag[s][1] | flag[s][2] | flag[s][3];
          if (__any_sync(FULL, (anyd & M01) != 0u)) {   // bit 0 of a flag byte is set only for double errors (0xFF)
            // neighbours held by the adjacent lane groups (pre-interpolation nibbles: SURVEY Q9, no cascading)
            uint32_t up[4], dn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              up[k] = __shfl_up_sync(FULL, nib[SPL - 1][k], C);
              dn[k] = __shfl_down_sync(FULL, nib[0][k], C);
              if (grp == 0) up[k] = carry[k];
            }
            // right neighbour of the page's last row = first row of the prefetched page (held raw by lane group 0)
            const bool e_next = grp == GROUPS - 1 && t0 + BS < ctx &&
                                ((flag[SPL - 1][0] | flag[SPL - 1][1] | flag[SPL - 1][2] | flag[SPL - 1][3]) & M01) != 0u;
            if (__any_sync(FULL, e_next)) {
              uint4 x;
              x.x = __shfl_sync(FULL, rawn[0].x, hbase + c); x.y = __shfl_sync(FULL, rawn[0].y, hbase + c);
              x.z = __shfl_sync(FULL, rawn[0].z, hbase + c); x.w = __shfl_sync(FULL, rawn[0].w, hbase + c);
              uint32_t nn[4], nf[4];
              decode_words4<CODEC>(x, nn, nf);
              if (grp == GROUPS - 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) dn[k] = nn[k];
              }
            }
            if (nvalid == BS) {
              // sequence ends: token 0 has no left neighbour, token ctx-1 no right neighbour -> the value itself
              if (grp == 0 && j == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) up[k] = nib[0][k];
              }
              if (grp == GROUPS - 1 && t0 + BS >= ctx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) dn[k] = nib[SPL - 1][k];
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t prev = up[k];
#pragma unroll
                for (int s = 0; s < SPL; ++s) {
                  const uint32_t self = nib[s][k];
                  const uint32_t right = s + 1 < SPL ? nib[s + 1][k] : dn[k];
                  const uint32_t avg = ((prev + right + M01) >> 1) & M0F;   // floor((l+r)*0.5+0.5), interpolation_triton.py:141-159
                  nib[s][k] = lop3<kLopSelect>(self, avg, flag[s][k]);
                  prev = self;
                }
              }
            } else {
              // ragged last page: clamp per token
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t prev = up[k];
#pragma unroll
                for (int s = 0; s < SPL; ++s) {
                  const int tok = t0 + slot0 + s;
                  const uint32_t self = nib[s][k];
                  const uint32_t left = tok == 0 ? self : prev;
                  const uint32_t right = tok + 1 >= ctx ? self : (s + 1 < SPL ? nib[s + 1][k] : dn[k]);
                  const uint32_t avg = ((left + right + M01) >> 1) & M0F;
                  nib[s][k] = lop3<kLopSelect>(self, avg, flag[s][k]);
                  prev = self;
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) carry[k] = last[k];
        }
        if (a.paged_nib) {
          // corrected INT4 pages of a one-layer cache with the same paged layout + a copy of the row scales
          const int64_t pg = (int64_t)blk * g.H + h;
          float* sco = kv ? a.scv_out : a.sck_out;
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            if (slot0 + s < nvalid) {
              st_stream(nout + pg * page_bytes + (slot0 + s) * D + c * 16, make_uint4(nib[s][0], nib[s][1], nib[s][2], nib[s][3]));
              if (c == 0) sco[pg * BS + slot0 + s] = sc[s];
            }
          }
        } else if (nvalid == BS) {
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            if (out) dequant_store16<TO>(out + obase + s * ostride, nib[s], sc[s]);
            if (nout) st_stream(nout + obase + s * ostride, make_uint4(nib[s][0], nib[s][1], nib[s][2], nib[s][3]));
          }
        } else {
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            if (slot0 + s < nvalid) {
              if (out) dequant_store16<TO>(out + obase + s * ostride, nib[s], sc[s]);
              if (nout) st_stream(nout + obase + s * ostride, make_uint4(nib[s][0], nib[s][1], nib[s][2], nib[s][3]));
            }
          }
        }
      } else if (kInterp) {
        // unmapped block inside the context: nothing is written, its rows read as zero nibbles for the neighbours
#pragma unroll
        for (int k = 0; k < 4; ++k) carry[k] = 0;
      }
      // ---- rotate
      if (kInterp) {
#pragma unroll
        for (int s = 0; s < SPL; ++s) { raw[s] = rawn[s]; sc[s] = scn[s]; }
      }
      blk = blk_n;
      blk_n = blk_nn;
      obase += (int64_t)BS * ostride;
    }
  }
  const uint32_t nd = fc.doubles();
  const int idx[3] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED, ECCQ_CTR_INTERPOLATED};
  const uint32_t val[3] = {fc.singles(), nd, kInterp ? nd : 0u};
  block_accumulate<3>(counters, idx, val);
}

}  // namespace eccq

// eccq_codecs.cu -- flat (non-paged) codec kernels: Hamming(7,4)/(8,4) encode + syndrome decode, Golay(24,12)
// triplet encode/decode, XOR-mask and Philox fault injection, temporal interpolation.
//
// Every kernel is HBM-bound byte/integer work: 128-bit coalesced loads/stores of packed codewords, a grid sized as a
// multiple of the SM count with a grid-stride loop, several independent 16-byte loads in flight per thread, and
// block-level counters (one atomicAdd per CTA).  Replaces the Triton kernels K1-K12 of SURVEY.md section 2.1.
#include "eccq_cache_geom.cuh"   // FastDiv

namespace eccq {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ------------------------------------------------------------------------------------------------ Hamming encode
template <int CODEC>
__device__ __forceinline__ uint32_t enc4(uint32_t x) { return CODEC == ECCQ_CODEC_H84 ? h84_encode4(x) : h74_encode4(x); }

template <int CODEC>
__global__ void __launch_bounds__(kThreads) hamming_encode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                   int64_t n, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec) {
    const int64_t n16 = n >> 4;
    for (int64_t i = tid; i < n16; i += nthreads * kUnroll) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) v[u] = ld_stream(in + (j << 4));
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) {
          uint4 o = make_uint4(enc4<CODEC>(v[u].x), enc4<CODEC>(v[u].y), enc4<CODEC>(v[u].z), enc4<CODEC>(v[u].w));
          st_stream(out + (j << 4), o);
        }
      }
    }
    done = n16 << 4;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) out[i] = (uint8_t)enc4<CODEC>(in[i]);
}

// ------------------------------------------------------------------------------------------------ Hamming decode
// hamming_decode4 (eccq_common.cuh) gives corrected nibbles and flag bytes (0x80 single, 0xFF double); the error-type
// byte of hamming84_triton.py:185-187 (0 none, 1 single, 2 double, 3 parity bit only) additionally needs "syndrome 0 with
// parity mismatch": one more PRMT look-up G = {clean 0x00, parity-only 0x02, single 0x81, double 0xFF}; its low two bits
// are the Gray code of the error type.
template <int CODEC, bool ZERO_DOUBLE, bool AUX>
__device__ __forceinline__ void dec4(uint32_t x, uint32_t& nib, uint32_t& aux, FlagCount& fc) {
  constexpr int HC = CODEC == ECCQ_CODEC_H74 ? ECCQ_CODEC_H74 : ECCQ_CODEC_H84;
  const uint32_t sel = hamming_sel4<HC>(x);
  const uint32_t flag = prmt(kFlagLutLo, kFlagLutHi, sel);
  nib = lop3<kLopXorAnd>(x, prmt(kFlipLutLo, kFlipLutHi, sel), M0F);
  fc.add(flag);
  if (CODEC == ECCQ_CODEC_H84) {
    if (ZERO_DOUBLE) nib &= ~((flag & M01) * 15u);   // fused_kernels.py:343-344 zeroes double-detected nibbles (Q4)
    if (AUX) {
      const uint32_t gcode = prmt(0x81818102u, 0x81818181u, sel) & 0x03030303u;   // 0, 2 (parity only), 1 (single), 3 (double)
      aux = gcode ^ ((gcode >> 1) & M01);                                           // Gray -> 0, 3, 1, 2
    }
  } else if (AUX) {
    aux = (flag >> 7) & M01;   // hamming74_triton.py:129-160: 1 = syndrome != 0 (corrected)
  }
}

template <int CODEC, bool ZERO_DOUBLE, bool AUX>
__global__ void __launch_bounds__(kThreads) hamming_decode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ nib_out,
                                                                   uint8_t* __restrict__ aux_out, uint64_t* counters,
                                                                   int64_t n, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  FlagCount fc;
  uint32_t ncorr = 0, ndet = 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n16 = n >> 4;
    for (int64_t i = tid; i < n16; i += nthreads * kUnroll) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) v[u] = ld_stream(in + (j << 4));
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) {
          uint4 o, a = make_uint4(0, 0, 0, 0);
          dec4<CODEC, ZERO_DOUBLE, AUX>(v[u].x, o.x, a.x, fc);
          dec4<CODEC, ZERO_DOUBLE, AUX>(v[u].y, o.y, a.y, fc);
          dec4<CODEC, ZERO_DOUBLE, AUX>(v[u].z, o.z, a.z, fc);
          dec4<CODEC, ZERO_DOUBLE, AUX>(v[u].w, o.w, a.w, fc);
          st_stream(nib_out + (j << 4), o);
          if (AUX) st_stream(aux_out + (j << 4), a);
        }
      }
      // the byte sums of FlagCount stay far below 2^31 over one flush interval
      ncorr += fc.singles(); ndet += fc.doubles(); fc = FlagCount();
    }
    done = n16 << 4;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    uint32_t o, a = 0;
    FlagCount f1;
    dec4<CODEC, ZERO_DOUBLE, AUX>(in[i], o, a, f1);   // upper three bytes are zero codewords: clean
    ncorr += f1.singles();
    ndet += f1.doubles();
    nib_out[i] = (uint8_t)o;
    if (AUX) aux_out[i] = (uint8_t)a;
  }
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {ncorr, ndet};
  block_accumulate<2>(counters, idx, val);
}

// ------------------------------------------------------------------------------------------------ Golay
constexpr int kGolayTile = 2048;  // codewords per CTA iteration: 6144 B of triplets <-> 8192 B of codewords

__device__ __forceinline__ uint32_t pack3(uint32_t b0, uint32_t b1, uint32_t b2) {
  return (b0 & 0xFu) | ((b1 & 0xFu) << 4) | ((b2 & 0xFu) << 8);
}

__global__ void __launch_bounds__(kThreads) golay_encode_kernel(const uint8_t* __restrict__ trip, int32_t* __restrict__ cw,
                                                                 int64_t n_cw, int vec) {
  __shared__ uint16_t s_lo[64], s_hi[64];
  __shared__ __align__(16) uint32_t s_in[kGolayTile * 3 / 4];  // 6144 B
  golay_load_parity(s_lo, s_hi);
  __syncthreads();
  if (!vec) {  // unaligned views: one codeword per thread, byte loads
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cw; i += nthreads)
      cw[i] = (int32_t)golay_encode1(s_lo, s_hi, pack3(trip[3 * i], trip[3 * i + 1], trip[3 * i + 2]));
    return;
  }
  const int64_t ntiles = (n_cw + kGolayTile - 1) / kGolayTile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t cw0 = tile * kGolayTile;
    const int64_t cnt = min((int64_t)kGolayTile, n_cw - cw0);
    const uint8_t* src = trip + cw0 * 3;
    if (cnt == kGolayTile) {
      // 384 x 16 B coalesced into shared memory, then 12 B (= 4 codewords) per thread at word stride 3: conflict-free
      for (int i = threadIdx.x; i < kGolayTile * 3 / 16; i += kThreads)
        reinterpret_cast<uint4*>(s_in)[i] = ld_stream(src + (i << 4));
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kGolayTile / 4 / kThreads; ++r) {
        const int c = r * kThreads + threadIdx.x;  // group of 4 codewords
        const uint32_t w0 = s_in[3 * c], w1 = s_in[3 * c + 1], w2 = s_in[3 * c + 2];
        uint4 o;
        o.x = golay_encode1(s_lo, s_hi, pack3(w0, w0 >> 8, w0 >> 16));
        o.y = golay_encode1(s_lo, s_hi, pack3(w0 >> 24, w1, w1 >> 8));
        o.z = golay_encode1(s_lo, s_hi, pack3(w1 >> 16, w1 >> 24, w2));
        o.w = golay_encode1(s_lo, s_hi, pack3(w2 >> 8, w2 >> 16, w2 >> 24));
        st_stream(cw + cw0 + 4 * (int64_t)c, o);
      }
      __syncthreads();
    } else {
      for (int64_t i = threadIdx.x; i < cnt; i += kThreads)
        cw[cw0 + i] = (int32_t)golay_encode1(s_lo, s_hi, pack3(src[3 * i], src[3 * i + 1], src[3 * i + 2]));
    }
  }
}

__global__ void __launch_bounds__(kThreads) golay_decode_kernel(const int32_t* __restrict__ cw, uint8_t* __restrict__ trip,
                                                                 uint8_t* __restrict__ errcnt, uint64_t* counters,
                                                                 int64_t n_cw, int vec) {
  __shared__ uint16_t s_lo[64], s_hi[64];
  __shared__ __align__(16) int32_t s_syn[4096];
  __shared__ __align__(16) uint32_t s_out[kGolayTile * 3 / 4];
  golay_load_parity(s_lo, s_hi);
  golay_load_syndrome(s_syn);
  __syncthreads();
  uint32_t nbits = 0, nunc = 0;
  auto one = [&](uint32_t w, uint32_t& c) {
    const uint32_t d = golay_decode1(s_lo, s_hi, s_syn, w, c);
    if (c == 4u) ++nunc; else nbits += c;
    return d;
  };
  if (!vec) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cw; i += nthreads) {
      uint32_t c;
      const uint32_t d = one((uint32_t)cw[i], c);
      trip[3 * i] = d & 0xF; trip[3 * i + 1] = (d >> 4) & 0xF; trip[3 * i + 2] = (d >> 8) & 0xF;
      if (errcnt) errcnt[i] = (uint8_t)c;
    }
  } else {
    const int64_t ntiles = (n_cw + kGolayTile - 1) / kGolayTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t cw0 = tile * kGolayTile;
      const int64_t cnt = min((int64_t)kGolayTile, n_cw - cw0);
      if (cnt == kGolayTile) {
        uint4 v[kGolayTile / 4 / kThreads];
#pragma unroll
        for (int r = 0; r < kGolayTile / 4 / kThreads; ++r)
          v[r] = ld_stream(cw + cw0 + 4 * (int64_t)(r * kThreads + threadIdx.x));
#pragma unroll
        for (int r = 0; r < kGolayTile / 4 / kThreads; ++r) {
          const int c = r * kThreads + threadIdx.x;
          uint32_t c0, c1, c2, c3;
          const uint32_t d0 = one(v[r].x, c0), d1 = one(v[r].y, c1), d2 = one(v[r].z, c2), d3 = one(v[r].w, c3);
          // 12 nibble bytes: d0:{0,1,2} d1:{3,4,5} d2:{6,7,8} d3:{9,10,11}
          const uint32_t n0 = d0 & 0xF, n1 = (d0 >> 4) & 0xF, n2 = d0 >> 8, n3 = d1 & 0xF, n4 = (d1 >> 4) & 0xF, n5 = d1 >> 8;
          const uint32_t n6 = d2 & 0xF, n7 = (d2 >> 4) & 0xF, n8 = d2 >> 8, n9 = d3 & 0xF, n10 = (d3 >> 4) & 0xF, n11 = d3 >> 8;
          s_out[3 * c] = n0 | (n1 << 8) | (n2 << 16) | (n3 << 24);
          s_out[3 * c + 1] = n4 | (n5 << 8) | (n6 << 16) | (n7 << 24);
          s_out[3 * c + 2] = n8 | (n9 << 8) | (n10 << 16) | (n11 << 24);
          if (errcnt) *reinterpret_cast<uint32_t*>(errcnt + cw0 + 4 * (int64_t)c) = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kGolayTile * 3 / 16; i += kThreads)
          st_stream(trip + cw0 * 3 + (i << 4), reinterpret_cast<const uint4*>(s_out)[i]);
        __syncthreads();
      } else {
        for (int64_t i = threadIdx.x; i < cnt; i += kThreads) {
          uint32_t c;
          const uint32_t d = one((uint32_t)cw[cw0 + i], c);
          uint8_t* t = trip + (cw0 + i) * 3;
          t[0] = d & 0xF; t[1] = (d >> 4) & 0xF; t[2] = (d >> 8) & 0xF;
          if (errcnt) errcnt[cw0 + i] = (uint8_t)c;
        }
      }
    }
  }
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {nbits, nunc};
  block_accumulate<2>(counters, idx, val);
}

// ------------------------------------------------------------------------------------------------ fault injection
// out = in ^ mask over 16-byte chunks; counters from popcount(mask restricted to the low n_bits of each element).
__global__ void __launch_bounds__(kThreads) xor_mask_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ mask,
                                                             uint8_t* __restrict__ out, int elem_size, uint32_t keep,
                                                             uint64_t* counters, int64_t nbytes, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  uint32_t flips = 0, affected = 0;
  auto count = [&](uint32_t m) {
    m &= keep;
    flips += __popc(m);
    if (elem_size == 4) affected += (m != 0);
    else { uint32_t t = m | (m >> 4); t |= t >> 2; t |= t >> 1; affected += __popc(t & M01); }
  };
  int64_t done = 0;
  if (vec) {
    const int64_t n16 = nbytes >> 4;
    for (int64_t i = tid; i < n16; i += nthreads * 2) {
      uint4 a[2], m[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) { a[u] = ld_stream(in + (j << 4)); m[u] = ld_stream(mask + (j << 4)); }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t j = i + (int64_t)u * nthreads;
        if (j < n16) {
          count(m[u].x); count(m[u].y); count(m[u].z); count(m[u].w);
          st_stream(out + (j << 4), make_uint4(a[u].x ^ m[u].x, a[u].y ^ m[u].y, a[u].z ^ m[u].z, a[u].w ^ m[u].w));
        }
      }
    }
    done = n16 << 4;
  }
  // tail (and unaligned views) element by element
  const int64_t nel = (nbytes - done) / elem_size;
  for (int64_t e = tid; e < nel; e += nthreads) {
    if (elem_size == 4) {
      uint32_t a, m;
      memcpy(&a, in + done + 4 * e, 4); memcpy(&m, mask + done + 4 * e, 4);
      count(m); a ^= m; memcpy(out + done + 4 * e, &a, 4);
    } else {
      const uint32_t m = mask[done + e];
      count(m); out[done + e] = in[done + e] ^ (uint8_t)m;
    }
  }
  const int idx[2] = {ECCQ_CTR_FLIPPED, ECCQ_CTR_AFFECTED};
  const uint32_t val[2] = {flips, affected};
  block_accumulate<2>(counters, idx, val);
}

// The reference's Triton Philox stream, element per thread (compat path: n_bits Philox-10 calls per element).
template <typename E>
__global__ void __launch_bounds__(kThreads) inject_ref_kernel(const E* __restrict__ in, E* __restrict__ out,
                                                               uint8_t* __restrict__ flipcnt, int n_bits, float ber,
                                                               uint32_t seed_lo, int vectorized, uint64_t* counters, int64_t n) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  uint32_t flips = 0, affected = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    const uint32_t m = vectorized ? ref_flip_mask_vec(seed_lo, (uint32_t)n, (uint32_t)i, n_bits, ber)
                                  : ref_flip_mask(seed_lo, (uint32_t)n, (uint32_t)i, n_bits, ber);
    out[i] = (E)((uint32_t)in[i] ^ m);
    const uint32_t c = __popc(m);
    if (flipcnt) flipcnt[i] = (uint8_t)c;
    flips += c;
    affected += (c != 0);
  }
  const int idx[2] = {ECCQ_CTR_FLIPPED, ECCQ_CTR_AFFECTED};
  const uint32_t val[2] = {flips, affected};
  block_accumulate<2>(counters, idx, val);
}

// Own stream: one 16-byte chunk per thread-iteration, flips placed by geometric gaps (see FastInject).
__global__ void __launch_bounds__(kThreads) inject_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                int elem_size, int n_bits, FastInject fi, uint64_t* counters,
                                                                int64_t nbytes, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int per_chunk = 16 / elem_size;
  uint32_t flips = 0, affected = 0;
  const int64_t nchunks = (nbytes + 15) >> 4;
  for (int64_t c = tid; c < nchunks; c += nthreads) {
    const int64_t off = c << 4;
    const int valid = (int)min((int64_t)16, nbytes - off);
    uint32_t w[4] = {0, 0, 0, 0};
    if (vec && valid == 16) {
      const uint4 v = ld_stream(in + off);
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
      for (int b = 0; b < valid; ++b) w[b >> 2] |= (uint32_t)in[off + b] << (8 * (b & 3));
    }
    uint32_t m[4] = {0, 0, 0, 0};
    fast_flip_positions(fi, (uint64_t)c, per_chunk * n_bits, [&](int pos) {
      const int el = pos / n_bits, bit = pos - el * n_bits;
      const int gb = el * elem_size * 8 + bit;  // bit index inside the 128-bit chunk
      m[gb >> 5] |= 1u << (gb & 31);
    });
    // elements past the end of the tensor take no flips
    if (valid < 16) {
      for (int b = valid; b < 16; ++b) m[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      flips += __popc(m[k]);
      if (elem_size == 4) affected += (m[k] != 0);
      else { uint32_t t = m[k] | (m[k] >> 4); t |= t >> 2; t |= t >> 1; affected += __popc(t & M01); }
      w[k] ^= m[k];
    }
    if (vec && valid == 16) st_stream(out + off, make_uint4(w[0], w[1], w[2], w[3]));
    else for (int b = 0; b < valid; ++b) out[off + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
  }
  const int idx[2] = {ECCQ_CTR_FLIPPED, ECCQ_CTR_AFFECTED};
  const uint32_t val[2] = {flips, affected};
  block_accumulate<2>(counters, idx, val);
}

// Windowed warp-cooperative version (warp_fill_flips, eccq_common.cuh) for the codeword formats of the path: every warp
// owns windows of kInjWindow x 32 consecutive 16-byte chunks; the chunk loads are issued first, the window's flips are
// generated into shared memory (one Philox call per eight flips), then masks are applied and stored.  The stream is keyed by
// (window, lane), so the result does not depend on the grid.  "Affected" elements are counted when a flip lands on an
// element whose mask was still empty (the atomicOr returns the previous word).
template <int ES, int NBITS>
__global__ void __launch_bounds__(kThreads) inject_fast2_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, FastInject fi,
                                                                 uint64_t* counters, int64_t nchunks) {
  constexpr int SW = 4;                               // mask words per 16-byte chunk
  constexpr uint32_t NPOS = (16 / ES) * NBITS;        // candidate bits per chunk
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ __align__(16) uint32_t s_flip[(kThreads / 32) * kInjWindow * 32 * SW];
  for (int i = threadIdx.x; i < (kThreads / 32) * kInjWindow * 32 * SW; i += blockDim.x) s_flip[i] = 0;
  __syncthreads();
  uint32_t* smask = s_flip + (threadIdx.x >> 5) * kInjWindow * 32 * SW;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nwin = (nchunks + kInjWindow * 32 - 1) / (kInjWindow * 32);
  uint32_t flips = 0, affected = 0;
  for (int64_t win = warp; win < nwin; win += nwarps) {
    const int64_t c0 = win * (kInjWindow * 32);
    const uint32_t nslot = (uint32_t)min((int64_t)kInjWindow * 32, nchunks - c0);
    uint4 v[kInjWindow];
#pragma unroll
    for (int w = 0; w < kInjWindow; ++w)
      if ((uint32_t)(w * 32 + lane) < nslot) v[w] = ld_stream(in + ((c0 + w * 32 + lane) << 4));
    GapRng4 rng;
    rng.init(fi, (uint32_t)win, (uint32_t)(win >> 32) * 32u + (uint32_t)lane);
    uint32_t first[4];
    rng.draw(first, true);
    uint32_t carry = __shfl_sync(FULL, first[0], 0) - 1u;
    uint32_t hit = 0;
    flips += warp_fill_flips<NPOS, SW>(rng, carry, smask, nslot * NPOS, lane, [&](uint32_t, uint32_t local, uint32_t& wd, uint32_t& b) {
      const uint32_t el = local / NBITS;
      const uint32_t idx = el * (8 * ES) + (local - el * NBITS);   // bit index inside the 128-bit chunk
      wd = idx >> 5; b = idx & 31u;
      return true;
    }, [&](uint32_t old, uint32_t b) {
      const uint32_t emask = ES == 4 ? 0xFFFFFFFFu : 0xFFu << (b & 24u);
      hit += (old & emask) == 0u;
    });
    affected += hit;
#pragma unroll
    for (int w = 0; w < kInjWindow; ++w) {
      if ((uint32_t)(w * 32 + lane) < nslot) {
        uint4* slot = reinterpret_cast<uint4*>(smask + (w * 32 + lane) * SW);
        const uint4 m = *slot;
        *slot = make_uint4(0, 0, 0, 0);
        st_stream(out + ((c0 + w * 32 + lane) << 4), make_uint4(v[w].x ^ m.x, v[w].y ^ m.y, v[w].z ^ m.z, v[w].w ^ m.w));
      }
    }
    __syncwarp();
  }
  const int idx[2] = {ECCQ_CTR_FLIPPED, ECCQ_CTR_AFFECTED};
  const uint32_t val[2] = {flips, affected};
  block_accumulate<2>(counters, idx, val);
}

// ------------------------------------------------------------------------------------------------ interpolation
// [outer, len, inner] view; neighbours at +-inner. VEC: inner % 16 == 0 and 16-byte aligned bases.
// FAST32: fewer than 2^31 16-byte chunks, so the token index comes from two multiply-shift divisions instead of two 64-bit
// divisions per chunk (the 64-bit ones cost more than the rest of the kernel).
template <bool VEC, bool FAST32>
__global__ void __launch_bounds__(kThreads) interpolate_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ et,
                                                                uint8_t* __restrict__ out, int64_t outer, int64_t len,
                                                                int64_t inner, FastDiv div_inner16, FastDiv div_len) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = outer * len * inner;
  // every output is clamp(v + 0.5, 0, 15) truncated (interpolation_triton.py:156-157): clean values above 15 saturate
  if (VEC) {
    const int64_t n16 = total >> 4;
    const int64_t inner16 = inner >> 4;
    for (int64_t i = tid; i < n16; i += nthreads) {
      int64_t t;
      if (FAST32) {
        const uint32_t row = div_inner16.div((uint32_t)i);
        t = (int64_t)(row - div_len.div(row) * (uint32_t)len);
      } else {
        t = (i / inner16) % len;
      }
      const uint4 v = ld_stream(q + (i << 4));
      const uint4 e = ld_stream(et + (i << 4));
      uint32_t vv[4] = {v.x, v.y, v.z, v.w};
      const uint32_t ee[4] = {e.x, e.y, e.z, e.w};
      uint32_t dbl[4], any = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // etype byte == 2  <=>  (e ^ 0x02) == 0 per byte
        const uint32_t x = ee[k] ^ 0x02020202u;
        uint32_t z = x | (x >> 4); z |= z >> 2; z |= z >> 1;
        dbl[k] = ~z & M01;
        any |= dbl[k];
      }
      // saturate: any byte > 15 -> 15
      uint32_t big = (vv[0] | vv[1] | vv[2] | vv[3]) & 0xF0F0F0F0u;
      if (big) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t hi = vv[k] & 0xF0F0F0F0u; hi |= hi >> 1; hi |= hi >> 2; hi = (hi >> 4) & M01;  // 1 where byte > 15
          vv[k] = (vv[k] & ~spread8(hi)) | (spread8(hi) & M0F);
        }
      }
      if (any) {
        const int64_t il = (t > 0) ? i - inner16 : i;
        const int64_t ir = (t + 1 < len) ? i + inner16 : i;
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(q + (il << 4))), r = __ldg(reinterpret_cast<const uint4*>(q + (ir << 4)));
        const uint32_t ll[4] = {l.x, l.y, l.z, l.w}, rr[4] = {r.x, r.y, r.z, r.w};
        if (((l.x | l.y | l.z | l.w | r.x | r.y | r.z | r.w) & 0xF0F0F0F0u) == 0) {
          // neighbours are INT4 codes (the only case the cache path produces): SWAR average, no byte loop
#pragma unroll
          for (int k = 0; k < 4; ++k) vv[k] = interp4(vv[k], ll[k], rr[k], dbl[k]);
        } else {
          const uint8_t* lb = reinterpret_cast<const uint8_t*>(ll);
          const uint8_t* rb = reinterpret_cast<const uint8_t*>(rr);
          uint8_t* ob = reinterpret_cast<uint8_t*>(vv);
          const uint8_t* db = reinterpret_cast<const uint8_t*>(dbl);
          for (int b = 0; b < 16; ++b)
            if (db[b]) { const uint32_t s = ((uint32_t)lb[b] + rb[b] + 1u) >> 1; ob[b] = (uint8_t)min(s, 15u); }
        }
      }
      st_stream(out + (i << 4), make_uint4(vv[0], vv[1], vv[2], vv[3]));
    }
  } else {
    for (int64_t i = tid; i < total; i += nthreads) {
      const int64_t t = (i / inner) % len;
      uint32_t v = q[i];
      if (et[i] == 2) {
        const uint32_t l = q[t > 0 ? i - inner : i], r = q[t + 1 < len ? i + inner : i];
        v = (l + r + 1u) >> 1;
      }
      out[i] = (uint8_t)min(v, 15u);
    }
  }
}

// inner == 1 (interpolation along the contiguous axis, the reference's 1-D / 2-D case) with len % 16 == 0: a lane owns 16
// consecutive positions of one row; the neighbour streams are the chunk shifted by one byte, completed by the byte before /
// after the chunk (index-clamped at the row ends).
__global__ void __launch_bounds__(kThreads) interpolate_rows_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ et,
                                                                     uint8_t* __restrict__ out, int64_t n16, FastDiv div_len16, uint32_t len16) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += nthreads) {
    const uint4 v = ld_stream(q + (i << 4));
    const uint4 e = ld_stream(et + (i << 4));
    uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    const uint32_t ee[4] = {e.x, e.y, e.z, e.w};
    uint32_t dbl[4], any = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x = ee[k] ^ 0x02020202u;
      uint32_t z = x | (x >> 4); z |= z >> 2; z |= z >> 1;
      dbl[k] = ~z & M01;
      any |= dbl[k];
    }
    if (any) {
      const uint32_t c = (uint32_t)i - div_len16.div((uint32_t)i) * len16;      // chunk index inside the row
      const uint32_t prev = c > 0 ? (uint32_t)__ldg(q + (i << 4) - 1) : (v.x & 0xFFu);
      const uint32_t next = c + 1 < len16 ? (uint32_t)__ldg(q + (i << 4) + 16) : (v.w >> 24);
      const uint32_t ll[4] = {(v.x << 8) | prev, __funnelshift_l(v.x, v.y, 8), __funnelshift_l(v.y, v.z, 8), __funnelshift_l(v.z, v.w, 8)};
      const uint32_t rr[4] = {__funnelshift_r(v.x, v.y, 8), __funnelshift_r(v.y, v.z, 8), __funnelshift_r(v.z, v.w, 8), (v.w >> 8) | (next << 24)};
      if (((v.x | v.y | v.z | v.w | (prev << 4) | (next << 4)) & 0xF0F0F0F0u) == 0 && prev < 16u && next < 16u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) vv[k] = interp4(vv[k], ll[k], rr[k], dbl[k]);
      } else {
        const uint8_t* lb = reinterpret_cast<const uint8_t*>(ll);
        const uint8_t* rb = reinterpret_cast<const uint8_t*>(rr);
        uint8_t* ob = reinterpret_cast<uint8_t*>(vv);
        const uint8_t* db = reinterpret_cast<const uint8_t*>(dbl);
        for (int b = 0; b < 16; ++b)
          if (db[b]) { const uint32_t s2 = ((uint32_t)lb[b] + rb[b] + 1u) >> 1; ob[b] = (uint8_t)s2; }
      }
    }
    // every output is clamp(v + 0.5, 0, 15) truncated: values above 15 saturate
    if ((vv[0] | vv[1] | vv[2] | vv[3]) & 0xF0F0F0F0u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t hi = vv[k] & 0xF0F0F0F0u; hi |= hi >> 1; hi |= hi >> 2; hi = (hi >> 4) & M01;
        vv[k] = (vv[k] & ~spread8(hi)) | (spread8(hi) & M0F);
      }
    }
    st_stream(out + (i << 4), make_uint4(vv[0], vv[1], vv[2], vv[3]));
  }
}

// ------------------------------------------------------------------------------------------------ launch helpers
static inline int grid_for(int64_t work_items, int per_block) {
  const int64_t want = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace eccq

using namespace eccq;

ECCQ_DEFINE_TU_TABLES(codecs)

extern "C" int eccq_h74_encode(const uint8_t* in, uint8_t* out, int64_t n, void* stream) {
  if (n < 0 || (n && (!in || !out))) return fail(ECCQ_EINVAL, "h74_encode: bad arguments");
  if (n == 0) return ECCQ_OK;
  const int vec = aligned16(in) && aligned16(out);
  hamming_encode_kernel<ECCQ_CODEC_H74><<<grid_for(n, kThreads * 16 * kUnroll), kThreads, 0, (cudaStream_t)stream>>>(in, out, n, vec);
  return check_launch("h74_encode");
}

extern "C" int eccq_h84_encode(const uint8_t* in, uint8_t* out, int64_t n, void* stream) {
  if (n < 0 || (n && (!in || !out))) return fail(ECCQ_EINVAL, "h84_encode: bad arguments");
  if (n == 0) return ECCQ_OK;
  const int vec = aligned16(in) && aligned16(out);
  hamming_encode_kernel<ECCQ_CODEC_H84><<<grid_for(n, kThreads * 16 * kUnroll), kThreads, 0, (cudaStream_t)stream>>>(in, out, n, vec);
  return check_launch("h84_encode");
}

extern "C" int eccq_h74_decode(const uint8_t* cw, uint8_t* nib, uint8_t* flag, uint64_t* counters, int64_t n, void* stream) {
  if (n < 0 || (n && (!cw || !nib))) return fail(ECCQ_EINVAL, "h74_decode: bad arguments");
  if (n == 0) return ECCQ_OK;
  const int vec = aligned16(cw) && aligned16(nib) && (!flag || aligned16(flag));
  const int grid = grid_for(n, kThreads * 16 * kUnroll);
  if (flag) hamming_decode_kernel<ECCQ_CODEC_H74, false, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(cw, nib, flag, counters, n, vec);
  else hamming_decode_kernel<ECCQ_CODEC_H74, false, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(cw, nib, flag, counters, n, vec);
  return check_launch("h74_decode");
}

extern "C" int eccq_h84_decode(const uint8_t* cw, uint8_t* nib, uint8_t* etype, uint64_t* counters, int64_t n, int zero_double,
                               void* stream) {
  if (n < 0 || (n && (!cw || !nib))) return fail(ECCQ_EINVAL, "h84_decode: bad arguments");
  if (n == 0) return ECCQ_OK;
  const int vec = aligned16(cw) && aligned16(nib) && (!etype || aligned16(etype));
  const int grid = grid_for(n, kThreads * 16 * kUnroll);
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_double) {
    if (etype) hamming_decode_kernel<ECCQ_CODEC_H84, true, true><<<grid, kThreads, 0, st>>>(cw, nib, etype, counters, n, vec);
    else hamming_decode_kernel<ECCQ_CODEC_H84, true, false><<<grid, kThreads, 0, st>>>(cw, nib, etype, counters, n, vec);
  } else {
    if (etype) hamming_decode_kernel<ECCQ_CODEC_H84, false, true><<<grid, kThreads, 0, st>>>(cw, nib, etype, counters, n, vec);
    else hamming_decode_kernel<ECCQ_CODEC_H84, false, false><<<grid, kThreads, 0, st>>>(cw, nib, etype, counters, n, vec);
  }
  return check_launch("h84_decode");
}

extern "C" int eccq_golay_encode(const uint8_t* trip, int32_t* cw, int64_t n_cw, void* stream) {
  if (n_cw < 0 || (n_cw && (!trip || !cw))) return fail(ECCQ_EINVAL, "golay_encode: bad arguments");
  if (n_cw == 0) return ECCQ_OK;
  if (int rc = ensure_tables()) return rc;
  const int vec = aligned16(trip) && aligned16(cw);
  golay_encode_kernel<<<grid_for(n_cw, kGolayTile), kThreads, 0, (cudaStream_t)stream>>>(trip, cw, n_cw, vec);
  return check_launch("golay_encode");
}

extern "C" int eccq_golay_decode(const int32_t* cw, uint8_t* trip, uint8_t* errcnt, uint64_t* counters, int64_t n_cw,
                                 void* stream) {
  if (n_cw < 0 || (n_cw && (!cw || !trip))) return fail(ECCQ_EINVAL, "golay_decode: bad arguments");
  if (n_cw == 0) return ECCQ_OK;
  if (int rc = ensure_tables()) return rc;
  const int vec = aligned16(cw) && aligned16(trip) && (!errcnt || (reinterpret_cast<uintptr_t>(errcnt) & 3u) == 0);
  const int64_t want = (n_cw + kGolayTile - 1) / kGolayTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 4));
  golay_decode_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(cw, trip, errcnt, counters, n_cw, vec);
  return check_launch("golay_decode");
}

extern "C" int eccq_xor_mask(const void* in, const void* mask, void* out, int elem_size, int n_bits, uint64_t* counters,
                             int64_t n, void* stream) {
  if (n < 0 || (elem_size != 1 && elem_size != 4) || n_bits < 1 || n_bits > 8 * elem_size || (n && (!in || !mask || !out)))
    return fail(ECCQ_EINVAL, "xor_mask: bad arguments");
  if (n == 0) return ECCQ_OK;
  const uint32_t low = n_bits >= 32 ? 0xFFFFFFFFu : ((1u << n_bits) - 1u);
  const uint32_t keep = elem_size == 4 ? low : low * M01;
  const int64_t nbytes = n * elem_size;
  const int vec = aligned16(in) && aligned16(mask) && aligned16(out);
  xor_mask_kernel<<<grid_for(nbytes, kThreads * 32), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)in, (const uint8_t*)mask, (uint8_t*)out, elem_size, keep, counters, nbytes, vec);
  return check_launch("xor_mask");
}

extern "C" int eccq_inject_philox_ref(const void* in, void* out, uint8_t* flipcnt, int elem_size, int n_bits, float ber,
                                      int64_t seed, int vectorized, uint64_t* counters, int64_t n, void* stream) {
  if (n < 0 || n > 0x7FFFFFFF || (elem_size != 1 && elem_size != 4) || n_bits < 1 || n_bits > (elem_size == 1 ? 8 : 24) ||
      (n && (!in || !out)))
    return fail(ECCQ_EINVAL, "inject_philox_ref: bad arguments");
  if (n == 0) return ECCQ_OK;
  const int grid = grid_for(n, kThreads);
  if (elem_size == 1)
    inject_ref_kernel<uint8_t><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const uint8_t*)in, (uint8_t*)out, flipcnt, n_bits, ber,
                                                                             (uint32_t)seed, vectorized, counters, n);
  else
    inject_ref_kernel<uint32_t><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const uint32_t*)in, (uint32_t*)out, flipcnt, n_bits,
                                                                              ber, (uint32_t)seed, vectorized, counters, n);
  return check_launch("inject_philox_ref");
}

extern "C" int eccq_inject_philox_fast(const void* in, void* out, int elem_size, int n_bits, float ber, uint64_t seed,
                                       uint64_t* counters, int64_t n, void* stream) {
  if (n < 0 || (elem_size != 1 && elem_size != 4) || n_bits < 1 || n_bits > 8 * elem_size || !(ber > 0.0f) ||
      (n && (!in || !out)))
    return fail(ECCQ_EINVAL, "inject_philox_fast: bad arguments (ber must be > 0)");
  if (n == 0) return ECCQ_OK;
  const FastInject fi = make_fast_inject(ber, seed);
  const int64_t nbytes = n * elem_size;
  const int vec = aligned16(in) && aligned16(out);
  if (vec && nbytes >= 16 && !(ber >= 1.0f)) {   // the codeword formats of the path: windowed warp-cooperative kernel
    const int64_t nchunks = nbytes >> 4;
    const int grid = grid_for(nchunks, (kThreads / 32) * kInjWindow * 32);
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t* ip = (const uint8_t*)in;
    uint8_t* op = (uint8_t*)out;
    bool done = true;
    if (elem_size == 1 && n_bits == 8) inject_fast2_kernel<1, 8><<<grid, kThreads, 0, st>>>(ip, op, fi, counters, nchunks);
    else if (elem_size == 1 && n_bits == 7) inject_fast2_kernel<1, 7><<<grid, kThreads, 0, st>>>(ip, op, fi, counters, nchunks);
    else if (elem_size == 1 && n_bits == 4) inject_fast2_kernel<1, 4><<<grid, kThreads, 0, st>>>(ip, op, fi, counters, nchunks);
    else if (elem_size == 4 && n_bits == 24) inject_fast2_kernel<4, 24><<<grid, kThreads, 0, st>>>(ip, op, fi, counters, nchunks);
    else done = false;
    if (done) {
      const int64_t tail = nbytes - (nchunks << 4);   // < 16 bytes: the per-chunk kernel on its own stream
      if (tail > 0) {
        const FastInject ft = make_fast_inject(ber, seed ^ 0x5BD1E9955BD1E995ull);
        inject_fast_kernel<<<1, 32, 0, st>>>(ip + (nchunks << 4), op + (nchunks << 4), elem_size, n_bits, ft, counters, tail, 0);
      }
      return check_launch("inject_philox_fast");
    }
  }
  inject_fast_kernel<<<grid_for(nbytes, kThreads * 16), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)in, (uint8_t*)out, elem_size, n_bits, fi, counters, nbytes, vec);
  return check_launch("inject_philox_fast");
}

extern "C" int eccq_interpolate(const uint8_t* q, const uint8_t* etype, uint8_t* out, int64_t outer, int64_t len, int64_t inner,
                                void* stream) {
  if (outer < 0 || len < 0 || inner < 0) return fail(ECCQ_EINVAL, "interpolate: bad arguments");
  const int64_t total = outer * len * inner;
  if (total == 0) return ECCQ_OK;
  if (!q || !etype || !out) return fail(ECCQ_EINVAL, "interpolate: null pointer");
  const bool vec = (inner % 16 == 0) && aligned16(q) && aligned16(etype) && aligned16(out);
  if (inner == 1 && len % 16 == 0 && (total >> 4) < 0x7FFFFFFFll && aligned16(q) && aligned16(etype) && aligned16(out)) {
    const uint32_t len16 = (uint32_t)(len >> 4);
    interpolate_rows_kernel<<<grid_for(total, kThreads * 16), kThreads, 0, (cudaStream_t)stream>>>(q, etype, out, total >> 4, FastDiv::make(len16), len16);
    return check_launch("interpolate");
  }
  const FastDiv d16 = FastDiv::make((uint32_t)std::max<int64_t>(1, std::min<int64_t>(inner >> 4, 0x7FFFFFFF)));
  const FastDiv dl = FastDiv::make((uint32_t)std::max<int64_t>(1, std::min<int64_t>(len, 0x7FFFFFFF)));
  if (vec && (total >> 4) < 0x7FFFFFFFll && len < 0x7FFFFFFFll)
    interpolate_kernel<true, true><<<grid_for(total, kThreads * 16), kThreads, 0, (cudaStream_t)stream>>>(q, etype, out, outer, len, inner, d16, dl);
  else if (vec)
    interpolate_kernel<true, false><<<grid_for(total, kThreads * 16), kThreads, 0, (cudaStream_t)stream>>>(q, etype, out, outer, len, inner, d16, dl);
  else
    interpolate_kernel<false, false><<<grid_for(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(q, etype, out, outer, len, inner, d16, dl);
  return check_launch("interpolate");
}

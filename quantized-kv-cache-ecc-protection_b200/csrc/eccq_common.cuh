// eccq_common.cuh -- device-side building blocks shared by every kernel of libeccq_sm100a.so.
//
// All codec arithmetic is SWAR ("SIMD within a register"): four codeword bytes ride in one 32-bit lane and the
// 8-entry Hamming syndrome table of the reference (ecc_codecs/triton_kernels/config.py:131-161) is folded into
// LOP3 boolean logic, so the "table" lives in the instruction stream / registers and nothing is gathered.
// The Golay tables (2 x 64-entry parity split table, 4096-entry syndrome table) live in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "eccq.h"

namespace eccq {

constexpr uint32_t M01 = 0x01010101u;
constexpr uint32_t M0F = 0x0F0F0F0Fu;

// ------------------------------------------------------------------------------------------------ memory
__device__ __forceinline__ uint4 ld_stream(const void* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st_stream(void* p, uint4 v) { __stcs(reinterpret_cast<uint4*>(p), v); }
__device__ __forceinline__ uint2 ld_stream8(const void* p) { return __ldcs(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ void st_stream8(void* p, uint2 v) { __stcs(reinterpret_cast<uint2*>(p), v); }

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cp.async: 16-byte global -> shared copies that bypass the register file (page staging)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// One atomicAdd per CTA per counter: lanes -> warp (redux) -> shared -> global.
template <int NV>
__device__ __forceinline__ void block_accumulate(uint64_t* counters, const int (&idx)[NV], const uint32_t (&val)[NV]) {
  __shared__ uint32_t s_acc[NV];
  if (threadIdx.x < NV) s_acc[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint32_t w = __reduce_add_sync(0xFFFFFFFFu, val[i]);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_acc[i], w);
  }
  __syncthreads();
  if (counters != nullptr && threadIdx.x < NV && s_acc[threadIdx.x])
    atomicAdd(reinterpret_cast<unsigned long long*>(counters) + idx[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ Hamming, 4 bytes/lane
// Encode: parity equations of hamming74_triton.py:62-72 / hamming84_triton.py:83-106 on the low nibble of each byte.
__device__ __forceinline__ uint32_t h74_encode4(uint32_t x) {
  x &= M0F;
  const uint32_t x1 = x >> 1, x2 = x >> 2, x3 = x >> 3;
  const uint32_t p0 = (x ^ x1 ^ x3) & M01;   // d0^d1^d3
  const uint32_t p1 = (x ^ x2 ^ x3) & M01;   // d0^d2^d3
  const uint32_t p2 = (x1 ^ x2 ^ x3) & M01;  // d1^d2^d3
  return x | (p0 << 4) | (p1 << 5) | (p2 << 6);
}
__device__ __forceinline__ uint32_t h84_encode4(uint32_t x) {
  x &= M0F;
  const uint32_t x1 = x >> 1, x2 = x >> 2, x3 = x >> 3;
  const uint32_t p0 = (x ^ x1 ^ x3) & M01;
  const uint32_t p1 = (x ^ x2 ^ x3) & M01;
  const uint32_t p2 = (x1 ^ x2 ^ x3) & M01;
  // overall parity of the 7 Hamming bits: p0^p1^p2 == d3, so it collapses to d0^d1^d2
  const uint32_t po = (x ^ x1 ^ x2) & M01;
  return x | (p0 << 4) | (p1 << 5) | (p2 << 6) | (po << 7);
}

struct Dec4 {
  uint32_t nib;  // corrected nibbles, one per byte
  uint32_t nz;   // 0x01 per byte: syndrome != 0
  uint32_t pm;   // 0xFF per byte: overall-parity mismatch (Hamming(8,4) only)
};

// Syndrome decode of hamming84_triton.py:116-209 on four codeword bytes at once.
//   y = x ^ x<<4 folds the data nibble onto the parity nibble (bit 4+i = c_i ^ c_{i+4}), z = x ^ x<<2 gives c0^c2 (bit 2)
//   and c1^c3 (bit 3), so on bits 4,5,6:  s0 = y0^(c1^c3), s1 = y1^(c0^c2)^(c1^c3), s2 = y2^(c1^c3);
//   overall parity mismatch = y0^y1^y2^y3 lands on bit 7.
//   Only LEFT shifts and constant multiplies are used to build these: they issue as IMAD/IMAD.SHL on the FMA pipe and keep
//   the ALU pipe -- the bottleneck of every decode-side kernel here -- for the boolean work.
//   The 8-entry syndrome table [-1,4,5,0,6,1,2,3] (config.py:131-161) only reaches a DATA bit for syndromes 3,5,6,7
//   (bits 0..3); it is applied as a register-resident byte LUT through PRMT: the four 3-bit syndromes are packed into one
//   selector and index the 8-byte tables {0,0,0,1,0,2,4,8} (bit to flip) and {0,1,1,1,1,1,1,1} (syndrome != 0).
//   PRMT's sign-replicate mode turns the parity bit into the 0xFF/0x00 byte mask that gates the correction
//   (only "syndrome != 0 AND parity mismatch" = single error is corrected; double errors keep the raw nibble).
__device__ __forceinline__ uint32_t hamming_syndrome_sel(uint32_t x, uint32_t& y) {
  y = x ^ (x * 16u);
  const uint32_t z = x ^ (x * 4u);
  const uint32_t a = (z & 0x08080808u) * 14u;      // c1^c3 on bits 4,5,6
  const uint32_t b = (z & 0x04040404u) * 8u;       // c0^c2 on bit 5
  const uint32_t s = (y & 0x70707070u) ^ a ^ b;    // 3-bit syndrome on bits 4..6 of each byte
  const uint32_t u = s >> 4;
  return __byte_perm(u | (u >> 4), 0u, 0x4420);    // low 16 bits: the four syndromes as PRMT selector nibbles
}
__device__ __forceinline__ Dec4 h84_decode4(uint32_t x) {
  uint32_t y;
  const uint32_t sel = hamming_syndrome_sel(x, y);
  const uint32_t t = y ^ (y * 4u);
  const uint32_t pe = (t ^ (t * 2u)) & 0x80808080u;
  Dec4 r;
  // msb -> 0xFF / 0x00 per byte: selector nibbles 8|i = "replicate the sign of byte i" (raw prmt; __byte_perm only
  // documents the low three selector bits)
  asm("prmt.b32 %0, %1, %2, 0xBA98;" : "=r"(r.pm) : "r"(pe), "r"(0u));
  const uint32_t flip = __byte_perm(0x01000000u, 0x08040200u, sel);   // syndrome -> data-bit one-hot
  r.nz = __byte_perm(0x01010100u, 0x01010101u, sel);
  r.nib = (x & M0F) ^ (flip & r.pm);
  return r;
}
// error type byte of hamming84_triton.py:185-187: 0 none, 1 single corrected, 2 double detected, 3 parity only
__device__ __forceinline__ uint32_t h84_etype4(const Dec4& d) { const uint32_t pe = d.pm & M01; return pe | ((d.nz ^ pe) << 1); }
__device__ __forceinline__ uint32_t h84_single4(const Dec4& d) { return d.nz & d.pm; }
__device__ __forceinline__ uint32_t h84_double4(const Dec4& d) { return d.nz & ~d.pm; }

// hamming74_triton.py:129-160: bit 7 is ignored, every non-zero syndrome is "corrected".
__device__ __forceinline__ Dec4 h74_decode4(uint32_t x) {
  uint32_t y;
  const uint32_t sel = hamming_syndrome_sel(x, y);
  Dec4 r;
  r.nib = (x & M0F) ^ __byte_perm(0x01000000u, 0x08040200u, sel);
  r.nz = __byte_perm(0x01010100u, 0x01010101u, sel);
  r.pm = 0;
  return r;
}

// ------------------------------------------------------------------------------------------------ Hamming, second formulation
// Pipe rates measured on B200 (tools/pipe_bench.cu): LOP3/PRMT/SHF/F2FP issue at 0.5 warp-instr/clk/SMSP (ALU pipe),
// IMAD/IDP.4A at 0.5 (FMA-heavy pipe), FP32 FFMA/FMUL/FADD/FHADD at ~0.93, and ALU+FMA co-issue reaches ~0.96.  The codec
// kernels are bound by the ALU pipe, so this formulation minimises ALU operations: 11 per four codewords (the first one
// needs 19) + 7 IMAD, with the flags counted by two IDP.4A.
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}
constexpr int kLopXorAnd = 0x28;   // (a ^ b) & c
constexpr int kLopXor3 = 0x96;     // a ^ b ^ c
constexpr int kLopSelect = 0xD8;   // a ^ ((a ^ b) & c) = c ? b : a
// raw PRMT: selector nibble bit 3 = "replicate the sign of the selected byte" (0xFF / 0x00)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Four codewords -> PRMT selector with one nibble per codeword: bits [2:0] = an invertible linear image of the syndrome,
//   i = s0 + 2 (s0^s1) + 4 (s0^s1^s2),  bit 3 = NOT(overall parity mismatch)  (Hamming(7,4): bit 3 = 0).
// With c_j the bits of a byte, yc = (x ^ x<<4) & 0xF0 folds data onto parity (bit 4+j = c_{4+j}^c_j), z = x ^ x<<2 gives
// u = c1^c3 (bit 3) and v = c0^c2 (bit 2).  W = yc ^ yc<<2 ^ u*0x30 ^ v*0x20 (^ 0x80) has bit4 = s0, bit5 = s1,
// bit6 = s0^s2 = y4^y6, bit7 = NOT(y5^y7); one more fold W ^ W<<1 yields the nibble above on bits 4..7 of every byte with
// the low nibble cleared, so the four nibbles pack with ONE multiply (x17: bytes 1 and 3 become the nibble pairs).
// Left shifts / multiplies issue as IMAD on the FMA pipe.
template <int CODEC>
__device__ __forceinline__ uint32_t hamming_sel4(uint32_t x) {
  constexpr uint32_t keep = CODEC == ECCQ_CODEC_H84 ? 0xF0F0F0F0u : 0x70707070u;
  const uint32_t yc = lop3<kLopXorAnd>(x, x * 16u, keep);
  const uint32_t z = x ^ (x * 4u);
  const uint32_t a2 = (z & 0x08080808u) * 6u + (CODEC == ECCQ_CODEC_H84 ? 0x80808080u : 0u);
  const uint32_t b2 = (z & 0x04040404u) * 8u;
  const uint32_t w = lop3<kLopXor3>(yc ^ (yc * 4u), a2, b2);
  const uint32_t selw = lop3<kLopXorAnd>(w, w * 2u, keep);
  return prmt(selw * 17u, 0u, 0x4431u);
}
// flip one-hot of the data bit per index i (0 for parity-bit syndromes); every byte has MSB 0, so in sign-replicate mode
// (parity matches: clean or double error) the lookup returns 0 and nothing is corrected.
constexpr uint32_t kFlipLutLo = 0x02040100u, kFlipLutHi = 0x00000800u;
// flag byte: 0x00 syndrome 0 (clean / parity-bit only), 0x80 single error corrected, 0xFF double error detected
// (entries 0x80 for i != 0: read as is when the parity mismatches, sign-replicated to 0xFF when it matches).
constexpr uint32_t kFlagLutLo = 0x80808000u, kFlagLutHi = 0x80808080u;

// hamming84_triton.py:116-209 / hamming74_triton.py:129-160 on four codewords: nib = corrected nibbles (one per byte,
// double errors keep the raw nibble), flag as above (Hamming(7,4): 0x80 = corrected, never 0xFF).
template <int CODEC>
__device__ __forceinline__ void hamming_decode4(uint32_t x, uint32_t& nib, uint32_t& flag) {
  const uint32_t sel = hamming_sel4<CODEC>(x);
  const uint32_t flip = prmt(kFlipLutLo, kFlipLutHi, sel);
  flag = prmt(kFlagLutLo, kFlagLutHi, sel);
  nib = lop3<kLopXorAnd>(x, flip, M0F);
}
// flag bytes are 0x80 (single) / 0xFF (double): unsigned byte sum = 128 ns + 255 nd, signed byte sum = -128 ns - nd
struct FlagCount {
  uint32_t tu = 0;
  int32_t ts = 0;
  __device__ __forceinline__ void add(uint32_t flag) {
    tu = __dp4a(flag, 0x01010101u, tu);
    ts = __dp4a((int32_t)flag, (int32_t)0x01010101, ts);
  }
  __device__ __forceinline__ uint32_t doubles() const { return (uint32_t)((int32_t)tu + ts) / 254u; }
  __device__ __forceinline__ uint32_t singles() const { return (tu - 255u * doubles()) >> 7; }
};
// error type byte of hamming84_triton.py:185-187 from (flag, parity): 0 none, 1 single, 2 double, 3 parity bit only
// (only the flat decode kernel materialises it)

// Encode four nibbles given as a PRMT selector (nibble e = value e): codeword(n) = LUT8[n & 7] ^ (n & 8 ? 0x78 : 0)
// (generator row of data bit 3 is 0x78 for both Hamming codes).  5 ALU operations per four codewords.
template <int CODEC>
__device__ __forceinline__ uint32_t hamming_encode_sel4(uint32_t sel) {
  constexpr uint32_t lo = CODEC == ECCQ_CODEC_H84 ? 0x63D2B100u : 0x63523100u;   // codewords of 0..3
  constexpr uint32_t hi = CODEC == ECCQ_CODEC_H84 ? 0x873655E4u : 0x07365564u;   // codewords of 4..7
  const uint32_t low = prmt(lo, hi, sel & 0x7777u);
  // selector nibble 8 -> sign of 0x80 replicated (0xFF), nibble 0 -> the byte 0x80 itself; "& 0x78" leaves 0x78 / 0
  const uint32_t top = prmt(0x80u, 0u, sel & 0x8888u);
  return low ^ (top & 0x78787878u);
}

// sum of the four byte counters of a SWAR accumulator (each byte < 256)
__device__ __forceinline__ uint32_t bytesum(uint32_t acc) { return (acc * M01) >> 24; }

// 0x01-per-byte flag word -> 0xFF-per-byte select mask
__device__ __forceinline__ uint32_t spread8(uint32_t f01) { return (f01 << 8) - f01; }

// Temporal interpolation of interpolation_triton.py:141-159 on nibble words: floor((l+r)*0.5+0.5) == (l+r+1)>>1 <= 15.
__device__ __forceinline__ uint32_t interp4(uint32_t self, uint32_t left, uint32_t right, uint32_t dbl01) {
  const uint32_t avg = ((left + right + M01) >> 1) & M0F;
  const uint32_t m = spread8(dbl01);
  return (self & ~m) | (avg & m);
}

// ------------------------------------------------------------------------------------------------ Golay(24,12)
// Parity of data d (12 bit) = XOR of B rows of the set bits (linear code) = T_lo[d & 63] ^ T_hi[d >> 6];
// B rows from config.py:329-347 (row i as a 12-bit mask, bit j = B[i][j]).  B is symmetric, so the same split table
// gives the syndrome: s = parity(d) ^ (cw >> 12)  (golay_triton.py:161-189 with H = [B^T | I]).
struct GolayTables {
  uint16_t par_lo[64];
  uint16_t par_hi[64];
  int32_t syn[4096];  // config.py:403-457; -1 = uncorrectable
};
// No relocatable device code: every translation unit that reads the tables owns a copy of the symbol and registers an
// uploader (ECCQ_DEFINE_TU_TABLES, eccq_internal.h); ensure_tables() fills all copies once per device.
static __device__ GolayTables g_golay;

__device__ __forceinline__ uint32_t golay_parity(const uint16_t* lo, const uint16_t* hi, uint32_t d) {
  return (uint32_t)lo[d & 63u] ^ (uint32_t)hi[(d >> 6) & 63u];
}
__device__ __forceinline__ uint32_t golay_encode1(const uint16_t* lo, const uint16_t* hi, uint32_t d12) {
  return d12 | (golay_parity(lo, hi, d12) << 12);
}
// returns corrected data (12 bit); cnt = bits corrected (0..3) or 4 when uncorrectable (data preserved,
// golay_triton.py:262-286)
__device__ __forceinline__ uint32_t golay_decode1(const uint16_t* lo, const uint16_t* hi, const int32_t* syn, uint32_t cw,
                                                  uint32_t& cnt) {
  const uint32_t d = cw & 0xFFFu;
  const uint32_t s = golay_parity(lo, hi, d) ^ ((cw >> 12) & 0xFFFu);
  const int32_t pat = syn[s];
  const bool ok = pat >= 0;
  cnt = ok ? (uint32_t)__popc(pat) : 4u;
  return ok ? ((cw ^ (uint32_t)pat) & 0xFFFu) : d;
}
__device__ __forceinline__ void golay_load_parity(uint16_t* s_lo, uint16_t* s_hi) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    s_lo[i] = g_golay.par_lo[i];
    s_hi[i] = g_golay.par_hi[i];
  }
}
__device__ __forceinline__ void golay_load_syndrome(int32_t* s_syn) {
  const int4* src = reinterpret_cast<const int4*>(g_golay.syn);
  int4* dst = reinterpret_cast<int4*>(s_syn);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------ Philox-4x32-10
// Exactly triton/language/random.py:philox_impl (uint32 branch): round constants 0xD2511F53 / 0xCD9E8D57,
// key schedule 0x9E3779B9 / 0xBB67AE85.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hiA = __umulhi(0xD2511F53u, c.x), loA = 0xD2511F53u * c.x;
    const uint32_t hiB = __umulhi(0xCD9E8D57u, c.z), loB = 0xCD9E8D57u * c.z;
    c = make_uint4(hiB ^ c.y ^ k.x, loB, hiA ^ c.w ^ k.y, loA);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// triton uint_to_uniform_float: int32 view, x<0 -> -x-1, * 4.6566127342e-10f
__device__ __forceinline__ float triton_uniform(uint32_t r) {
  int32_t x = (int32_t)r;
  x = x < 0 ? ~x : x;
  return __int2float_rn(x) * 4.6566127342e-10f;
}
// tl.rand(seed.to(int32), off.to(int32)): int32 seed sign-extended to 64 bit -> (k0 = lo, k1 = hi); counter (off,0,0,0)
__device__ __forceinline__ uint4 triton_philox(uint32_t key32, uint32_t off) {
  return philox4x32_10(make_uint4(off, 0u, 0u, 0u), make_uint2(key32, (key32 >> 31) ? 0xFFFFFFFFu : 0u));
}
// Flip mask the reference's per-bit kernels produce for element `off` of an N-element launch
// (fault_injection_triton.py:247-334): key = int32(seed*(N*n_bits) + off*n_bits + bit).
__device__ __forceinline__ uint32_t ref_flip_mask(uint32_t seed_lo, uint32_t n, uint32_t off, int n_bits, float ber) {
  const uint32_t base = seed_lo * (n * (uint32_t)n_bits) + off * (uint32_t)n_bits;
  uint32_t m = 0;
  for (int b = 0; b < n_bits; ++b) {
    const uint4 r = triton_philox(base + (uint32_t)b, off);
    m |= (triton_uniform(r.x) < ber ? 1u : 0u) << b;
  }
  return m;
}
// rand4x variant (fault_injection_triton.py:82-215): key = int32(seed*N + off + g*N), word j of call g -> bit 4g+j.
__device__ __forceinline__ uint32_t ref_flip_mask_vec(uint32_t seed_lo, uint32_t n, uint32_t off, int n_bits, float ber) {
  const uint32_t base = seed_lo * n + off;
  uint32_t m = 0;
  for (int g = 0; 4 * g < n_bits; ++g) {
    const uint4 r = triton_philox(base + (uint32_t)g * n, off);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * g + j < n_bits) m |= (triton_uniform(w[j]) < ber ? 1u : 0u) << (4 * g + j);
  }
  return m;
}

// Own stream: i.i.d. Bernoulli(p) over `npos` candidate bit positions of one chunk by geometric gap sampling.
// One Philox call = 4 gaps; expected calls per chunk = ceil((1 + npos*p)/4).  `hit(pos)` is invoked per flip.
struct FastInject {
  float inv_lg2_q;  // 1 / log2(1-p)  (negative); 0 => p >= 1 (flip everything)
  float ber;
  uint32_t seed_lo, seed_hi;
};
template <class Hit>
__device__ __forceinline__ void fast_flip_positions(const FastInject& fi, uint64_t chunk, int npos, Hit hit) {
  if (fi.inv_lg2_q == 0.0f) {
    for (int p = 0; p < npos; ++p) hit(p);
    return;
  }
  float pos = -1.0f;
  uint32_t call = 0;
  for (;;) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)chunk, (uint32_t)(chunk >> 32), call++, 0x45434351u),
                                  make_uint2(fi.seed_lo, fi.seed_hi));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = (float)((w[j] >> 8) + 1u) * 5.9604644775390625e-8f;  // (0,1], 24 bit
      pos += floorf(__log2f(u) * fi.inv_lg2_q) + 1.0f;
      if (!(pos < (float)npos)) return;
      hit((int)pos);
    }
  }
}

// ------------------------------------------------------------------------------------------------ own-stream injection, windowed
// One Bernoulli(p) process runs over the warp's bit stream of a WINDOW of up to kInjWindow warp-iterations: slot s (= iteration
// * 32 + lane) owns NPOS consecutive candidate positions.  Every round each lane draws FOUR geometric gaps from one
// Philox-4x32-10 call, a warp prefix sum turns the 128 gaps into 128 flip positions, and each position inside the window is
// OR-ed into its slot's mask in shared memory (SW words per slot, zero on entry).  The first position past the window is
// carried into the next window (later draws are discarded: the process is memoryless).  At BER 1e-2 a 4-iteration window
// of Hamming(8,4) rows holds ~330 flips = 3 rounds; the cost per value is a quarter of the per-iteration scheme's.
#ifndef ECCQ_INJ_WINDOW
#define ECCQ_INJ_WINDOW 4
#endif
constexpr int kInjWindow = ECCQ_INJ_WINDOW;
struct GapRng4 {
  uint32_t ctr, c0, c1, k0, k1;
  uint32_t spare0, spare1;   // upper halves of the last Philox call (every call feeds two rounds)
  float inv;
  __device__ __forceinline__ void init(const FastInject& fi, uint32_t a, uint32_t b) {
    ctr = 0; c0 = a; c1 = b; k0 = fi.seed_lo; k1 = fi.seed_hi; inv = fi.inv_lg2_q;
  }
  // gap = floor(log2(u) / log2(1-p)) >= 0 with u = 1 - h/65536 in (0,1] from 16 random bits h, saturated at 2^24;
  // increments returned are 1 + gap.  `m` is the float 2^23 + h (PRMT puts the 16 bits under the exponent byte 0x4B), so
  // u is ONE FFMA: (2^23 + h) * -2^-16 + 129.  16-bit uniforms bias the flip rate by < 0.05 % (the tail beyond
  // log(2^-16)/log(1-p) gaps is truncated), far inside the statistical tolerance the injector is held to, and halve the
  // Philox work per flip.
  __device__ __forceinline__ uint32_t inc(uint32_t m) const {
    const float u = fmaf(__uint_as_float(m), -1.52587890625e-5f, 129.0f);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));   // u >= 2^-16: never denormal
    return 1u + (uint32_t)fminf(lg * inv, 16777216.0f);
  }
  static __device__ __forceinline__ uint32_t lo16(uint32_t r) { return __byte_perm(r, 0x4B000000u, 0x7610); }
  static __device__ __forceinline__ uint32_t hi16(uint32_t r) { return __byte_perm(r, 0x4B000000u, 0x7632); }
  // `fresh` must be warp-uniform: true -> new Philox call, false -> the spare halves of the previous one
  __device__ __forceinline__ void draw(uint32_t (&i)[4], bool fresh) {
    if (fresh) {
      const uint4 r = philox4x32_10(make_uint4(c0, c1, ctr++, 0x45434351u), make_uint2(k0, k1));
      i[0] = inc(lo16(r.x)); i[1] = inc(lo16(r.y)); i[2] = inc(lo16(r.z)); i[3] = inc(lo16(r.w));
      spare0 = __byte_perm(r.x, r.y, 0x7632);   // {hi16(r.y), hi16(r.x)}
      spare1 = __byte_perm(r.z, r.w, 0x7632);
    } else {
      i[0] = inc(lo16(spare0)); i[1] = inc(hi16(spare0)); i[2] = inc(lo16(spare1)); i[3] = inc(hi16(spare1));
    }
  }
};
struct NoPlaceHook { __device__ __forceinline__ void operator()(uint32_t, uint32_t) const {} };
// on_place(old_word, bit) is called for every flip placed (old_word = the mask word before the OR)
template <uint32_t NPOS, int SW, class Place, class Hook = NoPlaceHook>
__device__ __forceinline__ uint32_t warp_fill_flips(GapRng4& rng, uint32_t& carry, uint32_t* smask, uint32_t W, int lane, Place place,
                                                    Hook on_place = Hook()) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  if (carry >= W) { carry -= W; return 0; }
  uint32_t base = carry, placed = 0;
  for (bool fresh = true;; fresh = !fresh) {   // every window starts on a fresh call: unused spares are simply dropped
    uint32_t inc[4];
    rng.draw(inc, fresh);
    const uint32_t mine = inc[0] + inc[1] + inc[2] + inc[3];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t f = base + incl - mine;   // position of this lane's first flip of the round
    uint32_t cand = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (f < W) {
        if constexpr (NPOS == SW * 32u) {   // dense layout: the window's bit index IS the mask bit index
          on_place(atomicOr(&smask[f >> 5], 1u << (f & 31u)), f & 31u);
          ++placed;
        } else {
          const uint32_t slot = f / NPOS;
          uint32_t w, b;
          if (place(slot, f - slot * NPOS, w, b)) {   // false: the position does not exist (padding of a ragged layout)
            on_place(atomicOr(&smask[slot * SW + w], 1u << b), b);
            ++placed;
          }
        }
      } else {
        cand = min(cand, f);
      }
      f += inc[q];
    }
    const uint32_t fb = min(__reduce_min_sync(FULL, cand), base + total);
    if (fb >= W) { carry = fb - W; break; }
    base = fb;
  }
  __syncwarp();
  return placed;
}

// ------------------------------------------------------------------------------------------------ float <-> INT4
template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float from(float v) { return v; }
  static __device__ __forceinline__ float to(float v) { return v; }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float from(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half to(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float from(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 to(float v) { return __float2bfloat16_rn(v); }
};

// scale of paged_cache_ecc.py:328-334 (zero_guard 0) or quantization_backends.py:253-257 (zero_guard 1)
__device__ __forceinline__ float make_scale(float amax, int zero_guard) {
  if (zero_guard) return __fdiv_rn(fmaxf(amax, 1e-8f), 7.0f);
  const float s = __fdiv_rn(amax, 7.0f);
  return s == 0.0f ? 1.0f : s;
}
// exact: clamp(rint(x / scale), -8, 7) + 8 with an IEEE divide (ecc_shim.py:575-580; torch.round = half-to-even)
__device__ __forceinline__ uint32_t quant_exact(float x, float scale) {
  float q = rintf(__fdiv_rn(x, scale));
  q = fminf(fmaxf(q, -8.0f), 7.0f);
  return (uint32_t)((int)q + 8) & 0xFu;
}
// Fast path: t = x * (1/scale) differs from the IEEE quotient by a few ulp, which can only change the rounded
// integer when the quotient sits within ~1e-5 of a tie; those (rare) elements are redone with quant_exact.
// Magic-constant rounding: (t + 1.5*2^23) rounds to nearest-even integer, result in the low mantissa bits.
constexpr float kMagic = 12582912.0f;
__device__ __forceinline__ uint32_t quant_fast(float x, float rscale, float scale, float& worst) {
  float t = x * rscale;
  t = fminf(fmaxf(t, -8.25f), 7.25f);
  const float m = t + kMagic;
  const float k = m - kMagic;
  worst = fmaxf(worst, fabsf(t - k));
  (void)scale;
  return (__float_as_uint(m) + 8u);  // low 4 bits valid after the caller masks with 0xF (k in [-8,7])
}
constexpr float kTieMargin = 0.5f - 2e-5f;

// nibble byte k (0..255) -> float(k) exactly via the 2^23 mantissa trick
__device__ __forceinline__ float byte_to_float_m8(uint32_t word, int byte_idx) {
  const uint32_t sel = 0x7440u | (uint32_t)byte_idx;  // {b, 0x00, 0x00, 0x4B} from (word, 0x4B000000)
  const uint32_t bits = __byte_perm(word, 0x4B000000u, sel);
  return __uint_as_float(bits) - 8388616.0f;  // (2^23 + n) - (2^23 + 8)
}

}  // namespace eccq

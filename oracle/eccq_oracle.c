/*
 * eccq_oracle.c -- C restatement of the reference's codec hot path for the CPU side of the benchmark.
 * TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  It is pinned against oracle/eccq_oracle.py (itself pinned against the
 * reference's golden vectors) in tests/test_oracle_c.py.
 *
 * What it restates (reference file:line):
 *   quantise      kv_cache/paged_cache_ecc.py:328-334 + kv_cache/ecc_shim.py:575-580 (absmax/7, 0 -> 1, IEEE divide, rint)
 *   H(8,4) encode ecc_codecs/triton_kernels/hamming84_triton.py:49-108
 *   H(8,4) decode ecc_codecs/triton_kernels/hamming84_triton.py:116-209 (double errors keep the raw nibble)
 *   interpolate   ecc_codecs/triton_kernels/interpolation_triton.py:120-159 along the token axis (ecc_shim.py:1048-1059)
 *   dequantise    ecc_shim.py:1070-1071
 * Rows are independent, so the loops are OpenMP-parallel over rows ("all the host threads it can use").
 *
 * Build: gcc -O3 -march=native -fopenmp -fPIC -shared oracle/eccq_oracle.c -o oracle/_build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static uint8_t ENC84[16];
static uint8_t DEC84_NIB[256], DEC84_TYPE[256];
static int tables_ready = 0;

static uint8_t encode84(uint8_t n) {
  uint8_t d0 = n & 1, d1 = (n >> 1) & 1, d2 = (n >> 2) & 1, d3 = (n >> 3) & 1;
  uint8_t p0 = d0 ^ d1 ^ d3, p1 = d0 ^ d2 ^ d3, p2 = d1 ^ d2 ^ d3;
  uint8_t h7 = (uint8_t)(d0 | (d1 << 1) | (d2 << 2) | (d3 << 3) | (p0 << 4) | (p1 << 5) | (p2 << 6));
  uint8_t p = h7 ^ (h7 >> 4);
  p ^= p >> 2;
  p ^= p >> 1;
  return (uint8_t)(h7 | ((p & 1) << 7));
}

static void decode84(uint8_t c, uint8_t* nib, uint8_t* etype) {
  static const int8_t lut[8] = {-1, 4, 5, 0, 6, 1, 2, 3}; /* config.py:149-161 */
  uint8_t h7 = c & 0x7F, stored = c >> 7;
  uint8_t b[7];
  for (int i = 0; i < 7; ++i) b[i] = (h7 >> i) & 1;
  int syn = (b[0] ^ b[1] ^ b[3] ^ b[4]) | ((b[0] ^ b[2] ^ b[3] ^ b[5]) << 1) | ((b[1] ^ b[2] ^ b[3] ^ b[6]) << 2);
  uint8_t p = h7 ^ (h7 >> 4);
  p ^= p >> 2;
  p ^= p >> 1;
  int perr = stored != (p & 1);
  uint8_t t = syn == 0 ? (perr ? 3 : 0) : (perr ? 1 : 2);
  if (t == 1 && lut[syn] >= 0) h7 ^= (uint8_t)(1u << lut[syn]);
  *nib = h7 & 0x0F;
  *etype = t;
}

void orc_init(void) {
  if (tables_ready) return;
  for (int i = 0; i < 16; ++i) ENC84[i] = encode84((uint8_t)i);
  for (int i = 0; i < 256; ++i) decode84((uint8_t)i, &DEC84_NIB[i], &DEC84_TYPE[i]);
  tables_ready = 1;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The thread count is set by the caller (bench.py passes the host's core count): OMP_NUM_THREADS, which torchrun sets to 1
 * for its workers, must not decide how many cores the CPU baseline uses. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* flat Hamming(8,4) encode of nibbles (hamming84_triton.py:49-108), the first kernel of the config-1 round trip */
void orc_encode_h84(const uint8_t* nib, uint8_t* cw, int64_t n) {
  orc_init();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) cw[i] = ENC84[nib[i] & 0x0F];
}

/* x [rows, D] fp32 -> codewords [rows, D] + scales [rows] */
void orc_quantize_encode_h84(const float* x, int64_t rows, int D, uint8_t* cw, float* scales) {
  orc_init();
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* row = x + r * D;
    float amax = 0.0f;
    for (int e = 0; e < D; ++e) {
      float a = fabsf(row[e]);
      amax = a > amax ? a : amax;
    }
    float s = amax / 7.0f;
    if (s == 0.0f) s = 1.0f;
    scales[r] = s;
    uint8_t* out = cw + r * D;
    for (int e = 0; e < D; ++e) {
      float q = rintf(row[e] / s); /* round-half-even under the default rounding mode == torch.round */
      q = q < -8.0f ? -8.0f : (q > 7.0f ? 7.0f : q);
      out[e] = ENC84[(int)q + 8];
    }
  }
}

/* out = cw ^ mask (the bit-exact injection route) */
void orc_xor_mask(const uint8_t* cw, const uint8_t* mask, uint8_t* out, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = cw[i] ^ mask[i];
}

/* plain decode: nibbles + error types + (corrected, detected) */
void orc_decode_h84(const uint8_t* cw, int64_t n, uint8_t* nib, uint8_t* etype, int64_t* counts) {
  orc_init();
  int64_t c = 0, d = 0;
#pragma omp parallel for schedule(static) reduction(+ : c, d)
  for (int64_t i = 0; i < n; ++i) {
    uint8_t t = DEC84_TYPE[cw[i]];
    nib[i] = DEC84_NIB[cw[i]];
    if (etype) etype[i] = t;
    c += t == 1;
    d += t == 2;
  }
  counts[0] = c;
  counts[1] = d;
}

/* cw [ctx, inner] (inner = Hkv*D, token-major), scales [ctx, Hkv] -> out [ctx, inner] fp32;
 * decode + optional temporal interpolation of double-detected values + dequantise. */
void orc_decode_interp_dequant_h84(const uint8_t* cw, const float* scales, int64_t ctx, int H, int D, int use_interp, float* out,
                                   int64_t* counts) {
  orc_init();
  const int64_t inner = (int64_t)H * D;
  int64_t c = 0, d = 0;
#pragma omp parallel for schedule(static) reduction(+ : c, d)
  for (int64_t t = 0; t < ctx; ++t) {
    const uint8_t* row = cw + t * inner;
    const uint8_t* left = cw + (t > 0 ? t - 1 : 0) * inner;
    const uint8_t* right = cw + (t + 1 < ctx ? t + 1 : ctx - 1) * inner;
    for (int h = 0; h < H; ++h) {
      const float s = scales[t * H + h];
      for (int e = 0; e < D; ++e) {
        const int64_t i = (int64_t)h * D + e;
        const uint8_t ty = DEC84_TYPE[row[i]];
        int n = DEC84_NIB[row[i]];
        c += ty == 1;
        d += ty == 2;
        if (use_interp && ty == 2) {
          float v = ((float)DEC84_NIB[left[i]] + (float)DEC84_NIB[right[i]]) * 0.5f + 0.5f;
          v = v < 0.0f ? 0.0f : (v > 15.0f ? 15.0f : v);
          n = (int)v;
        }
        out[t * inner + i] = ((float)n - 8.0f) * s;
      }
    }
  }
  counts[0] = c;
  counts[1] = d;
}

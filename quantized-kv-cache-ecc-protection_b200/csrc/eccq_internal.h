// eccq_internal.h -- host-side helpers shared by the translation units of libeccq_sm100a.so (error reporting, device info,
// one-time per-device upload of the Golay tables).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "eccq_common.cuh"

namespace eccq {

inline const char*& last_error_ref() {
  static thread_local const char* msg = "";
  return msg;
}
inline int fail(int code, const char* msg) {
  last_error_ref() = msg;
  return code;
}
inline int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    static thread_local char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    last_error_ref() = buf;
    return ECCQ_ECUDA;
  }
  return ECCQ_OK;
}

struct DeviceState {
  bool tables = false;
  int sms = 0;
  uint16_t* pair_tab = nullptr;   // 2 x 256-byte decode tables of the LUT attention kernel (H74, H84), see build_byte_table()
};
inline DeviceState& device_state() {
  static DeviceState st[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return st[dev & 63];
}
inline int sm_count() {
  DeviceState& s = device_state();
  if (s.sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    s.sms = n > 0 ? n : 148;
  }
  return s.sms;
}

// ---- Golay tables, built on the host exactly as the reference builds them ------------------------------------
// B rows: ecc_codecs/triton_kernels/config.py:329-347 (row i, bit j = B[i][j]).
inline const uint16_t* golay_b_rows() {
  static const uint16_t rows[12] = {
      // bit j of rows[i] = B[i][j]
      0xA3B, 0xD1D, 0xE8E, 0xB47, 0xDA3, 0xED1, 0xF68, 0xBB4, 0x9DA, 0x8ED, 0xC76, 0x7FF};
  return rows;
}
inline uint32_t host_golay_parity(uint32_t d) {
  // parity bit i = parity(d & column_i(B)); B is symmetric so column_i == row_i
  const uint16_t* b = golay_b_rows();
  uint32_t p = 0;
  for (int i = 0; i < 12; ++i) p |= (uint32_t)(__builtin_popcount(d & b[i]) & 1) << i;
  return p;
}
inline void build_golay_tables(GolayTables& t) {
  for (uint32_t i = 0; i < 64; ++i) {
    t.par_lo[i] = (uint16_t)host_golay_parity(i);
    t.par_hi[i] = (uint16_t)host_golay_parity(i << 6);
  }
  // config.py:403-457: weight 1, then 2, then 3; first writer wins
  auto syn = [](uint32_t e) { return host_golay_parity(e & 0xFFF) ^ ((e >> 12) & 0xFFF); };
  for (int i = 0; i < 4096; ++i) t.syn[i] = -1;
  t.syn[0] = 0;
  for (int i = 0; i < 24; ++i) t.syn[syn(1u << i)] = (int32_t)(1u << i);
  for (int i = 0; i < 24; ++i)
    for (int j = i + 1; j < 24; ++j) {
      const uint32_t e = (1u << i) | (1u << j);
      if (t.syn[syn(e)] == -1) t.syn[syn(e)] = (int32_t)e;
    }
  for (int i = 0; i < 24; ++i)
    for (int j = i + 1; j < 24; ++j)
      for (int k = j + 1; k < 24; ++k) {
        const uint32_t e = (1u << i) | (1u << j) | (1u << k);
        if (t.syn[syn(e)] == -1) t.syn[syn(e)] = (int32_t)e;
      }
}
inline const GolayTables& host_golay_tables() {
  static GolayTables t;
  static std::once_flag once;
  std::call_once(once, [] { build_golay_tables(t); });
  return t;
}
// ---- byte tables of the LUT attention kernel ---------------------------------------------------------------------
// entry[c] = hb(decode(c)) where decode() is the corrected nibble of the codeword byte (hamming84_triton.py:116-209: single
// errors corrected, double errors keep the raw nibble; hamming74_triton.py:129-160: every non-zero syndrome corrected, bit 7
// ignored) and hb(n) is the HIGH byte of the fp16 number n - 8 -- its low byte is zero for every integer in [-8, 7].
inline uint32_t host_hamming_decode_nibble(uint32_t c, bool h84) {
  static const int syn_to_bit[8] = {-1, 4, 5, 0, 6, 1, 2, 3};   // config.py:131-161
  const uint32_t b0 = c & 1, b1 = (c >> 1) & 1, b2 = (c >> 2) & 1, b3 = (c >> 3) & 1, b4 = (c >> 4) & 1, b5 = (c >> 5) & 1, b6 = (c >> 6) & 1;
  const uint32_t syn = (b4 ^ b0 ^ b1 ^ b3) | ((b5 ^ b0 ^ b2 ^ b3) << 1) | ((b6 ^ b1 ^ b2 ^ b3) << 2);
  const uint32_t perr = __builtin_popcount(c & 0xFF) & 1;   // overall parity mismatch (Hamming(8,4))
  uint32_t n = c & 0xF;
  const bool fix = h84 ? (syn != 0 && perr) : (syn != 0);
  if (fix && syn_to_bit[syn] >= 0 && syn_to_bit[syn] < 4) n ^= 1u << syn_to_bit[syn];
  return n;
}
inline void build_byte_table(uint8_t* t, bool h84) {
  static const uint8_t hb[16] = {0xC8, 0xC7, 0xC6, 0xC5, 0xC4, 0xC2, 0xC0, 0xBC, 0x00, 0x3C, 0x40, 0x42, 0x44, 0x45, 0x46, 0x47};
  for (uint32_t c = 0; c < 256; ++c) t[c] = hb[host_hamming_decode_nibble(c, h84)];
}
const uint16_t* ensure_pair_tables();   // device pointer to the two 256-byte tables [H74][H84]; defined in eccq_lib.cu

int ensure_tables();  // defined in eccq_lib.cu: uploads the Golay tables into every translation unit's copy, once per device

// A translation unit that reads g_golay (eccq_common.cuh) registers an uploader for ITS copy of the symbol.
using TableUploader = cudaError_t (*)(const GolayTables&);
void register_table_uploader(TableUploader f);   // eccq_lib.cu
#define ECCQ_DEFINE_TU_TABLES(name)                                                                                      \
  static cudaError_t eccq_upload_tables_##name(const eccq::GolayTables& t) {                                             \
    return cudaMemcpyToSymbol(eccq::g_golay, &t, sizeof(eccq::GolayTables), 0, cudaMemcpyHostToDevice);                  \
  }                                                                                                                      \
  static const int eccq_tables_registered_##name = (eccq::register_table_uploader(&eccq_upload_tables_##name), 0);

inline FastInject make_fast_inject(float ber, uint64_t seed) {
  FastInject fi;
  fi.seed_lo = (uint32_t)seed;
  fi.seed_hi = (uint32_t)(seed >> 32);
  fi.ber = ber;
  fi.inv_lg2_q = ber >= 1.0f ? 0.0f : (float)(1.0 / std::log2(1.0 - (double)ber));
  return fi;
}

}  // namespace eccq

// eccq_internal.h -- host-side helpers of the single translation unit eccq_lib.cu (error reporting, device info,
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
int ensure_tables();  // defined in eccq_lib.cu next to the __device__ symbol

inline FastInject make_fast_inject(float ber, uint64_t seed) {
  FastInject fi;
  fi.seed_lo = (uint32_t)seed;
  fi.seed_hi = (uint32_t)(seed >> 32);
  fi.ber = ber;
  fi.inv_lg2_q = ber >= 1.0f ? 0.0f : (float)(1.0 / std::log2(1.0 - (double)ber));
  return fi;
}

}  // namespace eccq

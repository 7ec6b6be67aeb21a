// eccq_lib.cu -- library-level state of libeccq_sm100a.so: table upload, device info, error string.  The kernels live in
// their own translation units (codecs, cache write, cache read, attention, tcgen05 attention), compiled in parallel for
// sm_100a only, without relocatable device code.
#include "eccq_internal.h"

namespace eccq {

static std::mutex g_mu;

static std::vector<TableUploader>& uploaders() {
  static std::vector<TableUploader> v;
  return v;
}
void register_table_uploader(TableUploader f) { uploaders().push_back(f); }

int ensure_tables() {
  DeviceState& s = device_state();
  if (s.tables) return ECCQ_OK;
  std::lock_guard<std::mutex> lock(g_mu);
  if (s.tables) return ECCQ_OK;
  const GolayTables& t = host_golay_tables();
  // synchronous on purpose: happens once per device, before the first kernel that needs the tables
  for (TableUploader up : uploaders()) {
    const cudaError_t e = up(t);
    if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
  }
  s.tables = true;
  return ECCQ_OK;
}

const uint16_t* ensure_pair_tables() {
  DeviceState& s = device_state();
  if (s.pair_tab) return s.pair_tab;
  std::lock_guard<std::mutex> lock(g_mu);
  if (s.pair_tab) return s.pair_tab;
  std::vector<uint16_t> h(256);   // 512 bytes
  build_byte_table(reinterpret_cast<uint8_t*>(h.data()), false);
  build_byte_table(reinterpret_cast<uint8_t*>(h.data()) + 256, true);
  uint16_t* d = nullptr;
  // synchronous on purpose: once per device, before the first launch that needs it
  if (cudaMalloc(&d, h.size() * sizeof(uint16_t)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
  s.pair_tab = d;
  return d;
}

}  // namespace eccq

extern "C" int eccq_abi_version(void) { return ECCQ_ABI_VERSION; }
extern "C" const char* eccq_last_error(void) { return eccq::last_error_ref(); }

extern "C" int eccq_device_info(int device, int* out_host) {
  if (!out_host) return eccq::fail(ECCQ_EINVAL, "device_info: null output");
  cudaDeviceProp p;
  const cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return eccq::fail(ECCQ_ECUDA, cudaGetErrorString(e));
  out_host[0] = p.multiProcessorCount;
  out_host[1] = p.major;
  out_host[2] = p.minor;
  out_host[3] = (int)(p.l2CacheSize >> 20);
  return ECCQ_OK;
}

extern "C" int eccq_golay_syndrome_table(int32_t* table_host) {
  if (!table_host) return eccq::fail(ECCQ_EINVAL, "golay_syndrome_table: null output");
  std::memcpy(table_host, eccq::host_golay_tables().syn, sizeof(int32_t) * 4096);
  return ECCQ_OK;
}

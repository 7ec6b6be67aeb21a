// eccq_cache_geom.cuh -- paged-cache geometry and launch helpers shared by the cache-write, cache-read and attention
// translation units (SimpleBlockManager layout, ecc_shim.py:271-316).
#pragma once
#include "eccq_internal.h"

namespace eccq {

struct CacheGeom {
  uint8_t* k_cache;
  uint8_t* v_cache;
  float* k_scales;
  float* v_scales;
  int L, H, D, bs, codec, cw;  // cw = codewords per token row
  int64_t nb;                  // num_blocks
};

__host__ __device__ __forceinline__ int64_t page_index(const CacheGeom& g, int blk, int layer, int h) {
  return ((int64_t)blk * g.L + layer) * g.H + h;
}

// floor(n / d) for n < 2^31 as umulhi + shift (divisors are launch constants: H, S, block_size)
struct FastDiv {
  uint32_t mul, shr, d;
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f; f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < d) ++lg;
      const uint32_t p = 31 + lg;
      f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
      f.shr = p - 32;
    }
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
};

static inline bool make_geom(const eccq_cache_desc* c, CacheGeom& g) {
  if (!c || !c->k_cache || !c->v_cache || !c->k_scales || !c->v_scales) return false;
  if (c->num_blocks <= 0 || c->num_layers <= 0 || c->num_kv_heads <= 0 || c->head_dim <= 0 || c->block_size <= 0) return false;
  if (c->codec < ECCQ_CODEC_H74 || c->codec > ECCQ_CODEC_INT4) return false;
  g.k_cache = (uint8_t*)c->k_cache; g.v_cache = (uint8_t*)c->v_cache;
  g.k_scales = c->k_scales; g.v_scales = c->v_scales;
  g.L = c->num_layers; g.H = c->num_kv_heads; g.D = c->head_dim; g.bs = c->block_size; g.codec = c->codec;
  g.cw = c->codec == ECCQ_CODEC_GOLAY ? (c->head_dim + 2) / 3 : c->head_dim;
  g.nb = c->num_blocks;
  return true;
}
static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int warps_grid(int64_t warps_needed) {
  const int64_t blocks = (warps_needed + 7) / 8;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count() * 8));
}


// grid = (resident CTAs per SM) x SMs for a grid-stride kernel, capped by the work; occupancy queried once per kernel
template <class K>
static int persistent_grid(K kern, int threads, int64_t ctas_needed) {
  static int per_sm = 0;   // one static per instantiation (= per kernel)
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, 0) != cudaSuccess || n <= 0) n = 2;
    per_sm = n;
  }
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)sm_count() * per_sm));
}

// tuning knobs of the streaming read (defaults chosen on B200; overridable for experiments)
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

}  // namespace eccq

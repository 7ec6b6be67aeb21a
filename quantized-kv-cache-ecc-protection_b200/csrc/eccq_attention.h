// eccq_attention.h -- launch arguments shared by the decode-attention translation units (general / mma.sync kernels in
// eccq_attention.cu, tcgen05 kernel in eccq_attention_tc.cu).
#pragma once
#include "eccq_cache_geom.cuh"

namespace eccq {

struct AttnArgs {
  const void* q;
  void* out;
  const int32_t* block_table;
  const int32_t* context_lens;
  float* ws;        // [nseq, Hq, nsplit, D + 2]  (acc[D], m, l)
  int max_blocks, nseq, Hq, layer, nsplit, chunk;
  float sm_scale;
  int interp, v_use_k_scales, count_errors, force_general;
  // prefill (q_len > 1, causal; eccq_paged_prefill_attention): q / out are [qbatch, Hq, Sq, D]; query t of batch row b sits at
  // position ctx_b - Sq + t; nseq == 1 with qbatch > 1 = one shared context (the reference's batch semantics)
  int Sq = 1, qbatch = 0;
};

// eccq_attention_tc.cu: 1 = the tcgen05 kernel took the launch (partials written, the caller merges), 0 = geometry outside it,
// < 0 = error.  May lower a.nsplit / raise a.chunk.
int try_attn_tc(const CacheGeom& g, AttnArgs& a, int q_dtype, uint64_t* counters, cudaStream_t st);

}  // namespace eccq

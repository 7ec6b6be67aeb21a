// eccq_attention_tc.cu -- ECC-fused paged decode attention on the Blackwell tensor cores (tcgen05 + TMEM + bulk TMA).
//
// Replaces paged_attention_ecc_kernel (kv_cache/attention_ecc.py:264-426) for head_dim 128.  The cache stores one INT4
// value per byte (a Hamming codeword whose corrected low nibble is the value), so after the syndrome decode a K or V page IS
// an unsigned 8-bit matrix operand -- no float conversion of K/V ever happens:
//
//   S_int[128 tokens x 16]  = K[128 x 128 dims]  . Qd^T      tcgen05.mma kind::i8, A = decoded K tile (u8, K-major, 128B swizzle),
//                                                            B = the GQA group's queries as 4 balanced base-128 digits (s8)
//   O_int[128 dims x 16]    = V^T[128 x 128 tok] . Wb        A = decoded V tile (u8, MN-major: the SAME [token][dim] page image),
//                                                            B = softmax weights p*v_scale as 32-bit fixed point, one byte per
//                                                            column (u8, [token][16] rows written with one 16-byte store)
//
// Both products are EXACT integer arithmetic (int32 accumulators in TMEM); the only roundings are the 28-bit query / 32-bit
// weight quantisation, far below the fp32 rounding of the reference's dot products.  The "-8" of (nibble - 8) is folded into
// per-head constants (QK^T) and into the running sum of weights (PV).
//
// One persistent CTA per SM, warp-specialised, all hand-offs through mbarriers:
//   warps 0, 3  producers (K / V): walk the block table, one cp.async.bulk (TMA) per page and scale row into a ring of
//               128-token stages (complete_tx on the stage's "full" barrier)
//   warp 1      MMA issuer (one thread): 4 + 4 tcgen05.mma per tile, tcgen05.commit to the barriers that hand S / O to the
//               softmax warps and the stage back to the producer
//   warp 2      TMEM allocation
//   warps 4-7   softmax: thread = token (TMEM lane) for S, thread = dim for O; online softmax in base 2, fp32 state
//   warps 8..   decode: LDS.128 raw codewords -> SWAR syndrome decode (eccq_common.cuh) -> STS.128 corrected nibbles IN PLACE
//               at the 128-byte-swizzled position the UMMA descriptors expect
// Partials (acc[D], m, l) per split go through the same workspace / attn_merge_kernel as the other attention kernels.
//
// Golay(24,12) caches (CODEC = ECCQ_CODEC_GOLAY, 43 codewords of 3 nibbles per 128-dim row) run in the same pipeline in
// ONE pass: a 2 752-byte Golay page cannot be decoded in place into a 2 048-byte operand page, so the producers only stage the
// scales and every decode warp reads the codewords of its eight token rows straight from the cache in global memory (one
// contiguous 1 376-byte run per part, the next tile's words in flight while this tile is decoded), corrects them with the
// 4 096-entry syndrome table held in shared memory and writes the nibbles into the tile at the swizzled byte positions.
//
// Temporal interpolation of double-detected values (ECCShimConfig.use_interpolation, kv_cache/ecc_shim.py:1038-1059 +
// interpolation_triton.py:120-159) runs in the SAME pass (template parameter INTERP): the decode warps leave bit 6 set on every
// double-detected byte and set bit (row) of their 16-byte column chunk's bitmap.  Interpolation runs along the token axis at a
// fixed dim, so the eight column chunks of a part are independent: once every row of the part is stored (the part's decode
// barrier), decode warp w < 8 rewrites the flagged cells of K chunk w, warp 8 + w those of V chunk w -- (left + right + 1) >> 1
// from the neighbouring rows' pre-interpolation nibbles, SWAR on 16 bytes, all reads of a 32-row group before its writes (no
// cascade) -- and arrives on the barrier the MMA issuer of the part waits for.  A neighbour outside the tile (row -1 / 128: the
// previous / next tile or split) is fetched from the cache in global memory and decoded on the spot (rare: one cell in ~300).
#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include "eccq_attention.h"

namespace eccq {
namespace tc {

constexpr int kTile = 128;                 // tokens per tile (= M of QK^T, K of PV)
constexpr int kD = 128;                    // head_dim served by this kernel
constexpr int kStages = 6;
constexpr int kPartBytes = kTile * kD;     // one K or V tile: 16 KB
constexpr int kStageBytes = 2 * kPartBytes + 2 * kTile * 4;   // K, V, k scales, v scales = 33 KB (multiple of 1024)
constexpr int kQopBytes = 16 * 128;        // 16 digit rows x 128 dims
constexpr int kWBytes = kTile * 16;        // [token][16 weight bytes]
constexpr int kFrontWarps = 12;            // roles 0-11: producer, S issuer, query/TMEM, PV issuer, 8 softmax
constexpr int kSoftmaxWarps = 8;
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
#ifndef ECCQ_TC_GROUPWAIT
#define ECCQ_TC_GROUPWAIT 0
#endif
#ifndef ECCQ_TC_REVERSE_ROLES
#define ECCQ_TC_REVERSE_ROLES 1            // roles are numbered from the LAST warp down: the issue arbiter favours high warp ids, and
#endif                                     // the latency-critical roles (issuers, softmax, producers) should win over the decode warps
constexpr int kHeadsPerPass = 4;           // q-heads per item: 4 heads x 4 digits = N 16
constexpr uint32_t kSuspendNs = 1000000u;   // try_wait suspend-time hint: sleep until the barrier is touched
constexpr uint32_t kSpinLimit = 1u << 24;   // x 4 probes, each sleeping up to the hint or until an arrival: seconds, then trap

// ---- shared-memory map (offsets from a 1024-byte aligned base) -------------------------------------------------------
constexpr int kOffStages = 0;
constexpr int kOffQop = kOffStages + kStages * kStageBytes;
constexpr int kQBufs = 3;                  // query-digit buffers: item i+2 is prepared while items i and i+1 are in flight
constexpr int kOffW = kOffQop + kQBufs * kQopBytes;
constexpr int kOffMisc = kOffW + 2 * kWBytes;
struct Misc {
  uint64_t full[kStages], kdec[kStages], vdec[kStages], empty[kStages];
  uint64_t kfix[kStages], vfix[kStages];   // INTERP: the part's flagged cells are rewritten (one arrival per column-chunk worker warp)
  uint64_t sready[2], sfree[2], wready[2], oready[2], qready[kQBufs];
  float qscale[kQBufs][kHeadsPerPass];        // 2^(e-27) * sm_scale * log2(e) per head (0 for unused heads)
  int32_t qc01[kQBufs][kHeadsPerPass];        // 8 * sum_d (d0*128 + d1)
  int32_t qc23[kQBufs][kHeadsPerPass];        // 8 * sum_d (d2*128 + d3)
  alignas(16) float red[2][kSoftmaxWarps][4];        // rare-path exchange of the softmax warps: max score of the 2 heads, max v scale
  alignas(16) float fin[kSoftmaxWarps][4];           // end-of-item reduction: l and sum(w) of the 2 heads per warp
  uint32_t page_mask[kStages];           // bit j: page j of the stage's tile was loaded
  alignas(16) int32_t rowtab[2][128];                // per producer: [2 segments][32] gather rows, then [2][32] pages (scale rows)
  alignas(16) uint32_t colflag[kStages][2][8][4];    // INTERP: [stage][part][16-byte column chunk]: bit r = token row r holds a double-detected value
  uint32_t pv_fresh[2];                  // 1: PV of this weight buffer starts a new accumulation in TMEM (written before wready)
  uint32_t tmem_base;
};
constexpr int kOffGolay = (kOffMisc + (int)sizeof(Misc) + 15) / 16 * 16;   // Golay variant: 4096 x u16 syndrome table + 2 x 64 x u16 parity tables
constexpr int kGolayBytes = 4096 * 2 + 2 * 64 * 2;
constexpr int kSmemBytes = kOffGolay + kGolayBytes + 1024;   // + alignment slack

struct alignas(64) Args {
  // 2-D tensor maps for tile::gather4: K / V caches as rows of 2048 bytes (256 x u64), scales as rows of block_size floats
  CUtensorMap tm_k, tm_v, tm_ks, tm_vs;
  CacheGeom g;
  const void* q;
  int q_dtype;
  const int32_t* block_table;
  const int32_t* context_lens;
  float* ws;
  const float* vscales;
  int max_blocks, nseq, Hq, layer, nsplit, chunk, npass, count_errors;
  float sm_scale;
};

// ---- optional pipeline trace (build with -DECCQ_TC_TRACE; tools/tc_trace.py): SM clock of every hand-off of CTA 0 ---------------
#ifdef ECCQ_TC_TRACE
__device__ long long g_tc_trace[512 * 16];
#define TC_TRACE(n, ev)                                                                                   \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (n) < 512) g_tc_trace[(n) * 16 + (ev)] = clock64();                            \
  } while (0)
#else
#define TC_TRACE(n, ev) do { } while (0)
#endif
enum { EV_LOAD_ISSUE = 0, EV_DEC_START = 1, EV_DEC_K = 2, EV_DEC_V = 3, EV_S_ISSUE = 4, EV_SM_START = 5, EV_SM_W = 6, EV_PV_ISSUE = 7,
       EV_O_READY = 8, EV_O_DONE = 9, EV_S_WAITK = 10, EV_LOAD_V = 11 };

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory"); }
// Every wait is bounded: a protocol bug traps instead of hanging the GPU.  try_wait carries a suspend-time hint: the warp
// sleeps in hardware (NANOSLEEP on the barrier) and is woken by arrivals, so a waiting role costs a handful of issue slots
// per wait.  Without the hint the probes spin through the issue slots the decode warps need (measured: 0.19 -> 0.27 ms); a
// time-based bound around the hinted probe cost ~10 instructions per wake-up, half of the softmax warps' work.
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.b32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(a), "r"(parity), "r"(kSuspendNs) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = sptr(b);
  for (uint32_t it = 0; it < kSpinLimit; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (mbar_try(a, parity)) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sptr(dst)), "l"(src), "r"(bytes), "r"(sptr(bar)) : "memory");
}
// four rows of a 2-D tensor (row indices r0..r3, column 0) -> 4 x row_bytes contiguous bytes at dst, ONE instruction
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(sptr(dst)), "l"(tm), "r"(sptr(bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(b)) : "memory"); }
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p; }" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// barrier + OR-reduction of a predicate over the `nthreads` threads of named barrier `id`
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
  uint32_t out;
  asm volatile("{ .reg .pred pi, po; setp.ne.u32 pi, %1, 0; bar.red.or.pred po, %2, %3, pi; selp.u32 %0, 1, 0, po; }"
               : "=r"(out) : "r"((uint32_t)pred), "r"(id), "r"(nthreads) : "memory");
  return out != 0;
}
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 @0, LBO >> 4 @16, SBO >> 4 @32, version 1 @46,
// layout type @61 (0 = no swizzle, 2 = 128-byte swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 @4, a/b signedness @7/@10, a/b major @15/@16, N>>3 @17, M>>4 @24
constexpr uint32_t kIdescS = (2u << 4) | (0u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // u8 x s8, K-major both
constexpr uint32_t kIdescO = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // u8 x u8, MN-major both

// ---- work items ---------------------------------------------------------------------------------------------------------
struct Item {
  int seq, kvh, pass, split, c0, c1, tile0, ntiles;
};
__device__ __forceinline__ bool get_item(const Args& a, int item, Item& it) {
  it.split = item % a.nsplit;
  int r = item / a.nsplit;
  it.pass = r % a.npass;
  r /= a.npass;
  it.kvh = r % a.g.H;
  it.seq = r / a.g.H;
  const int ctx = min(a.context_lens[it.seq], a.max_blocks * a.g.bs);
  it.c0 = it.split * a.chunk;
  it.c1 = min(ctx, it.c0 + a.chunk);
  it.tile0 = it.c0 / kTile;
  it.ntiles = it.c1 > it.c0 ? (it.c1 - it.c0 + kTile - 1) / kTile : 0;
  return it.ntiles > 0;
}

struct TileCursor {
  int item, t;
  Item it;
  bool valid;
};
__device__ __forceinline__ void cursor_seek(const Args& a, int nitems, TileCursor& c) {   // c.item = candidate: skip empty items
  while (c.item < nitems && !get_item(a, c.item, c.it)) c.item += gridDim.x;
  c.valid = c.item < nitems;
  c.t = 0;
}
__device__ __forceinline__ void cursor_next(const Args& a, int nitems, TileCursor& c) {
  if (++c.t == c.it.ntiles) {
    c.item += gridDim.x;
    cursor_seek(a, nitems, c);
  }
}

__device__ __forceinline__ float load_q(const void* q, int dtype, int64_t i) {
  if (dtype == ECCQ_F32) return reinterpret_cast<const float*>(q)[i];
  if (dtype == ECCQ_F16) return __half2float(reinterpret_cast<const __half*>(q)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(q)[i]);
}

template <int CODEC, bool COUNT>
__device__ __forceinline__ uint32_t decode_nibbles(uint32_t w, FlagCount& fc, bool counted) {
  if (CODEC == ECCQ_CODEC_H84 || CODEC == ECCQ_CODEC_H74) {
    constexpr int HC = CODEC == ECCQ_CODEC_H74 ? ECCQ_CODEC_H74 : ECCQ_CODEC_H84;
    const uint32_t sel = hamming_sel4<HC>(w);
    if (COUNT) fc.add(counted ? prmt(kFlagLutLo, kFlagLutHi, sel) : 0u);
    return lop3<kLopXorAnd>(w, prmt(kFlipLutLo, kFlipLutHi, sel), M0F);
  }
  return w & M0F;
}

// swizzled byte offset of 16-byte chunk c of token row r inside a K or V tile (128-byte swizzle, K-major / MN-major alike)
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
// INTERP: flip table with the double-error mark (see the decode warps): entries of the non-zero syndromes carry bit 7
constexpr uint32_t kFlipMarkLutLo = kFlipLutLo | 0x80808000u, kFlipMarkLutHi = kFlipLutHi | 0x80808080u;
constexpr uint32_t kMarkBits = 0x40404040u;
// pre-interpolation nibbles of four bytes of a decoded INTERP tile: a marked byte (bit 6) holds its nibble inverted
__device__ __forceinline__ uint32_t mark_nib(uint32_t w) { return (w ^ (((w >> 6) & M01) * 0x0Fu)) & M0F; }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 x;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(addr));
  return x;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 x) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
}

template <int CODEC, bool COUNT, int NDEC, bool INTERP>
__global__ void __launch_bounds__((kFrontWarps + NDEC) * 32, 1) attn_tc_kernel(const __grid_constant__ Args a, uint64_t* counters) {
  extern __shared__ __align__(16) unsigned char eccq_tc_smem[];
  const uint32_t raw = sptr(eccq_tc_smem);
  unsigned char* const sm = eccq_tc_smem + ((1024u - (raw & 1023u)) & 1023u);
  Misc& mi = *reinterpret_cast<Misc*>(sm + kOffMisc);
  const int lane = threadIdx.x & 31;
  const int warp = ECCQ_TC_REVERSE_ROLES ? (kFrontWarps + NDEC - 1) - (int)(threadIdx.x >> 5) : (int)(threadIdx.x >> 5);   // role index
  const int bs = a.g.bs;
  const int ppt = kTile / bs;                       // pages per tile
  const int lgbs = 31 - __clz(bs);                  // block_size is a power of two (16 .. 128)
  const int nitems = a.nseq * a.g.H * a.npass * a.nsplit;
  const int G = a.Hq / a.g.H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&mi.full[s], 2);   // one arrive.expect_tx per producer warp (K, V)
      mbar_init(&mi.kdec[s], NDEC);
      mbar_init(&mi.vdec[s], NDEC);
      mbar_init(&mi.empty[s], 1);
      mbar_init(&mi.kfix[s], 8);
      mbar_init(&mi.vfix[s], 8);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&mi.sready[b], 1);
      mbar_init(&mi.sfree[b], kSoftmaxThreads);
      mbar_init(&mi.wready[b], kSoftmaxThreads);
      mbar_init(&mi.oready[b], 1);
    }
    for (int b = 0; b < kQBufs; ++b) mbar_init(&mi.qready[b], 4);   // one arrival per preparing warp (= per q-head of the pass)
    for (int i = 0; i < kStages * 2 * 8 * 4; ++i) (&mi.colflag[0][0][0][0])[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr bool kGolay = CODEC == ECCQ_CODEC_GOLAY;
  uint16_t* const g_syn = reinterpret_cast<uint16_t*>(sm + kOffGolay);      // data-bit error pattern | weight << 12 | uncorrectable << 15
  uint16_t* const g_lo = g_syn + 4096;
  uint16_t* const g_hi = g_lo + 64;
  if (kGolay) {
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
      const int32_t pat = g_golay.syn[i];
      g_syn[i] = pat < 0 ? (uint16_t)0x8000u : (uint16_t)((pat & 0xFFF) | (__popc(pat) << 12));
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { g_lo[i] = g_golay.par_lo[i]; g_hi[i] = g_golay.par_hi[i]; }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sptr(&mi.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem = mi.tmem_base;   // columns [0,32): two S buffers, [32,48): the O accumulator

  if (warp == 0 || warp == 2) {
    // =============================================================== producers: role 0 = K codewords + K scales, role 2 = V
    // TMA tile::gather4: the four row coordinates of one instruction are four pages straight from the block table, so a
    // 128-token tile is 2 instructions for the codewords and 1-2 for the scales per producer (a plain cp.async.bulk per page
    // costs ~100 cycles of issue latency each -- 32 per tile made the producer the bottleneck).  Every warp of the CTA gets
    // about the same share of issue slots, so the producer's own instruction count per tile bounds the tile period: the
    // block table is therefore resolved for FOUR tiles at a time (one entry per lane, one coalesced load, the next segment's
    // entries fetched before this segment's copies go out) into a small row table in shared memory, and the per-tile work of
    // the issuing lane is two 16-byte reads of that table, the stage's barrier and the gathers.  Pages that are unmapped or
    // past the context are replaced by row 0 (loaded, then masked through page_mask), which keeps the byte count per stage
    // constant.
    const int part = warp == 0 ? 0 : 1;
    const CUtensorMap* tm_c = part ? &a.tm_v : &a.tm_k;
    const CUtensorMap* tm_s = part ? &a.tm_vs : &a.tm_ks;
    const float* scales = part ? a.vscales : a.g.k_scales;
    const int spp = bs >> 4;                        // 2048-byte rows per page
    const int lgspp = lgbs - 4;
    const int lh0 = a.layer * a.g.H, LH = a.g.L * a.g.H;
    const int tl = lane >> 3, pj = lane & 7;        // this lane's tile within the segment, page / row slot within the tile
    int32_t* rowtab = mi.rowtab[part];              // [2 segments][32]: 2048-byte rows; [2][32] after it: pages (scale rows)
    struct Seg { int item, t0, nt; Item it; bool valid; };
    auto seg_first = [&](Seg& sg, int item) {       // first segment of the next non-empty item at or after `item`
      sg.item = item;
      while (sg.item < nitems && !get_item(a, sg.item, sg.it)) sg.item += gridDim.x;
      sg.valid = sg.item < nitems;
      sg.t0 = 0;
      sg.nt = sg.valid ? min(4, sg.it.ntiles) : 0;
    };
    auto seg_next = [&](Seg& sg) {
      if (sg.t0 + sg.nt < sg.it.ntiles) { sg.t0 += sg.nt; sg.nt = min(4, sg.it.ntiles - sg.t0); }
      else seg_first(sg, sg.item + gridDim.x);
    };
    auto fetch = [&](const Seg& sg) -> int {
      if (!sg.valid || tl >= sg.nt || pj >= ppt) return -1;
      const int tok0 = (sg.it.tile0 + sg.t0 + tl) * kTile + pj * bs;
      const int lb = tok0 >> lgbs;
      return (tok0 < sg.it.c1 && lb < a.max_blocks) ? a.block_table[(int64_t)sg.it.seq * a.max_blocks + lb] : -1;
    };
    Seg sg;
    seg_first(sg, blockIdx.x);
    int blk = fetch(sg);
    int n = 0, segno = 0;                           // tiles issued, segments
    while (sg.valid) {
      Seg nx = sg;
      seg_next(nx);
      const int blk_next = fetch(nx);
      const uint32_t vm32 = __ballot_sync(0xFFFFFFFFu, blk >= 0);
      const int page = blk >= 0 ? blk * LH + lh0 + sg.it.kvh : 0;
      // lane (tile, j): 2048-byte row j of the tile = sub-row (j % spp) of page (j / spp) of that tile
      const int row = spp == 1 ? page : (__shfl_sync(0xFFFFFFFFu, page, (lane & ~7) | (pj >> lgspp)) << lgspp) + (pj & (spp - 1));
      int32_t* rt = rowtab + (segno & 1) * 32;
      rt[lane] = row;
      rt[64 + lane] = page;
      __syncwarp();
      if (lane == 0) {
        for (int ti = 0; ti < sg.nt; ++ti, ++n) {
          const int s = n % kStages;
          unsigned char* st = sm + kOffStages + (size_t)s * kStageBytes;
          const int4 ra = *reinterpret_cast<const int4*>(rt + ti * 8), rb = *reinterpret_cast<const int4*>(rt + ti * 8 + 4);
          mbar_wait(&mi.empty[s], ((n / kStages) & 1) ^ 1);
          TC_TRACE(n, part ? EV_LOAD_V : EV_LOAD_ISSUE);
          if (part == 0) mi.page_mask[s] = (vm32 >> (8 * ti)) & 0xFFu;
          if (kGolay) {
            mbar_arrive_tx(&mi.full[s], (uint32_t)(kTile * 4));      // scales only: the decode warps fetch the codewords themselves
          } else {
            mbar_arrive_tx(&mi.full[s], (uint32_t)(kPartBytes + kTile * 4));
            tma_gather4(st + part * kPartBytes, tm_c, ra.x, ra.y, ra.z, ra.w, &mi.full[s]);
            tma_gather4(st + part * kPartBytes + 8192, tm_c, rb.x, rb.y, rb.z, rb.w, &mi.full[s]);
          }
          unsigned char* sdst = st + 2 * kPartBytes + part * kTile * 4;
          if (ppt == 8) {
            tma_gather4(sdst, tm_s, ra.x, ra.y, ra.z, ra.w, &mi.full[s]);
            tma_gather4(sdst + 256, tm_s, rb.x, rb.y, rb.z, rb.w, &mi.full[s]);
          } else {
            const int4 pa = *reinterpret_cast<const int4*>(rt + 64 + ti * 8);
            if (ppt == 4) {
              tma_gather4(sdst, tm_s, pa.x, pa.y, pa.z, pa.w, &mi.full[s]);
            } else {
              bulk_g2s(sdst, scales + (int64_t)pa.x * bs, (uint32_t)(bs * 4), &mi.full[s]);
              if (ppt == 2) bulk_g2s(sdst + bs * 4, scales + (int64_t)pa.y * bs, (uint32_t)(bs * 4), &mi.full[s]);
            }
          }
        }
      }
      n = __shfl_sync(0xFFFFFFFFu, n, 0);
      sg = nx;
      blk = blk_next;
      ++segno;
    }
  } else if (warp == 1 || warp == 3) {
    // =============================================================== MMA issuers (one thread each): role 1 = QK^T, role 3 = PV
    // Two independent issue streams: S(n) goes out as soon as the K half of tile n is decoded and its TMEM buffer has been
    // read; PV(n) as soon as the weights of tile n are written and the V half is decoded.  Neither waits for the other.
    if (lane == 0) {
      struct Cur { int item, t, n, j, ntiles; bool valid; };
      auto seek = [&](Cur& c) {   // c.item points at a candidate: skip empty items
        Item it;
        while (c.item < nitems && !get_item(a, c.item, it)) c.item += gridDim.x;
        c.valid = c.item < nitems;
        if (c.valid) c.ntiles = it.ntiles;
      };
      auto advance = [&](Cur& c) {
        ++c.n;
        if (++c.t == c.ntiles) { c.t = 0; ++c.j; c.item += gridDim.x; seek(c); }
      };
      Cur c{(int)blockIdx.x, 0, 0, 0, 0, false};
      seek(c);
      // descriptors differ from tile to tile only in the start-address field (bits 0-13, units of 16 bytes)
      const uint64_t kdesc0 = smem_desc(sptr(sm + kOffStages), 16, 1024, 2), vdesc0 = smem_desc(sptr(sm + kOffStages + kPartBytes), 16, 1024, 2);
      const uint64_t qdesc0 = smem_desc(sptr(sm + kOffQop), 16, 1024, 2), wdesc0 = smem_desc(sptr(sm + kOffW), 128, 256, 0);
      if (warp == 1) {
        for (; c.valid; advance(c)) {
          const int s = c.n % kStages, b = c.n & 1, qb = c.j % kQBufs;
          if (c.t == 0) mbar_wait(&mi.qready[qb], (c.j / kQBufs) & 1);
          mbar_wait(&mi.sfree[b], ((c.n >> 1) & 1) ^ 1);          // S(n-2) has been read out of this TMEM buffer
          TC_TRACE(c.n, EV_S_WAITK);
          mbar_wait(INTERP ? &mi.kfix[s] : &mi.kdec[s], (c.n / kStages) & 1);
          TC_TRACE(c.n, EV_S_ISSUE);
          tc_after();
          const uint64_t ka = kdesc0 + (uint64_t)(s * (kStageBytes >> 4)), qa = qdesc0 + (uint64_t)(qb * (kQopBytes >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_i8(tmem + b * 16, ka + 2 * k, qa + 2 * k, kIdescS, k > 0);
          tc_commit(&mi.sready[b]);
        }
      } else {
        for (; c.valid; advance(c)) {
          const int s = c.n % kStages, b = c.n & 1;
          mbar_wait(&mi.wready[b], (c.n >> 1) & 1);               // W(n) and its fresh flag are in shared memory
          mbar_wait(INTERP ? &mi.vfix[s] : &mi.vdec[s], (c.n / kStages) & 1);
          TC_TRACE(c.n, EV_PV_ISSUE);
          tc_after();
          const uint32_t keep = mi.pv_fresh[b] ? 0u : 1u;          // accumulate onto the previous tiles' product, or start over
          const uint64_t va = vdesc0 + (uint64_t)(s * (kStageBytes >> 4)), wa = wdesc0 + (uint64_t)(b * (kWBytes >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_i8(tmem + 32, va + 256 * k, wa + 32 * k, kIdescO, k > 0 ? 1u : keep);
          tc_commit(&mi.oready[b]);
          tc_commit(&mi.empty[s]);       // every MMA that reads this stage has completed (S(n) was consumed before W(n) existed)
        }
        // the commits arrive asynchronously: see the last ones land before this CTA's shared memory is given up
        for (int back = min(c.n, kStages); back > 0; --back) mbar_wait(&mi.empty[(c.n - back) % kStages], ((c.n - back) / kStages) & 1);
      }
    }
  } else if (warp >= 4 && warp < 4 + kSoftmaxWarps) {
    // =============================================================== softmax / accumulate
    // Eight warps: group g = 0/1 owns q-heads 2g, 2g+1 of the pass (S columns 8g..8g+7, the same columns of O, bytes 8g..8g+7
    // of every weight row); thread = token row of S and dim row of O inside its TMEM lane quarter.
    const int wi = warp - 4;                              // 0..7
    const int g = wi >> 2;
    const int w = (int)(threadIdx.x >> 5) & 3;            // TMEM lane quarter = physical warp id mod 4
    const int r = w * 32 + lane;
    const uint32_t tlane = (uint32_t)(w * 32) << 16;
    const bool tracer = wi == 0 && lane == 0;
    // A wait that is likely to block: ONE warp of the group polls the mbarrier (each probe costs issue slots: the hinted
    // try_wait wakes every few dozen cycles), the other three sleep in a hardware named barrier, which costs nothing.
    auto group_wait = [&](uint64_t* bar, uint32_t parity) {
#if ECCQ_TC_GROUPWAIT
      if ((wi & 3) == 0) mbar_wait(bar, parity);
      named_bar_sync(2 + g, 128);
#else
      mbar_wait(bar, parity);
#endif
    };
    // ---- query digits.  Warp wi < 4 prepares q-head wi of a pass: lane = 4 consecutive dims; the digits go to the item's
    // buffer (three buffers: item i+2 is prepared when item i is finished, so the S issuer never waits for it and the
    // buffer's previous user, item i-1, is long done).  It used to run on a producer warp, where its ~1000 instructions
    // stalled the K loads for ~10 000 cycles at every item boundary.
    int qitem = blockIdx.x, qj = 0;        // next item to prepare (candidate), its index among the non-empty items
    auto prep_next_q = [&]() {
      Item qi;
      while (qitem < nitems && !get_item(a, qitem, qi)) qitem += gridDim.x;
      if (qitem >= nitems) return;
      if (wi < 4) {
        const int qb = qj % kQBufs;
        const int g0q = qi.pass * kHeadsPerPass;
        const bool used = g0q + wi < G;
        const int64_t qbase = ((int64_t)qi.seq * a.Hq + (int64_t)qi.kvh * G + g0q + wi) * kD + lane * 4;
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (used) {
          if (a.q_dtype == ECCQ_F32) {
            const float4 x = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.q) + qbase);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
          } else {
            const uint2 x = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(a.q) + qbase);
            if (a.q_dtype == ECCQ_F16) {
              const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&x.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&x.y));
              v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
            } else {
              v[0] = __uint_as_float(x.x << 16); v[1] = __uint_as_float(x.x & 0xFFFF0000u);
              v[2] = __uint_as_float(x.y << 16); v[3] = __uint_as_float(x.y & 0xFFFF0000u);
            }
          }
        }
        float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
        // amax = m * 2^ex with m in [0.5, 1): |q| * 2^(27-ex) < 2^27
        int ex = 0;
        if (amax > 0.0f && amax < INFINITY) (void)frexpf(amax, &ex);
        const float up = ldexpf(1.0f, 27 - ex);
        uint32_t pk[4] = {0, 0, 0, 0};   // four dims of digit 0..3 (balanced base-128 digits, most significant first)
        int32_t s01 = 0, s23 = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int32_t x = __float2int_rn(v[e] * up);
          const int32_t d3 = ((x + 64) & 127) - 64; x = (x - d3) >> 7;
          const int32_t d2 = ((x + 64) & 127) - 64; x = (x - d2) >> 7;
          const int32_t d1 = ((x + 64) & 127) - 64; x = (x - d1) >> 7;
          const int32_t d0 = x;
          s01 += d0 * 128 + d1;
          s23 += d2 * 128 + d3;
          pk[0] |= (uint32_t)(d0 & 0xFF) << (8 * e);
          pk[1] |= (uint32_t)(d1 & 0xFF) << (8 * e);
          pk[2] |= (uint32_t)(d2 & 0xFF) << (8 * e);
          pk[3] |= (uint32_t)(d3 & 0xFF) << (8 * e);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s01 += __shfl_xor_sync(0xFFFFFFFFu, s01, o);
          s23 += __shfl_xor_sync(0xFFFFFFFFu, s23, o);
        }
        unsigned char* qop = sm + kOffQop + qb * kQopBytes;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int row = wi * 4 + l;        // B operand row = head * 4 + digit; K-major, 128-byte swizzle
          *reinterpret_cast<uint32_t*>(qop + row * 128 + (((lane >> 2) ^ (row & 7)) << 4) + (lane & 3) * 4) = pk[l];
        }
        if (lane == 0) {
          mi.qscale[qb][wi] = used ? ldexpf(a.sm_scale * 1.4426950408889634f, ex - 27) : 0.0f;
          mi.qc01[qb][wi] = 8 * s01;
          mi.qc23[qb][wi] = 8 * s23;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&mi.qready[qb]);
      }
      qitem += gridDim.x;
      ++qj;
    };
    prep_next_q();
    prep_next_q();
    int n = 0, j = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      Item it;
      const bool live = get_item(a, item, it);
      const int g0 = it.pass * kHeadsPerPass;
      const int ng = min(kHeadsPerPass, G - g0);
      if (!live) {   // empty split: neutral partial
        for (int hh = 0; hh < 2; ++hh) {
          const int gi = 2 * g + hh;
          if (gi < ng) {
            float* wsrow = a.ws + (((int64_t)it.seq * a.Hq + it.kvh * G + g0 + gi) * a.nsplit + it.split) * (kD + 2);
            wsrow[r] = 0.0f;
            if (r == 0) { wsrow[kD] = -INFINITY; wsrow[kD + 1] = 0.0f; }
          }
        }
        continue;
      }
      const int qb = j % kQBufs;
      mbar_wait(&mi.qready[qb], (j / kQBufs) & 1);
      float qs[2]; int32_t c01[2], c23[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) { qs[hh] = mi.qscale[qb][2 * g + hh]; c01[hh] = mi.qc01[qb][2 * g + hh]; c23[hh] = mi.qc23[qb][2 * g + hh]; }
      // Running state.  The PV products of consecutive tiles ACCUMULATE in TMEM (exact int32) for as long as neither the running
      // maxima nor the weight scale move -- after the first tile of an item that is almost always -- so the steady state of a
      // tile is: read S, one barrier-with-OR ("does anyone exceed the running maximum?"), exponentials, one 8-byte store.
      // Lazy rescaling: the running maximum only moves when a score exceeds it by more than kSlack (base-2 units), so weights
      // may reach 2^kSlack; the 32-bit fixed-point format keeps kSlackBits of headroom for that and 28 bits below one.  (With
      // an exact running maximum every other tile of a 16-tile item set a new record for one of the four heads.)  The
      // headroom is paid in precision: every token's weight is rounded to the format's LSB, and over a 9 000-token context
      // those roundings add up like sqrt(N) -- 8 bits of headroom measured 7e-5 against a float64 evaluation, 4 bits 3e-6.
      // A finite sentinel instead of -inf keeps the arithmetic free of special cases.
      constexpr float kNegBig = -1.0e30f;
      constexpr int kSlackBits = 4;
      constexpr float kSlack = (float)kSlackBits;
      float m_run[2], l_thr[2], lw_thr[2], oacc[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) { m_run[hh] = kNegBig; l_thr[hh] = 0.0f; lw_thr[hh] = 0.0f; oacc[hh] = 0.0f; }
      uint32_t eb_run = 0;                 // biased exponent bound of the v scales seen so far
      float wscale = 0.0f, sigma = 0.0f;   // 2^(24-E), 2^(E-24)
      bool o_pending = false;              // TMEM holds products not yet folded into oacc
      int o_waited = n - 1;                // last tile whose PV completion (oready phase) has been consumed
      auto wait_o = [&](int nn) {          // every oready phase is consumed exactly once, in order
        if (o_waited < nn) {
          group_wait(&mi.oready[nn & 1], (nn >> 1) & 1);
          if (tracer) TC_TRACE(nn, EV_O_READY);
          o_waited = nn;
        }
      };
      auto drain_o = [&]() {               // TMEM accumulator (relative to m_run / sigma) -> oacc
        tc_after();
        uint32_t o[8];
        tmem_ld8(tmem + 32 + g * 8 + tlane, o);
        tc_before();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v = __int2float_rn((int32_t)o[4 * hh + 3]);
          v = fmaf(v, 256.0f, __int2float_rn((int32_t)o[4 * hh + 2]));
          v = fmaf(v, 256.0f, __int2float_rn((int32_t)o[4 * hh + 1]));
          v = fmaf(v, 256.0f, __int2float_rn((int32_t)o[4 * hh]));
          oacc[hh] = fmaf(v, sigma, oacc[hh]);
        }
        o_pending = false;
      };
      for (int t = 0; t < it.ntiles; ++t, ++n) {
        const int s = n % kStages, b = n & 1;
        const unsigned char* st = sm + kOffStages + (size_t)s * kStageBytes;
        mbar_wait(&mi.full[s], (n / kStages) & 1);
        const int tok = (it.tile0 + t) * kTile + r;
        const bool valid = tok < it.c1 && ((mi.page_mask[s] >> (r >> lgbs)) & 1u);
        const float ks = reinterpret_cast<const float*>(st + 2 * kPartBytes)[r];
        const float vs = valid ? reinterpret_cast<const float*>(st + 2 * kPartBytes + kTile * 4)[r] : 0.0f;
        group_wait(&mi.sready[b], (n >> 1) & 1);
        if (tracer) TC_TRACE(n, EV_SM_START);
        tc_after();
        uint32_t sr[8];
        tmem_ld8(tmem + b * 16 + g * 8 + tlane, sr);
        tc_before();
        mbar_arrive(&mi.sfree[b]);
        float sc[2];
        bool need = (__float_as_uint(vs) >> 23) > eb_run;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int32_t hi = (int32_t)sr[4 * hh] * 128 + (int32_t)sr[4 * hh + 1] - c01[hh];
          const int32_t lo = (int32_t)sr[4 * hh + 2] * 128 + (int32_t)sr[4 * hh + 3] - c23[hh];
          const float f = fmaf(__int2float_rn(hi), 16384.0f, __int2float_rn(lo));
          sc[hh] = valid ? f * (ks * qs[hh]) : kNegBig;
          need |= sc[hh] > m_run[hh] + kSlack;
        }
        bool fresh = t == 0;
        if (named_bar_or(1, kSoftmaxThreads, need)) {
          // ---- rare: a maximum or the weight scale moves (for any head: PV starts over for all 16 columns at once).  Exchange
          // the tile maxima, fold what TMEM holds into oacc under the old scales, rescale the running sums.
          const float wm0 = warp_max_f32(sc[0]), wm1 = warp_max_f32(sc[1]), wv = warp_max_f32(vs);
          if (lane == 0) *reinterpret_cast<float4*>(&mi.red[b][wi][0]) = make_float4(wm0, wm1, wv, 0.0f);
          named_bar_sync(1, kSoftmaxThreads);
          float tm0 = kNegBig, tm1 = kNegBig, vsmax = 0.0f;
#pragma unroll
          for (int ww = 0; ww < 4; ++ww) {
            const float4 x = *reinterpret_cast<const float4*>(&mi.red[b][g * 4 + ww][0]);
            tm0 = fmaxf(tm0, x.x); tm1 = fmaxf(tm1, x.y); vsmax = fmaxf(vsmax, x.z);
          }
          if (o_pending) {
            wait_o(n - 1);
            drain_o();
          }
          const float tm[2] = {tm0, tm1};
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float mn = fmaxf(m_run[hh], tm[hh]);
            const float al = exp2f(m_run[hh] - mn);     // sentinel - finite underflows to 0, sentinel - sentinel = 0 gives 1
            m_run[hh] = mn;
            l_thr[hh] *= al;
            lw_thr[hh] *= al;
            oacc[hh] *= al;
          }
          // weights p * vs < 2^(kSlackBits + E) (E - 1 = exponent bound of the v scales) as 32-bit fixed point:
          // wscale = 2^(32 - kSlackBits - E), sigma = 1 / wscale
          eb_run = max(eb_run, min(max(__float_as_uint(vsmax) >> 23, 32u - kSlackBits), 254u));
          wscale = __uint_as_float((285u - kSlackBits - eb_run) << 23);
          sigma = __uint_as_float((eb_run - 31u + kSlackBits) << 23);
          fresh = true;
        }
        uint32_t wfix[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float p = valid ? exp2f(sc[hh] - m_run[hh]) : 0.0f;
          l_thr[hh] += p;
          wfix[hh] = __float2uint_rn(p * vs * wscale);
          // the "-8" term uses the weight the tensor core sees, so a weight's rounding error multiplies (nibble - 8), not the nibble
          lw_thr[hh] = fmaf(__uint2float_rn(wfix[hh]), sigma, lw_thr[hh]);
        }
        *reinterpret_cast<uint2*>(sm + kOffW + b * kWBytes + r * 16 + g * 8) = make_uint2(wfix[0], wfix[1]);
        if (r == 0 && g == 0) mi.pv_fresh[b] = fresh ? 1u : 0u;
        fence_async_smem();
        mbar_arrive(&mi.wready[b]);
        if (tracer) TC_TRACE(n, EV_SM_W);
        o_pending = true;
        if (t > 0) wait_o(n - 1);          // long complete by now; keeps the barrier phases in step
      }
      wait_o(n - 1);
      drain_o();
      if (tracer) TC_TRACE(n - 1, EV_O_DONE);
      // ---- end of item: sums over the 128 token threads of the group, write the partial
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          l_thr[hh] += __shfl_xor_sync(0xFFFFFFFFu, l_thr[hh], o);
          lw_thr[hh] += __shfl_xor_sync(0xFFFFFFFFu, lw_thr[hh], o);
        }
      if (lane == 0) *reinterpret_cast<float4*>(&mi.fin[wi][0]) = make_float4(l_thr[0], l_thr[1], lw_thr[0], lw_thr[1]);
      named_bar_sync(1, kSoftmaxThreads);
      float L[2] = {0, 0}, LW[2] = {0, 0};
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float4 x = *reinterpret_cast<const float4*>(&mi.fin[g * 4 + ww][0]);
        L[0] += x.x; L[1] += x.y; LW[0] += x.z; LW[1] += x.w;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int gi = 2 * g + hh;
        if (gi < ng) {
          float* wsrow = a.ws + (((int64_t)it.seq * a.Hq + it.kvh * G + g0 + gi) * a.nsplit + it.split) * (kD + 2);
          wsrow[r] = fmaf(-8.0f, LW[hh], oacc[hh]);
          // the merge kernel works in natural-log units
          if (r == 0) { wsrow[kD] = (m_run[hh] <= kNegBig) ? -INFINITY : m_run[hh] * 0.6931471805599453f; wsrow[kD + 1] = L[hh]; }
        }
      }
      named_bar_sync(1, kSoftmaxThreads);          // mi.fin is reused by the next item; every S of this item has been consumed
      prep_next_q();                               // item j + 2 into the buffer item j - 1 used
      ++j;
    }
  } else if (warp >= kFrontWarps) {
    // =============================================================== decode, in place
    // Per tile and warp: the raw codewords of its K rows and V rows are loaded up front (LDS.128, one 512-byte run = 4 rows
    // per warp instruction), decoded in registers and stored back to the SAME rows at the 128-byte-swizzled chunk position the
    // UMMA descriptors expect; K is published first so QK^T overlaps the decoding of V.
    const int dw = warp - kFrontWarps;
    constexpr int OPS = 32 / NDEC;   // warp-wide 512-byte operations per part (K or V) per warp
    static_assert(NDEC > 0 && 32 % NDEC == 0, "decode warps must divide the 32 row groups of a part");
    uint32_t rd_off[OPS], wr_off[OPS];
#pragma unroll
    for (int i = 0; i < OPS; ++i) {
      const int row = (dw + i * NDEC) * 4 + (lane >> 3), c = lane & 7;
      rd_off[i] = (uint32_t)(row * 128 + (c << 4));
      wr_off[i] = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
    }
    const uint32_t stage0 = sptr(sm + kOffStages);
    FlagCount fc;
    // ---- INTERP: rewrite of the flagged cells of one (part, 16-byte column chunk) of a tile, after the tile's V part
    static_assert(!INTERP || NDEC == 16, "one worker warp per (part, 16-byte column chunk)");
    auto patch_chunk = [&](int s, int wpart, int wc, int wseq, int wkvh, int wc1, int wtokbase) {
          const uint32_t st = stage0 + (uint32_t)(s * kStageBytes);
          const uint4 bm4 = *reinterpret_cast<const uint4*>(&mi.colflag[s][wpart][wc][0]);
          if (bm4.x | bm4.y | bm4.z | bm4.w) {
            const uint32_t stp = st + (uint32_t)(wpart * kPartBytes);
            const uint32_t pmw = mi.page_mask[s];
            const unsigned char* cache = wpart ? a.g.v_cache : a.g.k_cache;
            // pre-interpolation nibbles of chunk wc of token `tok`, straight from the cache (neighbour outside the tile); the
            // context length is only needed here (one cell in ~300): loading it up front put a global round trip in every rewrite
            auto fetch = [&](int tok, uint32_t (&o)[4]) -> bool {
              if (tok < 0 || tok >= min(a.context_lens[wseq], a.max_blocks * bs)) return false;
              const int blk = a.block_table[(int64_t)wseq * a.max_blocks + (tok >> lgbs)];
              if (blk < 0) return false;
              const uint4 xr = *reinterpret_cast<const uint4*>(
                  cache + ((int64_t)(blk * (a.g.L * a.g.H) + a.layer * a.g.H + wkvh) * bs + (tok & (bs - 1))) * kD + wc * 16);
              uint32_t fl;
              hamming_decode4<ECCQ_CODEC_H84>(xr.x, o[0], fl); hamming_decode4<ECCQ_CODEC_H84>(xr.y, o[1], fl);
              hamming_decode4<ECCQ_CODEC_H84>(xr.z, o[2], fl); hamming_decode4<ECCQ_CODEC_H84>(xr.w, o[3], fl);
              return true;
            };
            auto row_ok = [&](int r) { return wtokbase + r < wc1 && ((pmw >> (r >> lgbs)) & 1u) != 0; };
            uint32_t carry[4] = {0, 0, 0, 0};   // pre-interpolation nibbles of row 32 q - 1 when the previous group rewrote it
            bool carry_set = false;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
              const uint32_t bmq = q == 0 ? bm4.x : q == 1 ? bm4.y : q == 2 ? bm4.z : bm4.w;
              if (bmq == 0) { carry_set = false; continue; }
              const int r = 32 * q + lane;
              const bool mine = (bmq >> lane) & 1u;
              uint32_t sw[4] = {0, 0, 0, 0}, up[4] = {0, 0, 0, 0}, dn[4] = {0, 0, 0, 0};
              bool upok = false, dnok = false;
              if (mine) {
                const uint4 x = lds128(stp + swz(r, wc));
                sw[0] = x.x; sw[1] = x.y; sw[2] = x.z; sw[3] = x.w;
                if (r == 0) {
                  upok = fetch(wtokbase - 1, up);
                } else if (lane == 0 && carry_set) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) up[k] = carry[k];
                  upok = row_ok(r - 1);
                } else {
                  const uint4 u = lds128(stp + swz(r - 1, wc));
                  up[0] = mark_nib(u.x); up[1] = mark_nib(u.y); up[2] = mark_nib(u.z); up[3] = mark_nib(u.w);
                  upok = row_ok(r - 1);
                }
                if (r == kTile - 1) {
                  dnok = fetch(wtokbase + kTile, dn);
                } else {
                  const uint4 d = lds128(stp + swz(r + 1, wc));
                  dn[0] = mark_nib(d.x); dn[1] = mark_nib(d.y); dn[2] = mark_nib(d.z); dn[3] = mark_nib(d.w);
                  dnok = row_ok(r + 1);
                }
              }
              __syncwarp();   // every read of the group precedes its writes
              carry_set = (bmq >> 31) & 1u;
#pragma unroll
              for (int k = 0; k < 4; ++k) carry[k] = __shfl_sync(0xFFFFFFFFu, mark_nib(sw[k]), 31);
              if (mine) {
                uint32_t nw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t sn4 = mark_nib(sw[k]);
                  nw[k] = interp4(sn4, upok ? up[k] : sn4, dnok ? dn[k] : sn4, (sw[k] >> 6) & M01);   // a missing neighbour = the value itself
                }
                sts128(stp + swz(r, wc), make_uint4(nw[0], nw[1], nw[2], nw[3]));
              }
              __syncwarp();
            }
            if (lane < 4) mi.colflag[s][wpart][wc][lane] = 0;
            fence_async_smem();
          }
          __syncwarp();
    };
    // Run by every decode warp after the tile's V part: warps 0-7 own K chunk dw (every warp stored its K rows a part ago, so
    // their wait is over), warps 8-15 V chunk dw - 8 (they wait for the slowest warp's V rows: 12 % of the kernel's stall samples).
    // Measured alternatives, all slower: the V rewrite deferred until after the next tile's K part (0.42 ms vs 0.29 ms), ONE
    // rotating V worker per tile that checks all eight chunks (0.38 ms), no waiting at all -- the warp that completes a part's
    // decode barrier (probe + atomicMax election, or a shared counter) rewrites all eight chunks (0.40 ms): whenever one warp
    // serialises the eight chunks, the issuer's barrier completes later and the stage is held longer.
    auto worker = [&](int wn, int wseq, int wkvh, int wc1, int wtokbase) {
      const int s = wn % kStages;
      const int wpart = dw >> 3;
      mbar_wait(wpart ? &mi.vdec[s] : &mi.kdec[s], (wn / kStages) & 1);   // every row of the part is stored (marks included)
      patch_chunk(s, wpart, dw & 7, wseq, wkvh, wc1, wtokbase);
      if (lane == 0) mbar_arrive(wpart ? &mi.vfix[s] : &mi.kfix[s]);
    };
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      Item it;
      if (!get_item(a, item, it)) continue;
      const bool do_count = COUNT && it.pass == 0;
      if constexpr (kGolay) {
        // ---- Golay: rows 8 dw .. 8 dw + 7 of every tile.  A work item is (row, group g of four codewords 4g .. 4g+3) = twelve
        // nibbles = dims 12g .. 12g+11 = THREE aligned 32-bit words of the operand row, each inside one 16-byte chunk (so the
        // swizzle never splits a word); group 10 is the row's tail (codewords 40-42, dims 120-127, two words).  88 items per
        // part and warp = 3 warp iterations.  The four codewords of an item are four 32-bit loads from the cache (stride 16 B
        // across the lanes: the four loads of an iteration touch every byte of the warp's 1 376-byte run once, L1 serves the
        // sectors); as soon as a part is decoded its register set is reloaded with the same part of the NEXT tile.
        static_assert(!kGolay || NDEC == 16, "eight token rows per decode warp");
        constexpr int CW = (kD + 2) / 3, GPR = 11, NITEM = 8 * GPR, NIT = (NITEM + 31) / 32;
        const int r0 = 8 * dw;
        const int LHg = a.g.L * a.g.H, lh0g = a.layer * a.g.H;
        uint32_t ncorr = 0, ndet = 0;
        uint32_t wk[NIT][4], wv[NIT][4];
        auto load_part = [&](int tt, const void* cache_base, uint32_t (&o)[NIT][4]) {
          const int tok0 = (it.tile0 + tt) * kTile + r0;
          const int lb = tok0 >> lgbs;
          const int blk = (tt < it.ntiles && tok0 < it.c1 && lb < a.max_blocks) ? a.block_table[(int64_t)it.seq * a.max_blocks + lb] : -1;
          const uint32_t* cache = reinterpret_cast<const uint32_t*>(cache_base) + ((int64_t)(blk * LHg + lh0g + it.kvh) * bs + (tok0 & (bs - 1))) * CW;
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            const int item = i * 32 + lane;
            const int rr = item / GPR, g = item - rr * GPR;
            const bool in = blk >= 0 && item < NITEM;
#pragma unroll
            for (int k = 0; k < 4; ++k) o[i][k] = (in && 4 * g + k < CW) ? __ldg(cache + rr * CW + 4 * g + k) : 0u;
          }
        };
        auto decode_part = [&](int part, const uint32_t (&w)[NIT][4], uint32_t st, uint32_t pm, int tokbase) {
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            const int item = i * 32 + lane;
            if (item < NITEM) {
              const int rr = item / GPR, g = item - rr * GPR;
              const int row = r0 + rr;
              const bool cnt = COUNT && do_count && tokbase + row < it.c1 && ((pm >> (row >> lgbs)) & 1u);
              uint32_t sp[4];   // the three corrected nibbles of each codeword, one per byte
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t cw = w[i][k];
                uint32_t d = cw & 0xFFFu;
                const uint32_t syn = ((uint32_t)g_lo[d & 63u] ^ (uint32_t)g_hi[d >> 6]) ^ ((cw >> 12) & 0xFFFu);
                const uint32_t e = g_syn[syn];
                d ^= e & 0xFFFu;                               // uncorrectable words (bit 15) carry a zero pattern: data preserved
                if (COUNT && cnt && 4 * g + k < CW) { if (e & 0x8000u) ++ndet; else ncorr += (e >> 12) & 7u; }
                sp[k] = (d & 0xFu) | ((d & 0xF0u) << 4) | ((d & 0xF00u) << 8);
              }
              const uint32_t w0 = sp[0] | (sp[1] << 24), w1 = (sp[1] >> 8) | (sp[2] << 16), w2 = (sp[2] >> 16) | (sp[3] << 8);
              const uint32_t rowbase = st + (uint32_t)(part * kPartBytes + row * 128);
              const int d0 = 12 * g;                           // multiple of 4: every word lies inside one 16-byte chunk
              const uint32_t sx = (uint32_t)(row & 7);
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(rowbase + ((((uint32_t)d0 >> 4) ^ sx) << 4) + ((uint32_t)d0 & 15u)), "r"(w0) : "memory");
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(rowbase + ((((uint32_t)(d0 + 4) >> 4) ^ sx) << 4) + ((uint32_t)(d0 + 4) & 15u)), "r"(w1) : "memory");
              if (g < GPR - 1)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(rowbase + ((((uint32_t)(d0 + 8) >> 4) ^ sx) << 4) + ((uint32_t)(d0 + 8) & 15u)), "r"(w2) : "memory");
            }
          }
        };
        load_part(0, a.g.k_cache, wk);
        load_part(0, a.g.v_cache, wv);
        for (int t = 0; t < it.ntiles; ++t, ++n) {
          const int s = n % kStages;
          const uint32_t st = stage0 + (uint32_t)(s * kStageBytes);
          mbar_wait(&mi.full[s], (n / kStages) & 1);           // the stage is free (and its scales have landed)
          if (dw == 0 && lane == 0) TC_TRACE(n, EV_DEC_START);
          const uint32_t pm = COUNT ? mi.page_mask[s] : 0u;
          const int tokbase = (it.tile0 + t) * kTile;
          decode_part(0, wk, st, pm, tokbase);
          load_part(t + 1, a.g.k_cache, wk);                   // the next tile's K words into the set just consumed
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mi.kdec[s]);
          if (dw == 0 && lane == 0) TC_TRACE(n, EV_DEC_K);
          decode_part(1, wv, st, pm, tokbase);
          load_part(t + 1, a.g.v_cache, wv);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mi.vdec[s]);
          if (dw == 0 && lane == 0) TC_TRACE(n, EV_DEC_V);
        }
        if (COUNT) {
          ncorr = __reduce_add_sync(0xFFFFFFFFu, ncorr); ndet = __reduce_add_sync(0xFFFFFFFFu, ndet);
          if (lane == 0 && counters != nullptr) {
            if (ncorr) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + ECCQ_CTR_CORRECTED, (unsigned long long)ncorr);
            if (ndet) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + ECCQ_CTR_DETECTED, (unsigned long long)ndet);
          }
        }
        continue;
      }
      for (int t = 0; t < it.ntiles; ++t, ++n) {
        const int s = n % kStages;
        const uint32_t st = stage0 + (uint32_t)(s * kStageBytes);
        mbar_wait(&mi.full[s], (n / kStages) & 1);
        if (dw == 0 && lane == 0) TC_TRACE(n, EV_DEC_START);
        const uint32_t pm = COUNT ? mi.page_mask[s] : 0u;
        const int tokbase = (it.tile0 + t) * kTile;
        uint4 x[2][OPS];
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int i = 0; i < OPS; ++i)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[part][i].x), "=r"(x[part][i].y), "=r"(x[part][i].z), "=r"(x[part][i].w)
                         : "r"(st + part * kPartBytes + rd_off[i]));
        __syncwarp();
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
          for (int i = 0; i < OPS; ++i) {
            bool cnt = false;
            if (COUNT) {
              const int row = (dw + i * NDEC) * 4 + (lane >> 3);
              cnt = do_count && tokbase + row < it.c1 && ((pm >> (row >> lgbs)) & 1u);
            }
            uint32_t y0, y1, y2, y3;
            if constexpr (INTERP) {
              // ONE look-up serves the correction and the double-error mark: the flip table with bit 7 set on every entry of a
              // non-zero syndrome.  Parity mismatch (single error): the entry is read as it is, flip bit | 0x80; parity match (PRMT
              // sign-replicate mode): 0x00 for syndrome 0, 0xFF for a non-zero syndrome = double error.  (x & 0x0F) ^ (v & 0x4F)
              // then corrects a single error, and for a double error leaves bit 6 set on the byte with its nibble INVERTED (the
              // byte is rewritten by the interpolation anyway; mark_nib() undoes the inversion where the old value is needed).
              // Two LOP3 + the OR below per word on top of the plain decode (the first version spent a second PRMT, a shift and
              // two LOP3 on a separate flag byte).
              const uint32_t w4[4] = {x[part][i].x, x[part][i].y, x[part][i].z, x[part][i].w};
              uint32_t y4[4], any = 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t sel = hamming_sel4<ECCQ_CODEC_H84>(w4[k]);
                if (COUNT) fc.add(cnt ? prmt(kFlagLutLo, kFlagLutHi, sel) : 0u);
                const uint32_t v = prmt(kFlipMarkLutLo, kFlipMarkLutHi, sel);
                y4[k] = lop3<kLopXorAnd>(w4[k] & M0F, v, 0x4F4F4F4Fu);
                any |= y4[k];
              }
              if (any & kMarkBits) {
                const int row = (dw + i * NDEC) * 4 + (lane >> 3);
                atomicOr(&mi.colflag[s][part][lane & 7][row >> 5], 1u << (row & 31));
              }
              y0 = y4[0]; y1 = y4[1]; y2 = y4[2]; y3 = y4[3];
            } else {
              y0 = decode_nibbles<CODEC, COUNT>(x[part][i].x, fc, cnt);
              y1 = decode_nibbles<CODEC, COUNT>(x[part][i].y, fc, cnt);
              y2 = decode_nibbles<CODEC, COUNT>(x[part][i].z, fc, cnt);
              y3 = decode_nibbles<CODEC, COUNT>(x[part][i].w, fc, cnt);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st + part * kPartBytes + wr_off[i]), "r"(y0), "r"(y1), "r"(y2), "r"(y3) : "memory");
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(part == 0 ? &mi.kdec[s] : &mi.vdec[s]);
          if (dw == 0 && lane == 0) TC_TRACE(n, part == 0 ? EV_DEC_K : EV_DEC_V);
        }
        if constexpr (INTERP) {
          worker(n, it.seq, it.kvh, it.c1, tokbase);
        }
      }
    }
    if (COUNT) {
      const uint32_t ns = __reduce_add_sync(0xFFFFFFFFu, fc.singles()), nd = __reduce_add_sync(0xFFFFFFFFu, fc.doubles());
      if (lane == 0 && counters != nullptr) {
        if (ns) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + ECCQ_CTR_CORRECTED, (unsigned long long)ns);
        if (nd) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + ECCQ_CTR_DETECTED, (unsigned long long)nd);
        if (INTERP && nd) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + ECCQ_CTR_INTERPOLATED, (unsigned long long)nd);
      }
    }
  }
  tc_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// ---- host: tensor maps (encoded once per (buffer, geometry), cached) -----------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// rows x row_bytes bytes at `base`, box = one row; elem_bytes 8 (codeword rows, 256 elements) or 4 (scale rows)
static bool tensor_map_rows(CUtensorMap& out, const void* base, uint64_t rows, uint32_t row_bytes, uint32_t elem_bytes) {
  struct Entry { const void* base; uint64_t rows; uint32_t row_bytes; CUtensorMap tm; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& e : cache)
    if (e.base == base && e.rows == rows && e.row_bytes == row_bytes) { out = e.tm; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {row_bytes / elem_bytes, rows};
  const cuuint64_t gstr[1] = {row_bytes};
  const cuuint32_t box[2] = {row_bytes / elem_bytes, 1};
  const cuuint32_t estr[2] = {1, 1};
  Entry e{base, rows, row_bytes, {}};
  if (fn(&e.tm, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() >= 64) cache.erase(cache.begin());
  cache.push_back(e);
  out = e.tm;
  return true;
}

template <int CODEC, bool COUNT, int NDEC, bool INTERP = false>
static int launch(const Args& a, uint64_t* counters, cudaStream_t st) {
  auto kern = attn_tc_kernel<CODEC, COUNT, NDEC, INTERP>;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);   // per device
  if (e != cudaSuccess) return fail(ECCQ_ECUDA, cudaGetErrorString(e));
  const int nitems = a.nseq * a.g.H * a.npass * a.nsplit;
  kern<<<std::min(nitems, sm_count()), (kFrontWarps + NDEC) * 32, kSmemBytes, st>>>(a, counters);
  return check_launch("paged_decode_attention (tcgen05)");
}

}  // namespace tc

#ifdef ECCQ_TC_TRACE
extern "C" int eccq_tc_trace_read(long long* out_host) {
  return cudaMemcpyFromSymbol(out_host, tc::g_tc_trace, sizeof(long long) * 512 * 16) == cudaSuccess ? 0 : -2;
}
#endif

// returns 1 when the tcgen05 kernel took the launch (partials written; the caller runs the merge), 0 when the geometry is
// outside it, < 0 on error.  May lower a.nsplit / raise a.chunk (fewer, longer splits; the workspace was sized for more).
int try_attn_tc(const CacheGeom& g, AttnArgs& a, int q_dtype, uint64_t* counters, cudaStream_t st) {
  static const int env = [] { const char* v = std::getenv("ECCQ_ATTN_TC"); return v && *v ? std::atoi(v) : 1; }();
  static const int golay_env = [] { const char* v = std::getenv("ECCQ_ATTN_TC_GOLAY"); return v && *v ? std::atoi(v) : 1; }();
  const bool golay = g.codec == ECCQ_CODEC_GOLAY;
  if (!env || (golay && !golay_env) || (a.interp && g.codec != ECCQ_CODEC_H84) || g.D != tc::kD) return 0;
  if (g.bs < 16 || g.bs > tc::kTile || (tc::kTile % g.bs) != 0) return 0;
  if (!aligned16(g.k_cache) || !aligned16(g.v_cache) || !aligned16(g.k_scales) || !aligned16(g.v_scales)) return 0;
  const int G = a.Hq / g.H;
  tc::Args t;
  const uint64_t pages = (uint64_t)g.nb * g.L * g.H;
  if (pages * (uint64_t)(g.bs / 16) >= 0x7FFFFFFFull) return 0;
  // Golay: the codewords are read by the decode warps themselves (no tensor map; the two slots alias the scale maps)
  if ((!golay && (!tc::tensor_map_rows(t.tm_k, g.k_cache, pages * (g.bs / 16), 2048, 8) ||
                  !tc::tensor_map_rows(t.tm_v, g.v_cache, pages * (g.bs / 16), 2048, 8))) ||
      (golay && (!tc::tensor_map_rows(t.tm_k, g.k_scales, pages, (uint32_t)g.bs * 4, 4) || !tc::tensor_map_rows(t.tm_v, g.k_scales, pages, (uint32_t)g.bs * 4, 4))) ||
      !tc::tensor_map_rows(t.tm_ks, g.k_scales, pages, (uint32_t)g.bs * 4, 4) ||
      !tc::tensor_map_rows(t.tm_vs, a.v_use_k_scales ? g.k_scales : g.v_scales, pages, (uint32_t)g.bs * 4, 4))
    return 0;
  t.g = g; t.q = a.q; t.q_dtype = q_dtype; t.block_table = a.block_table; t.context_lens = a.context_lens; t.ws = a.ws;
  t.vscales = a.v_use_k_scales ? g.k_scales : g.v_scales;
  t.max_blocks = a.max_blocks; t.nseq = a.nseq; t.Hq = a.Hq; t.layer = a.layer; t.sm_scale = a.sm_scale;
  t.npass = (G + tc::kHeadsPerPass - 1) / tc::kHeadsPerPass;
  t.count_errors = a.count_errors;
  // splits: multiples of the 128-token tile, at least 2048 tokens, never more splits than the workspace was planned for
  const int64_t maxctx = (int64_t)a.max_blocks * g.bs;
  int chunk = (int)(((maxctx + a.nsplit - 1) / a.nsplit + tc::kTile - 1) / tc::kTile * tc::kTile);
  static const int min_chunk = [] { const char* v = std::getenv("ECCQ_ATTN_TC_CHUNK"); return v && *v ? std::atoi(v) : 2048; }();
  chunk = std::max(chunk, std::max(min_chunk / tc::kTile * tc::kTile, tc::kTile));
  a.chunk = chunk;
  a.nsplit = (int)((maxctx + chunk - 1) / chunk);
  t.chunk = a.chunk; t.nsplit = a.nsplit;
  static const int ndec = [] { const char* v = std::getenv("ECCQ_ATTN_TC_NDEC"); return v && *v ? std::atoi(v) : 16; }();
  int rc;
#define ECCQ_TC_CASE(C, CNT)                                            \
  rc = ndec == 8 ? tc::launch<C, CNT, 8>(t, counters, st) : tc::launch<C, CNT, 16>(t, counters, st)
  if (golay) {
    rc = a.count_errors ? tc::launch<ECCQ_CODEC_GOLAY, true, 16>(t, counters, st) : tc::launch<ECCQ_CODEC_GOLAY, false, 16>(t, counters, st);
  } else if (a.interp) {
    rc = a.count_errors ? tc::launch<ECCQ_CODEC_H84, true, 16, true>(t, counters, st) : tc::launch<ECCQ_CODEC_H84, false, 16, true>(t, counters, st);
  } else if (a.count_errors) {
    switch (g.codec) {
      case ECCQ_CODEC_H84: ECCQ_TC_CASE(ECCQ_CODEC_H84, true); break;
      case ECCQ_CODEC_H74: ECCQ_TC_CASE(ECCQ_CODEC_H74, true); break;
      default: ECCQ_TC_CASE(ECCQ_CODEC_INT4, false); break;
    }
  } else {
    switch (g.codec) {
      case ECCQ_CODEC_H84: ECCQ_TC_CASE(ECCQ_CODEC_H84, false); break;
      case ECCQ_CODEC_H74: ECCQ_TC_CASE(ECCQ_CODEC_H74, false); break;
      default: ECCQ_TC_CASE(ECCQ_CODEC_INT4, false); break;
    }
  }
#undef ECCQ_TC_CASE
  return rc ? rc : 1;
}

}  // namespace eccq

ECCQ_DEFINE_TU_TABLES(attention_tc)

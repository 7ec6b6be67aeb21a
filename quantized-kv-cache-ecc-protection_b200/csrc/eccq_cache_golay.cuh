// eccq_cache_golay.cuh -- fused cache read for the Golay(24,12) codec (included by eccq_cache.cu).
//
// Replaces the Golay branch of ECCBackend.attend's read half (kv_cache/ecc_shim.py:996-1009: gather pages, golay_decode of
// ctx*Hkv*43 codewords through a global-memory table, view [ctx,Hkv,129][:,:,:128], dequantise) with one launch.
// One warp per (tensor, sequence, block, head) page.  A page is block_size rows of cw = ceil(D/3) int32 codewords
// (per-row packing): it is read as one flat run of 16-byte chunks (98 % lane efficiency instead of the 67 % of a
// row-per-iteration mapping), every codeword is decoded against tables resident in shared memory -- the 2 x 64-entry
// split parity table (conflict-free) and the 4096-entry syndrome table, stored with the error weight packed beside the
// pattern so that corrected-bit / uncorrectable counters are one add per codeword and no POPC -- and the dequantised row
// segments are staged in shared memory so that the output leaves in coalesced 16-byte stores.
#pragma once

namespace eccq {

struct GolayReadArgs {
  FastDiv div_cw;        // codeword index inside the page -> row
  int chunks;            // 16-byte chunks per page = block_size * cw / 4
  int out_chunks_row;    // 16-byte chunks per output row = D * sizeof(TO) / 16
};

// syndrome table entry: bits 0..23 error pattern, bits 24..25 its weight; uncorrectable: 0x80000000 (data preserved,
// golay_triton.py:262-286)
__device__ __forceinline__ void golay_load_syndrome_packed(uint32_t* s_syn) {
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    const int32_t e = g_golay.syn[i];
    s_syn[i] = e < 0 ? 0x80000000u : ((uint32_t)e | ((uint32_t)__popc(e) << 24));
  }
}

template <typename TO>
__global__ void __launch_bounds__(256, 3) cache_read_golay_kernel(CacheGeom g, ReadArgs a, GolayReadArgs ga, int nblk_logical,
                                                                 uint64_t* counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_syn = reinterpret_cast<uint32_t*>(smem_raw);                 // 16 KB
  uint16_t* s_lo = reinterpret_cast<uint16_t*>(smem_raw + 16384);          // 128 B
  uint16_t* s_hi = s_lo + 64;                                              // 128 B
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = g.bs * g.D * (int)sizeof(TO) + g.bs * 4;         // dequantised rows + row scales
  unsigned char* wsm = smem_raw + 16384 + 256 + (size_t)warp_in_cta * stage_bytes;
  TO* s_out = reinterpret_cast<TO*>(wsm);
  float* s_sc = reinterpret_cast<float*>(wsm + g.bs * g.D * sizeof(TO));
  golay_load_parity(s_lo, s_hi);
  golay_load_syndrome_packed(s_syn);
  __syncthreads();

  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = (int64_t)2 * a.nseq * nblk_logical * g.H;
  const int cw = g.cw, D = g.D, bs = g.bs;
  const int64_t page_bytes = (int64_t)bs * cw * 4;
  constexpr int MAXCH = 8;   // chunks per lane held in registers (pages up to 256 chunks = 4 KB)
  uint32_t ncorr = 0, ndet = 0;

  for (int64_t u = warp; u < units; u += nwarps) {
    // u = ((i*nblk + j)*H + h)*2 + kv
    const int kv = (int)(u & 1);
    const uint32_t uh = (uint32_t)(u >> 1);
    const uint32_t ij = uh / (uint32_t)g.H;
    const int h = (int)(uh - ij * (uint32_t)g.H);
    const int i = (int)(ij / (uint32_t)nblk_logical);
    const int j = (int)(ij - (uint32_t)i * (uint32_t)nblk_logical);
    const int ctx = min(a.context_lens[i], a.T);
    const int t0 = j * bs;
    if (t0 >= ctx || j >= a.max_blocks) continue;
    const int blk = a.block_table[(int64_t)i * a.max_blocks + j];
    if (blk < 0) continue;
    const int nvalid = min(ctx - t0, bs);
    const int64_t page = page_index(g, blk, a.layer, h);
    const unsigned char* pbase = (kv ? g.v_cache : g.k_cache) + page * page_bytes;
    const float* scales = (kv ? g.v_scales : g.k_scales) + page * bs;

    uint4 v[MAXCH];
#pragma unroll
    for (int r = 0; r < MAXCH; ++r) {
      const int q = r * 32 + lane;
      if (q < ga.chunks) v[r] = ld_stream(pbase + (q << 4));
    }
    if (lane < bs) s_sc[lane] = __ldg(scales + lane);
    __syncwarp();

    uint32_t acc = 0;   // bits 0..6 sum of weights (<= 3 * 32), bits 7.. uncorrectable count
#pragma unroll
    for (int r = 0; r < MAXCH; ++r) {
      const int q = r * 32 + lane;
      if (q < ga.chunks) {
        const uint32_t w4[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
        const uint32_t wi0 = (uint32_t)q * 4u;
        uint32_t t = ga.div_cw.div(wi0);
        uint32_t jj = wi0 - t * (uint32_t)cw;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if ((int)t < nvalid) {
            const uint32_t c = w4[k];
            const uint32_t d = c & 0xFFFu;
            const uint32_t s = ((uint32_t)s_lo[d & 63u] ^ (uint32_t)s_hi[d >> 6]) ^ ((c >> 12) & 0xFFFu);
            const uint32_t e = s_syn[s];
            acc += e >> 24;
            const uint32_t dd = (d ^ e) & 0xFFFu;
            const float sc = s_sc[t];
            TO* dst = s_out + t * D + 3 * jj;
            const int e0 = 3 * (int)jj;
            // (n - 8) * scale with one rounding to the output type, as ((float)n - 8) * scale .to(dtype)
            dst[0] = Cvt<TO>::to((__uint_as_float(0x4B000000u | (dd & 0xFu)) - 8388616.0f) * sc);
            if (e0 + 1 < D) dst[1] = Cvt<TO>::to((__uint_as_float(0x4B000000u | ((dd >> 4) & 0xFu)) - 8388616.0f) * sc);
            if (e0 + 2 < D) dst[2] = Cvt<TO>::to((__uint_as_float(0x4B000000u | (dd >> 8)) - 8388616.0f) * sc);
          }
          if (++jj == (uint32_t)cw) { jj = 0; ++t; }
        }
      }
    }
    ncorr += acc & 0x7Fu;
    ndet += acc >> 7;
    __syncwarp();

    // ---- coalesced write-out: rows of D outputs, 16 bytes per lane
    TO* out = reinterpret_cast<TO*>(kv ? a.v_out : a.k_out);
    if (out) {
      const int total = nvalid * ga.out_chunks_row;
      const int64_t obase = (((int64_t)i * a.T + t0) * g.H + h) * D;
      const int64_t ostride = (int64_t)g.H * D;
      for (int c = lane; c < total; c += 32) {
        const int row = c / ga.out_chunks_row, col = c - row * ga.out_chunks_row;
        const uint4 x = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(s_out) + ((size_t)row * D * sizeof(TO)) + (col << 4));
        st_stream(reinterpret_cast<unsigned char*>(out + obase + row * ostride) + (col << 4), x);
      }
    }
    __syncwarp();
  }
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {ncorr, ndet};
  block_accumulate<2>(counters, idx, val);
}

constexpr int kGolay2TableBytes = 16384 + 8192;   // syndrome table + full parity table

// ---------------------------------------------------------------------------------------------- production shapes
// head_dim 128 / 64 with 16-token blocks.  The page (BS rows x CW codewords) is staged RAW in shared memory with cp.async
// (double buffered: the next page of the warp streams in while this one is decoded).  A lane-task is (row, group of 8
// consecutive codewords) = 24 consecutive values of one row: its outputs are 48 contiguous bytes (fp16), so the
// dequantised values leave as 16-byte stores straight from registers -- no output staging, no 2-byte shared stores.
// The last group of a row is shorter (D=128: 3 codewords -> values 120..127; the 129th nibble is the zero pad of
// ecc_shim.py:669-682).  Tasks per page = BS * ceil(CW/8) = 96 for D=128: exactly three per lane.
template <typename TO, int D, int BS>
__global__ void __launch_bounds__(256, 3) cache_read_golay2_kernel(CacheGeom g, ReadArgs a, int nblk_logical, uint64_t* counters) {
  constexpr int CW = (D + 2) / 3;                 // codewords per row
  constexpr int NG = (CW + 7) / 8;                // groups per row
  constexpr int PAGE_WORDS = BS * CW;             // 688 for D=128
  constexpr int CHUNKS = PAGE_WORDS / 4;          // 16-byte chunks per page
  constexpr int STAGE_WORDS = PAGE_WORDS + BS;    // raw codewords + row scales
  constexpr int TASKS = BS * NG;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  static_assert(PAGE_WORDS % 4 == 0 && (STAGE_WORDS * 4) % 16 == 0, "page must be a whole number of 16-byte chunks");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_syn = reinterpret_cast<uint32_t*>(smem_raw);                 // 16 KB: syndrome -> pattern | weight << 24
  uint16_t* s_par = reinterpret_cast<uint16_t*>(smem_raw + 16384);         // 8 KB: parity of every 12-bit data word
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* wsm = reinterpret_cast<uint32_t*>(smem_raw + kGolay2TableBytes) + (size_t)warp_in_cta * 2 * STAGE_WORDS;
  // The decode is bound by the ALU pipe, not by shared-memory wavefronts: one (bank-conflicted) look-up in the full
  // 4096-entry parity table costs fewer ALU operations than two conflict-free look-ups in the 2 x 64 split table.
  for (int d = threadIdx.x; d < 4096; d += blockDim.x) s_par[d] = g_golay.par_lo[d & 63] ^ g_golay.par_hi[d >> 6];
  golay_load_syndrome_packed(s_syn);
  __syncthreads();

  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = (int64_t)2 * a.nseq * nblk_logical * g.H;
  constexpr int64_t page_bytes = (int64_t)PAGE_WORDS * 4;
  uint32_t ncorr = 0, ndet = 0;

  struct Unit { int i, j, h, kv, blk, nvalid; };
  auto locate = [&](int64_t u, Unit& un) {   // u = ((i*nblk + j)*H + h)*2 + kv ; returns false for pages outside the context
    if (u >= units) return false;
    un.kv = (int)(u & 1);
    const uint32_t uh = (uint32_t)(u >> 1);
    const uint32_t ij = uh / (uint32_t)g.H;
    un.h = (int)(uh - ij * (uint32_t)g.H);
    un.i = (int)(ij / (uint32_t)nblk_logical);
    un.j = (int)(ij - (uint32_t)un.i * (uint32_t)nblk_logical);
    const int ctx = min(a.context_lens[un.i], a.T);
    const int t0 = un.j * BS;
    if (t0 >= ctx || un.j >= a.max_blocks) return false;
    un.blk = a.block_table[(int64_t)un.i * a.max_blocks + un.j];
    un.nvalid = min(ctx - t0, BS);
    return un.blk >= 0;
  };
  auto stage = [&](const Unit& un, int st) {
    const int64_t page = page_index(g, un.blk, a.layer, un.h);
    const unsigned char* src = (un.kv ? g.v_cache : g.k_cache) + page * page_bytes;
    uint32_t* dst = wsm + (size_t)st * STAGE_WORDS;
    for (int q = lane; q < CHUNKS; q += 32) cp_async16(dst + 4 * q, src + (q << 4));
    if (lane < BS / 4) cp_async16(dst + PAGE_WORDS + 4 * lane, (un.kv ? g.v_scales : g.k_scales) + page * BS + 4 * lane);
  };

  Unit cur, nxt;
  int64_t u = warp;
  bool cur_ok = locate(u, cur);
  if (cur_ok) stage(cur, 0);
  cp_async_commit();
  int st = 0;
  for (; u < units; u += nwarps, st ^= 1) {
    const bool nxt_ok = locate(u + nwarps, nxt);
    if (nxt_ok) stage(nxt, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    if (cur_ok) {
      const uint32_t* s_raw = wsm + (size_t)st * STAGE_WORDS;
      const float* s_sc = reinterpret_cast<const float*>(s_raw + PAGE_WORDS);
      TO* out = reinterpret_cast<TO*>(cur.kv ? a.v_out : a.k_out);
      const int64_t obase = (((int64_t)cur.i * a.T + (int64_t)cur.j * BS) * g.H + cur.h) * D;
      const int64_t ostride = (int64_t)g.H * D;
      uint32_t acc = 0;   // bits 0..6 sum of error weights (<= 3 * 8 * 3), bits 7.. uncorrectable count
#pragma unroll
      for (int r = 0; r < (TASKS + 31) / 32; ++r) {
        const int id = r * 32 + lane;
        const int t = id / NG, gi = id - t * NG;
        if (id < TASKS && t < cur.nvalid) {
          const int ncw = min(8, CW - 8 * gi);
          const uint32_t* src = s_raw + t * CW + 8 * gi;
          const float sc = s_sc[t];
          // nibble k of the corrected word sits at bits 4k: 2^23 | (word & 0xF << 4k) is the float 2^23 + 16^k * n, so
          // subtracting 2^23 + 8 * 16^k and multiplying by scale / 16^k gives (n - 8) * scale with the same single
          // rounding (a power-of-two factor is exact) -- no shifts.  Scales so small that scale/256 would be denormal
          // take the shifted form.
          const bool tiny = !(sc >= 1e-30f);
          const float sc1 = sc * 0.0625f, sc2 = sc * 0.00390625f;
          float f[24];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t c = k < ncw ? src[k] : 0u;   // missing codewords of the short last group decode clean
            const uint32_t e = s_syn[(uint32_t)s_par[c & 0xFFFu] ^ ((c >> 12) & 0xFFFu)];
            acc = __dp4a(e, 0x01000000u, acc);           // += weight byte (bits 24..31 of the table entry)
            const uint32_t x = c ^ e;
            if (!tiny) {
              f[3 * k] = (__uint_as_float(0x4B000000u | (x & 0x00Fu)) - 8388616.0f) * sc;
              f[3 * k + 1] = (__uint_as_float(0x4B000000u | (x & 0x0F0u)) - 8388736.0f) * sc1;
              f[3 * k + 2] = (__uint_as_float(0x4B000000u | (x & 0xF00u)) - 8390656.0f) * sc2;
            } else {
              f[3 * k] = (__uint_as_float(0x4B000000u | (x & 0xFu)) - 8388616.0f) * sc;
              f[3 * k + 1] = (__uint_as_float(0x4B000000u | ((x >> 4) & 0xFu)) - 8388616.0f) * sc;
              f[3 * k + 2] = (__uint_as_float(0x4B000000u | ((x >> 8) & 0xFu)) - 8388616.0f) * sc;
            }
          }
          if (a.paged_nib) {
            // eccq_cache_correct_pages: corrected nibbles as INT4 pages of a one-layer cache (row = D bytes) + row scales
            uint8_t* pages = cur.kv ? a.nib_v : a.nib_k;
            const int64_t pg = (int64_t)cur.blk * g.H + cur.h;
            uint32_t w[6] = {0, 0, 0, 0, 0, 0};   // 24 nibble bytes (f[] holds scaled values: the nibbles come from the codewords again)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t c = k < ncw ? src[k] : 0u;
              const uint32_t e = s_syn[(uint32_t)s_par[c & 0xFFFu] ^ ((c >> 12) & 0xFFFu)];
              const uint32_t x = c ^ e;
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                const int b = 3 * k + q;                       // byte index inside the 24-byte run
                w[b >> 2] |= ((x >> (4 * q)) & 0xFu) << (8 * (b & 3));
              }
            }
            const int nval = min(24, D - 24 * gi);
            uint2* dst = reinterpret_cast<uint2*>(pages + pg * ((int64_t)BS * D) + t * D + 24 * gi);
            dst[0] = make_uint2(w[0], w[1]);
            if (nval > 8) dst[1] = make_uint2(w[2], w[3]);
            if (nval > 16) dst[2] = make_uint2(w[4], w[5]);
            if (gi == 0) (cur.kv ? a.scv_out : a.sck_out)[pg * BS + t] = sc;
          } else if (out) {
            TO* dst = out + obase + t * ostride + 24 * gi;
            constexpr int VPC = 16 / (int)sizeof(TO);   // values per 16-byte chunk
            const int nval = min(24, D - 24 * gi);
#pragma unroll
            for (int ch = 0; ch < 24 / VPC; ++ch) {
              if (ch * VPC < nval) {
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if constexpr (sizeof(TO) == 4) o[q] = __float_as_uint(f[ch * 4 + q]);
                  else o[q] = Pack2<TO>::pack(f[ch * 8 + 2 * q], f[ch * 8 + 2 * q + 1]);
                }
                st_stream(dst + ch * VPC, make_uint4(o[0], o[1], o[2], o[3]));
              }
            }
          }
        }
      }
      ncorr += acc & 0x7Fu;
      ndet += acc >> 7;
    }
    __syncwarp();
    cur = nxt;
    cur_ok = nxt_ok;
  }
  cp_async_wait<0>();
  const int idx[2] = {ECCQ_CTR_CORRECTED, ECCQ_CTR_DETECTED};
  const uint32_t val[2] = {ncorr, ndet};
  block_accumulate<2>(counters, idx, val);
  (void)FULL;
}

}  // namespace eccq

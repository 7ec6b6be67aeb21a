/*
 * eccq.h  --  C ABI of libeccq_sm100a.so: the B200-native (sm_100a) ECC-protected INT4 KV-cache path.
 *
 * This is the drop-in boundary for ONE path of indrajeetadityaroy9/quantized-kv-cache-ecc-protection.
 * The reference has no FFI: its boundary is Python (Triton wrappers called by kv_cache/ecc_shim.py).  Each
 * entry point below names the reference interface it replaces (file:line inside the reference tree); the
 * Python host layer in quantized-kv-cache-ecc-protection_b200/ binds these with ctypes and re-creates the
 * reference's function names, argument meaning and error behaviour on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing here allocates,
 *     synchronises the host, or throws; kernels are enqueued and the call returns;
 *   - return value: ECCQ_OK (0) or a negative ECCQ_E* code; eccq_last_error() gives a static message;
 *   - `counters` is a device uint64_t[ECCQ_NCOUNTERS] accumulated with one atomicAdd per CTA, or NULL;
 *   - float dtypes are named by ECCQ_F32 / ECCQ_F16 / ECCQ_BF16; codecs by ECCQ_CODEC_*.
 *   - results are bit-exact with the reference for all integer outputs (codewords, nibbles, error types,
 *     counters, scales) and for dequantised values (single IEEE rounding, no fast-math).
 */
#ifndef ECCQ_H_
#define ECCQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECCQ_ABI_VERSION 1

enum { ECCQ_OK = 0, ECCQ_EINVAL = -1, ECCQ_ECUDA = -2, ECCQ_EUNSUPPORTED = -3 };
enum { ECCQ_F32 = 0, ECCQ_F16 = 1, ECCQ_BF16 = 2 };
/* SUPPORTED_CODECS of the ECC-protected modes, kv_cache/ecc_shim.py:163 (+ int4 = no ECC, nibble per byte) */
enum { ECCQ_CODEC_H74 = 0, ECCQ_CODEC_H84 = 1, ECCQ_CODEC_GOLAY = 2, ECCQ_CODEC_INT4 = 3 };
/* counters[]: semantics per codec follow the reference:
 *   CORRECTED  H84: #codewords of error type 1; H74: #codewords with syndrome != 0;
 *              Golay: sum of popcount(error pattern) over correctable codewords (bits)
 *   DETECTED   H84: #codewords of type 2; Golay: #uncorrectable codewords
 *   FLIPPED    bits flipped by an injector;  AFFECTED  codewords with >= 1 flipped bit
 *   VALUES     INT4 values written by eccq_cache_write (K and V) */
enum { ECCQ_CTR_CORRECTED = 0, ECCQ_CTR_DETECTED = 1, ECCQ_CTR_FLIPPED = 2, ECCQ_CTR_AFFECTED = 3,
       ECCQ_CTR_VALUES = 4, ECCQ_CTR_INTERPOLATED = 5, ECCQ_NCOUNTERS = 8 };
/* fault-injection modes of the fused cache write */
enum { ECCQ_INJECT_NONE = 0,
       ECCQ_INJECT_MASK = 1,        /* XOR an externally supplied flip mask (bit-exact parity runs)          */
       ECCQ_INJECT_PHILOX_REF = 2,  /* the reference's Philox stream, one launch-equivalent per (token,head):
                                       fault_injection_triton.py:227-334 keyed as ecc_shim.py:643-651,:713-721 */
       ECCQ_INJECT_PHILOX_FAST = 3  /* own Philox-4x32-10 stream + geometric gap sampling (throughput runs)   */ };

int         eccq_abi_version(void);
const char* eccq_last_error(void);
/* device properties the host layer sizes grids with; out_host = int[4] {sm_count, cc_major, cc_minor, l2_bytes/1MiB} */
int eccq_device_info(int device, int* out_host);

/* ---- flat codecs ---------------------------------------------------------------------------------- */
/* hamming74_encode, ecc_codecs/triton_kernels/hamming74_triton.py:170-201 (kernel :47-91) */
int eccq_h74_encode(const uint8_t* nibbles, uint8_t* cw, int64_t n, void* stream);
/* hamming74_decode, hamming74_triton.py:218-277 (kernel :99-162); flag nullable */
int eccq_h74_decode(const uint8_t* cw, uint8_t* nibbles, uint8_t* flag, uint64_t* counters, int64_t n, void* stream);
/* hamming84_encode, hamming84_triton.py:217-254 (kernel :49-108) */
int eccq_h84_encode(const uint8_t* nibbles, uint8_t* cw, int64_t n, void* stream);
/* hamming84_decode, hamming84_triton.py:281-351 (kernel :116-209); etype nullable.
 * zero_double != 0 zeroes the nibble of double-detected codewords as fused_kernels.py:343-344 does (Q4). */
int eccq_h84_decode(const uint8_t* cw, uint8_t* nibbles, uint8_t* etype, uint64_t* counters, int64_t n,
                    int zero_double, void* stream);
/* golay_encode, golay_triton.py:382-422 (kernel :98-157): triplets uint8[n_cw*3] -> int32[n_cw] */
int eccq_golay_encode(const uint8_t* triplets, int32_t* cw, int64_t n_cw, void* stream);
/* golay_decode, golay_triton.py:425-498 (kernel :212-295, table config.py:403-457); errcnt nullable (4 = uncorrectable) */
int eccq_golay_decode(const int32_t* cw, uint8_t* triplets, uint8_t* errcnt, uint64_t* counters, int64_t n_cw, void* stream);
/* copy the 4096-entry syndrome table (int32, -1 = uncorrectable) built on the host into table_host */
int eccq_golay_syndrome_table(int32_t* table_host);

/* ---- fault injection --------------------------------------------------------------------------------- */
/* out = in ^ mask (elem_size 1 or 4); counters FLIPPED/AFFECTED from popcount(mask & low n_bits) */
int eccq_xor_mask(const void* in, const void* mask, void* out, int elem_size, int n_bits, uint64_t* counters,
                  int64_t n, void* stream);
/* inject_bit_errors_triton (vectorized=0) / _vectorized (=1), fault_injection_triton.py:337-424,:434-496.
 * Bit-exact with the reference's Triton Philox stream for the same (n, n_bits, seed, ber). flipcnt nullable uint8[n]. */
int eccq_inject_philox_ref(const void* in, void* out, uint8_t* flipcnt, int elem_size, int n_bits, float ber,
                           int64_t seed, int vectorized, uint64_t* counters, int64_t n, void* stream);
/* own stream: Philox-4x32-10 keyed (seed, 16-byte chunk index), geometric gap sampling; i.i.d. Bernoulli(ber)
 * on the low n_bits of every element; statistically checked only */
int eccq_inject_philox_fast(const void* in, void* out, int elem_size, int n_bits, float ber, uint64_t seed,
                            uint64_t* counters, int64_t n, void* stream);

/* ---- interpolation / quantisation ---------------------------------------------------------------------- */
/* interpolate_double_errors, interpolation_triton.py:162-265 (kernel :120-159) on a [outer, len, inner] view:
 * neighbours are len-1 / len+1 at stride `inner` (inner = Hkv*D for the shim's seq_dim=0 call, 1 for seq_dim=-1) */
int eccq_interpolate(const uint8_t* q, const uint8_t* etype, uint8_t* out, int64_t outer, int64_t len, int64_t inner,
                     void* stream);
/* compute_quantization_scales + quantise, kv_cache/paged_cache_ecc.py:302-334, kv_cache/ecc_shim.py:572-580.
 * q nullable (scales only). zero_guard: 0 = reference shim (scale==0 -> 1.0), 1 = registry style
 * clamp(absmax, 1e-8) (ecc_codecs/quantization_backends.py:253-257). codec >= 0 additionally encodes
 * (fused_quantize_encode_hamming84/74, fused_kernels.py:97-160,:222-269, with the torch rounding of Q1);
 * codec < 0 stores raw nibbles. */
int eccq_quantize_rows(const void* x, int dtype, uint8_t* q, float* scales, int64_t rows, int row_len, int zero_guard,
                       int codec, void* stream);
/* (nibble-8)*scale -> out dtype; codec < 0: input is raw nibbles, else Hamming codewords decoded first
 * (fused_decode_dequantize_hamming84, fused_kernels.py:372-437 with zero_double=1) */
int eccq_dequantize_rows(const uint8_t* q, const float* scales, void* out, int out_dtype, int64_t rows, int row_len,
                         int codec, int zero_double, uint64_t* counters, void* stream);

/* ---- paged cache: fused write and fused read ------------------------------------------------------------- */
typedef struct eccq_cache_desc {
  void*  k_cache;      /* [num_blocks, num_layers, num_kv_heads, block_size*cw_per_token] uint8 | int32 (golay) */
  void*  v_cache;
  float* k_scales;     /* [num_blocks, num_layers, num_kv_heads, block_size] fp32 */
  float* v_scales;
  int32_t num_blocks, num_layers, num_kv_heads, head_dim, block_size;
  int32_t codec;       /* ECCQ_CODEC_* ; cw_per_token = head_dim, or (head_dim+2)/3 for golay (ecc_shim.py:260-263) */
} eccq_cache_desc;

/* ECCBackend.write, kv_cache/ecc_shim.py:458-737 as ONE launch: quantise + encode + inject + paged scatter of K and V.
 * k, v: [nseq, S, Hkv, D] `dtype`, contiguous.  Token s of sequence i goes to position start_pos[i]+s (start_pos NULL = 0)
 * through block_table[i*max_blocks + pos/block_size]; rows whose block is < 0 are skipped.
 * inject: ECCQ_INJECT_*; k_mask/v_mask: [nseq,S,Hkv,cw_per_token] uint8|int32 for ECCQ_INJECT_MASK;
 * seed/inj_base: ECCQ_INJECT_PHILOX_REF uses seed_K = seed + inj_base + ((i*S+s)*Hkv+h), seed_V = seed_K+1. */
int eccq_cache_write(const eccq_cache_desc* c, const void* k, const void* v, int dtype, const int32_t* block_table,
                     int max_blocks, const int32_t* start_pos, int nseq, int S, int layer, int inject,
                     const void* k_mask, const void* v_mask, float ber, int64_t seed, int64_t inj_base,
                     uint64_t* counters, void* stream);

/* The read half of ECCBackend.attend, ecc_shim.py:931-1071, as ONE launch: gather pages + syndrome decode + correct +
 * flag + [temporal interpolation of double-detected values] + dequantise, K and V.
 * k_out, v_out: [nseq, out_stride_tokens, Hkv, D] `out_dtype`; tokens >= context_lens[i] are left untouched.
 * nib_out_k/v (nullable): corrected (and interpolated) INT4 codes uint8 [same shape] for parity checks. */
int eccq_cache_read(const eccq_cache_desc* c, const int32_t* block_table, int max_blocks, const int32_t* context_lens,
                    int nseq, int out_stride_tokens, int layer, int use_interpolation, int out_dtype, void* k_out, void* v_out,
                    uint8_t* nib_out_k, uint8_t* nib_out_v, uint64_t* counters, void* stream);

/* Decode [+ temporally interpolate] the pages of ONE layer into corrected INT4 pages of a one-layer cache with the same
 * paged layout (k_pages / v_pages: uint8 [num_blocks, 1, Hkv, block_size*D], one nibble per byte; k/v_page_scales: fp32
 * [num_blocks, 1, Hkv, block_size], copies of the row scales).  This is the decode + interpolate half of ECCBackend.attend
 * (ecc_shim.py:979-1038) for q_len == 1: paged decode attention with interpolation runs its INT4 tile path over the corrected
 * pages (neighbours along the token axis span pages and tiles, so they are resolved once here, by the streaming read kernel).
 * Counters: CORRECTED / DETECTED / INTERPOLATED as eccq_cache_read.  ECCQ_EUNSUPPORTED for Golay or geometries other than
 * head_dim 128/64 with 16/32-token blocks (the caller then uses eccq_paged_decode_attention's general path). */
int eccq_cache_correct_pages(const eccq_cache_desc* c, const int32_t* block_table, int max_blocks, const int32_t* context_lens,
                             int nseq, int layer, int use_interpolation, uint8_t* k_pages, uint8_t* v_pages,
                             float* k_page_scales, float* v_page_scales, uint64_t* counters, void* stream);

/* ---- unprotected comparison modes: paged scatter / gather of raw rows ----------------------------------------- */
/* The "fp16" and "fp8" codecs of the shim (kv_cache/ecc_shim.py:511-570 write, :802-872 read): rows are stored as they are
 * (fp16) or as float8_e4m3fn bytes (fp8), no scales.  k, v: [nseq, S, Hkv, row_bytes] raw bytes (row_bytes % 16 == 0);
 * cache [num_blocks, L, Hkv, block_size * row_bytes].  inject = ECCQ_INJECT_PHILOX_REF flips bits of the low n_bits of
 * every BYTE with the reference's per-(token, head) streams (N = row_bytes, seed_K = seed + inj_base + row, seed_V = +1;
 * ecc_shim.py:547-556); other injectors are rejected. */
int eccq_paged_rows_write(void* k_cache, void* v_cache, const void* k, const void* v, int row_bytes, int num_layers,
                          int num_kv_heads, int block_size, const int32_t* block_table, int max_blocks,
                          const int32_t* start_pos, int nseq, int S, int layer, int inject, int n_bits, float ber,
                          int64_t seed, int64_t inj_base, uint64_t* counters, void* stream);
/* gather the rows of whole contexts: k_out, v_out [nseq, out_stride_tokens, Hkv, row_bytes]; tokens past a context untouched */
int eccq_paged_rows_read(const void* k_cache, const void* v_cache, void* k_out, void* v_out, int row_bytes, int num_layers,
                         int num_kv_heads, int block_size, const int32_t* block_table, int max_blocks,
                         const int32_t* context_lens, int nseq, int out_stride_tokens, int layer, void* stream);

/* ---- ECC-fused paged decode attention ------------------------------------------------------------------- */
/* paged_attention_ecc, kv_cache/attention_ecc.py:620-780 (kernel :264-426), generalised to GQA (kv_head = q_head/group,
 * SURVEY Q13), to hamming74/golay, to temporal interpolation, and split over the context (FlashDecoding).
 * q, out: [nseq, Hq, D] q_dtype. workspace: device scratch of eccq_attention_workspace_bytes(); empty context -> zeros (Q14).
 * count_errors != 0 accumulates CORRECTED/DETECTED for the codewords it decodes (the reference fast path counts none, Q12). */
int64_t eccq_attention_workspace_bytes(int nseq, int Hq, int D, int max_context);
int eccq_paged_decode_attention(const eccq_cache_desc* c, const void* q, void* out, int q_dtype, const int32_t* block_table,
                                int max_blocks, const int32_t* context_lens, int nseq, int Hq, int layer, float sm_scale,
                                int use_interpolation, int v_use_k_scales, int count_errors, void* workspace,
                                int64_t workspace_bytes, uint64_t* counters, void* stream);

/* ---- ECC-fused causal prefill attention (q_len > 1) ------------------------------------------------------ */
/* Replaces the read half of ECCBackend.attend for q_len > 1 -- gather pages, decode, interpolate, dequantise to fp32 K/V
 * [ctx, Hkv, D], repeat_interleave, SDPA(is_causal) -- kv_cache/ecc_shim.py:739-1089 and _run_attention :1138-1164, with ONE
 * kernel that consumes the codewords directly (ECC decode in the K/V prologue, no K/V materialisation, no library kernel).
 * q, out: [batch, Hq, Sq, D] q_dtype.  nseq == batch: batch row b attends over sequence b; nseq == 1: every batch row
 * attends over the one shared context (the reference's write semantics).  Query t sits at position
 * context_len - Sq + t and sees the tokens at or before it (equal to SDPA's top-left is_causal mask when context_len == Sq,
 * the reference's only case; chunked prefill in append mode otherwise).  Rows with no visible token -> zeros.
 * count_errors != 0 accumulates CORRECTED / DETECTED once per decoded codeword of the context (as the reference's decode does). */
int eccq_paged_prefill_attention(const eccq_cache_desc* c, const void* q, void* out, int q_dtype, const int32_t* block_table,
                                 int max_blocks, const int32_t* context_lens, int nseq, int batch, int Hq, int Sq, int layer,
                                 float sm_scale, int use_interpolation, int count_errors, uint64_t* counters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECCQ_H_ */

"""
CPU oracle for the ECC-protected INT4 KV-cache path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, in plain numpy, the arithmetic of the reference
(indrajeetadityaroy9/quantized-kv-cache-ecc-protection) for the one hot path this
repository re-implements in CUDA.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package never does (it fails loudly when its CUDA library is missing).

Parity status: PINNED.  Every function below is checked in ``tests/test_oracle_golden.py``
against fixtures under ``tests/golden/`` that were produced by importing the reference
itself in the authoring container (its Triton kernels run on CPU tensors under
``TRITON_INTERPRET=1``; generator: ``tests/golden/make_golden.py``) and against the
known answers recorded in SURVEY.md section 2.2.

Citations are ``path:line`` inside the reference tree.
"""

from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------
# Constants (ecc_codecs/triton_kernels/config.py)
# ----------------------------------------------------------------------------------------

# syndrome -> bit position to flip, config.py:131-161
HAMMING_SYNDROME_TO_BIT = np.array([-1, 4, 5, 0, 6, 1, 2, 3], dtype=np.int8)

# config.py:329-347 (row i = parity contributions of data bit i)
GOLAY_B = np.array(
    [
        [1, 1, 0, 1, 1, 1, 0, 0, 0, 1, 0, 1],
        [1, 0, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1],
        [0, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1, 1],
        [1, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 1],
        [1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 1, 1],
        [1, 0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1],
        [0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1],
        [0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 0, 1],
        [0, 1, 0, 1, 1, 0, 1, 1, 1, 0, 0, 1],
        [1, 0, 1, 1, 0, 1, 1, 1, 0, 0, 0, 1],
        [0, 1, 1, 0, 1, 1, 1, 0, 0, 0, 1, 1],
        [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0],
    ],
    dtype=np.uint8,
)

ERR_NONE, ERR_SINGLE, ERR_DOUBLE, ERR_PARITY = 0, 1, 2, 3  # config.py:197-214


def golay_b_cols() -> list[int]:
    """12-bit column masks of B: parity bit i = parity(data & col_i). golay_triton.py:59-70."""
    return [int(sum(int(GOLAY_B[j, i]) << j for j in range(12))) for i in range(12)]


def golay_h_masks() -> list[int]:
    """24-bit rows of H = [B^T | I]. config.py:354-379, golay_triton.py:332-374."""
    cols = golay_b_cols()
    return [cols[i] | (1 << (12 + i)) for i in range(12)]


def _parity32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x = x ^ (x >> 16)
    x = x ^ (x >> 8)
    x = x ^ (x >> 4)
    x = x ^ (x >> 2)
    x = x ^ (x >> 1)
    return x & 1


def golay_syndrome(cw: np.ndarray) -> np.ndarray:
    """12-bit syndrome, bit i = parity(cw & H_i). golay_triton.py:161-189."""
    cw = np.asarray(cw).astype(np.int64).astype(np.uint32) & np.uint32(0xFFFFFFFF)
    s = np.zeros(cw.shape, dtype=np.uint32)
    for i, m in enumerate(golay_h_masks()):
        s |= _parity32(cw & np.uint32(m)) << np.uint32(i)
    return s


_SYN_TABLE = None


def golay_syndrome_table() -> np.ndarray:
    """4096-entry syndrome -> error pattern table, -1 = uncorrectable. config.py:403-457.

    Enumeration order (weight 1, then 2, then 3, first writer wins for weight >= 2) is kept
    although patterns of weight <= 3 never collide for this code.
    """
    global _SYN_TABLE
    if _SYN_TABLE is not None:
        return _SYN_TABLE
    masks = golay_h_masks()

    def syn(e: int) -> int:
        s = 0
        for i, m in enumerate(masks):
            s |= (bin(e & m).count("1") & 1) << i
        return s

    single = [syn(1 << i) for i in range(24)]
    table = [-1] * 4096
    table[0] = 0
    for i in range(24):
        table[single[i]] = 1 << i
    for i in range(24):
        for j in range(i + 1, 24):
            s = single[i] ^ single[j]
            if table[s] == -1:
                table[s] = (1 << i) | (1 << j)
    for i in range(24):
        for j in range(i + 1, 24):
            for k in range(j + 1, 24):
                s = single[i] ^ single[j] ^ single[k]
                if table[s] == -1:
                    table[s] = (1 << i) | (1 << j) | (1 << k)
    _SYN_TABLE = np.array(table, dtype=np.int32)
    return _SYN_TABLE


# ----------------------------------------------------------------------------------------
# Hamming(7,4) / Hamming(8,4)
# ----------------------------------------------------------------------------------------


def _h7(n: np.ndarray) -> np.ndarray:
    n = np.asarray(n).astype(np.uint8)
    d0, d1, d2, d3 = (n >> 0) & 1, (n >> 1) & 1, (n >> 2) & 1, (n >> 3) & 1
    p0 = d0 ^ d1 ^ d3
    p1 = d0 ^ d2 ^ d3
    p2 = d1 ^ d2 ^ d3
    return (d0 | (d1 << 1) | (d2 << 2) | (d3 << 3) | (p0 << 4) | (p1 << 5) | (p2 << 6)).astype(np.uint8)


def h74_encode(n: np.ndarray) -> np.ndarray:
    """hamming74_triton.py:47-91. Only the low 4 bits of the input are used."""
    return _h7(n)


def h84_encode(n: np.ndarray) -> np.ndarray:
    """hamming84_triton.py:49-108. Bit 7 = parity of the 7 Hamming bits."""
    h7 = _h7(n)
    p = h7 ^ (h7 >> 4)
    p = p ^ (p >> 2)
    p = p ^ (p >> 1)
    return (h7 | ((p & 1) << 7)).astype(np.uint8)


def _syndrome3(c: np.ndarray) -> np.ndarray:
    b = [(c >> i) & 1 for i in range(7)]
    s0 = b[0] ^ b[1] ^ b[3] ^ b[4]
    s1 = b[0] ^ b[2] ^ b[3] ^ b[5]
    s2 = b[1] ^ b[2] ^ b[3] ^ b[6]
    return (s0 | (s1 << 1) | (s2 << 2)).astype(np.int32)


def h74_decode(cw: np.ndarray):
    """hamming74_triton.py:99-162 -> (nibble, flag); flag = syndrome != 0; bit 7 is ignored."""
    c = np.asarray(cw).astype(np.uint8)
    syn = _syndrome3(c)
    pos = HAMMING_SYNDROME_TO_BIT[syn].astype(np.int32)
    corr = np.where(pos >= 0, np.left_shift(1, np.maximum(pos, 0)), 0).astype(np.uint8)
    data = ((c ^ corr) & 0x0F).astype(np.uint8)
    return data, (syn != 0).astype(np.uint8)


def h84_decode(cw: np.ndarray):
    """hamming84_triton.py:116-209 -> (nibble, error_type). Double errors keep the raw nibble."""
    c = np.asarray(cw).astype(np.uint8)
    h7 = c & 0x7F
    stored = (c >> 7) & 1
    syn = _syndrome3(h7)
    p = h7 ^ (h7 >> 4)
    p = p ^ (p >> 2)
    p = p ^ (p >> 1)
    perr = stored != (p & 1)
    sz = syn == 0
    etype = np.where(sz, np.where(perr, 3, 0), np.where(perr, 1, 2)).astype(np.uint8)
    pos = HAMMING_SYNDROME_TO_BIT[syn].astype(np.int32)
    fix = (etype == 1) & (pos >= 0)
    corr = np.where(fix, np.left_shift(1, np.maximum(pos, 0)), 0).astype(np.uint8)
    data = ((h7 ^ corr) & 0x0F).astype(np.uint8)
    return data, etype


def h84_counters(etype: np.ndarray):
    """(corrected, detected) as hamming84_triton.py:341-342."""
    return int((etype == 1).sum()), int((etype == 2).sum())


# ----------------------------------------------------------------------------------------
# Golay(24,12)
# ----------------------------------------------------------------------------------------


def golay_encode(triplets: np.ndarray) -> np.ndarray:
    """golay_triton.py:98-157. triplets uint8 [N,3] -> int32 [N]."""
    t = np.asarray(triplets).astype(np.uint32).reshape(-1, 3)
    d = (t[:, 0] & 0xF) | ((t[:, 1] & 0xF) << 4) | ((t[:, 2] & 0xF) << 8)
    par = np.zeros_like(d)
    for i, c in enumerate(golay_b_cols()):
        par |= _parity32(d & np.uint32(c)) << np.uint32(i)
    return (d | (par << 12)).astype(np.int32)


def golay_decode(cw: np.ndarray):
    """golay_triton.py:212-295 -> (triplets uint8 [N,3], errcnt uint8 [N]); errcnt 4 = uncorrectable
    (data preserved)."""
    c = np.asarray(cw).astype(np.int32).reshape(-1)
    syn = golay_syndrome(c)
    pat = golay_syndrome_table()[syn]
    ok = pat >= 0
    fixed = np.where(ok, c ^ pat, c)
    d = fixed & 0xFFF
    trip = np.stack([d & 0xF, (d >> 4) & 0xF, (d >> 8) & 0xF], axis=1).astype(np.uint8)
    return trip, _popcnt_pat(pat)


def _popcnt_pat(pat: np.ndarray) -> np.ndarray:
    p = np.where(pat >= 0, pat, 0).astype(np.uint32)
    p = (p & 0x55555555) + ((p >> 1) & 0x55555555)
    p = (p & 0x33333333) + ((p >> 2) & 0x33333333)
    p = (p & 0x0F0F0F0F) + ((p >> 4) & 0x0F0F0F0F)
    p = (p & 0x00FF00FF) + ((p >> 8) & 0x00FF00FF)
    p = (p & 0x0000FFFF) + ((p >> 16) & 0x0000FFFF)
    return np.where(pat >= 0, p, 4).astype(np.uint8)


def golay_counters(errcnt: np.ndarray):
    """(bits_corrected, n_uncorrectable) as golay_triton.py:491-493."""
    ok = errcnt < 4
    return int(errcnt[ok].astype(np.int64).sum()), int((~ok).sum())


# ----------------------------------------------------------------------------------------
# Fault injection: the reference's Philox stream (triton.language.random, triton 3.6.0) and
# its call sites fault_injection_triton.py:227-334 (per-bit) and :56-224 (rand4x).
# ----------------------------------------------------------------------------------------

_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 exactly as triton/language/random.py:philox_impl (uint32 branch)."""
    A = np.uint64(0xD2511F53)
    B = np.uint64(0xCD9E8D57)
    KA = np.uint64(0x9E3779B9)
    KB = np.uint64(0xBB67AE85)
    c0, c1, c2, c3, k0, k1 = (np.asarray(v).astype(np.uint64) & _M32 for v in (c0, c1, c2, c3, k0, k1))
    for _ in range(10):
        pb = B * c2
        pa = A * c0
        n0 = (pb >> np.uint64(32)) ^ c1 ^ k0
        n2 = (pa >> np.uint64(32)) ^ c3 ^ k1
        c1 = pb & _M32
        c3 = pa & _M32
        c0, c2 = n0 & _M32, n2 & _M32
        k0 = (k0 + KA) & _M32
        k1 = (k1 + KB) & _M32
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _uniform(u32: np.ndarray) -> np.ndarray:
    """uint_to_uniform_float: x<0 -> -x-1, times 4.6566127342e-10 in fp32."""
    x = u32.astype(np.uint32).view(np.int32).astype(np.int64)
    x = np.where(x < 0, -x - 1, x)
    return x.astype(np.float32) * np.float32(4.6566127342e-10)


def _seed_key(key32: np.ndarray):
    """tl.rand(seed.to(int32), ...): the int32 seed is sign-extended to uint64, lo -> k0, hi -> k1."""
    k = key32.astype(np.uint32)
    hi = np.where((k >> np.uint32(31)) != 0, np.uint32(0xFFFFFFFF), np.uint32(0))
    return k, hi


def inject_mask_ref(n: int, ber: float, n_bits: int, seed: int, vectorized: bool = False):
    """Flip mask the reference injector would XOR into an n-element tensor.

    per-bit kernels (the ones ECCBackend.write uses), fault_injection_triton.py:247-249,:322-326:
        key = int32(seed*(N*n_bits) + off*n_bits + bit), counter = (off,0,0,0), word 0 of Philox.
    rand4x kernels, :82-85,:109 / :165-215: key = int32(seed*N + off + g*N), words 0..3 -> bits 4g..4g+3.
    Returns (mask uint32 [n], count uint8 [n]).
    """
    off = np.arange(n, dtype=np.int64)
    ber32 = np.float32(ber)
    mask = np.zeros(n, dtype=np.uint32)
    zero = np.zeros(n, dtype=np.uint32)
    if not vectorized:
        base = seed * (n * n_bits) + off * n_bits
        for bit in range(n_bits):
            k0, k1 = _seed_key(((base + bit) & 0xFFFFFFFF).astype(np.uint32))
            r0, _, _, _ = philox4x32_10(off & 0xFFFFFFFF, zero, zero, zero, k0, k1)
            mask |= (_uniform(r0) < ber32).astype(np.uint32) << np.uint32(bit)
    else:
        base = seed * n + off
        for g in range((n_bits + 3) // 4):
            k0, k1 = _seed_key(((base + g * n) & 0xFFFFFFFF).astype(np.uint32))
            r = philox4x32_10(off & 0xFFFFFFFF, zero, zero, zero, k0, k1)
            for j in range(4):
                bit = 4 * g + j
                if bit < n_bits:
                    mask |= (_uniform(r[j]) < ber32).astype(np.uint32) << np.uint32(bit)
    cnt = _popcnt_pat(mask.astype(np.int64).astype(np.int32) & 0x7FFFFFFF)
    return mask, cnt


def inject_ref(data: np.ndarray, ber: float, n_bits: int, seed: int, vectorized: bool = False):
    """inject_bit_errors_triton / _vectorized, fault_injection_triton.py:337-424, :434-496.
    Returns (corrupted, (total_flips, elements_affected))."""
    data = np.asarray(data)
    if ber <= 0:
        return data, (0, 0)
    flat = data.reshape(-1)
    mask, cnt = inject_mask_ref(flat.size, ber, n_bits, seed, vectorized)
    if flat.dtype == np.uint8:
        out = flat ^ mask.astype(np.uint8)
    elif flat.dtype == np.int32:
        out = flat ^ mask.astype(np.int32)
    else:
        raise ValueError(f"Unsupported dtype: {flat.dtype}. Use uint8 or int32.")
    return out.reshape(data.shape), (int(cnt.sum()), int((cnt > 0).sum()))


# ----------------------------------------------------------------------------------------
# Temporal interpolation of double-detected values (interpolation_triton.py:120-159)
# ----------------------------------------------------------------------------------------


def interpolate_double_errors(q: np.ndarray, etype: np.ndarray, seq_axis: int = -1) -> np.ndarray:
    """For etype==2: floor((q[max(i-1,0)] + q[min(i+1,L-1)])*0.5 + 0.5), neighbours taken from the
    pre-interpolation tensor; every output is clamp(.+0.5, 0, 15) truncated, so clean values above
    15 saturate exactly as in the reference."""
    q = np.asarray(q).astype(np.uint8)
    etype = np.asarray(etype).astype(np.uint8)
    if q.size == 0:
        return q.copy()
    qm = np.moveaxis(q, seq_axis, -1).astype(np.float32)
    em = np.moveaxis(etype, seq_axis, -1)
    L = qm.shape[-1]
    idx = np.arange(L)
    left = qm[..., np.maximum(idx - 1, 0)]
    right = qm[..., np.minimum(idx + 1, L - 1)]
    interp = (left + right) * np.float32(0.5)
    res = np.where(em == ERR_DOUBLE, interp, qm)
    res = np.maximum(np.float32(0.0), np.minimum(np.float32(15.0), res + np.float32(0.5)))
    return np.moveaxis(res.astype(np.uint8), -1, seq_axis)


# ----------------------------------------------------------------------------------------
# INT4 quantisation (kv_cache/paged_cache_ecc.py:328-334, kv_cache/ecc_shim.py:572-580)
# ----------------------------------------------------------------------------------------


def quant_scales(x: np.ndarray) -> np.ndarray:
    """absmax over the last axis / 7.0 in fp32, 0 -> 1.0."""
    a = np.abs(np.asarray(x).astype(np.float32)).max(axis=-1)
    s = (a / np.float32(7.0)).astype(np.float32)
    return np.where(s == 0, np.float32(1.0), s).astype(np.float32)


def quantize_int4(x: np.ndarray, scales: np.ndarray | None = None):
    """clamp(round_half_even(x/scale), -8, 7) + 8 -> uint8; IEEE fp32 divide."""
    xf = np.asarray(x).astype(np.float32)
    if scales is None:
        scales = quant_scales(xf)
    q = np.rint(xf / scales[..., None].astype(np.float32))
    q = np.clip(q, -8, 7) + 8
    return q.astype(np.uint8), scales


def dequantize_int4(q: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """(q - 8) * scale in fp32 (ecc_shim.py:1070-1071)."""
    return (q.astype(np.float32) - np.float32(8.0)) * scales[..., None].astype(np.float32)


def golay_cw_per_row(head_dim: int) -> int:
    return (head_dim + 2) // 3  # ecc_shim.py:260-263, :623-624


def encode_rows(q: np.ndarray, codec: str) -> np.ndarray:
    """INT4 rows [..., D] -> codeword rows; golay pads each row to a multiple of 3 with code 0
    (ecc_shim.py:669-682) -> int32 [..., (D+2)//3]."""
    if codec == "hamming84":
        return h84_encode(q)
    if codec == "hamming74":
        return h74_encode(q)
    if codec == "int4":
        return q.astype(np.uint8)
    if codec == "golay":
        D = q.shape[-1]
        ncw = golay_cw_per_row(D)
        pad = np.zeros(q.shape[:-1] + (ncw * 3,), dtype=np.uint8)
        pad[..., :D] = q
        return golay_encode(pad.reshape(-1, 3)).reshape(q.shape[:-1] + (ncw,))
    raise ValueError(f"Unknown codec: {codec}")


CODEC_BITS = {"hamming74": 7, "hamming84": 8, "golay": 24, "int4": 4}


# ----------------------------------------------------------------------------------------
# Paged cache write / read (kv_cache/ecc_shim.py:458-737, :739-1089)
# ----------------------------------------------------------------------------------------


class PagedCache:
    """SimpleBlockManager storage (ecc_shim.py:271-316): caches [NB, L, Hkv, bs*cw_per_token],
    scales [NB, L, Hkv, bs] fp32."""

    def __init__(self, num_blocks, block_size, num_layers, num_kv_heads, head_dim, codec):
        self.nb, self.bs, self.L, self.H, self.D, self.codec = num_blocks, block_size, num_layers, num_kv_heads, head_dim, codec
        self.cw = golay_cw_per_row(head_dim) if codec == "golay" else head_dim
        dt = np.int32 if codec == "golay" else np.uint8
        self.k_cache = np.zeros((num_blocks, num_layers, num_kv_heads, block_size * self.cw), dtype=dt)
        self.v_cache = np.zeros_like(self.k_cache)
        self.k_scales = np.zeros((num_blocks, num_layers, num_kv_heads, block_size), dtype=np.float32)
        self.v_scales = np.zeros_like(self.k_scales)


def cache_write(cache: PagedCache, k, v, layer, block_table_row, ber=0.0, seed=42, injection_count=0,
                k_mask=None, v_mask=None):
    """ECCBackend.write for hamming74/hamming84/golay/int4 (ecc_shim.py:572-737).

    k, v: [B, S, Hkv*D]. Every batch row writes positions [0,S) of the same sequence (Q11), so the
    last row wins; the injection counter advances once per (b, pos, head) with seed_K = seed+count,
    seed_V = seed_K+1 (Q10). k_mask/v_mask [B,S,Hkv,cw] replace the Philox stream when given.
    Returns the new injection_count.
    """
    B, S, _ = k.shape
    H, D, bs, cw = cache.H, cache.D, cache.bs, cache.cw
    kq, ks = quantize_int4(np.asarray(k).reshape(B, S, H, D))
    vq, vs = quantize_int4(np.asarray(v).reshape(B, S, H, D))
    kc = encode_rows(kq, cache.codec)
    vc = encode_rows(vq, cache.codec)
    nbits = CODEC_BITS[cache.codec]
    for b in range(B):
        for pos in range(S):
            blk = int(block_table_row[pos // bs])
            if blk < 0:
                continue
            slot = pos % bs
            for h in range(H):
                ke, ve = kc[b, pos, h], vc[b, pos, h]
                if k_mask is not None:
                    ke = ke ^ k_mask[b, pos, h].astype(ke.dtype)
                    ve = ve ^ v_mask[b, pos, h].astype(ve.dtype)
                elif ber > 0:
                    sd = seed + injection_count
                    injection_count += 1
                    ke, _ = inject_ref(ke, ber, nbits, sd)
                    ve, _ = inject_ref(ve, ber, nbits, sd + 1)
                cache.k_cache[blk, layer, h, slot * cw:(slot + 1) * cw] = ke
                cache.v_cache[blk, layer, h, slot * cw:(slot + 1) * cw] = ve
                cache.k_scales[blk, layer, h, slot] = ks[b, pos, h]
                cache.v_scales[blk, layer, h, slot] = vs[b, pos, h]
    return injection_count


def cache_gather(cache: PagedCache, layer, block_table_row, context_len):
    """Gather [ctx, Hkv, cw] codewords and [ctx, Hkv] scales for one sequence (ecc_shim.py:931-977)."""
    bs, cw = cache.bs, cache.cw
    kl, vl, ksl, vsl = [], [], [], []
    for blk_i in range((context_len + bs - 1) // bs):
        blk = int(block_table_row[blk_i])
        if blk < 0:
            continue
        for slot in range(min(bs, context_len - blk_i * bs)):
            kl.append(cache.k_cache[blk, layer, :, slot * cw:(slot + 1) * cw])
            vl.append(cache.v_cache[blk, layer, :, slot * cw:(slot + 1) * cw])
            ksl.append(cache.k_scales[blk, layer, :, slot])
            vsl.append(cache.v_scales[blk, layer, :, slot])
    if not kl:
        z = np.zeros((0, cache.H, cw), dtype=cache.k_cache.dtype)
        zs = np.zeros((0, cache.H), dtype=np.float32)
        return z, z.copy(), zs, zs.copy()
    return np.stack(kl), np.stack(vl), np.stack(ksl), np.stack(vsl)


def decode_rows(cw_rows: np.ndarray, codec: str, head_dim: int, use_interpolation: bool = False):
    """[ctx, H, cw] -> (INT4 [ctx, H, D], corrected, detected) with the reference's counter
    semantics (Q6) and, for hamming84+interpolation, interpolation along axis 0 (Q9)."""
    ctx, H, _ = cw_rows.shape
    if codec == "hamming74":
        d, f = h74_decode(cw_rows)
        return d, int(f.sum()), 0
    if codec == "hamming84":
        d, e = h84_decode(cw_rows)
        c, det = h84_counters(e)
        if use_interpolation and (e == ERR_DOUBLE).any():
            d = interpolate_double_errors(d, e, seq_axis=0)
        return d, c, det
    if codec == "golay":
        t, cnt = golay_decode(cw_rows.reshape(-1))
        bits, unc = golay_counters(cnt)
        d = t.reshape(ctx, H, -1)[:, :, :head_dim]
        return d, bits, unc
    if codec == "int4":
        return cw_rows.astype(np.uint8), 0, 0
    raise ValueError(codec)


def cache_read(cache: PagedCache, layer, block_table_row, context_len, use_interpolation=False):
    """Decode + [interpolate] + dequantise the whole context: returns
    (k fp32 [ctx,H,D], v fp32 [ctx,H,D], corrected, detected). ecc_shim.py:979-1071."""
    kc, vc, ks, vs = cache_gather(cache, layer, block_table_row, context_len)
    kd, c1, d1 = decode_rows(kc, cache.codec, cache.D, use_interpolation)
    vd, c2, d2 = decode_rows(vc, cache.codec, cache.D, use_interpolation)
    return dequantize_int4(kd, ks), dequantize_int4(vd, vs), c1 + c2, d1 + d2


# ----------------------------------------------------------------------------------------
# Attention over the decoded cache
# ----------------------------------------------------------------------------------------


def attention_prefill(q, k_float, v_float, num_kv_groups, causal):
    """_run_attention (ecc_shim.py:1138-1164) in fp32: q [B,Hq,Sq,D], k/v [ctx,Hkv,D]; GQA by
    repeat_interleave; SDPA top-left causal mask when causal."""
    q = np.asarray(q).astype(np.float32)
    k = np.repeat(k_float, num_kv_groups, axis=1).transpose(1, 0, 2)  # [Hq, ctx, D]
    v = np.repeat(v_float, num_kv_groups, axis=1).transpose(1, 0, 2)
    D = q.shape[-1]
    s = np.einsum("bhqd,hcd->bhqc", q, k) / np.sqrt(np.float32(D))
    if causal:
        Sq, C = s.shape[-2:]
        m = np.tril(np.ones((Sq, C), dtype=bool))
        s = np.where(m, s, -np.inf)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True)
    return np.einsum("bhqc,hcd->bhqd", p, v).astype(np.float32)


def paged_decode_attention(query, cache: PagedCache, block_table, context_lens, layer, sm_scale=None,
                           use_interpolation=False, v_use_k_scales=False, precision=np.float32):
    """reference_attention_ecc (attention_ecc.py:783-909) generalised to GQA (kv_head = q_head //
    group) and to separate K/V scales (Q15). query [B,Hq,D] -> fp32 [B,Hq,D]; an empty
    context yields zeros (Q14). Also returns (corrected, detected) summed over K and V.
    precision=np.float64 evaluates the same formula (fp32 dequantised K/V, as the reference has them) with float64
    products and sums: the yardstick for contexts long enough that fp32 summation order shows at the 1e-4 level."""
    query = np.asarray(query).astype(np.float32)
    B, Hq, D = query.shape
    g = Hq // cache.H
    if sm_scale is None:
        sm_scale = 1.0 / np.sqrt(D)
    out = np.zeros((B, Hq, D), dtype=np.float32)
    corr = det = 0
    for b in range(B):
        ctx = int(context_lens[b])
        if ctx <= 0:
            continue
        kc, vc, ks, vs = cache_gather(cache, layer, block_table[b], ctx)
        if kc.shape[0] == 0:
            continue
        kd, c1, d1 = decode_rows(kc, cache.codec, D, use_interpolation)
        vd, c2, d2 = decode_rows(vc, cache.codec, D, use_interpolation)
        corr += c1 + c2
        det += d1 + d2
        kf = dequantize_int4(kd, ks)
        vf = dequantize_int4(vd, ks if v_use_k_scales else vs)
        kf = np.repeat(kf, g, axis=1).transpose(1, 0, 2).astype(precision)
        vf = np.repeat(vf, g, axis=1).transpose(1, 0, 2).astype(precision)
        s = np.einsum("hd,hcd->hc", query[b].astype(precision), kf) * precision(sm_scale)
        s = s - s.max(axis=-1, keepdims=True)
        p = np.exp(s)
        p = p / p.sum(axis=-1, keepdims=True)
        out[b] = np.einsum("hc,hcd->hd", p, vf)
    return out, corr, det

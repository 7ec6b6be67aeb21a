"""
Flat codec entry points of the CUDA backend.  Same names, argument meaning, return tuples and error
behaviour as the reference's Triton wrappers; each one is a thin shim over one C-ABI call.

Statistics: the reference turns error counts into Python ints with a device sum per call
(hamming84_triton.py:341-342 ...).  Here the kernels accumulate block-level counters on the device
(one atomicAdd per CTA) and the wrapper reads the 64-byte counter vector once.
"""
from __future__ import annotations

import torch

from ... import _lib
from ..._lib import (CTR_AFFECTED, CTR_CORRECTED, CTR_DETECTED, CTR_FLIPPED, check, device_of, load, new_counters, ptr,
                     require_cuda, stream_of)
from .config import (DecodeResult, ErrorType, GolayDecodeResult, GOLAY_B_MATRIX, GOLAY_H_ROW_MASKS, GOLAY_UNCORRECTABLE,
                     HAMMING74_G, HAMMING74_H, HAMMING84_G, HAMMING84_H, SYNDROME_LUT_HAMMING74, SYNDROME_LUT_HAMMING84,
                     build_golay_syndrome_table)


def _flat_u8(t: torch.Tensor) -> torch.Tensor:
    return t.flatten().to(torch.uint8).contiguous()


# ----------------------------------------------------------------------------------------------- Hamming(7,4)
def hamming74_encode(int4_values: torch.Tensor) -> torch.Tensor:
    """hamming74_triton.py:170-201."""
    require_cuda(int4_values)
    flat = _flat_u8(int4_values)
    out = torch.empty_like(flat)
    with device_of(flat):
        check(load().eccq_h74_encode(ptr(flat), ptr(out), flat.numel(), stream_of(flat)), "hamming74_encode")
    return out.view(int4_values.shape)


def hamming74_decode(codewords: torch.Tensor, return_error_detected: bool = False):
    """hamming74_triton.py:218-277 -> (decoded, (n_corrected,)) | (decoded, flags, (n_corrected,))."""
    require_cuda(codewords)
    flat = _flat_u8(codewords)
    out = torch.empty_like(flat)
    flags = torch.empty_like(flat) if return_error_detected else None
    ctr = new_counters(flat.device)
    with device_of(flat):
        check(load().eccq_h74_decode(ptr(flat), ptr(out), ptr(flags), ptr(ctr), flat.numel(), stream_of(flat)), "hamming74_decode")
    n = int(ctr[CTR_CORRECTED])
    out = out.view(codewords.shape)
    if return_error_detected:
        return out, flags.view(codewords.shape), (n,)
    return out, (n,)


class Hamming74:
    """hamming74_triton.py:285-363: accepts tensors on any device, computes on ``device``."""
    G, H, SYNDROME_TO_POSITION = HAMMING74_G, HAMMING74_H, SYNDROME_LUT_HAMMING74

    def __init__(self, device: str = "cuda"):
        self.device = device
        self._G, self._H = self.G.to(device), self.H.to(device)
        self._syndrome_lut = self.SYNDROME_TO_POSITION.to(device)

    def encode(self, int4_values):
        return hamming74_encode(int4_values.to(self.device))

    def decode(self, codewords):
        data, flags, _ = hamming74_decode(codewords.to(self.device), return_error_detected=True)
        return data, flags.bool()

    encode_batch = encode
    decode_batch = decode


# ----------------------------------------------------------------------------------------------- Hamming(8,4)
def hamming84_encode(int4_values: torch.Tensor) -> torch.Tensor:
    """hamming84_triton.py:217-254."""
    require_cuda(int4_values)
    flat = _flat_u8(int4_values)
    out = torch.empty_like(flat)
    with device_of(flat):
        check(load().eccq_h84_encode(ptr(flat), ptr(out), flat.numel(), stream_of(flat)), "hamming84_encode")
    return out.view(int4_values.shape)


def hamming84_decode(codewords: torch.Tensor, return_error_types: bool = False):
    """hamming84_triton.py:281-351 -> (decoded, (corrected, detected)) | (decoded, error_types, (..)).
    Double-detected codewords keep their raw nibble (Q4)."""
    require_cuda(codewords)
    flat = _flat_u8(codewords)
    out = torch.empty_like(flat)
    et = torch.empty_like(flat) if return_error_types else None
    ctr = new_counters(flat.device)
    with device_of(flat):
        check(load().eccq_h84_decode(ptr(flat), ptr(out), ptr(et), ptr(ctr), flat.numel(), 0, stream_of(flat)), "hamming84_decode")
    c = ctr.tolist()
    stats = (c[CTR_CORRECTED], c[CTR_DETECTED])
    out = out.view(codewords.shape)
    if return_error_types:
        return out, et.view(codewords.shape), stats
    return out, stats


class Hamming84:
    """hamming84_triton.py:359-451."""
    G_74, H_74, SYNDROME_TO_POSITION = HAMMING84_G, HAMMING84_H, SYNDROME_LUT_HAMMING84

    def __init__(self, device: str = "cuda", on_double_error: str = "zero"):
        self.device = device
        self.on_double_error = on_double_error  # accepted and ignored, as in the reference (:394-403)
        self._G, self._H = self.G_74.to(device), self.H_74.to(device)
        self._syndrome_lut = self.SYNDROME_TO_POSITION.to(device)

    def encode(self, int4_values):
        return hamming84_encode(int4_values.to(self.device))

    def decode(self, codewords) -> DecodeResult:
        data, et, (c, d) = hamming84_decode(codewords.to(self.device), return_error_types=True)
        return DecodeResult(data=data, error_type=et, corrected_count=c, detected_count=d)


# ----------------------------------------------------------------------------------------------- Golay(24,12)
def golay_encode(triplets: torch.Tensor) -> torch.Tensor:
    """golay_triton.py:382-422: uint8 [N,3] (or [3]) -> int32 [N]."""
    require_cuda(triplets)
    if triplets.dim() == 1:
        triplets = triplets.unsqueeze(0)
    n = triplets.shape[0]
    flat = triplets.contiguous().flatten().to(torch.uint8)
    cw = torch.empty(n, dtype=torch.int32, device=triplets.device)
    with device_of(flat):
        check(load().eccq_golay_encode(ptr(flat), ptr(cw), n, stream_of(flat)), "golay_encode")
    return cw


def golay_decode(codewords: torch.Tensor, return_error_counts: bool = False):
    """golay_triton.py:425-498 -> (decoded [N,3], (bits_corrected, n_uncorrectable)) | (+ per-codeword counts)."""
    require_cuda(codewords)
    flat = codewords.flatten().to(torch.int32).contiguous()
    n = flat.numel()
    dec = torch.empty(n * 3, dtype=torch.uint8, device=flat.device)
    cnt = torch.empty(n, dtype=torch.uint8, device=flat.device) if return_error_counts else None
    ctr = new_counters(flat.device)
    with device_of(flat):
        check(load().eccq_golay_decode(ptr(flat), ptr(dec), ptr(cnt), ptr(ctr), n, stream_of(flat)), "golay_decode")
    c = ctr.tolist()
    stats = (c[CTR_CORRECTED], c[CTR_DETECTED])
    dec = dec.view(n, 3)
    if return_error_counts:
        return dec, cnt, stats
    return dec, stats


class Golay2412:
    """golay_triton.py:506-619."""
    UNCORRECTABLE = GOLAY_UNCORRECTABLE

    def __init__(self, device: str = "cuda"):
        self.device = device
        B = GOLAY_B_MATRIX.to(device)
        eye = torch.eye(12, dtype=torch.uint8, device=device)
        self.G, self.H, self.P = torch.cat([eye, B], dim=1), torch.cat([B.T, eye], dim=1), B
        self.syndrome_table = build_golay_syndrome_table().to(device)
        self.h_row_masks = torch.tensor(GOLAY_H_ROW_MASKS, dtype=torch.int64, device=device)

    def encode(self, triplets):
        return golay_encode(triplets.to(self.device)).to(torch.int64)

    def decode(self, codewords) -> GolayDecodeResult:
        data, (bits, unc) = golay_decode(codewords.to(torch.int32).to(self.device))
        return GolayDecodeResult(data=data, errors_corrected=bits, uncorrectable_count=unc)

    def verify_properties(self) -> bool:
        return bool(((self.G.float() @ self.H.T.float()) % 2).sum() == 0)


# ----------------------------------------------------------------------------------------------- fault injection
def _elem(data: torch.Tensor) -> int:
    if data.dtype == torch.uint8:
        return 1
    if data.dtype == torch.int32:
        return 4
    raise ValueError(f"Unsupported dtype: {data.dtype}. Use uint8 or int32.")


def _inject(data, ber, n_bits, seed, return_stats, vectorized):
    require_cuda(data)
    if ber <= 0:  # the reference returns the input tensor itself (fault_injection_triton.py:371-374)
        return (data, (0, 0)) if return_stats else data
    flat = data.flatten().contiguous()
    es = _elem(flat)
    out = torch.empty_like(flat)
    ctr = new_counters(flat.device) if return_stats else None
    with device_of(flat):
        check(load().eccq_inject_philox_ref(ptr(flat), ptr(out), None, es, int(n_bits), float(ber), int(seed), int(vectorized),
                                            ptr(ctr), flat.numel(), stream_of(flat)), "inject_bit_errors")
    out = out.view(data.shape)
    if return_stats:
        c = ctr.tolist()
        return out, (c[CTR_FLIPPED], c[CTR_AFFECTED])
    return out


def inject_bit_errors_triton(data, ber, n_bits, seed=0, return_stats=False):
    """fault_injection_triton.py:337-424.  Name kept for drop-in use; the flips are BIT-IDENTICAL to the
    reference's Triton Philox stream (key = int32(seed*N*n_bits + off*n_bits + bit), counter = off)."""
    return _inject(data, ber, n_bits, seed, return_stats, False)


def inject_bit_errors_triton_vectorized(data, ber, n_bits, seed=0, return_stats=False):
    """fault_injection_triton.py:434-496 (rand4x stream), bit-identical."""
    return _inject(data, ber, n_bits, seed, return_stats, True)


def inject_bit_errors_triton_batched(data, ber, n_bits, seed=0):
    """fault_injection_triton.py:427-431."""
    out, (total, _) = inject_bit_errors_triton(data, ber, n_bits, seed, return_stats=True)
    return out, total


inject_bit_errors = inject_bit_errors_triton


def inject_bit_errors_fast(data, ber, n_bits, seed=0, return_stats=False):
    """New in the CUDA backend: own Philox-4x32-10 stream with geometric gap sampling (one Philox call per
    16 bytes on average instead of n_bits calls per element).  Same distribution, different bits."""
    require_cuda(data)
    if ber <= 0:
        return (data, (0, 0)) if return_stats else data
    flat = data.flatten().contiguous()
    es = _elem(flat)
    out = torch.empty_like(flat)
    ctr = new_counters(flat.device) if return_stats else None
    with device_of(flat):
        check(load().eccq_inject_philox_fast(ptr(flat), ptr(out), es, int(n_bits), float(ber), int(seed) & (2**64 - 1), ptr(ctr),
                                             flat.numel(), stream_of(flat)), "inject_bit_errors_fast")
    out = out.view(data.shape)
    if return_stats:
        c = ctr.tolist()
        return out, (c[CTR_FLIPPED], c[CTR_AFFECTED])
    return out


def xor_flip_mask(data, mask, n_bits=None, return_stats=False):
    """New: apply an externally supplied flip mask (bit-exact parity runs)."""
    require_cuda(data)
    require_cuda(mask, "Mask")
    if data.shape != mask.shape or data.dtype != mask.dtype:
        raise ValueError("mask must match data in shape and dtype")
    flat, m = data.flatten().contiguous(), mask.flatten().contiguous()
    es = _elem(flat)
    out = torch.empty_like(flat)
    ctr = new_counters(flat.device) if return_stats else None
    with device_of(flat):
        check(load().eccq_xor_mask(ptr(flat), ptr(m), ptr(out), es, int(n_bits or 8 * es), ptr(ctr), flat.numel(), stream_of(flat)),
              "xor_flip_mask")
    out = out.view(data.shape)
    if return_stats:
        c = ctr.tolist()
        return out, (c[CTR_FLIPPED], c[CTR_AFFECTED])
    return out


# ----------------------------------------------------------------------------------------------- interpolation
def interpolate_double_errors(q, error_type, original_shape=None, seq_dim=-1):
    """interpolation_triton.py:162-265.  No permute/contiguous copies and no `.any()` host sync: the kernel
    walks the [outer, seq, inner] view in place (neighbours at stride `inner`)."""
    require_cuda(q)
    require_cuda(error_type, "Error type")
    assert q.shape == error_type.shape, "Shape mismatch between q and error_type"
    if q.numel() == 0:
        return q.clone()
    # dims <= 2 always interpolate along the last axis (the reference ignores seq_dim there, :205-212)
    sd = q.dim() - 1 if q.dim() <= 2 else (seq_dim if seq_dim >= 0 else q.dim() + seq_dim)
    qc, ec = q.contiguous().to(torch.uint8), error_type.contiguous().to(torch.uint8)
    length = q.shape[sd] if q.dim() else 1
    outer = 1
    for s in q.shape[:sd]:
        outer *= s
    inner = 1
    for s in q.shape[sd + 1:]:
        inner *= s
    out = torch.empty_like(qc)
    with device_of(qc):
        check(load().eccq_interpolate(ptr(qc), ptr(ec), ptr(out), outer, length, inner, stream_of(qc)), "interpolate_double_errors")
    return out


def interpolate_double_errors_1d(q, error_type):
    return interpolate_double_errors(q, error_type, seq_dim=-1)


interpolate_double_errors_autotuned = interpolate_double_errors


# ----------------------------------------------------------------------------------------------- fused quantise / dequantise
def _rows(x: torch.Tensor):
    row = x.shape[-1]
    flat = x.unsqueeze(0) if x.dim() == 1 else x.reshape(-1, row)
    return flat.contiguous(), flat.shape[0], row


def _fused_qe(x: torch.Tensor, codec: int):
    require_cuda(x)
    flat, rows, row = _rows(x)
    if flat.dtype not in (torch.float16, torch.float32, torch.bfloat16):
        flat = flat.float()
    cw = torch.empty(rows, row, dtype=torch.uint8, device=x.device)
    scales = torch.empty(rows, dtype=torch.float32, device=x.device)
    with device_of(flat):
        check(load().eccq_quantize_rows(ptr(flat), _lib.dtype_code(flat.dtype), ptr(cw), ptr(scales), rows, row, 0, codec,
                                        stream_of(flat)), "fused_quantize_encode")
    if x.dim() == 1:
        return cw.squeeze(0), scales
    return cw.view(x.shape), scales.view(x.shape[:-1])


def fused_quantize_encode_hamming84(input_tensor):
    """fused_kernels.py:97-160.  Rounding follows the torch path the shim uses (IEEE divide + round-half-even,
    SURVEY Q1), so it equals quantise-then-encode exactly (the Triton kernel is only ~99.9% equal, Q2)."""
    return _fused_qe(input_tensor, _lib.CODEC_H84)


def fused_quantize_encode_hamming74(input_tensor):
    """fused_kernels.py:222-269."""
    return _fused_qe(input_tensor, _lib.CODEC_H74)


def fused_decode_dequantize_hamming84(codewords, scales, output_dtype=torch.float32):
    """fused_kernels.py:372-437 -> (dequantised, n_corrected); double-detected nibbles are zeroed (Q4)."""
    require_cuda(codewords)
    require_cuda(scales, "Scales")
    flat, rows, row = _rows(codewords)
    sc = scales.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty(rows, row, dtype=output_dtype, device=codewords.device)
    ctr = new_counters(codewords.device)
    with device_of(flat):
        check(load().eccq_dequantize_rows(ptr(flat.to(torch.uint8)), ptr(sc), ptr(out), _lib.dtype_code(output_dtype), rows, row,
                                          _lib.CODEC_H84, 1, ptr(ctr), stream_of(flat)), "fused_decode_dequantize")
    out = out.squeeze(0) if codewords.dim() == 1 else out.view(codewords.shape)
    return out, int(ctr[CTR_CORRECTED])

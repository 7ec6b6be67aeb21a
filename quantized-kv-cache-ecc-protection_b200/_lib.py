"""
ctypes binding of libeccq_sm100a.so (C ABI declared in include/eccq.h).

PyTorch is plumbing only: it owns device memory and streams; every call below passes raw
``tensor.data_ptr()`` values and the current CUDA stream handle to the C ABI.  There is no CPU or
eager-PyTorch fallback: if the shared library is missing or a tensor is not on a CUDA device the
call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("ECCQ_LIB", _PKG / "lib" / "libeccq_sm100a.so"))

ECCQ_F32, ECCQ_F16, ECCQ_BF16 = 0, 1, 2
CODEC_H74, CODEC_H84, CODEC_GOLAY, CODEC_INT4 = 0, 1, 2, 3
CODEC_IDS = {"hamming74": CODEC_H74, "hamming84": CODEC_H84, "golay": CODEC_GOLAY, "int4": CODEC_INT4}
CODEC_BITS = {"hamming74": 7, "hamming84": 8, "golay": 24, "int4": 4}
ECCQ_EUNSUPPORTED = -3
CTR_CORRECTED, CTR_DETECTED, CTR_FLIPPED, CTR_AFFECTED, CTR_VALUES, CTR_INTERPOLATED, NCOUNTERS = 0, 1, 2, 3, 4, 5, 8
INJECT_NONE, INJECT_MASK, INJECT_PHILOX_REF, INJECT_PHILOX_FAST = 0, 1, 2, 3
INJECT_MODES = {"none": INJECT_NONE, "mask": INJECT_MASK, "philox_ref": INJECT_PHILOX_REF, "philox_fast": INJECT_PHILOX_FAST}

_DTYPES = {torch.float32: ECCQ_F32, torch.float16: ECCQ_F16, torch.bfloat16: ECCQ_BF16}


class EccqError(RuntimeError):
    pass


class CacheDesc(ctypes.Structure):
    """struct eccq_cache_desc (include/eccq.h)."""
    _fields_ = [
        ("k_cache", c_void_p), ("v_cache", c_void_p), ("k_scales", c_void_p), ("v_scales", c_void_p),
        ("num_blocks", c_int32), ("num_layers", c_int32), ("num_kv_heads", c_int32), ("head_dim", c_int32),
        ("block_size", c_int32), ("codec", c_int32),
    ]


_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)   -- one entry per prototype in include/eccq.h
    "eccq_abi_version": (c_int, []),
    "eccq_last_error": (c_char_p, []),
    "eccq_device_info": (c_int, [c_int, POINTER(c_int)]),
    "eccq_h74_encode": (c_int, [_P, _P, c_int64, _P]),
    "eccq_h74_decode": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "eccq_h84_encode": (c_int, [_P, _P, c_int64, _P]),
    "eccq_h84_decode": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "eccq_golay_encode": (c_int, [_P, _P, c_int64, _P]),
    "eccq_golay_decode": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "eccq_golay_syndrome_table": (c_int, [POINTER(c_int32)]),
    "eccq_xor_mask": (c_int, [_P, _P, _P, c_int, c_int, _P, c_int64, _P]),
    "eccq_inject_philox_ref": (c_int, [_P, _P, _P, c_int, c_int, c_float, c_int64, c_int, _P, c_int64, _P]),
    "eccq_inject_philox_fast": (c_int, [_P, _P, c_int, c_int, c_float, c_uint64, _P, c_int64, _P]),
    "eccq_interpolate": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P]),
    "eccq_quantize_rows": (c_int, [_P, c_int, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "eccq_dequantize_rows": (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P, _P]),
    "eccq_cache_write": (c_int, [POINTER(CacheDesc), _P, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P,
                                 c_float, c_int64, c_int64, _P, _P]),
    "eccq_cache_read": (c_int, [POINTER(CacheDesc), _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "eccq_cache_correct_pages": (c_int, [POINTER(CacheDesc), _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "eccq_paged_rows_write": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_int64, c_int64, _P, _P]),
    "eccq_paged_rows_read": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, _P]),
    "eccq_attention_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "eccq_paged_decode_attention": (c_int, [POINTER(CacheDesc), _P, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_float,
                                            c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "eccq_paged_prefill_attention": (c_int, [POINTER(CacheDesc), _P, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float,
                                             c_int, c_int, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library once; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EccqError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or sh {_PKG / 'build.sh'}). This package has no CPU / eager fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.eccq_abi_version() != 1:
        raise EccqError(f"ABI mismatch: library reports {lib.eccq_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().eccq_last_error()
        raise EccqError(f"{what or 'eccq'} failed (code {rc}): {msg.decode() if msg else '?'}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise ValueError(f"Unsupported dtype: {dt}. Use float32, float16 or bfloat16.") from None


def require_cuda(t: torch.Tensor, name: str = "Input"):
    # same exception type and wording as the reference wrappers (e.g. hamming84_triton.py:237)
    assert t.is_cuda, f"{name} must be on CUDA device"


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_of(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class device_of:
    """Make the tensor's device current for the duration of a call (per-device table upload, SM count)."""

    def __init__(self, t: torch.Tensor):
        # already current (the single-GPU-per-process case): nothing to switch, nothing to restore
        self.guard = None if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        return self.guard.__exit__(*exc) if self.guard is not None else False


def new_counters(device) -> torch.Tensor:
    """Device-side uint64[NCOUNTERS] (stored as int64) the kernels accumulate into; read lazily."""
    return torch.zeros(NCOUNTERS, dtype=torch.int64, device=device)


def cache_desc(k_cache, v_cache, k_scales, v_scales, head_dim: int, block_size: int, codec: str) -> CacheDesc:
    nb, nl, nh, _ = k_cache.shape
    return CacheDesc(k_cache.data_ptr(), v_cache.data_ptr(), k_scales.data_ptr(), v_scales.data_ptr(),
                     nb, nl, nh, head_dim, block_size, CODEC_IDS[codec])

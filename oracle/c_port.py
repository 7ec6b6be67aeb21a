"""ctypes loader for oracle/eccq_oracle.c (CPU checker / CPU baseline only -- see that file's header)."""
from __future__ import annotations

import ctypes
import subprocess
from ctypes import c_float, c_int, c_int64, c_void_p
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "eccq_oracle.c"
_lib = None


def _cpu_tag() -> str:
    """The library is compiled with -march=native, and the build directory travels with the repository to other machines (the GPU
    box): the file name carries a digest of this CPU's feature flags, so a host with different flags compiles its own copy instead
    of loading code it may not be able to execute."""
    import hashlib
    try:
        flags = next(line for line in open("/proc/cpuinfo") if line.startswith("flags"))
    except (OSError, StopIteration):
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


LIB = HERE / "_build" / f"liboracle-{_cpu_tag()}.so"


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", str(SRC), "-o", str(LIB), "-lm"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(build()))
        lib.orc_num_threads.restype = c_int
        lib.orc_set_num_threads.argtypes = [c_int]
        lib.orc_encode_h84.argtypes = [c_void_p, c_void_p, c_int64]
        lib.orc_quantize_encode_h84.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p]
        lib.orc_xor_mask.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
        lib.orc_decode_h84.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        lib.orc_decode_interp_dequant_h84.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]
        lib.orc_init()
        _lib = lib
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def num_threads() -> int:
    return int(load().orc_num_threads())


def set_num_threads(n: int) -> int:
    """Use `n` OpenMP threads from now on, whatever OMP_NUM_THREADS says; returns the count in effect."""
    load().orc_set_num_threads(int(n))
    return num_threads()


def encode_h84(nib: np.ndarray, out: np.ndarray | None = None):
    nib = np.ascontiguousarray(nib, dtype=np.uint8)
    out = np.empty_like(nib) if out is None else out
    load().orc_encode_h84(_p(nib), _p(out), nib.size)
    return out


def quantize_encode_h84(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, D = x.reshape(-1, x.shape[-1]).shape
    cw = np.empty((rows, D), dtype=np.uint8)
    sc = np.empty(rows, dtype=np.float32)
    load().orc_quantize_encode_h84(_p(x), rows, D, _p(cw), _p(sc))
    return cw.reshape(x.shape), sc.reshape(x.shape[:-1])


def xor_mask(cw: np.ndarray, mask: np.ndarray, out: np.ndarray | None = None):
    out = np.empty_like(cw) if out is None else out
    load().orc_xor_mask(_p(cw), _p(mask), _p(out), cw.size)
    return out


def decode_h84(cw: np.ndarray):
    cw = np.ascontiguousarray(cw, dtype=np.uint8)
    nib, et = np.empty_like(cw), np.empty_like(cw)
    counts = np.zeros(2, dtype=np.int64)
    load().orc_decode_h84(_p(cw), cw.size, _p(nib), _p(et), _p(counts))
    return nib, et, (int(counts[0]), int(counts[1]))


def decode_interp_dequant_h84(cw: np.ndarray, scales: np.ndarray, use_interp: bool, out: np.ndarray | None = None):
    """cw [ctx, H, D] uint8, scales [ctx, H] fp32 -> (fp32 [ctx, H, D], (corrected, detected))."""
    ctx, H, D = cw.shape
    out = np.empty((ctx, H, D), dtype=np.float32) if out is None else out
    counts = np.zeros(2, dtype=np.int64)
    load().orc_decode_interp_dequant_h84(_p(cw), _p(scales), ctx, H, D, int(use_interp), _p(out), _p(counts))
    return out, (int(counts[0]), int(counts[1]))

"""
Codec constants of the CUDA backend, same public names as the reference's
``ecc_codecs/triton_kernels/config.py`` (cited per symbol).  The device kernels do not read these
tensors: the Hamming table is folded into boolean logic and the Golay tables are built in C++ by the
same enumeration (csrc/eccq_internal.h); these objects exist for callers and for algebraic checks.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple

import torch

# launch geometry of the CUDA kernels (the reference's *_BLOCK_SIZE were Triton tile sizes, config.py:34-38)
HAMMING74_BLOCK_SIZE = HAMMING84_BLOCK_SIZE = FAULT_INJECTION_BLOCK_SIZE = INTERPOLATION_BLOCK_SIZE = 256 * 16 * 4
GOLAY_BLOCK_SIZE = 2048

_PHYSICAL = {"hamming74": torch.uint8, "hamming84": torch.uint8, "golay": torch.int32, "int4": torch.uint8,
             "none": torch.float16}
_CW_BITS = {"hamming74": 7, "hamming84": 8, "golay": 24}
_DATA_BITS = {"hamming74": 4, "hamming84": 4, "golay": 12}


def _lookup(table, codec):
    if codec not in table:
        raise ValueError(f"Unknown codec: {codec}")
    return table[codec]


def get_physical_dtype(codec: str) -> torch.dtype:      # config.py:41-70
    return _lookup(_PHYSICAL, codec)


def get_codeword_bits(codec: str) -> int:               # config.py:73-92
    return _lookup(_CW_BITS, codec)


def get_data_bits(codec: str) -> int:                   # config.py:95-114
    return _lookup(_DATA_BITS, codec)


class ErrorType:                                        # config.py:197-214
    NO_ERROR = 0
    SINGLE_CORRECTED = 1
    DOUBLE_DETECTED = 2
    PARITY_ONLY = 3


class DecodeResult(NamedTuple):                         # config.py:222-242
    data: torch.Tensor
    error_type: torch.Tensor
    corrected_count: int
    detected_count: int


class GolayDecodeResult(NamedTuple):                    # config.py:245-267
    data: torch.Tensor
    errors_corrected: int
    uncorrectable_count: int


# Hamming(7,4): codeword bit order [d0 d1 d2 d3 p0 p1 p2]; p0=d0^d1^d3, p1=d0^d2^d3, p2=d1^d2^d3 (config.py:285-310)
_PARITY_OF = ((0, 1, 3), (0, 2, 3), (1, 2, 3))
HAMMING74_G = torch.tensor([[int(c == r) for c in range(4)] + [int(r in p) for p in _PARITY_OF] for r in range(4)],
                           dtype=torch.uint8)
HAMMING74_H = torch.tensor([[int(c in p) for c in range(4)] + [int(c == r) for c in range(3)] for r, p in enumerate(_PARITY_OF)],
                           dtype=torch.uint8)
HAMMING84_G, HAMMING84_H = HAMMING74_G, HAMMING74_H


def _syndrome_to_position():
    lut = [-1] * 8
    for bit in range(7):
        col = HAMMING74_H[:, bit].tolist()
        lut[col[0] | (col[1] << 1) | (col[2] << 2)] = bit
    return torch.tensor(lut, dtype=torch.int8)


SYNDROME_LUT_HAMMING74 = _syndrome_to_position()        # [-1,4,5,0,6,1,2,3], config.py:131-143
SYNDROME_LUT_HAMMING84 = SYNDROME_LUT_HAMMING74.clone() # config.py:149-161

# Golay B matrix (config.py:329-347): rows as 12-bit masks, bit j of row i = B[i][j]; symmetric, B.B = I (mod 2)
_GOLAY_B_ROWS = (0xA3B, 0xD1D, 0xE8E, 0xB47, 0xDA3, 0xED1, 0xF68, 0xBB4, 0x9DA, 0x8ED, 0xC76, 0x7FF)
GOLAY_B_MATRIX = torch.tensor([[(r >> j) & 1 for j in range(12)] for r in _GOLAY_B_ROWS], dtype=torch.uint8)
GOLAY_UNCORRECTABLE = 0xFFFFFFFF                        # config.py:351
GOLAY_B_COLS = tuple(int(sum(int(GOLAY_B_MATRIX[j, i]) << j for j in range(12))) for i in range(12))
GOLAY_H_ROW_MASKS = tuple(GOLAY_B_COLS[i] | (1 << (12 + i)) for i in range(12))   # config.py:354-379


def build_golay_syndrome_table() -> torch.Tensor:       # config.py:403-457
    """4096-entry syndrome -> error-pattern table (int32, -1 = uncorrectable): the very table the
    kernels stage in shared memory, fetched from the library."""
    from ... import _lib
    buf = (ctypes.c_int32 * 4096)()
    _lib.check(_lib.load().eccq_golay_syndrome_table(buf), "golay_syndrome_table")
    return torch.tensor(list(buf), dtype=torch.int32)

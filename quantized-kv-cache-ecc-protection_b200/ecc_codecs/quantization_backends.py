"""
The quantiser registry entry the CUDA backend adds (north_star: "ecc_codecs/quantization_backends ... gains the new
backend").  Interface of ecc_codecs/quantization_backends.py:83-163 / :672-711 (QuantizedTensor, QuantizerBackend,
QUANTIZER_BACKENDS, get_quantizer, list_backends); the reference's seven torch strategies are out of scope, the
registry below holds the CUDA per-token strategy only and INTEGRATION.md shows the one-line registration in the
reference tree.

Arithmetic = PerTokenQuantizer (:241-269): scale = clamp(absmax, 1e-8) / 7, q = clamp(round(x/scale) + 8, 0, 15),
computed by eccq_quantize_rows (zero_guard=1) / eccq_dequantize_rows on the GPU.  Inputs on any device are accepted
and results are returned on the caller's device (the reference's registry tests pass CPU fp16 tensors,
tests/test_quantization_backends.py:88-91): that is a host<->device copy around the kernel, not a CPU compute path.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum

import torch

from .. import _lib
from .._lib import check, device_of, load, ptr, stream_of


class QuantizationMode(Enum):
    KEY = "key"
    VALUE = "value"
    GENERIC = "generic"


@dataclass
class QuantizationConfig:
    bits: int = 4
    symmetric: bool = True
    block_size: int = 32
    group_size: int = None
    dtype: torch.dtype = torch.float16
    residual_length: int = 128
    channel_axis: int = -1
    device: str = "cuda"


@dataclass
class QuantizedTensor:
    data: torch.Tensor
    scales: torch.Tensor
    zero_points: torch.Tensor = None
    original_shape: tuple = None
    mode: QuantizationMode = QuantizationMode.GENERIC
    metadata: dict = None

    def to(self, device):
        zp = self.zero_points.to(device) if self.zero_points is not None else None
        return QuantizedTensor(self.data.to(device), self.scales.to(device), zp, self.original_shape, self.mode, self.metadata)


class QuantizerBackend(ABC):
    QMIN, QMAX, ZERO_POINT = 0, 15, 8

    def __init__(self, config=None):
        self.config = config or QuantizationConfig()

    @abstractmethod
    def quantize(self, x, mode=QuantizationMode.GENERIC): ...

    @abstractmethod
    def dequantize(self, q): ...

    def quantize_simple(self, x):
        r = self.quantize(x)
        return r.data, r.scales

    def dequantize_simple(self, q, scales):
        return self.dequantize(QuantizedTensor(data=q, scales=scales))

    @property
    def name(self):
        return self.__class__.__name__


class CudaPerTokenQuantizer(QuantizerBackend):
    """Per-token symmetric INT4 on the GPU (one warp per row)."""

    def _device(self, x):
        return x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def quantize(self, x, mode=QuantizationMode.GENERIC):
        home = x.device
        xd = x.to(self._device(x))
        if xd.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            xd = xd.float()
        xd = xd.contiguous()
        row = xd.shape[-1]
        rows = xd.numel() // max(row, 1)
        q = torch.empty(xd.shape, dtype=torch.uint8, device=xd.device)
        scales = torch.empty(rows, dtype=torch.float32, device=xd.device)
        if rows:
            with device_of(xd):
                check(load().eccq_quantize_rows(ptr(xd), _lib.dtype_code(xd.dtype), ptr(q), ptr(scales), rows, row, 1, -1, stream_of(xd)),
                      "CudaPerTokenQuantizer.quantize")
        scales = scales.view(xd.shape[:-1]).to(x.dtype if x.dtype.is_floating_point else torch.float32)
        return QuantizedTensor(data=q.to(home), scales=scales.to(home), original_shape=x.shape, mode=mode)

    def dequantize(self, qt):
        home = qt.data.device
        qd = qt.data.to(self._device(qt.data)).to(torch.uint8).contiguous()
        sc = qt.scales.to(qd.device).float().reshape(-1).contiguous()
        out_dtype = self.config.dtype if self.config.dtype in (torch.float32, torch.float16, torch.bfloat16) else torch.float16
        out = torch.empty(qd.shape, dtype=out_dtype, device=qd.device)
        row = qd.shape[-1]
        rows = qd.numel() // max(row, 1)
        if rows:
            with device_of(qd):
                check(load().eccq_dequantize_rows(ptr(qd), ptr(sc), ptr(out), _lib.dtype_code(out_dtype), rows, row, -1, 0, None,
                                                  stream_of(qd)), "CudaPerTokenQuantizer.dequantize")
        return out.to(home)


QUANTIZER_BACKENDS = {"cuda_per_token": CudaPerTokenQuantizer}


def get_quantizer(backend="cuda_per_token", config=None):
    key = backend.lower().replace("-", "_")
    if key not in QUANTIZER_BACKENDS:
        raise ValueError(f"Unknown quantizer backend '{backend}'. Available backends: {', '.join(QUANTIZER_BACKENDS)}")
    return QUANTIZER_BACKENDS[key](config)


def list_backends():
    return list(QUANTIZER_BACKENDS)

"""kv_cache/paged_cache_ecc.py:302-397 on the CUDA backend."""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import check, device_of, load, ptr, require_cuda, stream_of
from ..ecc_codecs.cuda_kernels import golay_encode, hamming74_encode, hamming84_encode


def _rows(t: torch.Tensor, dim: int):
    if dim not in (-1, t.dim() - 1):
        t = t.movedim(dim, -1)
    if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        t = t.float()
    return t.contiguous()


def compute_quantization_scales(tensor: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """absmax over `dim` / 7 in fp32, zero -> 1.0 (paged_cache_ecc.py:302-334); one warp per row on the GPU."""
    require_cuda(tensor)
    t = _rows(tensor, dim)
    rows, row = t.numel() // max(t.shape[-1], 1), t.shape[-1]
    scales = torch.empty(rows, dtype=torch.float32, device=t.device)
    with device_of(t):
        check(load().eccq_quantize_rows(ptr(t), _lib.dtype_code(t.dtype), None, ptr(scales), rows, row, 0, -1, stream_of(t)),
              "compute_quantization_scales")
    return scales.view(t.shape[:-1])


def quantize_int4(tensor: torch.Tensor, zero_guard: int = 0):
    """INT4 quantise per last-dim row (ecc_shim.py:572-580): returns (uint8 nibbles, fp32 scales)."""
    require_cuda(tensor)
    t = _rows(tensor, -1)
    rows, row = t.numel() // max(t.shape[-1], 1), t.shape[-1]
    q = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    scales = torch.empty(rows, dtype=torch.float32, device=t.device)
    with device_of(t):
        check(load().eccq_quantize_rows(ptr(t), _lib.dtype_code(t.dtype), ptr(q), ptr(scales), rows, row, zero_guard, -1, stream_of(t)),
              "quantize_int4")
    return q, scales.view(t.shape[:-1])


def write_kv_to_cache_simple(kv: torch.Tensor, codec="hamming84", scale=None):
    """paged_cache_ecc.py:337-397: non-paged quantise + encode; inputs on any device are moved to the GPU as
    the reference does (:376, :388)."""
    kv = kv if kv.is_cuda else kv.cuda()
    if scale is None:
        q, scale = quantize_int4(kv)
    else:
        q = (torch.round(kv / scale.to(kv.device).unsqueeze(-1)).clamp(-8, 7) + 8).to(torch.uint8)
    if codec == "hamming84":
        return hamming84_encode(q), scale
    if codec == "hamming74":
        return hamming74_encode(q), scale
    if codec == "golay":
        flat = q.flatten()
        pad = (3 - flat.numel() % 3) % 3
        if pad:
            flat = torch.cat([flat, flat.new_zeros(pad)])
        return golay_encode(flat.view(-1, 3)), scale
    return q, scale

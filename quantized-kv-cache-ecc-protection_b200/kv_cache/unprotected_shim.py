"""
The unprotected INT4 KV cache of the reference (kv_cache/unprotected_shim.py:1-574) on the CUDA backend: the same INT4
quantisation and paged storage as the ECC modes, no code; bit errors are injected AFTER the cache read and BEFORE
dequantisation (unprotected_shim.py:276-302), one Philox stream over the whole gathered context per attend() call.

Same public surface: ``UnprotectedDummyCache``, ``UnprotectedShimConfig``, ``UnprotectedBackend``,
``UnprotectedPagedAttentionShim``, ``patch_model_with_unprotected_attention``, ``reset_unprotected_cache``,
``get_unprotected_stats``.  The write is the fused cache-write kernel with the ``int4`` codec (one launch per layer instead of
S*Hkv slice assignments), the read the fused gather kernel; the flip stream is bit-identical with the reference's.

Reference quirk kept on purpose: ``_run_attention`` passes ``is_causal=True`` also for q_len == 1 (unprotected_shim.py:323-328),
i.e. SDPA's top-left mask lets a single decode query see only key 0.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch
import torch.nn.functional as F

from ..ecc_codecs.cuda_kernels import inject_bit_errors_triton
from .attention_ecc import read_cache
from .ecc_shim import (ECCBackend, ECCDummyCache, ECCPagedAttentionShim, ECCShimConfig, SimpleBlockManager, _find_rotary_embedding,
                       _get_attention_params, _model_layout)


class UnprotectedDummyCache(ECCDummyCache):
    """unprotected_shim.py:32-80 (the same placeholder as ECCDummyCache)."""


class UnprotectedShimConfig:
    """unprotected_shim.py:83-108 (+ ``backend`` / ``append`` as ECCShimConfig)."""

    def __init__(self, ber=0.0, block_size=16, num_blocks=256, inject_errors=False, seed=42, backend="cuda", append=False):
        if backend != "cuda":
            raise ValueError(f"Unsupported backend: '{backend}'. This package implements backend='cuda' only.")
        self.ber, self.block_size, self.num_blocks = ber, block_size, num_blocks
        self.inject_errors, self.seed = inject_errors, seed
        self.codec = "int4"          # SimpleBlockManager layout without ECC
        self.backend, self.append = backend, append


class UnprotectedBackend:
    """unprotected_shim.py:111-333."""

    def __init__(self, manager, config, num_heads):
        self.manager, self.config, self.num_heads = manager, config, num_heads
        self.num_kv_heads, self.head_dim = manager.num_kv_heads, manager.head_dim
        self.num_kv_groups = num_heads // self.num_kv_heads
        self._injection_count = 0
        self._bits_flipped = 0
        self._total_bits = 0
        # the write IS the int4 cache write of the ECC backend without injection (same kernel, same cache image)
        self._writer = ECCBackend(manager, ECCShimConfig(codec="int4", ber=0.0, block_size=config.block_size, num_blocks=config.num_blocks,
                                                          inject_errors=False, seed=config.seed, append=config.append), num_heads)

    def reset_stats(self):
        self._injection_count = self._bits_flipped = self._total_bits = 0

    def write(self, k, v, layer_idx, seq_id=0):
        self._writer.write(k, v, layer_idx, seq_id=seq_id)

    def attend(self, q, layer_idx, seq_id=0):
        m, cfg = self.manager, self.config
        ctx = m.seq_layer_len.get((seq_id, layer_idx), 0) if cfg.append else m.get_context_len(seq_id)
        if ctx == 0:
            return torch.zeros_like(q)
        bt = m.block_table[seq_id:seq_id + 1]
        cl = torch.tensor([ctx], dtype=torch.int32).to(q.device, non_blocking=True)
        # corrected == raw nibbles for the int4 codec; [1, ctx, Hkv, D]
        _, _, kn, vn = read_cache(m.k_cache, m.v_cache, m.k_scales, m.v_scales, bt, cl, layer_idx, m.block_size, self.head_dim, "int4",
                                  return_nibbles=True, max_context=ctx)
        nblk = -(-ctx // m.block_size)
        blocks = bt[0, :nblk].long()
        ks = m.k_scales[blocks, layer_idx].permute(0, 2, 1).reshape(-1, self.num_kv_heads)[:ctx]      # [ctx, Hkv]
        vs = m.v_scales[blocks, layer_idx].permute(0, 2, 1).reshape(-1, self.num_kv_heads)[:ctx]
        kn, vn = kn[0], vn[0]
        if cfg.inject_errors and cfg.ber > 0:
            seed = cfg.seed + self._injection_count
            self._injection_count += 1
            self._total_bits += kn.numel() * 4 * 2
            kf, kst = inject_bit_errors_triton(kn.flatten(), cfg.ber, n_bits=4, seed=seed, return_stats=True)
            vf, vst = inject_bit_errors_triton(vn.flatten(), cfg.ber, n_bits=4, seed=seed + 1, return_stats=True)
            self._bits_flipped += kst[0] + vst[0]
            kn, vn = kf.view(kn.shape), vf.view(vn.shape)
        k_float = (kn.float() - 8.0) * ks.unsqueeze(-1)
        v_float = (vn.float() - 8.0) * vs.unsqueeze(-1)
        return self._run_attention(q, k_float, v_float)

    def _run_attention(self, q, k_float, v_float):
        k = k_float.permute(1, 0, 2).unsqueeze(0).to(q.dtype)
        v = v_float.permute(1, 0, 2).unsqueeze(0).to(q.dtype)
        if k.shape[0] != q.shape[0]:
            k, v = k.expand(q.shape[0], -1, -1, -1), v.expand(q.shape[0], -1, -1, -1)
        return F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=self.num_kv_groups > 1)


class UnprotectedPagedAttentionShim(ECCPagedAttentionShim):
    """unprotected_shim.py:336-452: the same wrapper as the ECC shim around an UnprotectedBackend (LLaMA and GPT-2 layouts)."""

    def __init__(self, original_attn, layer_idx, backend, rotary_emb, model_type="llama"):
        super().__init__(original_attn, layer_idx, backend, rotary_emb, model_type)


@contextmanager
def patch_model_with_unprotected_attention(model, config, num_blocks=256):
    """unprotected_shim.py:455-547."""
    layers, model_type, attr = _model_layout(model)
    num_heads, num_kv_heads, head_dim = _get_attention_params(getattr(layers[0], attr))
    manager = SimpleBlockManager(num_blocks=num_blocks, block_size=config.block_size, num_layers=len(layers), num_kv_heads=num_kv_heads,
                                 head_dim=head_dim, device=next(model.parameters()).device, codec="int4")
    backend = UnprotectedBackend(manager, config, num_heads)
    rotary = None if model_type == "gpt2" else _find_rotary_embedding(model, layers)
    originals = {}
    try:
        for idx, layer in enumerate(layers):
            originals[idx] = getattr(layer, attr)
            setattr(layer, attr, UnprotectedPagedAttentionShim(originals[idx], idx, backend, rotary, model_type))
        model._unprotected_block_manager, model._unprotected_backend = manager, backend
        yield model
    finally:
        for idx, orig in originals.items():
            setattr(layers[idx], attr, orig)
        for name in ("_unprotected_block_manager", "_unprotected_backend"):
            if hasattr(model, name):
                delattr(model, name)


def reset_unprotected_cache(model):
    """unprotected_shim.py:550-553 (the manager only: the reference keeps the backend's flip statistics running)."""
    if hasattr(model, "_unprotected_block_manager"):
        model._unprotected_block_manager.reset()


def get_unprotected_stats(model):
    """unprotected_shim.py:556-574, same keys."""
    stats = {}
    if hasattr(model, "_unprotected_block_manager"):
        m = model._unprotected_block_manager
        stats["allocated_blocks"] = sum(len(b) for b in m.seq_to_blocks.values())
        stats["free_blocks"] = len(m.free_blocks)
        stats["sequences"] = len(m.seq_to_blocks)
    if hasattr(model, "_unprotected_backend"):
        be = model._unprotected_backend
        stats["injection_count"] = be._injection_count
        stats["bits_flipped"] = be._bits_flipped
        stats["total_bits"] = be._total_bits
        stats["actual_ber"] = be._bits_flipped / be._total_bits if be._total_bits > 0 else 0.0
    return stats

"""
The plugin boundary of the ECC-protected INT4 KV cache on the CUDA (sm_100a) backend.

Same public surface as the reference's ``kv_cache/ecc_shim.py`` -- ``ECCShimConfig``, ``SimpleBlockManager``,
``ECCBackend``, ``ECCPagedAttentionShim``, ``patch_model_with_ecc_attention``, ``reset_ecc_cache``,
``get_ecc_stats`` (+ ``ECCDummyCache``, ``compute_injection_seed``) -- with the same cache layout, seeds,
counter semantics and quirks, but:

  * ``ECCBackend.write``  is ONE kernel launch per layer (quantise + encode + inject + paged scatter, K and V)
    instead of ~4*S*Hkv Triton launches each behind a blocking ``.item()``  (reference :458-737);
  * ``ECCBackend.attend`` is ONE launch (fused paged decode attention, q_len == 1) or ONE launch + SDPA
    (prefill: fused gather + decode + correct + flag + interpolate + dequantise)       (reference :739-1089);
  * error statistics stay on the device (uint64 counters, one atomicAdd per CTA) and become Python ints only in
    ``get_ecc_stats`` -- no host sync on the hot path.

There is no Triton, no dispatch to other backends and no CPU fallback in this package.
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from .._lib import (CTR_CORRECTED, CTR_DETECTED, CTR_FLIPPED, CTR_INTERPOLATED, INJECT_MASK, INJECT_MODES, INJECT_NONE, check,
                    device_of, load, new_counters, ptr, stream_of)
from .attention_ecc import paged_attention_ecc, paged_prefill_attention_ecc, read_cache

# Longest context / largest problem served by the fused causal prefill kernel (fp32 CUDA-core dot products, ~13 warp-instructions
# per (query, token) pair); beyond them the attention is GEMM-shaped and goes through the fused read + the library's flash SDPA.
# Measured on a B200 (layer call, Hamming(8,4)+interpolation, fp32 queries): GPT-2 geometry (12 x 64) 256 tokens 0.048 ms fused vs
# 0.049 ms library, 1 024 tokens 0.22 vs 0.12 ms; LLaMA geometry (32 x 128, 8 kv-heads) 2 048 tokens 2.0 ms fused vs 2.8 ms library
# (SDPA has no flash kernel for fp32 and the read materialises fp32 K/V).  Hence: up to 2^22 (query, token) pairs always, up to 2^27
# for fp32 queries at head_dim >= 128.
_PREFILL_FUSED_MAX_CTX = int(os.environ.get("ECCQ_PREFILL_FUSED_MAX_CTX", "2048"))
_PREFILL_FUSED_MAX_PAIRS = int(os.environ.get("ECCQ_PREFILL_FUSED_MAX_PAIRS", str(1 << 22)))
_PREFILL_FUSED_MAX_PAIRS_F32 = int(os.environ.get("ECCQ_PREFILL_FUSED_MAX_PAIRS_F32", str(1 << 27)))


def compute_injection_seed(base_seed: int, layer_idx: int, injection_count: int) -> int:
    """ecc_shim.py:57-80 (exported by the reference, unused by its backend)."""
    return base_seed + layer_idx * 10000 + injection_count


class ECCDummyCache:
    """Placeholder satisfying the transformers cache protocol (ecc_shim.py:83-131): the shim owns the real cache."""

    def __init__(self, num_layers=0):
        self.key_cache, self.value_cache = [], []
        self._num_layers, self._seen_tokens = num_layers, 0

    def __len__(self):
        return self._num_layers

    def __iter__(self):
        return iter(zip(self.key_cache, self.value_cache))

    def __getitem__(self, layer_idx):
        if layer_idx < len(self.key_cache):
            return self.key_cache[layer_idx], self.value_cache[layer_idx]
        return None, None

    def to_legacy_cache(self):
        return ()

    def get_seq_length(self, layer_idx=0):
        return self._seen_tokens

    def get_max_length(self):
        return None

    def get_usable_length(self, new_seq_length, layer_idx=0):
        return self._seen_tokens

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        self._seen_tokens += key_states.shape[-2]
        return key_states, value_states

    @property
    def seen_tokens(self):
        return self._seen_tokens


class ECCShimConfig:
    """ecc_shim.py:134-186 plus the arguments the CUDA backend adds:

    backend:  "cuda" (the only backend this package serves; the reference never consults the argument, Q17)
    injector: "philox_ref"  -- the reference's Triton Philox stream, bit-identical flips (default), or
              "philox_fast" -- own stream, ~n_bits x fewer Philox calls, same distribution (throughput runs)
    append:   False keeps the reference's write semantics (positions [0,S) of sequence 0, later batch rows
              overwrite earlier ones, Q11); True writes batch row b to sequence b at its current length.
    """

    SUPPORTED_CODECS = {"fp16", "fp8", "int4", "hamming74", "hamming84", "golay"}
    CUDA_CODECS = ("fp16", "fp8", "int4", "hamming74", "hamming84", "golay")
    RAW_CODECS = ("fp16", "fp8")      # unprotected comparison modes: rows stored unencoded, no scales (ecc_shim.py:240-247)

    def __init__(self, codec="hamming84", ber=0.0, block_size=16, num_blocks=256, inject_errors=False, seed=42,
                 use_interpolation=False, backend="cuda", injector="philox_ref", append=False):
        if codec not in self.SUPPORTED_CODECS:
            raise ValueError(f"Unsupported codec: '{codec}'. Supported codecs: {sorted(self.SUPPORTED_CODECS)}")
        if backend != "cuda":
            raise ValueError(f"Unsupported backend: '{backend}'. This package implements backend='cuda' only.")
        if injector not in ("philox_ref", "philox_fast"):
            raise ValueError(f"Unsupported injector: '{injector}'")
        self.codec, self.ber, self.block_size, self.num_blocks = codec, ber, block_size, num_blocks
        self.inject_errors, self.seed, self.use_interpolation = inject_errors, seed, use_interpolation
        self.backend, self.injector, self.append = backend, injector, append


class SimpleBlockManager:
    """Paged storage + FIFO block allocator (ecc_shim.py:189-361), identical layout:
    k_cache/v_cache [num_blocks, L, Hkv, block_size*cw_per_token] (uint8, int32 for golay),
    k_scales/v_scales [num_blocks, L, Hkv, block_size] fp32, block_table [32, num_blocks] int32 (-1 = unmapped)."""

    def __init__(self, num_blocks, block_size, num_layers, num_kv_heads, head_dim, device="cuda", codec="hamming84"):
        if codec not in ECCShimConfig.CUDA_CODECS:
            raise ValueError(f"Unknown codec: {codec}")
        self.num_blocks, self.block_size, self.num_layers = num_blocks, block_size, num_layers
        self.num_kv_heads, self.head_dim, self.device, self.codec = num_kv_heads, head_dim, device, codec
        cw_per_token = (head_dim + 2) // 3 if codec == "golay" else head_dim
        self.values_per_head = self.codewords_per_head = block_size * cw_per_token
        self.cache_dtype = {"golay": torch.int32, "fp16": torch.float16, "fp8": torch.float8_e4m3fn}.get(codec, torch.uint8)
        self.needs_scales = codec not in ECCShimConfig.RAW_CODECS
        shape = (num_blocks, num_layers, num_kv_heads, self.codewords_per_head)
        self.k_cache = torch.zeros(shape, dtype=self.cache_dtype, device=device)
        self.v_cache = torch.zeros(shape, dtype=self.cache_dtype, device=device)
        # the reference allocates the scale arrays for every codec (ecc_shim.py:292-307); the raw codecs never touch them
        self.k_scales = torch.zeros(num_blocks, num_layers, num_kv_heads, block_size, dtype=torch.float32, device=device)
        self.v_scales = torch.zeros_like(self.k_scales)
        self._desc = None
        self.free_blocks = list(range(num_blocks))
        self.seq_to_blocks, self.seq_to_len = {}, {}
        self.seq_layer_len = {}      # (seq_id, layer) -> tokens that layer has appended (ECCShimConfig(append=True))
        self.max_seqs, self.max_blocks_per_seq = 32, num_blocks
        self.block_table = torch.full((self.max_seqs, self.max_blocks_per_seq), -1, dtype=torch.int32, device=device)

    def allocate(self, seq_id, num_tokens):
        have = self.seq_to_blocks.get(seq_id, [])
        need = max(0, -(-num_tokens // self.block_size) - len(have))
        if need > len(self.free_blocks):
            raise RuntimeError(f"Out of blocks: need {need}, have {len(self.free_blocks)}")
        fresh, self.free_blocks = self.free_blocks[:need], self.free_blocks[need:]
        blocks = have + fresh
        self.seq_to_blocks[seq_id] = blocks
        self.seq_to_len[seq_id] = num_tokens
        if fresh:  # one small H2D copy for the new tail instead of one per block
            self.block_table[seq_id, len(have):len(blocks)] = torch.tensor(fresh, dtype=torch.int32).to(self.block_table.device,
                                                                                                         non_blocking=True)
        return self.block_table[seq_id], num_tokens

    def get_block_table(self, seq_id):
        return self.block_table[seq_id]

    def get_context_len(self, seq_id):
        return self.seq_to_len.get(seq_id, 0)

    def reset(self):
        for blocks in self.seq_to_blocks.values():
            self.free_blocks.extend(blocks)
        self.seq_to_blocks.clear()
        self.seq_to_len.clear()
        self.seq_layer_len.clear()
        self.block_table.fill_(-1)
        for t in (self.k_cache, self.v_cache, self.k_scales, self.v_scales):
            t.zero_()

    def desc(self):
        if self._desc is None:      # the four buffers live as long as the manager
            self._desc = _lib.cache_desc(self.k_cache, self.v_cache, self.k_scales, self.v_scales, self.head_dim, self.block_size,
                                         self.codec)
        return self._desc


class ECCBackend:
    """ecc_shim.py:363-1164: write() / attend() between the model's attention and the protected cache."""

    def __init__(self, manager, config, num_heads):
        self.manager, self.config, self.num_heads = manager, config, num_heads
        self.num_kv_heads, self.head_dim = manager.num_kv_heads, manager.head_dim
        self.num_kv_groups = num_heads // self.num_kv_heads
        self._counters = new_counters(manager.k_cache.device)
        self._injection_count = 0
        self._total_values = 0
        self._flip_masks = None
        self._ctx_lens = None          # (host list, device int32 tensor) of the last attend()
        # the reference's Triton decode fast path (hamming84, no interpolation) updates no counters (Q12); keep its numbers
        self.count_decode_errors = not (config.codec == "hamming84" and not config.use_interpolation)

    # -- statistics: device counters, read lazily ---------------------------------------------------------
    @property
    def _errors_corrected(self):
        return int(self._counters[CTR_CORRECTED])

    @property
    def _errors_detected(self):
        return int(self._counters[CTR_DETECTED])

    @property
    def _bits_flipped(self):
        return int(self._counters[CTR_FLIPPED])

    @property
    def _values_interpolated(self):
        return int(self._counters[CTR_INTERPOLATED])

    def reset_stats(self):
        self._injection_count = 0
        self._total_values = 0
        self._counters.zero_()

    def set_flip_masks(self, k_mask, v_mask):
        """Use externally supplied flip masks ([B,S,Hkv,cw_per_token], cache dtype) for the next write() instead of
        a Philox stream -- the bit-exact parity route of BASELINE.md section 5."""
        self._flip_masks = (k_mask, v_mask)

    # -- write -----------------------------------------------------------------------------------------------
    def write(self, k, v, layer_idx, seq_id=0):
        B, S, _ = k.shape
        m, cfg = self.manager, self.config
        H, D = self.num_kv_heads, self.head_dim
        self._total_values += 2 * B * S * H * D
        if k.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            k, v = k.float(), v.float()
        if v.dtype != k.dtype:
            v = v.to(k.dtype)
        k = k.reshape(B, S, H, D).contiguous()
        v = v.reshape(B, S, H, D).contiguous()
        if cfg.codec in ECCShimConfig.RAW_CODECS:
            return self._write_raw(k, v, layer_idx, seq_id)
        inject, km, vm = INJECT_NONE, None, None
        if self._flip_masks is not None:
            inject, (km, vm) = INJECT_MASK, self._flip_masks
            km, vm = km.contiguous(), vm.contiguous()
            self._flip_masks = None
        elif cfg.inject_errors and cfg.ber > 0:
            inject = INJECT_MODES[cfg.injector]
        inj_base = self._injection_count
        if cfg.append:
            # every layer appends at ITS OWN length: the layers of one forward pass write the same positions one after another
            starts = [m.seq_layer_len.get((seq_id + b, layer_idx), 0) for b in range(B)]
            for b in range(B):
                m.allocate(seq_id + b, max(m.get_context_len(seq_id + b), starts[b] + S))
                m.seq_layer_len[(seq_id + b, layer_idx)] = starts[b] + S
            start = torch.tensor(starts, dtype=torch.int32)
            bt, start_dev, nseq = m.block_table[seq_id:seq_id + B], start.to(k.device, non_blocking=True), B
        else:
            # reference semantics (Q11): every batch row lands on positions [0,S) of `seq_id`, so only the last row
            # survives; it carries the injection seeds of rows (B-1)*S*H ...
            if m.get_context_len(seq_id) < S:
                m.allocate(seq_id, S)
            bt, start_dev, nseq = m.block_table[seq_id:seq_id + 1], None, 1
            inj_base += (B - 1) * S * H
            k, v = k[B - 1:], v[B - 1:]
            if km is not None:
                km, vm = km[B - 1:].contiguous(), vm[B - 1:].contiguous()
        with device_of(k):
            check(load().eccq_cache_write(m.desc(), ptr(k), ptr(v), _lib.dtype_code(k.dtype), ptr(bt), bt.shape[1], ptr(start_dev),
                                          nseq, S, int(layer_idx), inject, ptr(km), ptr(vm), float(cfg.ber), int(cfg.seed),
                                          int(inj_base), ptr(self._counters), stream_of(k)), "ECCBackend.write")
        if inject not in (INJECT_NONE, INJECT_MASK):
            self._injection_count += B * S * H  # one per (batch, pos, head), ecc_shim.py:643-651

    def _rows_and_table(self, B, S, seq_id, layer_idx):
        """Block table rows / start positions of a write: the reference's semantics (Q11: every batch row lands on positions
        [0, S) of `seq_id`, the last one survives) or, with append=True, row b appended to sequence seq_id + b."""
        m, cfg = self.manager, self.config
        if cfg.append:
            starts = [m.seq_layer_len.get((seq_id + b, layer_idx), 0) for b in range(B)]
            for b in range(B):
                m.allocate(seq_id + b, max(m.get_context_len(seq_id + b), starts[b] + S))
                m.seq_layer_len[(seq_id + b, layer_idx)] = starts[b] + S
            start = torch.tensor(starts, dtype=torch.int32).to(m.block_table.device, non_blocking=True)
            return m.block_table[seq_id:seq_id + B], start, B, 0
        if m.get_context_len(seq_id) < S:
            m.allocate(seq_id, S)
        return m.block_table[seq_id:seq_id + 1], None, 1, B - 1

    def _write_raw(self, k, v, layer_idx, seq_id):
        """Codecs "fp16" / "fp8" (ecc_shim.py:511-570): rows stored unencoded; fp8 rows take the reference's per-(token, head)
        fault injection on their raw bytes (seed_K = seed + count, seed_V = +1, count += 1)."""
        B, S, H, D = k.shape
        m, cfg = self.manager, self.config
        bt, start_dev, nseq, first = self._rows_and_table(B, S, seq_id, layer_idx)
        inj_base = self._injection_count + first * S * H
        k, v = k[first:], v[first:]
        inject, n_bits = INJECT_NONE, 8
        if cfg.codec == "fp16":
            kr, vr = k.to(torch.float16).contiguous(), v.to(torch.float16).contiguous()
        else:
            kr, vr = k.to(torch.float8_e4m3fn).contiguous(), v.to(torch.float8_e4m3fn).contiguous()
            if cfg.inject_errors and cfg.ber > 0:
                inject = _lib.INJECT_PHILOX_REF
        row_bytes = D * kr.element_size()
        with device_of(kr):
            check(load().eccq_paged_rows_write(ptr(m.k_cache), ptr(m.v_cache), ptr(kr), ptr(vr), row_bytes, m.num_layers, H, m.block_size,
                                               ptr(bt), bt.shape[1], ptr(start_dev), nseq, S, int(layer_idx), inject, n_bits, float(cfg.ber),
                                               int(cfg.seed), int(inj_base), ptr(self._counters), stream_of(kr)), "ECCBackend.write")
        if inject != INJECT_NONE:
            self._injection_count += B * S * H

    def _read_raw(self, bt, cl, nseq, T, layer_idx):
        m = self.manager
        k = torch.zeros(nseq, T, m.num_kv_heads, m.head_dim, dtype=m.cache_dtype, device=m.k_cache.device)
        v = torch.zeros_like(k)
        with device_of(k):
            check(load().eccq_paged_rows_read(ptr(m.k_cache), ptr(m.v_cache), ptr(k), ptr(v), m.head_dim * k.element_size(), m.num_layers,
                                              m.num_kv_heads, m.block_size, ptr(bt), bt.shape[1], ptr(cl), nseq, T, int(layer_idx),
                                              stream_of(k)), "ECCBackend.attend")
        return k, v

    # -- attend ----------------------------------------------------------------------------------------------
    def attend(self, q, layer_idx, seq_id=0):
        B, Hq, Sq, D = q.shape
        m, cfg = self.manager, self.config
        if cfg.append:
            lens = [m.seq_layer_len.get((seq_id + b, layer_idx), 0) for b in range(B)]
        else:
            lens = [m.get_context_len(seq_id)]
        if max(lens) == 0:
            return torch.zeros_like(q)
        nseq = len(lens)
        if self._ctx_lens is None or self._ctx_lens[0] != lens:      # every layer of one forward pass asks for the same lengths
            self._ctx_lens = (lens, torch.tensor(lens, dtype=torch.int32).to(q.device, non_blocking=True))
        cl = self._ctx_lens[1]
        bt = m.block_table[seq_id:seq_id + nseq]
        if cfg.codec in ECCShimConfig.RAW_CODECS:
            # ecc_shim.py:802-872: gather the rows, (fp8 -> fp16,) SDPA; no decode step, also for q_len == 1
            k, v = self._read_raw(bt, cl, nseq, max(lens), layer_idx)
            if cfg.codec == "fp8":
                k, v = k.to(torch.float16), v.to(torch.float16)
            return self._run_attention(q, k, v, self._append_mask(q, cl, lens, k.shape[1]))
        if Sq == 1:
            # fused paged decode attention for every codec (the reference: hamming84 without interpolation only)
            qd = q[:, :, 0, :]
            shared = nseq == 1 and B > 1
            if shared:
                # reference layout: ONE shared context, B query rows.  Fold the batch into the GQA group
                # ([B,Hkv,G,D] -> [1, Hkv*(B*G), D]) so the context is decoded (and counted) once, as the reference does.
                G = self.num_kv_groups
                qd = qd.reshape(B, self.num_kv_heads, G, D).permute(1, 0, 2, 3).reshape(1, self.num_kv_heads * B * G, D)
            out = paged_attention_ecc(qd, m.k_cache, m.v_cache, bt, cl, m.k_scales, layer_idx, m.block_size, codec=cfg.codec,
                                      v_scales=m.v_scales, use_interpolation=cfg.use_interpolation,
                                      counters=self._counters if self.count_decode_errors else None,
                                      ber_hint=cfg.ber if cfg.inject_errors else 0.0)
            if shared:
                out = out.reshape(self.num_kv_heads, B, G, D).permute(1, 0, 2, 3).reshape(B, Hq, D)
            return out.unsqueeze(2)
        pairs = B * Hq * Sq * max(lens)
        if max(lens) <= _PREFILL_FUSED_MAX_CTX and D <= 256 and (
                pairs <= _PREFILL_FUSED_MAX_PAIRS or (q.dtype == torch.float32 and D >= 128 and pairs <= _PREFILL_FUSED_MAX_PAIRS_F32)):
            # ECC-fused causal prefill: ONE hand-written kernel consumes the codeword pages (decode + interpolation in the K/V
            # prologue, fp32 online softmax); no K/V tensor is materialised and no library attention kernel runs
            return paged_prefill_attention_ecc(q, m.k_cache, m.v_cache, bt, cl, m.k_scales, m.v_scales, layer_idx, m.block_size,
                                               codec=cfg.codec, use_interpolation=cfg.use_interpolation, counters=self._counters)
        # long contexts: the contraction is GEMM-shaped work for the library's flash kernels (fused read -> K/V in q's dtype -> SDPA)
        dt = q.dtype if q.dtype in (torch.float32, torch.float16, torch.bfloat16) else torch.float32
        k, v = read_cache(m.k_cache, m.v_cache, m.k_scales, m.v_scales, bt, cl, layer_idx, m.block_size, D, cfg.codec,
                          use_interpolation=cfg.use_interpolation, out_dtype=dt, counters=self._counters, max_context=max(lens))
        return self._run_attention(q, k, v, self._append_mask(q, cl, lens, k.shape[1]))

    def _append_mask(self, q, cl, lens, ctx):
        """Chunked prefill / ragged batch in append mode: query t of row b sits at position lens[b]-Sq+t and sees keys <= that
        position (bottom-right aligned causal mask; it also hides the zero padding past a shorter sequence).  SDPA's
        is_causal is top-left aligned and is only right for a one-shot prefill of equal lengths (the reference's only case)."""
        Sq = q.shape[2]
        if not self.config.append or (min(lens) == max(lens) and lens[0] == Sq):
            return None
        key_pos = torch.arange(ctx, device=q.device)
        q_pos = cl.to(torch.int64)[:, None] - Sq + torch.arange(Sq, device=q.device)[None, :]
        return (key_pos[None, None, :] <= q_pos[:, :, None])[:, None]                  # [B, 1, Sq, ctx]

    def _run_attention(self, q, k, v, attn_mask=None):
        """ecc_shim.py:1138-1164: GQA expand + SDPA, causal (top-left) when q_len > 1. k, v: [nseq, ctx, Hkv, D]."""
        k = k.permute(0, 2, 1, 3)
        v = v.permute(0, 2, 1, 3)
        if k.shape[0] != q.shape[0]:
            k, v = k.expand(q.shape[0], -1, -1, -1), v.expand(q.shape[0], -1, -1, -1)
        # the GQA expansion happens inside SDPA (no Hq-sized copy of K/V, which the reference makes with repeat_interleave)
        if attn_mask is not None:
            return F.scaled_dot_product_attention(q, k.to(q.dtype), v.to(q.dtype), attn_mask=attn_mask,
                                                  enable_gqa=self.num_kv_groups > 1)
        return F.scaled_dot_product_attention(q, k.to(q.dtype), v.to(q.dtype), is_causal=q.shape[2] > 1,
                                              enable_gqa=self.num_kv_groups > 1)


def _get_attention_params(attn):
    """(num_heads, num_kv_heads, head_dim) of an HF attention module (ecc_shim.py:1556-1611)."""
    cfg = getattr(attn, "config", None)

    def first(obj, names):
        for n in names:
            if obj is not None and hasattr(obj, n):
                return getattr(obj, n)
        return None

    num_heads = first(attn, ("num_heads", "num_attention_heads")) or first(cfg, ("num_attention_heads", "num_heads", "n_head"))
    head_dim = first(attn, ("head_dim", "head_size")) or first(cfg, ("head_dim",))
    if head_dim is None and cfg is not None and hasattr(cfg, "hidden_size") and num_heads:
        head_dim = cfg.hidden_size // num_heads
    num_kv = first(attn, ("num_key_value_heads", "num_kv_heads")) or first(cfg, ("num_key_value_heads", "num_kv_heads"))
    if (num_heads is None or head_dim is None) and hasattr(getattr(attn, "q_proj", None), "weight"):
        out_features = attn.q_proj.weight.shape[0]
        for cand in (128, 64, 32, 96, 256):
            if out_features % cand == 0:
                head_dim = head_dim or cand
                num_heads = num_heads or out_features // cand
                break
    if num_heads is None:
        raise ValueError("Could not determine num_heads from attention module")
    if head_dim is None:
        raise ValueError("Could not determine head_dim from attention module")
    return num_heads, num_kv or num_heads, head_dim


class _PlainRotary(nn.Module):
    def __init__(self, dim, base=10000.0):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)), persistent=False)

    def forward(self, x, position_ids):
        f = torch.outer(position_ids[0].float(), self.inv_freq.to(position_ids.device))
        e = torch.cat((f, f), dim=-1)
        return e.cos()[None, None], e.sin()[None, None]


def _find_rotary_embedding(model, layers):
    """Locate (or build) the rotary embedding of a LLaMA-style model (ecc_shim.py:1484-1553)."""
    for holder in (layers[0].self_attn, layers[0], getattr(model, "model", None), model):
        if holder is not None and hasattr(holder, "rotary_emb"):
            return holder.rotary_emb
    cfg = getattr(model, "config", None) or getattr(getattr(model, "model", None), "config", None)
    device = next(model.parameters()).device
    if cfg is not None:
        try:
            from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding
            return LlamaRotaryEmbedding(config=cfg, device=device)
        except Exception:
            pass
    return _PlainRotary(_get_attention_params(layers[0].self_attn)[2]).to(device)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class ECCPagedAttentionShim(nn.Module):
    """Drop-in attention module (ecc_shim.py:1167-1392): keeps the original projections, routes K/V through the
    ECC cache.  GPT-2 (c_attn/c_proj) and LLaMA/Mistral (q/k/v/o_proj + RoPE) layouts."""

    def __init__(self, original_attn, layer_idx, backend, rotary_emb, model_type="llama"):
        super().__init__()
        self.model_type = model_type
        names = ("c_attn", "c_proj") if model_type == "gpt2" else ("q_proj", "k_proj", "v_proj", "o_proj")
        for name in names:
            proj = getattr(original_attn, name, None)
            if proj is None:
                available = [a for a in dir(original_attn) if not a.startswith("_")]
                raise ValueError(f"Attention module {type(original_attn).__name__} missing '{name}'. Available attributes: {available}")
            setattr(self, name, proj)
        for name in ("attn_dropout", "resid_dropout"):
            if hasattr(original_attn, name):
                setattr(self, name, getattr(original_attn, name))
        self.num_heads, self.num_kv_heads, self.head_dim = _get_attention_params(original_attn)
        self.hidden_size = self.num_heads * self.head_dim
        if hasattr(original_attn, "split_size"):
            self.split_size = original_attn.split_size
        self.backend, self.layer_idx, self.rotary_emb = backend, layer_idx, rotary_emb
        self.scale = 1.0 / math.sqrt(self.head_dim)

    def _split(self, x, heads):
        b, s, _ = x.shape
        return x.view(b, s, heads, self.head_dim).transpose(1, 2)

    def _cache_round_trip(self, q, k, v):
        b, _, s, _ = k.shape
        self.backend.write(k.transpose(1, 2).reshape(b, s, -1), v.transpose(1, 2).reshape(b, s, -1), self.layer_idx, seq_id=0)
        o = self.backend.attend(q, self.layer_idx, seq_id=0)
        return o.transpose(1, 2).reshape(b, s, self.hidden_size)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None, output_attentions=False,
                use_cache=False, cache_position=None, layer_past=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, position_embeddings=None, **kwargs):
        if self.model_type == "gpt2":
            q, k, v = self.c_attn(hidden_states).split(self.split_size, dim=2)
            q, k, v = self._split(q, self.num_heads), self._split(k, self.num_kv_heads), self._split(v, self.num_kv_heads)
            out = self.c_proj(self._cache_round_trip(q, k, v))
            if hasattr(self, "resid_dropout"):
                out = self.resid_dropout(out)
            res = (out, (k, v) if use_cache else None)
            return res + (None,) if output_attentions else res
        b, s, _ = hidden_states.shape
        q = self._split(self.q_proj(hidden_states), self.num_heads)
        k = self._split(self.k_proj(hidden_states), self.num_kv_heads)
        v = self._split(self.v_proj(hidden_states), self.num_kv_heads)
        if position_embeddings is not None:      # recent transformers hand the layer (cos, sin) of the right positions
            cos, sin = position_embeddings
        else:
            if position_ids is None:
                position_ids = torch.arange(s, device=hidden_states.device).unsqueeze(0).expand(b, -1)
            cos, sin = self.rotary_emb(v, position_ids)
        while cos.dim() < 4:
            cos, sin = cos.unsqueeze(0 if cos.dim() < 2 else 1), sin.unsqueeze(0 if sin.dim() < 2 else 1)
        if cos.shape[1] not in (1, q.shape[1]):
            cos, sin = cos[:, :1], sin[:, :1]
        q, k = q * cos + _rotate_half(q) * sin, k * cos + _rotate_half(k) * sin
        return self.o_proj(self._cache_round_trip(q, k, v)), None


def _model_layout(model):
    if hasattr(model, "model") and hasattr(model.model, "layers"):
        return model.model.layers, "llama", "self_attn"
    if hasattr(model, "layers"):
        return model.layers, "llama", "self_attn"
    if hasattr(model, "transformer") and hasattr(model.transformer, "h"):
        return model.transformer.h, "gpt2", "attn"
    raise ValueError("Unsupported model architecture")


@contextmanager
def patch_model_with_ecc_attention(model, config, num_blocks=256):
    """ecc_shim.py:1395-1481: swap every attention module for the ECC shim for the duration of the `with` block."""
    layers, model_type, attr = _model_layout(model)
    num_heads, num_kv_heads, head_dim = _get_attention_params(getattr(layers[0], attr))
    manager = SimpleBlockManager(num_blocks=num_blocks, block_size=config.block_size, num_layers=len(layers),
                                 num_kv_heads=num_kv_heads, head_dim=head_dim, device=next(model.parameters()).device,
                                 codec=config.codec)
    backend = ECCBackend(manager, config, num_heads)
    rotary = None if model_type == "gpt2" else _find_rotary_embedding(model, layers)
    originals = {}
    try:
        for idx, layer in enumerate(layers):
            originals[idx] = getattr(layer, attr)
            setattr(layer, attr, ECCPagedAttentionShim(originals[idx], idx, backend, rotary, model_type))
        model._ecc_block_manager, model._ecc_backend = manager, backend
        model._ecc_model_type, model._ecc_attn_attr = model_type, attr
        yield model
    finally:
        for idx, orig in originals.items():
            setattr(layers[idx], attr, orig)
        for name in ("_ecc_block_manager", "_ecc_backend", "_ecc_model_type", "_ecc_attn_attr"):
            if hasattr(model, name):
                delattr(model, name)


def reset_ecc_cache(model):
    """ecc_shim.py:1614-1624."""
    if hasattr(model, "_ecc_block_manager"):
        model._ecc_block_manager.reset()
    if hasattr(model, "_ecc_backend"):
        model._ecc_backend.reset_stats()


def get_ecc_stats(model, all_reduce=False):
    """ecc_shim.py:1627-1642, same keys.  This is the one place the device counters are read (one 64-byte D2H copy).
    all_reduce=True sums the counter vector over the torch.distributed default group first (NCCL over NVLink;
    the only collective of the whole path)."""
    stats = {}
    if hasattr(model, "_ecc_block_manager"):
        m = model._ecc_block_manager
        stats["allocated_blocks"] = sum(len(b) for b in m.seq_to_blocks.values())
        stats["free_blocks"] = len(m.free_blocks)
        stats["sequences"] = len(m.seq_to_blocks)
    if hasattr(model, "_ecc_backend"):
        be = model._ecc_backend
        vec = torch.cat([be._counters, torch.tensor([be._injection_count, be._total_values], dtype=torch.int64,
                                                    device=be._counters.device)])
        if all_reduce:
            from ..distributed import all_reduce_counters
            vec = all_reduce_counters(vec)
        c = vec.tolist()
        stats["injection_count"] = c[-2]
        stats["errors_corrected"] = c[CTR_CORRECTED]
        stats["errors_detected"] = c[CTR_DETECTED]
        stats["total_values"] = c[-1]
    return stats

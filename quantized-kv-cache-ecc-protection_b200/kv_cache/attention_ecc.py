"""
ECC-fused paged decode attention on the CUDA backend: ``paged_attention_ecc`` and
``reference_attention_ecc`` with the reference's signatures (kv_cache/attention_ecc.py:620-635, :783-795).

Differences from the reference, all deliberate and tested:
  * GQA caches are indexed with kv_head = q_head // group (the Triton kernel derives strides from the query
    head count and is only correct for MHA);
  * an empty context yields zeros (the Triton kernel returns -8.0, Q14);
  * hamming74 / golay / hamming84+interpolation also run fused (the reference falls back to Python loops);
  * the launch is split over the context (FlashDecoding) and never recompiles for a new max_blocks.
"""
from __future__ import annotations

import math
import os

import torch

from .. import _lib
from .._lib import check, device_of, load, ptr, require_cuda, stream_of

_workspaces: dict = {}


_page_scratch: dict = {}


def _corrected_pages(device, num_blocks, hkv, block_size, head_dim):
    """Scratch one-layer INT4 cache (pages + scales for K and V) for decode attention with interpolation."""
    # keyed by stream as well: two streams of one device must not share (and race on) one set of corrected pages
    key = (device, torch.cuda.current_stream(device).cuda_stream, num_blocks, hkv, block_size, head_dim)
    t = _page_scratch.get(key)
    if t is None:
        for k in [k for k in _page_scratch if k[:2] == key[:2]]:   # one geometry at a time per (device, stream): the scratch
            del _page_scratch[k]                                    # is as large as one layer of the cache
        # zero-filled like the cache itself: slots past a ragged context are staged with their tile (and masked), so their
        # scales must be finite
        t = (torch.zeros(num_blocks, 1, hkv, block_size * head_dim, dtype=torch.uint8, device=device),
             torch.zeros(num_blocks, 1, hkv, block_size * head_dim, dtype=torch.uint8, device=device),
             torch.zeros(num_blocks, 1, hkv, block_size, dtype=torch.float32, device=device),
             torch.zeros(num_blocks, 1, hkv, block_size, dtype=torch.float32, device=device))
        _page_scratch[key] = t
    return t


def _workspace(device, nbytes: int) -> torch.Tensor:
    # one workspace per (device, stream): concurrent callers on different streams never share split-KV partials, and a
    # buffer that is outgrown is only ever replaced by its own stream (stream-ordered free in torch's caching allocator)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _head_dim(k_cache, block_size, codec, query):
    if codec == "golay":
        return query.shape[-1]
    return k_cache.shape[-1] // block_size


def paged_attention_ecc(query, k_cache, v_cache, block_table, context_lens, k_scales, layer_idx, block_size, sm_scale=None,
                        codec="hamming84", syndrome_table=None, use_tiled=False, block_m=4, v_scales=None, *,
                        use_interpolation=False, counters=None, ber_hint=None):
    """query [B, Hq, D] -> [B, Hq, D] in query.dtype.  `syndrome_table`, `use_tiled`, `block_m` are accepted for
    signature compatibility and ignored (the tables live in the library; there is one kernel).  `counters`
    (device int64[8]) receives corrected/detected counts when given.  `ber_hint`: the caller's bit error rate if it knows it
    (the shim does); interpolation at BER >= 5e-3 -- a double error in every third row -- takes the two-pass route, which is
    faster there (C5 shard at 1e-2: 0.39 ms against 0.44 ms; at 1e-3: 0.36 ms against 0.28 ms)."""
    assert query.is_cuda, "Query must be on CUDA"
    if codec not in ("hamming84", "hamming74", "golay", "int4"):
        raise ValueError(f"Unknown codec: {codec}")
    want = torch.int32 if codec == "golay" else torch.uint8
    assert k_cache.dtype == want, f"{codec} K cache must be {want}"
    assert v_cache.dtype == want, f"{codec} V cache must be {want}"
    v_use_k = v_scales is None
    if v_scales is None:
        v_scales = k_scales
    B, Hq, D = query.shape
    qd = query if query.dtype in (torch.float32, torch.float16, torch.bfloat16) else query.float()
    qd = qd.contiguous()
    head_dim = _head_dim(k_cache, block_size, codec, query)
    if head_dim != D:
        raise ValueError(f"query head_dim {D} does not match the cache ({head_dim})")
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    bt = block_table.to(torch.int32).contiguous()
    cl = context_lens.to(torch.int32).contiguous()
    max_blocks = bt.shape[1]
    desc = _lib.cache_desc(k_cache, v_cache, k_scales, v_scales, D, block_size, codec)
    out = torch.empty_like(qd)
    interp = bool(use_interpolation) and codec == "hamming84"     # the reference interpolates hamming84 only
    with device_of(qd):
        lib = load()
        nbytes = lib.eccq_attention_workspace_bytes(B, Hq, D, max_blocks * block_size)
        ws = _workspace(qd.device, nbytes)
        # Hamming(8,4) + interpolation at head_dim 128: ONE pass -- the tcgen05 kernel rewrites double-detected values from
        # their temporal neighbours inside the tile pipeline (csrc/eccq_attention_tc.cu, INTERP variant).
        single_pass = interp and D == 128 and block_size in (16, 32, 64, 128) and os.environ.get("ECCQ_ATTN_TC", "1") != "0" \
            and os.environ.get("ECCQ_ATTN_INTERP_TWO_PASS", "0") == "0" and not (ber_hint is not None and ber_hint >= 5e-3)
        # Golay at head_dim 128: also one pass (the decode warps of the tcgen05 kernel read the codewords from global memory)
        golay_single = codec == "golay" and D == 128 and block_size in (16, 32, 64, 128) and os.environ.get("ECCQ_ATTN_TC", "1") != "0" \
            and os.environ.get("ECCQ_ATTN_TC_GOLAY", "1") != "0"
        if (interp and not single_pass) or (codec == "golay" and not golay_single):
            # temporal neighbours span pages (and Golay codewords span values): resolve them once with the streaming read
            # kernels into corrected INT4 pages (same paged layout, one layer), then attend over those with the INT4 tile
            # path.  Falls through to the general kernel for geometries the streaming kernels do not cover.
            nb, _, hkv, _ = k_cache.shape
            kp, vp, ksp, vsp = _corrected_pages(qd.device, nb, hkv, block_size, D)
            rc = lib.eccq_cache_correct_pages(desc, ptr(bt), max_blocks, ptr(cl), B, int(layer_idx), int(interp), ptr(kp), ptr(vp), ptr(ksp),
                                              ptr(vsp), ptr(counters), stream_of(qd))
            if rc == 0:
                pdesc = _lib.cache_desc(kp, vp, ksp, ksp if v_use_k else vsp, D, block_size, "int4")
                check(lib.eccq_paged_decode_attention(pdesc, ptr(qd), ptr(out), _lib.dtype_code(qd.dtype), ptr(bt), max_blocks, ptr(cl),
                                                      B, Hq, 0, float(sm_scale), 0, int(v_use_k), 0, ptr(ws), ws.numel(), None,
                                                      stream_of(qd)), "paged_attention_ecc")
                return out if out.dtype == query.dtype else out.to(query.dtype)
            if rc != _lib.ECCQ_EUNSUPPORTED:
                check(rc, "cache_correct_pages")
        use_interpolation = interp
        check(lib.eccq_paged_decode_attention(desc, ptr(qd), ptr(out), _lib.dtype_code(qd.dtype), ptr(bt), max_blocks, ptr(cl), B, Hq,
                                              int(layer_idx), float(sm_scale), int(bool(use_interpolation)), int(v_use_k),
                                              int(counters is not None), ptr(ws), ws.numel(), ptr(counters), stream_of(qd)),
              "paged_attention_ecc")
    return out if out.dtype == query.dtype else out.to(query.dtype)


def paged_prefill_attention_ecc(query, k_cache, v_cache, block_table, context_lens, k_scales, v_scales, layer_idx, block_size,
                                sm_scale=None, codec="hamming84", *, use_interpolation=False, counters=None):
    """ECC-fused causal prefill attention: query [B, Hq, Sq, D] -> [B, Hq, Sq, D] in query.dtype, straight from the codeword
    pages (the decode + interpolation + dequantisation of ECCBackend.attend, kv_cache/ecc_shim.py:979-1089, and the causal
    SDPA of `_run_attention`, :1138-1164, in one kernel; no K/V tensor is materialised).  block_table [nseq, max_blocks],
    context_lens [nseq] with nseq == B, or nseq == 1 for one context shared by all B query rows (the reference's batch
    semantics).  Query t sits at position context_len - Sq + t."""
    assert query.is_cuda, "Query must be on CUDA"
    if codec not in ("hamming84", "hamming74", "golay", "int4"):
        raise ValueError(f"Unknown codec: {codec}")
    B, Hq, Sq, D = query.shape
    qd = query if query.dtype in (torch.float32, torch.float16, torch.bfloat16) else query.float()
    qd = qd.contiguous()
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    bt = block_table.to(torch.int32).contiguous()
    if bt.dim() == 1:
        bt = bt.unsqueeze(0)
    cl = context_lens.to(torch.int32).contiguous()
    nseq = bt.shape[0]
    if nseq not in (1, B):
        raise ValueError(f"block_table rows ({nseq}) must be 1 or the query batch ({B})")
    desc = _lib.cache_desc(k_cache, v_cache, k_scales, v_scales, D, block_size, codec)
    out = torch.empty_like(qd)
    with device_of(qd):
        check(load().eccq_paged_prefill_attention(desc, ptr(qd), ptr(out), _lib.dtype_code(qd.dtype), ptr(bt), bt.shape[1], ptr(cl), nseq, B,
                                                  Hq, Sq, int(layer_idx), float(sm_scale), int(bool(use_interpolation) and codec == "hamming84"),
                                                  int(counters is not None), ptr(counters), stream_of(qd)), "paged_prefill_attention_ecc")
    return out if out.dtype == query.dtype else out.to(query.dtype)


def read_cache(k_cache, v_cache, k_scales, v_scales, block_table, context_lens, layer_idx, block_size, head_dim, codec,
               use_interpolation=False, out_dtype=torch.float32, counters=None, return_nibbles=False, max_context=None, out=None):
    """Fused gather + decode + [interpolate] + dequantise of whole contexts: returns K, V
    [nseq, max_ctx, Hkv, D] (tokens past context_len are zero).  The read half of ECCBackend.attend.

    max_context: longest context if the caller knows it (saves the one host sync used to size the outputs).
    out:         optional preallocated (k_out, v_out); tokens past a sequence's context are then left untouched."""
    require_cuda(k_cache)
    bt = block_table.to(torch.int32).contiguous()
    if bt.dim() == 1:
        bt = bt.unsqueeze(0)
    cl = context_lens.to(torch.int32).contiguous()
    nseq, max_blocks = bt.shape
    if max_context is None:
        lens = cl.tolist()
        T, ragged = (max(lens), min(lens) != max(lens)) if lens else (0, False)
    else:
        T, ragged = int(max_context), nseq > 1      # one sequence: its context IS the longest
    H = k_cache.shape[2]
    dev = k_cache.device
    if out is not None:
        k_out, v_out = out
        T = k_out.shape[1]
    else:
        alloc = torch.zeros if ragged else torch.empty
        k_out = alloc(nseq, T, H, head_dim, dtype=out_dtype, device=dev)
        v_out = alloc(nseq, T, H, head_dim, dtype=out_dtype, device=dev)
    nk = nv = None
    if return_nibbles:
        nk = torch.zeros(nseq, T, H, head_dim, dtype=torch.uint8, device=dev)
        nv = torch.zeros_like(nk)
    if T > 0:
        desc = _lib.cache_desc(k_cache, v_cache, k_scales, v_scales, head_dim, block_size, codec)
        with device_of(k_cache):
            check(load().eccq_cache_read(desc, ptr(bt), max_blocks, ptr(cl), nseq, T, int(layer_idx), int(bool(use_interpolation)),
                                         _lib.dtype_code(k_out.dtype), ptr(k_out), ptr(v_out), ptr(nk), ptr(nv), ptr(counters),
                                         stream_of(k_cache)), "cache_read")
    if return_nibbles:
        return k_out, v_out, nk, nv
    return k_out, v_out


def reference_attention_ecc(query, k_cache, v_cache, block_table, context_lens, scales, layer_idx, block_size, head_dim,
                            sm_scale=None, codec="hamming84"):
    """attention_ecc.py:783-909 (one `scales` tensor for K and V, Q15; fp32 output): decode the whole context with
    the fused read kernel, then a plain fp32 softmax(QK^T)V in torch.  A second, independent route to the
    same numbers as `paged_attention_ecc`, used by the tests."""
    assert query.is_cuda, "Query must be on CUDA"
    B, Hq, _ = query.shape
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(head_dim)
    k, v = read_cache(k_cache, v_cache, scales, scales, block_table, context_lens, layer_idx, block_size, head_dim, codec)
    Hkv = k.shape[2]
    out = torch.zeros(B, Hq, head_dim, device=query.device)
    lens = context_lens.tolist()
    for b in range(B):
        c = int(lens[b])
        if c <= 0:
            continue
        kb = k[b, :c].repeat_interleave(Hq // Hkv, dim=1).permute(1, 0, 2)
        vb = v[b, :c].repeat_interleave(Hq // Hkv, dim=1).permute(1, 0, 2)
        w = torch.softmax(torch.einsum("hd,hcd->hc", query[b].float(), kb) * sm_scale, dim=-1)
        out[b] = torch.einsum("hc,hcd->hd", w, vb)
    return out

"""ECC-fused causal prefill attention (eccq_paged_prefill_attention: q_len > 1, ECC decode + interpolation in the K/V prologue, no
K/V materialisation) against the CPU oracle: the oracle's cache_read (decode / interpolate / dequantise, ecc_shim.py:979-1071)
followed by a numpy causal softmax attention with the bottom-right aligned mask (== SDPA's is_causal when ctx == Sq, the
reference's only case, ecc_shim.py:1138-1164)."""
import numpy as np
import pytest
import torch

from tests.test_cuda_attention_tc import build_cache, dev


def ref_prefill(oracle, q, oc, bt, lens, layer, interp):
    """q [B, Hq, Sq, D]; one context per batch row (len(lens) == B) or one shared context (len(lens) == 1)."""
    B, Hq, Sq, D = q.shape
    out = np.zeros_like(q, dtype=np.float32)
    corr = det = 0
    cache = {}
    for b in range(B):
        i = 0 if len(lens) == 1 else b
        ctx = int(lens[i])
        if ctx == 0:
            continue
        if i not in cache:
            cache[i] = oracle.cache_read(oc, layer, bt[i], ctx, use_interpolation=interp)
            corr += cache[i][2]
            det += cache[i][3]
        k, v = cache[i][0], cache[i][1]                     # [ctx, Hkv, D] fp32
        g = Hq // k.shape[1]
        kk = np.repeat(k, g, axis=1).transpose(1, 0, 2).astype(np.float64)
        vv = np.repeat(v, g, axis=1).transpose(1, 0, 2).astype(np.float64)
        s = np.einsum("hqd,hcd->hqc", q[b].astype(np.float64), kk) / np.sqrt(D)
        qpos = ctx - Sq + np.arange(Sq)
        vis = np.arange(ctx)[None, :] <= qpos[:, None]
        s = np.where(vis[None], s, -np.inf)
        with np.errstate(invalid="ignore"):
            p = np.exp(s - s.max(axis=-1, keepdims=True))
            p = np.where(vis[None], p, 0.0)
            den = p.sum(axis=-1, keepdims=True)
            o = np.einsum("hqc,hcd->hqd", p, vv) / np.where(den > 0, den, 1.0)
        out[b] = np.where(den > 0, o, 0.0)
    return out, corr, det


@pytest.mark.gpu
@pytest.mark.parametrize("codec,interp", [("hamming84", False), ("hamming84", True), ("hamming74", False), ("golay", False), ("int4", False)])
@pytest.mark.parametrize("D,bs,Hq,Hkv", [(64, 16, 4, 4), (128, 16, 8, 2), (128, 32, 4, 1), (32, 16, 6, 2), (96, 16, 4, 2)])   # 96: warp-per-token kernel
def test_one_shot_prefill_matches_oracle(oracle, codec, interp, D, bs, Hq, Hkv):
    """ctx == Sq per batch row (the reference's case: top-left causal mask), ragged batch, flips at BER 0.02, counters counted once."""
    from eccq_b200.kv_cache import paged_prefill_attention_ecc
    lens = [77, 130]
    outs = []
    for b, n in enumerate(lens):       # one-shot prefill needs Sq == ctx: one call per row
        oc, bt = build_cache(oracle, codec, Hkv, D, bs, [n], seed=D + bs + b, ber=0.02)
        q = np.random.default_rng(b).standard_normal((1, Hq, n, D)).astype(np.float32)
        ref, corr, det = ref_prefill(oracle, q, oc, bt, [n], 1, interp)
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        out = paged_prefill_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array([n], np.int32)), dev(oc.k_scales),
                                          dev(oc.v_scales), 1, bs, codec=codec, use_interpolation=interp, counters=ctr)
        assert np.abs(out.cpu().numpy() - ref).max() < 2e-5, (codec, interp, n)
        assert (int(ctr[0]), int(ctr[1])) == (corr, det)
        outs.append(out)
    qh = dev(q).half()
    oh = paged_prefill_attention_ecc(qh, dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array([n], np.int32)), dev(oc.k_scales),
                                     dev(oc.v_scales), 1, bs, codec=codec, use_interpolation=interp)
    refh, _, _ = ref_prefill(oracle, qh.float().cpu().numpy(), oc, bt, [n], 1, interp)
    assert oh.dtype == torch.float16 and np.abs(oh.float().cpu().numpy() - refh).max() < 2e-3 * max(1.0, np.abs(refh).max())


@pytest.mark.gpu
def test_chunked_prefill_and_shared_context(oracle):
    """ctx > Sq (append mode: the queries are the LAST Sq positions, bottom-right mask), a ragged batch with an empty row, and the
    reference's shared-context layout (one context, B query rows)."""
    from eccq_b200.kv_cache import paged_prefill_attention_ecc
    Hq, Hkv, D, bs, Sq = 8, 2, 128, 16, 9
    lens = [300, 9, 0, 40]
    oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=5, ber=0.02)
    q = np.random.default_rng(7).standard_normal((len(lens), Hq, Sq, D)).astype(np.float32)
    ref, corr, det = ref_prefill(oracle, q, oc, bt, lens, 1, True)
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = paged_prefill_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales),
                                      dev(oc.v_scales), 1, bs, use_interpolation=True, counters=ctr)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5
    assert not out[2].any()
    assert (int(ctr[0]), int(ctr[1])) == (corr, det)
    # one shared context for all query rows: decoded (and counted) once
    ref1, corr1, det1 = ref_prefill(oracle, q, oc, bt[:1], lens[:1], 1, False)
    ctr.zero_()
    out1 = paged_prefill_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt[:1]), dev(np.array(lens[:1], np.int32)), dev(oc.k_scales),
                                       dev(oc.v_scales), 1, bs, counters=ctr)
    assert np.abs(out1.cpu().numpy() - ref1).max() < 2e-5
    assert (int(ctr[0]), int(ctr[1])) == (corr1, det1)


@pytest.mark.gpu
def test_backend_prefill_takes_the_fused_kernel_and_matches_the_library_route(monkeypatch):
    """ECCBackend.attend for q_len > 1: the fused kernel (default for contexts up to 2 048 tokens) against the round-1 route
    (fused read -> fp K/V -> library SDPA) on the same cache image: same outputs to fp32 rounding, same counters."""
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    import eccq_b200.kv_cache.ecc_shim as shim
    torch.manual_seed(0)
    Hq, Hkv, D, S, L = 8, 2, 64, 200, 2
    k = torch.randn(1, S, Hkv * D, device="cuda")
    v = torch.randn(1, S, Hkv * D, device="cuda")
    q = torch.randn(1, Hq, S, D, device="cuda")
    res = []
    for limit in (2048, 0):
        monkeypatch.setattr(shim, "_PREFILL_FUSED_MAX_CTX", limit)
        cfg = ECCShimConfig(codec="hamming84", ber=0.02, inject_errors=True, block_size=16, num_blocks=32, seed=3, use_interpolation=True,
                            backend="cuda")
        mgr = SimpleBlockManager(32, 16, L, Hkv, D, device="cuda", codec="hamming84")
        be = ECCBackend(mgr, cfg, Hq)
        be.write(k, v, layer_idx=1)
        out = be.attend(q, layer_idx=1)
        torch.cuda.synchronize()
        res.append((out, be._counters.clone()))
    assert (res[0][0] - res[1][0]).abs().max() < 2e-5
    assert torch.equal(res[0][1], res[1][1]) and int(res[0][1][1]) > 0

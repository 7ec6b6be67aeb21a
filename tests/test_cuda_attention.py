"""ECC-fused paged decode attention on the CUDA backend vs the reference's Triton kernel outputs (golden), the
reference's Python path (golden) and the CPU oracle (GQA, ragged contexts, all codecs, interpolation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4     # the reference's own bar for fp32 queries (tests/test_paged_attention_ecc.py:115-116)
TOL_F16 = 2e-3     # fp16 queries/outputs: |out| <= 8*scale ~ O(1), half has 11 bits -> 1e-3 relative
TOL_BF16 = 1.6e-2  # bf16: 8 bits


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_triton_outputs(golden):
    from eccq_b200.kv_cache import paged_attention_ecc, reference_attention_ecc
    g = golden("attention_vectors")
    Bq, H, D, L, bs, nb, mb = (int(v) for v in g["dims"])
    args = (dev(g["q"]), dev(g["k_cache"]), dev(g["v_cache"]), dev(g["block_table"]), dev(g["context_lens"]), dev(g["k_scales"]), 1, bs)
    out = paged_attention_ecc(*args, v_scales=dev(g["v_scales"]))
    assert out.shape == (Bq, H, D) and out.dtype == torch.float32
    assert np.abs(out.cpu().numpy() - g["out_triton_kv"]).max() < TOL_F32
    out = paged_attention_ecc(*args)                       # v_scales defaults to k_scales (attention_ecc.py:666-668)
    assert np.abs(out.cpu().numpy() - g["out_triton_kscale_only"]).max() < TOL_F32
    assert np.abs(out.cpu().numpy() - g["out_pyref_kscale_only"]).max() < TOL_F32
    ref = reference_attention_ecc(*args, D)
    assert np.abs(ref.cpu().numpy() - g["out_pyref_kscale_only"]).max() < TOL_F32
    Dg, ncw = (int(v) for v in g["g_dims"])
    out = paged_attention_ecc(dev(g["g_q"]), dev(g["g_k_cache"]), dev(g["g_v_cache"]), dev(g["block_table"]), dev(g["context_lens"]),
                              dev(g["k_scales"]), 0, bs, codec="golay")
    assert np.abs(out.cpu().numpy() - g["g_out"]).max() < TOL_F32
    with pytest.raises(ValueError):
        paged_attention_ecc(*args, codec="bogus")
    with pytest.raises(AssertionError):
        paged_attention_ecc(args[0].cpu(), *args[1:])


def _build(oracle, codec, B, Hq, Hkv, D, bs, lens, seed, ber=0.03):
    rng = np.random.default_rng(seed)
    L = 2
    nblk = sum(-(-max(n, 1) // bs) for n in lens) + 2
    mb = max(-(-max(n, 1) // bs) for n in lens) + 1
    oc = oracle.PagedCache(nblk, bs, L, Hkv, D, codec)
    perm = rng.permutation(nblk)
    bt = np.full((B, mb), -1, dtype=np.int32)
    p = 0
    for b, n in enumerate(lens):
        nb_ = -(-n // bs)
        bt[b, :nb_] = perm[p:p + nb_]
        p += nb_
        if n:
            k = rng.standard_normal((1, n, Hkv * D)).astype(np.float32)
            v = rng.standard_normal((1, n, Hkv * D)).astype(np.float32)
            cw = oc.cw
            nbits = oracle.CODEC_BITS[codec]
            fl = rng.random((1, n, Hkv, cw, nbits)) < ber
            m1 = (fl * (1 << np.arange(nbits))).sum(-1).astype(oc.k_cache.dtype)
            oracle.cache_write(oc, k, v, 1, bt[b], k_mask=m1, v_mask=np.roll(m1, 3, axis=3))
    q = rng.standard_normal((B, Hq, D)).astype(np.float32)
    return oc, bt, q


@pytest.mark.parametrize("codec", ["hamming84", "hamming74", "golay", "int4"])
@pytest.mark.parametrize("Hq,Hkv,D", [(4, 4, 64), (32, 8, 128), (8, 1, 128), (6, 2, 32), (4, 2, 256)])
def test_vs_oracle_gqa_ragged(oracle, codec, Hq, Hkv, D):
    from eccq_b200.kv_cache import paged_attention_ecc
    bs = 16
    lens = [300, 1, 0, 16, 33]
    B = len(lens)
    oc, bt, q = _build(oracle, codec, B, Hq, Hkv, D, bs, lens, seed=Hq * 7 + D)
    for interp in ((False, True) if codec == "hamming84" else (False,)):
        ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, use_interpolation=interp)
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales),
                                  1, bs, codec=codec, v_scales=dev(oc.v_scales), use_interpolation=interp, counters=ctr)
        assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32, (codec, interp)
        assert not out[2].any()                                    # empty context -> zeros (Q14)
        assert (int(ctr[0]), int(ctr[1])) == (corr, det)
        for dt, tol in ((torch.float16, TOL_F16), (torch.bfloat16, TOL_BF16)):
            qh = dev(q).to(dt)
            refh, _, _ = oracle.paged_decode_attention(qh.float().cpu().numpy(), oc, bt, np.array(lens), 1, use_interpolation=interp)
            outh = paged_attention_ecc(qh, dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales),
                                       1, bs, codec=codec, v_scales=dev(oc.v_scales), use_interpolation=interp)
            assert outh.dtype == dt and np.abs(outh.float().cpu().numpy() - refh).max() < tol


def test_uniform_softmax_known_answer():
    """tests/test_paged_attention_ecc.py:511-594: q = 0 -> uniform weights -> output = mean of V."""
    from eccq_b200.ecc_codecs.cuda_kernels import hamming84_encode
    from eccq_b200.kv_cache import paged_attention_ecc
    H, D, bs, nb = 2, 64, 16, 4
    k = hamming84_encode(torch.full((nb, 1, H, bs * D), 9, dtype=torch.uint8, device="cuda"))
    vals = torch.tensor([8, 9, 10, 11], dtype=torch.uint8, device="cuda").repeat_interleave(D).repeat(4)   # tokens cycle 0,1,2,3
    v = hamming84_encode(vals.view(1, 1, 1, bs * D).expand(nb, 1, H, bs * D).contiguous())
    sc = torch.ones(nb, 1, H, bs, device="cuda")
    bt = torch.arange(nb, dtype=torch.int32, device="cuda").view(1, nb)
    out = paged_attention_ecc(torch.zeros(1, H, D, device="cuda"), k, v, bt, torch.tensor([64], dtype=torch.int32, device="cuda"), sc, 0, bs)
    assert torch.allclose(out, torch.full_like(out, 1.5), atol=1e-5)


def test_long_context_split_merge(oracle):
    """A context long enough for many chunks (split-KV + merge) against a torch fp32 softmax over the fused read."""
    from eccq_b200.kv_cache import paged_attention_ecc, reference_attention_ecc
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    Hq, Hkv, D, bs, S = 32, 8, 128, 16, 5000
    nblk = -(-S // bs)
    cfg = ECCShimConfig(codec="hamming84", ber=1e-2, inject_errors=True, block_size=bs, num_blocks=nblk, injector="philox_fast")
    mgr = SimpleBlockManager(nblk, bs, 1, Hkv, D, device="cuda", codec="hamming84")
    mgr.free_blocks = [int(b) for b in np.random.default_rng(0).permutation(nblk)]
    be = ECCBackend(mgr, cfg, Hq)
    torch.manual_seed(0)
    be.write(torch.randn(1, S, Hkv * D, device="cuda", dtype=torch.float16), torch.randn(1, S, Hkv * D, device="cuda", dtype=torch.float16), 0)
    q = torch.randn(1, Hq, D, device="cuda")
    cl = torch.tensor([S], dtype=torch.int32, device="cuda")
    out = paged_attention_ecc(q, mgr.k_cache, mgr.v_cache, mgr.block_table[:1], cl, mgr.k_scales, 0, bs, v_scales=mgr.v_scales)
    ref = reference_attention_ecc(q, mgr.k_cache, mgr.v_cache, mgr.block_table[:1], cl, mgr.k_scales, 0, bs, D)
    # reference_attention_ecc uses k_scales for V too (Q15) -> compare with the same convention
    out_k = paged_attention_ecc(q, mgr.k_cache, mgr.v_cache, mgr.block_table[:1], cl, mgr.k_scales, 0, bs)
    assert (out_k - ref).abs().max() < TOL_F32
    assert torch.isfinite(out).all()

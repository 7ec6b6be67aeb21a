"""The unprotected comparison modes on the CUDA backend against fixtures produced by the REFERENCE's own code
(tests/golden/make_golden.py `unprotected`): ECCBackend with codec fp16 / fp8 / int4 (kv_cache/ecc_shim.py:511-625, :802-928)
and kv_cache/unprotected_shim.py's UnprotectedBackend, BER 0 and 0.05 -- cache images bit-exact (incl. the per-(token, head)
Philox injection on fp8 bytes and INT4 nibbles), attention outputs within 2e-5, statistics exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-5


def close(a, b, tol=TOL):
    """max |a - b| < tol (relative to the largest value) where both are finite.  A bit flip can turn an fp8 value into NaN
    (0x7F / 0xFF), in the reference as here -- the cache images are compared bit for bit; how far a NaN then spreads through
    the attention output is a property of the SDPA implementation (masked weights times NaN, NaN scores inside a softmax
    row ...: the reference's CPU math path and the CUDA kernels differ), so only the finite values are compared."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both = np.isfinite(a) & np.isfinite(b)
    if not np.isnan(b).any():
        assert not np.isnan(a).any()
    scale = max(1.0, float(np.abs(b[both]).max()))      # flipped exponent bits make values of a few hundred
    return both.any() and np.abs(a - b)[both].max() < tol * scale


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("codec", ["fp16", "fp8", "int4"])
@pytest.mark.parametrize("ber", [0.0, 0.05])
def test_ecc_backend_unprotected_codecs(golden, codec, ber):
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    g = golden("unprotected_vectors")
    B, S, Hq, Hkv, D, L, bs, nblk = (int(x) for x in g["dims"])
    tag = f"{codec}_ber{ber}"
    cfg = ECCShimConfig(codec=codec, ber=ber, block_size=bs, num_blocks=nblk, inject_errors=ber > 0, seed=42, backend="cuda")
    mgr = SimpleBlockManager(nblk, bs, L, Hkv, D, device="cuda", codec=codec)
    assert mgr.cache_dtype == {"fp16": torch.float16, "fp8": torch.float8_e4m3fn, "int4": torch.uint8}[codec]
    assert mgr.needs_scales == (codec == "int4")
    be = ECCBackend(mgr, cfg, Hq)
    be.write(T(g["k"]), T(g["v"]), layer_idx=1, seq_id=0)
    raw = (lambda t: t.view(torch.uint8).cpu().numpy()) if codec == "fp8" else (lambda t: t.cpu().numpy())
    assert np.array_equal(raw(mgr.k_cache), g[tag + "_k_cache"]), "K cache image differs from the reference's"
    assert np.array_equal(raw(mgr.v_cache), g[tag + "_v_cache"])
    assert np.array_equal(mgr.k_scales.cpu().numpy(), g[tag + "_k_scales"]) and np.array_equal(mgr.v_scales.cpu().numpy(), g[tag + "_v_scales"])
    out = be.attend(T(g["q_prefill"]), layer_idx=1, seq_id=0)
    assert close(out.cpu().numpy(), g[tag + "_out_prefill"])
    out = be.attend(T(g["q_decode"]), layer_idx=1, seq_id=0)
    assert close(out.cpu().numpy(), g[tag + "_out_decode"])
    inj, corr, det, total = (int(x) for x in g[tag + "_stats"])
    assert (be._injection_count, be._errors_corrected, be._errors_detected, be._total_values) == (inj, corr, det, total)


@pytest.mark.parametrize("ber", [0.0, 0.05])
def test_unprotected_backend(golden, ber):
    from eccq_b200.kv_cache import SimpleBlockManager, UnprotectedBackend, UnprotectedShimConfig
    g = golden("unprotected_vectors")
    B, S, Hq, Hkv, D, L, bs, nblk = (int(x) for x in g["dims"])
    tag = f"unprotected_ber{ber}"
    cfg = UnprotectedShimConfig(ber=ber, block_size=bs, num_blocks=nblk, inject_errors=ber > 0, seed=42)
    assert cfg.codec == "int4"
    mgr = SimpleBlockManager(nblk, bs, L, Hkv, D, device="cuda", codec="int4")
    be = UnprotectedBackend(mgr, cfg, Hq)
    be.write(T(g["k"]), T(g["v"]), layer_idx=1, seq_id=0)
    assert np.array_equal(mgr.k_cache.cpu().numpy(), g[tag + "_k_cache"]) and np.array_equal(mgr.v_cache.cpu().numpy(), g[tag + "_v_cache"])
    assert np.array_equal(mgr.k_scales.cpu().numpy(), g[tag + "_k_scales"]) and np.array_equal(mgr.v_scales.cpu().numpy(), g[tag + "_v_scales"])
    for q, key in ((g["q_prefill"], "_out_prefill"), (g["q_prefill"], "_out_prefill2"), (g["q_decode"], "_out_decode")):
        out = be.attend(T(q), layer_idx=1, seq_id=0)       # every call draws the next seed (unprotected_shim.py:279-280)
        assert close(out.cpu().numpy(), g[tag + key]), key
    assert (be._injection_count, be._bits_flipped, be._total_bits) == tuple(int(x) for x in g[tag + "_stats"])


def test_unprotected_patch_and_stats():
    """patch_model_with_unprotected_attention / reset / stats on a random-init 2-layer GPT-2: BER 0 reproduces the int4 ECC-shim
    mode (same quantiser, same storage, top-left causal mask), BER > 0 flips about ber * total_bits bits."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from eccq_b200.kv_cache import (ECCShimConfig, UnprotectedShimConfig, get_unprotected_stats, patch_model_with_ecc_attention,
                                    patch_model_with_unprotected_attention, reset_ecc_cache, reset_unprotected_cache)
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(n_layer=2, n_head=2, n_embd=64, vocab_size=100, n_positions=64, resid_pdrop=0.0, embd_pdrop=0.0,
                                       attn_pdrop=0.0)).eval().cuda()
    ids = torch.randint(0, 100, (1, 40), device="cuda")
    with torch.no_grad():
        with patch_model_with_ecc_attention(model, ECCShimConfig(codec="int4", block_size=16, num_blocks=8), num_blocks=8):
            reset_ecc_cache(model)
            ref = model(ids).logits
        with patch_model_with_unprotected_attention(model, UnprotectedShimConfig(block_size=16, num_blocks=8), num_blocks=8):
            reset_unprotected_cache(model)
            out = model(ids).logits
            st = get_unprotected_stats(model)
        # same cache image, same arithmetic; the ECC shim's prefill runs in the fused kernel, this one through SDPA: fp32 rounding apart
        assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5)
        assert st == {"allocated_blocks": 3, "free_blocks": 5, "sequences": 1, "injection_count": 0, "bits_flipped": 0, "total_bits": 0,
                      "actual_ber": 0.0}
        with patch_model_with_unprotected_attention(model, UnprotectedShimConfig(ber=0.02, inject_errors=True, block_size=16, num_blocks=8),
                                                    num_blocks=8):
            reset_unprotected_cache(model)
            noisy = model(ids).logits
            st = get_unprotected_stats(model)
        assert st["injection_count"] == 2 and st["total_bits"] == 2 * 2 * 40 * 2 * 32 * 4
        assert abs(st["actual_ber"] - 0.02) < 0.006 and torch.isfinite(noisy).all() and not torch.equal(noisy, ref)
    assert not hasattr(model, "_unprotected_backend")

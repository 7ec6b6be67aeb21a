"""Config 2 in miniature on the GPU: a random-init 2-layer GPT-2 (weights from the fixture) under
patch_model_with_ecc_attention(backend="cuda"), every ECC mode, BER 0 and 1e-2 -- logits, loss and get_ecc_stats against
the numbers the REFERENCE produced on the same weights and token ids (tests/golden/make_golden.py gpt2)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODES = (("hamming74", False), ("hamming84", False), ("hamming84", True), ("golay", False))


@pytest.fixture(scope="module")
def gpt2(golden):
    from transformers import GPT2Config, GPT2LMHeadModel
    g = golden("gpt2_shim_vectors")
    cfg = GPT2Config(n_layer=2, n_head=2, n_embd=32, vocab_size=100, n_positions=64, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    model = GPT2LMHeadModel(cfg).eval()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w::")}
    model.load_state_dict(sd)
    model = model.cuda()
    ids = torch.from_numpy(g["input_ids"]).cuda()
    with torch.no_grad():
        base = model(ids).logits
    assert np.abs(base.cpu().numpy() - g["baseline_logits"]).max() < 1e-4      # same model on both sides
    return model, ids, g


@pytest.mark.parametrize("codec,interp", MODES)
@pytest.mark.parametrize("ber", [0.0, 1e-2])
def test_patched_gpt2_matches_reference(gpt2, codec, interp, ber):
    from eccq_b200.kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention, reset_ecc_cache
    model, ids, g = gpt2
    tag = f"{codec}{'_interp' if interp else ''}_ber{ber}"
    cfg = ECCShimConfig(codec=codec, ber=ber, inject_errors=ber > 0, seed=42, use_interpolation=interp, block_size=16, num_blocks=8,
                        backend="cuda")
    orig = model.transformer.h[0].attn
    with patch_model_with_ecc_attention(model, cfg, num_blocks=8):
        assert hasattr(model, "_ecc_backend") and model.transformer.h[0].attn is not orig
        reset_ecc_cache(model)
        with torch.no_grad():
            r = model(ids, labels=ids)
        st = get_ecc_stats(model)
    assert model.transformer.h[0].attn is orig and not hasattr(model, "_ecc_backend")       # restored on exit
    ref = g[tag + "_stats"]
    got = [st[k] for k in ("injection_count", "errors_corrected", "errors_detected", "total_values", "allocated_blocks",
                           "free_blocks", "sequences")]
    assert got == ref.tolist(), (tag, got, ref.tolist())           # counters are bit-exact (same Philox stream)
    # logits: GPU fp32 GEMMs vs CPU fp32 GEMMs -> small float noise only (the cache contents are identical)
    assert np.abs(r.logits.cpu().numpy() - g[tag + "_logits"]).max() < 2e-4, tag
    assert abs(float(r.loss) - float(g[tag + "_loss"])) < 1e-4


def test_stats_keys_and_reset(gpt2):
    """tests/test_ecc_shim.py:586-622, :737-868: stats dict keys, reset zeroes counters, stats accumulate across texts."""
    from eccq_b200.kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention, reset_ecc_cache
    model, ids, _ = gpt2
    cfg = ECCShimConfig(codec="hamming84", ber=0.05, inject_errors=True, num_blocks=8)
    with patch_model_with_ecc_attention(model, cfg, num_blocks=8):
        with torch.no_grad():
            model(ids)
            s1 = get_ecc_stats(model)
            model(ids)
            s2 = get_ecc_stats(model)
        assert set(s1) == {"allocated_blocks", "free_blocks", "sequences", "injection_count", "errors_corrected", "errors_detected",
                           "total_values"}
        assert s2["total_values"] == 2 * s1["total_values"] and s2["errors_corrected"] > s1["errors_corrected"] > 0
        reset_ecc_cache(model)
        s3 = get_ecc_stats(model)
        assert s3["errors_corrected"] == s3["errors_detected"] == s3["total_values"] == s3["injection_count"] == 0
        assert s3["allocated_blocks"] == 0 and s3["free_blocks"] == 8
        assert not torch.isnan(model(ids).logits).any()


def test_llama_style_gqa_model_runs():
    """LLaMA layout (q/k/v/o_proj + RoPE, GQA 4:2): patched forward is finite and close to the unpatched model at BER 0
    (INT4 quantisation noise only)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from eccq_b200.kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=128, max_position_embeddings=128)
    model = LlamaForCausalLM(cfg).eval().cuda()
    ids = torch.randint(0, 128, (1, 33), device="cuda")
    with torch.no_grad():
        base = model(ids).logits
        for codec in ("hamming84", "golay"):
            with patch_model_with_ecc_attention(model, ECCShimConfig(codec=codec, num_blocks=8), num_blocks=8):
                out = model(ids).logits
                st = get_ecc_stats(model)
            assert torch.isfinite(out).all() and st["total_values"] == 2 * 2 * 33 * 2 * 16
            assert (out - base).abs().max() < 0.25 * base.abs().max()


def test_quantizer_registry_entry():
    """north_star: quantization_backends gains the CUDA backend; reference registry semantics (PerTokenQuantizer)."""
    from eccq_b200.ecc_codecs import QuantizationConfig, QuantizedTensor, get_quantizer, list_backends
    assert "cuda_per_token" in list_backends()
    q = get_quantizer("cuda-per-token")
    x = torch.randn(4, 16, 64, dtype=torch.float16)                  # CPU fp16 in, as tests/test_quantization_backends.py:88-91
    qt = q.quantize(x)
    assert isinstance(qt, QuantizedTensor) and qt.data.device == x.device and qt.data.dtype == torch.uint8
    # the kernel computes in fp32 (like the shim); PerTokenQuantizer's formula in fp32 must match bit for bit,
    # the same formula evaluated in fp16 (what the reference does for fp16 inputs) differs on rounding ties only
    xf = x.float()
    ref_scale = xf.abs().amax(dim=-1, keepdim=True).clamp(min=1e-8) / 7.0
    ref_q = (torch.round(xf / ref_scale) + 8).clamp(0, 15).to(torch.uint8)
    assert torch.equal(qt.data, ref_q) and torch.equal(qt.scales.float(), ref_scale.squeeze(-1).to(torch.float16).float())
    h_scale = x.abs().amax(dim=-1, keepdim=True).clamp(min=1e-8) / 7.0
    h_q = (torch.round(x / h_scale) + 8).clamp(0, 15).to(torch.uint8)
    assert (qt.data == h_q).float().mean() > 0.99
    xr = q.dequantize(qt)
    assert xr.shape == x.shape and (x.float() - xr.float()).pow(2).mean() < 0.02
    z = q.quantize(torch.zeros(2, 8))
    assert (z.data == 8).all()
    with pytest.raises(ValueError):
        get_quantizer("nope")
    q32 = get_quantizer("cuda_per_token", QuantizationConfig(dtype=torch.float32))
    assert q32.dequantize(q32.quantize(x.float().cuda())).dtype == torch.float32


def test_cuda_phase_timer_reports_bandwidth():
    from eccq_b200.ecc_codecs.cuda_kernels import hamming84_encode
    from eccq_b200.evaluation import TimingStats, cuda_phase_timer, cuda_timer
    x = torch.randint(0, 16, (1 << 24,), dtype=torch.uint8, device="cuda")
    st = TimingStats(values_processed=x.numel())
    hamming84_encode(x)          # first call in a process pays module load / allocator warm-up
    with cuda_phase_timer(st, "encode", bytes_moved=2 * x.numel()):
        hamming84_encode(x)
    assert st.encode_ns > 0 and st.gbps("encode") > 1.0
    assert cuda_timer(lambda: hamming84_encode(x), warmup=2, repeats=5) > 0


def test_sweep_driver_on_the_cuda_backend(gpt2):
    """the sweep that drives the path (evaluation/sweep.py of the reference) with backend='cuda' -- every
    protected mode x BER cell runs, counters behave (none at BER 0, corrections grow with the BER), a cell reproduces the
    reference's loss on the same weights / ids / seed, seeds aggregate."""
    from eccq_b200.evaluation.sweep import SweepConfig, generate_clean_logits, run_single_trial, run_sweep
    model, ids, g = gpt2
    texts = [ids[:1], ids[:1, :40]]
    clean = generate_clean_logits(model, None, texts, max_length=64)
    cfg = SweepConfig(ber_levels=[0.0, 1e-3, 1e-2], seeds=[42, 101], max_length=64, num_blocks=8, clean_logits=clean)
    res = run_sweep(model, None, texts, cfg)
    assert len(res.trials) == 4 * 3 * 2
    for mode in cfg.cache_modes:
        cells = res.aggregated[mode]
        assert cells[0.0].errors_corrected_mean == 0 and cells[0.0].errors_detected_mean == 0
        assert cells[0.0].perplexity_std == 0.0                         # no injection: seeds cannot matter
        assert cells[1e-2].errors_corrected_mean > cells[1e-3].errors_corrected_mean > 0
        assert 0.0 <= cells[1e-2].top5_accuracy_mean <= 1.0 and cells[1e-2].kl_divergence_mean >= 0.0
        assert math.isfinite(cells[1e-2].perplexity_mean)
    # the full-length text alone, hamming84 @ 1e-2, seed 42 is the golden cell of test_patched_gpt2_matches_reference
    cfg1 = SweepConfig(max_length=64, num_blocks=8)
    t = run_single_trial(model, None, [ids[:1]], "int4-hamming84", 1e-2, 42, cfg1)
    assert abs(math.log(t.perplexity) - float(g["hamming84_ber0.01_loss"])) < 1e-4
    ref = g["hamming84_ber0.01_stats"]
    assert [t.injection_count, t.errors_corrected, t.errors_detected, t.total_values] == ref[:4].tolist()
    from eccq_b200.evaluation.monte_carlo import MonteCarloConfig, run_monte_carlo_experiment
    mc = MonteCarloConfig(cache_modes=["int4-hamming84"], ber_levels=[1e-2], seeds=[42], max_length=64, num_blocks=8)
    r = run_monte_carlo_experiment(model, None, mc, [ids[:1]], verbose=False)
    assert r.aggregated["int4-hamming84"][1e-2].perplexity_mean == t.perplexity and r.trials[0].kl_divergence > 0
    with pytest.raises(ValueError):
        run_single_trial(model, None, texts, "fp16", 0.0, 42, cfg1)       # unprotected comparison codecs: out of scope


@pytest.mark.parametrize("codec,interp", MODES)
def test_incremental_decode_equals_one_shot_prefill(gpt2, codec, interp):
    """SURVEY 8 f3 on a model: with `append=True` a prefill followed by single-token steps (fused decode attention over the
    protected cache, q_len == 1) reproduces the logits of one prefill over the whole sequence (fused read + SDPA) -- the
    per-(token, head) quantisation makes the cache image identical, so only fp32 summation order differs."""
    from eccq_b200.kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention, reset_ecc_cache
    model, ids, _ = gpt2
    S, T = ids.shape[1], ids.shape[1] - 5
    cfg = ECCShimConfig(codec=codec, num_blocks=8, use_interpolation=interp, append=True)
    with torch.no_grad(), patch_model_with_ecc_attention(model, cfg, num_blocks=8):
        full = model(ids[:1]).logits
        reset_ecc_cache(model)
        steps = [model(ids[:1, :T]).logits]
        for t in range(T, S):
            steps.append(model(ids[:1, t:t + 1], position_ids=torch.tensor([[t]], device="cuda")).logits)
        st = get_ecc_stats(model)
    inc = torch.cat(steps, dim=1)
    assert inc.shape == full.shape
    assert (inc - full).abs().max().item() < 2e-4 * max(1.0, full.abs().max().item())
    assert st["errors_corrected"] == 0 and st["errors_detected"] == 0
    assert st["total_values"] == 2 * 2 * S * 2 * 16            # K+V x layers x tokens x heads x head_dim: every token written once


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2), (torch.float16, 1e-2)])
def test_llama_gqa_incremental_decode_equals_prefill(dtype, tol):
    """The same on the LLaMA layout (RoPE positions, GQA 4:2, fused decode attention with two query heads per kv head), in
    fp32 and in the half-precision dtypes a served model runs in (K/V reach the cache, and come back, in the model's dtype)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from eccq_b200.kv_cache import ECCShimConfig, patch_model_with_ecc_attention, reset_ecc_cache
    torch.manual_seed(1)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=128, max_position_embeddings=128)
    model = LlamaForCausalLM(cfg).eval().cuda().to(dtype)
    S, T = 37, 30
    ids = torch.randint(0, 128, (1, S), device="cuda")
    for codec, interp in (("hamming84", False), ("hamming84", True), ("golay", False)):
        ecc = ECCShimConfig(codec=codec, num_blocks=8, use_interpolation=interp, append=True)
        with torch.no_grad(), patch_model_with_ecc_attention(model, ecc, num_blocks=8):
            full = model(ids).logits
            reset_ecc_cache(model)
            steps = [model(ids[:, :T]).logits]
            for t in range(T, S):
                steps.append(model(ids[:, t:t + 1], position_ids=torch.tensor([[t]], device="cuda")).logits)
        inc = torch.cat(steps, dim=1).float()
        assert inc.dtype == torch.float32 and steps[0].dtype == dtype and torch.isfinite(inc).all()
        assert (inc - full.float()).abs().max().item() < tol * max(1.0, full.float().abs().max().item()), codec


def test_batched_append_equals_per_sequence_runs(gpt2):
    """append=True maps batch row b to sequence b: a batched prefill + batched single-token steps give, row by row, the logits
    of running each sequence alone (the reference's write lets later rows overwrite earlier ones, Q11)."""
    from eccq_b200.kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention, reset_ecc_cache
    model, ids, _ = gpt2
    g = torch.Generator().manual_seed(5)
    B, S, T = 3, 40, 36
    batch = torch.randint(0, 100, (B, S), generator=g).cuda()
    pos = lambda t: torch.full((B, 1), t, device="cuda")
    for codec, interp in (("hamming84", True), ("golay", False)):
        cfg = ECCShimConfig(codec=codec, num_blocks=16, use_interpolation=interp, append=True)
        with torch.no_grad(), patch_model_with_ecc_attention(model, cfg, num_blocks=16):
            together = torch.cat([model(batch[:, :T]).logits] + [model(batch[:, t:t + 1], position_ids=pos(t)).logits for t in range(T, S)], dim=1)
            st = get_ecc_stats(model)
            assert st["sequences"] == B and st["total_values"] == 2 * 2 * B * S * 2 * 16
            alone = []
            for b in range(B):
                reset_ecc_cache(model)
                alone.append(model(batch[b:b + 1]).logits)
        assert (together - torch.cat(alone)).abs().max().item() < 2e-4, codec

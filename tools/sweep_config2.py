#!/usr/bin/env python
"""BASELINE.json configs[1] ("config 2") end to end on the CUDA backend: GPT-2 small (12 layers x 12 heads x 64, random-init
weights -- no checkpoints in the sandbox), ECC-shim Monte Carlo sweep over hamming74 / hamming84 / hamming84-interp / golay at
BER 0, 1e-4, 1e-3, 1e-2 with seeds 42 / 101 / 997, synthetic token sequences.  Prints one row per (mode, BER) cell and the
wall time; the numbers are meaningful relative to each other (random weights), the point is that the reference's sweep runs
unchanged on backend="cuda".   Usage: python tools/sweep_config2.py [n_texts] [seq_len]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from transformers import GPT2Config, GPT2LMHeadModel  # noqa: E402

from eccq_b200.evaluation.sweep import SweepConfig, generate_clean_logits, run_sweep  # noqa: E402

n_texts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seq_len = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
model = GPT2LMHeadModel(GPT2Config(resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)).eval().cuda()
gen = torch.Generator().manual_seed(1)
texts = [torch.randint(0, 50257, (1, seq_len), generator=gen) for _ in range(n_texts)]
clean = generate_clean_logits(model, None, texts, max_length=seq_len)
cfg = SweepConfig(seeds=[42, 101, 997], max_length=seq_len, num_blocks=256, clean_logits=clean)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = run_sweep(model, None, texts, cfg)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"GPT-2 small, random init, {n_texts} texts x {seq_len} tokens, {len(res.trials)} trials in {dt:.1f} s "
      f"({dt / len(res.trials) * 1e3:.0f} ms per trial = {n_texts} forward passes of 12 patched layers)")
print(f"{'mode':24s} {'BER':>7s} {'ppl mean':>12s} {'ppl std':>10s} {'KL':>10s} {'top5':>7s} {'corrected':>12s} {'detected':>10s}")
for mode in cfg.cache_modes:
    for ber in cfg.ber_levels:
        a = res.aggregated[mode][ber]
        print(f"{mode:24s} {ber:7.0e} {a.perplexity_mean:12.2f} {a.perplexity_std:10.3f} {a.kl_divergence_mean:10.3e} "
              f"{a.top5_accuracy_mean:7.4f} {a.errors_corrected_mean:12.0f} {a.errors_detected_mean:10.0f}")

"""Monte Carlo BER experiment on the CUDA backend -- the `backend` pass-through of SURVEY section 8 row f1.

Mirrors the driving part of the reference's evaluation/experiments/monte_carlo.py: `MonteCarloConfig` (:31-72, with
`backend="cuda"` where the reference says "triton") -> `SweepConfig` -> `run_sweep` (:75-127).  The reference loads WikiText-2
and a pretrained checkpoint here; neither exists in the sandbox, so the caller passes `texts` (strings with a tokenizer, or
token-id tensors) and the model.  LaTeX / table formatting (:130-) is reporting, out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from .sweep import BER_LEVELS_EXTENDED, CACHE_MODE_ORDER, SEEDS, SweepConfig, SweepResults, generate_clean_logits, run_sweep


@dataclass
class MonteCarloConfig:
    model_name: str = "gpt2"
    cache_modes: Optional[list] = None
    ber_levels: Optional[list] = None
    seeds: Optional[list] = None
    max_samples: int = 50
    max_length: int = 256
    device: str = "auto"
    backend: str = "cuda"
    num_blocks: int = 2048
    block_size: int = 16
    injector: str = "philox_ref"
    compute_kl_divergence: bool = True
    compute_top5: bool = True
    compute_catastrophic: bool = True
    catastrophic_threshold: float = 1000.0

    def __post_init__(self):
        if self.cache_modes is None:
            self.cache_modes = CACHE_MODE_ORDER.copy()
        if self.ber_levels is None:
            self.ber_levels = BER_LEVELS_EXTENDED.copy()
        if self.seeds is None:
            self.seeds = SEEDS.copy()

    def to_sweep_config(self, clean_logits=None) -> SweepConfig:
        return SweepConfig(cache_modes=self.cache_modes, ber_levels=self.ber_levels, seeds=self.seeds, max_length=self.max_length,
                           block_size=self.block_size, num_blocks=self.num_blocks, device=self.device, aggregate_seeds=True,
                           compute_kl_divergence=self.compute_kl_divergence, compute_top5=self.compute_top5,
                           compute_catastrophic=self.compute_catastrophic, catastrophic_threshold=self.catastrophic_threshold,
                           clean_logits=clean_logits, backend=self.backend, injector=self.injector)


def run_monte_carlo_experiment(model, tokenizer, config: MonteCarloConfig, texts, verbose: bool = True) -> SweepResults:
    texts = list(texts)[:config.max_samples]
    clean = None
    if config.compute_kl_divergence:
        device = "cuda" if config.device == "auto" else config.device
        clean = generate_clean_logits(model, tokenizer, texts, max_length=config.max_length, device=device)

    def progress(message, current, total):
        if verbose:
            print(f"  [{(current / total * 100 if total else 0):5.1f}%] {message}")

    return run_sweep(model, tokenizer, texts, config.to_sweep_config(clean), progress)

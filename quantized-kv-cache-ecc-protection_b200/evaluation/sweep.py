"""BER sweep over the ECC shim on the CUDA backend -- SURVEY section 8 row f1 (the direct caller of the hot path).

Mirrors the part of the reference's evaluation/sweep.py that drives the path (SweepConfig :64-111, TrialResult :113-154,
run_single_trial :352-369, run_sweep :372-420, run_sweep_single_seed :423-445, the trial body :448-626) with
`SweepConfig.backend = "cuda"`: every (cache mode, BER, seed) cell patches the model with
`patch_model_with_ecc_attention`, runs each text once and reads `get_ecc_stats`.  The ECC-protected INT4 modes of
constants.py:157-165 are supported (`int4`, `int4-hamming`, `int4-hamming84`, `int4-hamming84-interp`, `int12-golay`);
the unprotected fp16 / fp8 comparison codecs are outside this repository's scope.

Two deliberate differences, both because the sandbox has no datasets or tokenizer files: `texts` may be token-id tensors
(shape [1, S] or [S]) instead of strings, in which case `tokenizer` may be None; and the quality metrics are computed from
the logits of the SAME forward pass that produced the loss (the reference re-runs the model per metric,
evaluation/metrics.py:146-228) -- same definitions: KL(clean || corrupted) of metrics.py:111-118 averaged over texts,
top-5 accuracy of :146-186, catastrophic rate of :136-143 over per-text perplexities.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from ..kv_cache import ECCShimConfig, get_ecc_stats, patch_model_with_ecc_attention, reset_ecc_cache

# constants.py:157-165
MODE_CONFIG = {
    "int4": {"codec": "int4", "use_interpolation": False},
    "int4-hamming": {"codec": "hamming74", "use_interpolation": False},
    "int4-hamming84": {"codec": "hamming84", "use_interpolation": False},
    "int4-hamming84-interp": {"codec": "hamming84", "use_interpolation": True},
    "int12-golay": {"codec": "golay", "use_interpolation": False},
}
CACHE_MODE_ORDER = ["int4-hamming", "int4-hamming84", "int4-hamming84-interp", "int12-golay"]
BER_LEVELS = [0.0, 1e-4, 1e-3, 1e-2]
BER_LEVELS_EXTENDED = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2]      # constants.py:93
SEEDS = [42, 101, 997, 1999, 4999, 7919, 10007, 15073, 21001, 31337]


def get_mode_config(cache_mode: str) -> dict:
    if cache_mode not in MODE_CONFIG:
        raise ValueError(f"Unknown cache mode: {cache_mode}. Valid modes: {list(MODE_CONFIG.keys())}")
    return MODE_CONFIG[cache_mode].copy()


@dataclass
class SweepConfig:
    cache_modes: List[str] = field(default_factory=lambda: CACHE_MODE_ORDER.copy())
    ber_levels: List[float] = field(default_factory=lambda: BER_LEVELS.copy())
    seeds: List[int] = field(default_factory=lambda: [42])
    aggregate_seeds: bool = True
    max_length: int = 256
    block_size: int = 16
    num_blocks: int = 2048
    device: str = "cuda"
    compute_kl_divergence: bool = True
    compute_top5: bool = True
    compute_catastrophic: bool = True
    catastrophic_threshold: float = 1000.0
    clean_logits: Optional[List[torch.Tensor]] = None
    backend: str = "cuda"
    injector: str = "philox_ref"      # the reference's Philox stream (bit-exact cache images); "philox_fast" = own stream

    @classmethod
    def default(cls) -> "SweepConfig":
        return cls()

    @classmethod
    def full(cls) -> "SweepConfig":
        return cls(seeds=SEEDS.copy())


@dataclass
class TrialResult:
    cache_mode: str
    ber: float
    seed: int
    perplexity: float
    errors_corrected: int = 0
    errors_detected: int = 0
    total_values: int = 0
    kl_divergence: float = 0.0
    top5_accuracy: float = 1.0
    catastrophic_rate: float = 0.0
    injection_count: int = 0
    correction_rate: float = 0.0
    detection_rate: float = 0.0
    silent_corruption_rate: float = 0.0


@dataclass
class AggregatedResult:
    """Mean / sample standard deviation over the seeds of one (mode, BER) cell."""
    cache_mode: str
    ber: float
    n_trials: int
    perplexity_mean: float
    perplexity_std: float
    kl_divergence_mean: float
    top5_accuracy_mean: float
    catastrophic_rate_mean: float
    errors_corrected_mean: float
    errors_detected_mean: float

    @classmethod
    def from_trials(cls, trials: Sequence[TrialResult]) -> "AggregatedResult":
        n = len(trials)
        mean = lambda xs: sum(xs) / n
        ppl = [t.perplexity for t in trials]
        m = mean(ppl)
        std = math.sqrt(sum((p - m) ** 2 for p in ppl) / (n - 1)) if n > 1 else 0.0
        return cls(trials[0].cache_mode, trials[0].ber, n, m, std, mean([t.kl_divergence for t in trials]),
                   mean([t.top5_accuracy for t in trials]), mean([t.catastrophic_rate for t in trials]),
                   mean([t.errors_corrected for t in trials]), mean([t.errors_detected for t in trials]))


@dataclass
class SweepResults:
    config: SweepConfig
    trials: List[TrialResult] = field(default_factory=list)
    aggregated: Dict[str, Dict[float, AggregatedResult]] = field(default_factory=dict)


def _token_ids(text, tokenizer, max_length: int, device):
    if isinstance(text, torch.Tensor):
        ids = text.view(1, -1)[:, :max_length]
        return ids.to(device)
    if not text.strip():
        return None
    return tokenizer(text, return_tensors="pt", truncation=True, max_length=max_length)["input_ids"].to(device)


def generate_clean_logits(model, tokenizer, texts, max_length: int = 256, device="cuda",
                          keep_on_device: bool = True) -> List[Optional[torch.Tensor]]:
    """Logits of the unpatched model per text (metrics.py:230-256); pass as SweepConfig.clean_logits.

    The reference parks them on the CPU and re-uploads [S, vocab] fp32 (51 MB for GPT-2 at S=256) per text per trial, which
    costs more than the forward pass; 180 GB of HBM holds them for any sample count the sweep uses, so they stay resident
    unless `keep_on_device=False`."""
    out = []
    with torch.no_grad():
        for text in texts:
            ids = _token_ids(text, tokenizer, max_length, device)
            if ids is None or ids.size(1) < 2:
                out.append(None)
                continue
            logits = model(ids, use_cache=False).logits[0].float()
            out.append(logits if keep_on_device else logits.cpu())
    return out


def run_single_trial(model, tokenizer, texts, cache_mode: str, ber: float, seed: int, config: SweepConfig) -> TrialResult:
    if config.backend != "cuda":
        raise ValueError(f"this package implements backend='cuda' only, got {config.backend!r}")
    if config.device not in ("cuda", "auto"):
        raise RuntimeError("backend='cuda' requires a CUDA device")
    mode = get_mode_config(cache_mode)
    ecc = ECCShimConfig(codec=mode["codec"], ber=ber, inject_errors=ber > 0, seed=seed, num_blocks=config.num_blocks,
                        block_size=config.block_size, use_interpolation=mode["use_interpolation"], backend="cuda",
                        injector=config.injector)
    device = next(model.parameters()).device
    # per-text results stay on the device until the trial ends: one host sync per trial instead of four per text, so the
    # host keeps launching the next forward pass while the GPU finishes this one
    rows, lens, positions = [], [], 0
    zero = torch.zeros((), device=device)
    # token ids go up before the first forward pass: a pageable host-to-device copy blocks until the stream drains
    all_ids = [_token_ids(text, tokenizer, config.max_length, device) for text in texts]
    with torch.no_grad(), patch_model_with_ecc_attention(model, ecc, num_blocks=ecc.num_blocks):
        for n, ids in enumerate(all_ids):
            if ids is None or ids.size(1) < 2:
                continue
            reset_ecc_cache(model)
            out = model(ids, labels=ids, use_cache=True)
            seq_len = ids.size(1)
            logits = out.logits[0].float()
            hit, kl = zero, zero.new_full((), float("nan"))
            if config.compute_top5:
                top5 = torch.topk(logits[:-1], k=5, dim=-1).indices
                hit = (top5 == ids[0, 1:, None]).any(-1).sum().float()
            if config.compute_kl_divergence and config.clean_logits is not None and config.clean_logits[n] is not None:
                clean = config.clean_logits[n].to(device)
                m = min(clean.size(0), logits.size(0))
                lp, lq = F.log_softmax(clean[:m], dim=-1), F.log_softmax(logits[:m], dim=-1)
                kl = (lp.exp() * (lp - lq)).sum(-1).mean()
            rows.append(torch.stack((out.loss.float(), hit, kl)))
            lens.append(seq_len)
        stats = get_ecc_stats(model)
    total_loss, total_tokens, hits = 0.0, 0, 0
    kls, sample_ppls = [], []
    for (loss, hit, kl), seq_len in zip(torch.stack(rows).tolist() if rows else [], lens):
        if math.isnan(loss) or math.isinf(loss):          # such a text counts as catastrophic and nothing else
            sample_ppls.append(float("inf"))
            continue
        total_loss += loss * seq_len
        total_tokens += seq_len
        sample_ppls.append(float(math.exp(min(loss, 700.0))))
        if config.compute_top5:
            hits += int(hit)
            positions += seq_len - 1
        if not (math.isnan(kl) or math.isinf(kl)):
            kls.append(kl)
    ppl = float("inf") if total_tokens == 0 else float(math.exp(min(total_loss / total_tokens, 700.0)))
    corrected, detected = int(stats.get("errors_corrected", 0)), int(stats.get("errors_detected", 0))
    events = corrected + detected
    cat = 0.0
    if config.compute_catastrophic and sample_ppls:
        cat = sum(1 for p in sample_ppls if p > config.catastrophic_threshold or math.isinf(p)) / len(sample_ppls)
    return TrialResult(cache_mode=cache_mode, ber=ber, seed=seed, perplexity=ppl, errors_corrected=corrected,
                       errors_detected=detected, total_values=int(stats.get("total_values", 0)),
                       kl_divergence=sum(kls) / len(kls) if kls else 0.0,
                       top5_accuracy=hits / positions if positions else (1.0 if not config.compute_top5 else 0.0),
                       catastrophic_rate=cat, injection_count=int(stats.get("injection_count", 0)),
                       correction_rate=corrected / events if events else 0.0, detection_rate=detected / events if events else 0.0)


def run_sweep(model, tokenizer, texts, config: SweepConfig = None,
              progress_callback: Callable[[str, int, int], None] = None) -> SweepResults:
    config = config or SweepConfig.default()
    results = SweepResults(config=config)
    total = len(config.cache_modes) * len(config.ber_levels) * len(config.seeds)
    current = 0
    for cache_mode in config.cache_modes:
        for ber in config.ber_levels:
            cell = []
            for seed in config.seeds:
                if progress_callback:
                    progress_callback(f"{cache_mode} @ BER={ber:.0e} seed={seed}", current, total)
                trial = run_single_trial(model, tokenizer, texts, cache_mode, ber, seed, config)
                results.trials.append(trial)
                cell.append(trial)
                current += 1
            if config.aggregate_seeds and cell:
                results.aggregated.setdefault(cache_mode, {})[ber] = AggregatedResult.from_trials(cell)
    return results


def run_sweep_single_seed(model, tokenizer, texts, config: SweepConfig = None, seed: int = 42,
                          progress_callback: Callable[[str, int, int], None] = None) -> Dict[str, Dict[float, TrialResult]]:
    config = config or SweepConfig.default()
    config.seeds, config.aggregate_seeds = [seed], False
    out: Dict[str, Dict[float, TrialResult]] = {}
    for trial in run_sweep(model, tokenizer, texts, config, progress_callback).trials:
        out.setdefault(trial.cache_mode, {})[trial.ber] = trial
    return out

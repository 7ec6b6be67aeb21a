"""Timing infrastructure of the CUDA backend and the BER sweep that drives the path (`evaluation.sweep`); the rest
of the reference's evaluation/ package is out of scope."""
from .timing import (AggregatedTimingStats, TimingContext, TimingStats, cuda_phase_timer, cuda_timer, cuda_transfer_timer,  # noqa: F401
                     measured_hbm_peak_gbs, run_warmup)

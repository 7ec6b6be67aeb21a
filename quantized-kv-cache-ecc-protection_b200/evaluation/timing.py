"""
Timing containers of the reference (evaluation/timing.py:8-239: TimingStats, AggregatedTimingStats, TimingContext,
cuda_transfer_timer, run_warmup; same field names) plus what the CUDA backend adds (north_star: "evaluation/timing
... gains the new timings"): phases for the fused kernels, a bytes-moved field per phase so GB/s and the HBM
roofline fraction fall out, and a CUDA-event timer bound to the launching stream.
"""
from __future__ import annotations

import json
import time
from contextlib import contextmanager
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional

import torch

# algorithmic bytes per INT4 value; D = head_dim
ALGORITHMIC_BYTES = {
    "encode": lambda D=128: 2.0, "decode": lambda D=128: 3.0, "decode_data_only": lambda D=128: 2.0,
    "inject_mask": lambda D=128: 3.0, "inject_philox": lambda D=128: 2.0,
    "fused_write": lambda D=128: 3.0 + 4.0 / D, "fused_read": lambda D=128: 3.0 + 4.0 / D,
    "golay_encode": lambda D=128: 7.0 / 3.0, "golay_decode": lambda D=128: 8.0 / 3.0,
}


def measured_hbm_peak_gbs(default: float = 6650.0):
    """(GB/s, provenance): MEASURED_PEAKS.json at the repo root (driver-written) or the profiling guide's fallback."""
    p = Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    except Exception:
        return default, "fallback"


@dataclass
class TimingStats:
    cpu_to_gpu_transfer_ns: int = 0
    gpu_to_cpu_transfer_ns: int = 0
    quantize_ns: int = 0
    encode_ns: int = 0
    decode_ns: int = 0
    dequantize_ns: int = 0
    values_processed: int = 0
    bytes_transferred: int = 0
    errors_corrected: int = 0
    errors_detected: int = 0
    # --- CUDA backend additions
    fused_write_ns: int = 0
    fused_read_ns: int = 0
    fused_attention_ns: int = 0
    inject_ns: int = 0
    bytes_moved: Dict[str, int] = field(default_factory=dict)   # algorithmic bytes per phase

    @property
    def total_codec_ns(self) -> int:
        return self.encode_ns + self.decode_ns + self.fused_write_ns + self.fused_read_ns

    @property
    def total_transfer_ns(self) -> int:
        return self.cpu_to_gpu_transfer_ns + self.gpu_to_cpu_transfer_ns

    @property
    def total_ns(self) -> int:
        return (self.total_transfer_ns + self.quantize_ns + self.encode_ns + self.decode_ns + self.dequantize_ns +
                self.fused_write_ns + self.fused_read_ns + self.fused_attention_ns + self.inject_ns)

    @property
    def encode_time_ms(self) -> float:
        return self.encode_ns / 1e6

    @property
    def decode_time_ms(self) -> float:
        return self.decode_ns / 1e6

    @property
    def throughput_mvalues_sec(self) -> float:
        return 0.0 if self.total_codec_ns == 0 else (self.values_processed / 1e6) / (self.total_codec_ns / 1e9)

    @property
    def transfer_overhead_pct(self) -> float:
        return 0.0 if self.total_ns == 0 else 100.0 * self.total_transfer_ns / self.total_ns

    def gbps(self, phase: str) -> float:
        """Algorithmic GB/s of a phase = bytes_moved[phase] / its accumulated time."""
        ns = getattr(self, TimingContext.PHASE_ATTRS[phase])
        return 0.0 if ns == 0 else self.bytes_moved.get(phase, 0) / ns

    def roofline_fraction(self, phase: str, peak_gbs: Optional[float] = None) -> float:
        peak = peak_gbs or measured_hbm_peak_gbs()[0]
        return self.gbps(phase) / peak


@dataclass
class AggregatedTimingStats:
    total_cpu_to_gpu_ns: int = 0
    total_gpu_to_cpu_ns: int = 0
    total_quantize_ns: int = 0
    total_encode_ns: int = 0
    total_decode_ns: int = 0
    total_dequantize_ns: int = 0
    total_values: int = 0
    total_bytes: int = 0
    total_errors_corrected: int = 0
    total_errors_detected: int = 0
    encode_times_ns: List[int] = field(default_factory=list)
    decode_times_ns: List[int] = field(default_factory=list)
    n_operations: int = 0
    is_cpu_bound: bool = False   # the CUDA backend has no CPU stage
    total_fused_write_ns: int = 0
    total_fused_read_ns: int = 0
    total_fused_attention_ns: int = 0

    def add(self, s: TimingStats) -> None:
        self.total_cpu_to_gpu_ns += s.cpu_to_gpu_transfer_ns
        self.total_gpu_to_cpu_ns += s.gpu_to_cpu_transfer_ns
        self.total_quantize_ns += s.quantize_ns
        self.total_encode_ns += s.encode_ns
        self.total_decode_ns += s.decode_ns
        self.total_dequantize_ns += s.dequantize_ns
        self.total_fused_write_ns += s.fused_write_ns
        self.total_fused_read_ns += s.fused_read_ns
        self.total_fused_attention_ns += s.fused_attention_ns
        self.total_values += s.values_processed
        self.total_bytes += s.bytes_transferred
        self.total_errors_corrected += s.errors_corrected
        self.total_errors_detected += s.errors_detected
        self.encode_times_ns.append(s.encode_ns + s.fused_write_ns)
        self.decode_times_ns.append(s.decode_ns + s.fused_read_ns)
        self.n_operations += 1

    def reset(self) -> None:
        self.__init__(is_cpu_bound=self.is_cpu_bound)

    @staticmethod
    def _mean_std_ms(xs):
        if not xs:
            return 0.0, 0.0
        mean = sum(xs) / len(xs)
        std = (sum((x - mean) ** 2 for x in xs) / len(xs)) ** 0.5 if len(xs) > 1 else 0.0
        return mean / 1e6, std / 1e6

    @property
    def mean_encode_ms(self):
        return self._mean_std_ms(self.encode_times_ns)[0]

    @property
    def std_encode_ms(self):
        return self._mean_std_ms(self.encode_times_ns)[1]

    @property
    def mean_decode_ms(self):
        return self._mean_std_ms(self.decode_times_ns)[0]

    @property
    def std_decode_ms(self):
        return self._mean_std_ms(self.decode_times_ns)[1]

    @property
    def throughput_mvalues_sec(self) -> float:
        ns = self.total_encode_ns + self.total_decode_ns + self.total_fused_write_ns + self.total_fused_read_ns
        return 0.0 if ns == 0 else (self.total_values / 1e6) / (ns / 1e9)

    @property
    def transfer_overhead_pct(self) -> float:
        xfer = self.total_cpu_to_gpu_ns + self.total_gpu_to_cpu_ns
        total = (xfer + self.total_quantize_ns + self.total_encode_ns + self.total_decode_ns + self.total_dequantize_ns +
                 self.total_fused_write_ns + self.total_fused_read_ns + self.total_fused_attention_ns)
        return 0.0 if total == 0 else 100.0 * xfer / total

    def to_dict(self) -> Dict:
        return {
            "n_operations": self.n_operations, "total_values": self.total_values, "is_cpu_bound": self.is_cpu_bound,
            "encode_ms_mean": self.mean_encode_ms, "encode_ms_std": self.std_encode_ms,
            "decode_ms_mean": self.mean_decode_ms, "decode_ms_std": self.std_decode_ms,
            "throughput_mvalues_sec": self.throughput_mvalues_sec, "transfer_overhead_pct": self.transfer_overhead_pct,
            "errors_corrected": self.total_errors_corrected, "errors_detected": self.total_errors_detected,
        }


class TimingContext:
    """Wall-clock phase timer (evaluation/timing.py:186-217) with the CUDA backend's extra phases."""
    PHASE_ATTRS = {
        "cpu_to_gpu": "cpu_to_gpu_transfer_ns", "gpu_to_cpu": "gpu_to_cpu_transfer_ns", "quantize": "quantize_ns",
        "encode": "encode_ns", "decode": "decode_ns", "dequantize": "dequantize_ns",
        "fused_write": "fused_write_ns", "fused_read": "fused_read_ns", "fused_attention": "fused_attention_ns",
        "inject": "inject_ns",
    }

    def __init__(self, stats: Optional[TimingStats], phase: str):
        if phase not in self.PHASE_ATTRS:
            raise ValueError(f"Unknown phase: {phase}. Valid: {list(self.PHASE_ATTRS.keys())}")
        self.stats, self.phase, self.start_time = stats, phase, 0

    def __enter__(self):
        if self.stats is not None:
            self.start_time = time.perf_counter_ns()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.stats is not None:
            attr = self.PHASE_ATTRS[self.phase]
            setattr(self.stats, attr, getattr(self.stats, attr) + time.perf_counter_ns() - self.start_time)
        return False


@contextmanager
def cuda_transfer_timer(stats: Optional[TimingStats], phase: str, bytes_moved: int = 0):
    """CUDA-event timer on the CURRENT stream (evaluation/timing.py:220-239), also valid for kernel phases;
    `bytes_moved` (algorithmic bytes) is accumulated so `stats.gbps(phase)` works."""
    if stats is None:
        yield
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    yield
    end.record()
    end.synchronize()
    attr = TimingContext.PHASE_ATTRS[phase]
    setattr(stats, attr, getattr(stats, attr) + int(start.elapsed_time(end) * 1e6))
    if bytes_moved:
        stats.bytes_moved[phase] = stats.bytes_moved.get(phase, 0) + int(bytes_moved)


cuda_phase_timer = cuda_transfer_timer


def cuda_timer(fn, warmup: int = 10, repeats: int = 100):
    """Median latency in ms over `repeats` runs after `warmup` (kv_cache/benchmark_harness.py:42-57 used the mean of one
    event pair; the median of per-run events is what BASELINE.md section 3 specifies)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(repeats):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2]


def run_warmup(func, n_iterations: int = 3, *args, **kwargs) -> None:
    for _ in range(n_iterations):
        func(*args, **kwargs)
    torch.cuda.synchronize()

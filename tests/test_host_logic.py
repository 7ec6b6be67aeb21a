"""CPU-only tests: the C-ABI library loads and exports every symbol include/eccq.h declares, the host-side mirror of
the reference API validates its arguments like the reference, and the multi-process plumbing works (gloo, world 2)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from eccq_b200 import _lib
    header = (ROOT / "include" / "eccq.h").read_text()
    declared = set(re.findall(r"\b(eccq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/eccq.h"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()                       # raises if the .so is missing: there is no fallback
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.eccq_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (eccq_[a-z0-9_]+)", out))
    assert declared <= exported


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from eccq_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.EccqError, match="no CPU / eager fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = ROOT / "quantized-kv-cache-ecc-protection_b200"
    for f in pkg.rglob("*.py"):
        assert "oracle" not in f.read_text().replace("the oracle", ""), f
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        assert "oracle/" not in f.read_text(), f


def test_host_tables_without_a_gpu():
    from eccq_b200.ecc_codecs.cuda_kernels import config as c
    assert c.SYNDROME_LUT_HAMMING84.tolist() == [-1, 4, 5, 0, 6, 1, 2, 3]
    assert ((c.HAMMING74_G.int() @ c.HAMMING74_H.int().T) % 2 == 0).all()          # G.H^T = 0 (verification.py:143-212)
    B = c.GOLAY_B_MATRIX.int()
    assert torch.equal(B, B.T) and torch.equal((B @ B) % 2, torch.eye(12, dtype=torch.int32))
    t = c.build_golay_syndrome_table()
    assert t.shape == (4096,) and int((t >= 0).sum()) == 2325 and int(t[0xFFF]) == 8390656
    assert c.get_physical_dtype("golay") == torch.int32 and c.get_codeword_bits("hamming84") == 8 and c.get_data_bits("golay") == 12
    with pytest.raises(ValueError):
        c.get_physical_dtype("bogus")


def test_config_and_wrappers_reject_bad_input_like_the_reference():
    from eccq_b200.ecc_codecs import cuda_kernels as ck
    from eccq_b200.kv_cache import ECCCacheConfig, ECCShimConfig, create_block_table
    with pytest.raises(ValueError, match="Unsupported codec"):
        ECCShimConfig(codec="adaptive")                       # tests/test_sweep_config.py:81-105
    cfg = ECCShimConfig()
    assert (cfg.codec, cfg.ber, cfg.block_size, cfg.num_blocks, cfg.inject_errors, cfg.seed, cfg.use_interpolation, cfg.backend) == \
        ("hamming84", 0.0, 16, 256, False, 42, False, "cuda")
    with pytest.raises(AssertionError, match="CUDA"):
        ck.hamming84_encode(torch.zeros(4, dtype=torch.uint8))           # CPU tensor -> AssertionError, as the reference
    with pytest.raises(AssertionError):
        ck.golay_decode(torch.zeros(4, dtype=torch.int32))
    c = ECCCacheConfig(num_heads=8, head_size=128, num_layers=2, codec="golay")     # tests/test_kv_cache.py:6-39
    assert c.dtype == torch.int32 and c.values_per_block == 2048 and c.codewords_per_block == 683
    bt = create_block_table(2, 100, 16, device="cpu")
    assert bt.shape == (2, 7) and (bt == -1).all()


def test_monte_carlo_config_maps_onto_a_cuda_sweep():
    """evaluation/experiments/monte_carlo.py:31-72 of the reference with backend="cuda"."""
    from eccq_b200.evaluation.monte_carlo import MonteCarloConfig
    from eccq_b200.evaluation.sweep import SweepConfig, get_mode_config
    mc = MonteCarloConfig(seeds=[42, 101, 997], ber_levels=[0.0, 1e-2])
    sc = mc.to_sweep_config(clean_logits=["x"])
    assert isinstance(sc, SweepConfig) and sc.backend == "cuda" and sc.seeds == [42, 101, 997] and sc.aggregate_seeds
    assert sc.cache_modes == ["int4-hamming", "int4-hamming84", "int4-hamming84-interp", "int12-golay"] and sc.clean_logits == ["x"]
    assert MonteCarloConfig().ber_levels == [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2] and len(MonteCarloConfig().seeds) == 10
    assert get_mode_config("int4-hamming84-interp") == {"codec": "hamming84", "use_interpolation": True}
    with pytest.raises(ValueError, match="Unknown cache mode"):
        get_mode_config("fp16")


def test_timing_containers():
    from eccq_b200.evaluation import AggregatedTimingStats, TimingContext, TimingStats
    s = TimingStats(values_processed=1_000_000)
    with TimingContext(s, "fused_write"):
        pass
    with TimingContext(s, "decode"):
        pass
    assert s.fused_write_ns > 0 and s.decode_ns > 0 and s.total_codec_ns == s.fused_write_ns + s.decode_ns
    s.bytes_moved["fused_write"] = 3_000_000
    assert s.gbps("fused_write") > 0 and 0 < s.roofline_fraction("fused_write", 6542.0)
    with pytest.raises(ValueError):
        TimingContext(s, "nonsense")
    agg = AggregatedTimingStats()
    agg.add(s)
    d = agg.to_dict()
    assert d["n_operations"] == 1 and d["total_values"] == 1_000_000 and d["throughput_mvalues_sec"] > 0


def test_shard_range_partitions_exactly():
    from eccq_b200.distributed import shard_range
    for n, w in ((16, 8), (8, 3), (64, 8), (5, 8)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from eccq_b200.distributed import all_reduce_counters, init_from_env, max_over_ranks, shard_range
rank, world, local = init_from_env(backend="gloo")
lo, hi = shard_range(16, rank, world)                       # C3: 16 sequences sharded by batch
vec = torch.tensor([10 * (rank + 1), hi - lo, 0, 0, 0, 0, 0, 0, 1, 2], dtype=torch.int64)
tot = all_reduce_counters(vec)
assert tot[0].item() == 10 * world * (world + 1) // 2 and tot[1].item() == 16, tot
assert max_over_ranks(float(rank)) == world - 1
print("ok", rank, flush=True)
"""


def test_two_rank_counter_allreduce_gloo(tmp_path):
    """The N>1 path on CPU: two ranks, gloo backend, rendezvous on 127.0.0.1."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)

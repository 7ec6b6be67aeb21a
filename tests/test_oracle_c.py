"""Pin the C restatement (oracle/eccq_oracle.c, the CPU baseline of bench.py) against the numpy oracle."""
import numpy as np


def test_c_port_matches_numpy_oracle(oracle):
    from oracle import c_port
    rng = np.random.default_rng(0)
    x = rng.standard_normal((200, 4, 64)).astype(np.float32)
    x[3] = 0
    x[5, 0, :8] = np.array([3.5, 2.5, 1.5, 0.5, -0.5, -1.5, -2.5, -3.5]) * (x[5, 0].__abs__().max() / 7)
    cw, sc = c_port.quantize_encode_h84(x)
    q, s = oracle.quantize_int4(x)
    assert np.array_equal(sc.view(np.int32), s.view(np.int32)) and np.array_equal(cw, oracle.h84_encode(q))
    mask = ((rng.random(cw.shape + (8,)) < 0.03) * (1 << np.arange(8))).sum(-1).astype(np.uint8)
    noisy = c_port.xor_mask(cw, mask)
    assert np.array_equal(noisy, cw ^ mask)
    nib, et, st = c_port.decode_h84(noisy)
    od, oe = oracle.h84_decode(noisy)
    assert np.array_equal(nib, od) and np.array_equal(et, oe) and st == oracle.h84_counters(oe)
    for interp in (False, True):
        out, st = c_port.decode_interp_dequant_h84(noisy, sc, interp)
        d = oracle.interpolate_double_errors(od, oe, seq_axis=0) if interp else od
        assert np.array_equal(out, oracle.dequantize_int4(d, s)) and st == oracle.h84_counters(oe)
    assert c_port.num_threads() >= 1


def test_c_port_flat_encode_and_thread_count(oracle):
    """The config-1 leg of the CPU baseline: flat encode, and the thread count set by the caller (OMP_NUM_THREADS, which
    torchrun sets to 1 for its workers, must not decide how many cores the baseline uses)."""
    from oracle import c_port
    rng = np.random.default_rng(3)
    nib = rng.integers(0, 16, 100_003, dtype=np.uint8)
    assert np.array_equal(c_port.encode_h84(nib), oracle.h84_encode(nib))
    assert c_port.set_num_threads(3) == 3
    assert np.array_equal(c_port.encode_h84(nib), oracle.h84_encode(nib))
    assert c_port.set_num_threads(1) == 1


def test_torch_cpu_codec_matches_numpy_oracle(oracle):
    """oracle/torch_cpu.py (the pure-torch CPU codec path bench.py times beside the C port) against the numpy oracle:
    exhaustive over the 256 codewords and on a random stream with flips."""
    import torch
    from oracle import torch_cpu
    allcw = np.arange(256, dtype=np.uint8)
    nib, et, (c, d) = torch_cpu.hamming84_decode(torch.from_numpy(allcw))
    onib, oet = oracle.h84_decode(allcw)
    assert np.array_equal(nib.numpy(), onib) and np.array_equal(et.numpy(), oet)
    assert (c, d) == oracle.h84_counters(oet)
    rng = np.random.default_rng(5)
    x = rng.integers(0, 16, 50_000, dtype=np.uint8)
    cw = torch_cpu.hamming84_encode(torch.from_numpy(x))
    assert np.array_equal(cw.numpy(), oracle.h84_encode(x))
    g = torch.Generator().manual_seed(1)
    noisy = torch_cpu.inject_bit_errors(cw, 0.02, 8, generator=g)
    flips = np.unpackbits((noisy ^ cw).numpy()).sum()
    assert abs(flips / (8 * x.size) - 0.02) < 0.002
    n2, e2, (c2, d2) = torch_cpu.hamming84_decode(noisy)
    on2, oe2 = oracle.h84_decode(noisy.numpy())
    assert np.array_equal(n2.numpy(), on2) and np.array_equal(e2.numpy(), oe2) and (c2, d2) == oracle.h84_counters(oe2)
    assert torch_cpu.inject_bit_errors(cw, 0.0) is cw


def test_bench_cpu_leg_ignores_omp_num_threads(monkeypatch):
    """bench.py's CPU arm sets the OpenMP thread count itself: the same cores at every N of a torchrun launch."""
    import bench
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    cores = bench.host_cores()
    assert cores == len(__import__("os").sched_getaffinity(0))
    r = bench.cpu_leg(1, 1, sample_seqs=1)
    assert r["cores"] == cores and r["value"] > 0

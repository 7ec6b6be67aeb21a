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

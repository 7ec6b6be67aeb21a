#!/usr/bin/env python
"""Times the REFERENCE's own Triton kernels (unmodified, from baseline/_ref -- see tools/install_reference.sh) next to this
repository's CUDA kernels on the same B200, with the reference's timer (`cuda_timer`, kv_cache/benchmark_harness.py:42-57:
warm-up, then one event pair around `repeat` back-to-back calls).

Sizes: C1 (GPT-2 KV shape, 18 874 368 values) for the flat codecs and the three-kernel round trip of
kv_cache/benchmark_harness.py:174-210; 2^26 values for the bandwidth-bound regime; decode attention on an MHA cache with
the C5 shard's byte count (8 sequences x 32 768 tokens x 8 heads x 128: the Triton kernel derives its strides from the QUERY
head count and is only correct for MHA).  Usage: python tools/ref_triton_bench.py [--json out.json] [--ctx N]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = ROOT / "baseline" / "_ref"


def cuda_timer(torch, func, warmup=10, repeat=100):
    for _ in range(warmup):
        func()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(repeat):
        func()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1000.0 / repeat      # us


def run(ctx=32768, attention=True):
    import torch
    if not REF.exists():
        return {"unavailable": "baseline/_ref missing (run tools/install_reference.sh where /root/reference exists)"}
    sys.path.insert(0, str(REF))
    try:
        from ecc_codecs.triton_kernels import (golay_decode, golay_encode, hamming84_decode, hamming84_encode, inject_bit_errors_triton,
                                               interpolate_double_errors)
        from kv_cache.attention_ecc import paged_attention_ecc as ref_attention
    finally:
        sys.path.remove(str(REF))
    from eccq_b200 import _lib
    from eccq_b200.ecc_codecs import cuda_kernels as ck
    from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc

    dev = torch.device("cuda", torch.cuda.current_device())
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        peak = 6650.0
    rows = []

    def add(name, n_values, bytes_per_value, ref_fn, our_fn, repeat=50):
        r = cuda_timer(torch, ref_fn, 5, repeat)
        o = cuda_timer(torch, our_fn, 5, repeat)
        rows.append({"op": name, "values": n_values, "reference_triton_us": r, "ours_us": o, "speedup": r / o,
                     "reference_gbs": n_values * bytes_per_value / r / 1e3, "ours_gbs": n_values * bytes_per_value / o / 1e3,
                     "ours_frac_of_hbm": n_values * bytes_per_value / o / 1e3 / peak})
        print(f"{name:44s} n={n_values:>11d}  reference Triton {r:10.1f} us   ours {o:9.1f} us   x{r / o:7.2f}   ours {rows[-1]['ours_gbs']:7.0f} GB/s", flush=True)

    for label, n in (("C1", 18_874_368), ("64M", 1 << 26)):
        x = torch.randint(0, 16, (n,), dtype=torch.uint8, device=dev)
        cw = hamming84_encode(x)
        assert torch.equal(cw, ck.hamming84_encode(x))
        add(f"hamming84_encode [{label}]", n, 2, lambda: hamming84_encode(x), lambda: ck.hamming84_encode(x))
        noisy = inject_bit_errors_triton(cw, 1e-2, 8, seed=42)
        assert torch.equal(noisy, ck.inject_bit_errors_triton(cw, 1e-2, 8, seed=42))     # bit-identical Philox stream
        add(f"inject_bit_errors (ber 1e-2, ref stream) [{label}]", n, 2, lambda: inject_bit_errors_triton(cw, 1e-2, 8, seed=42),
            lambda: ck.inject_bit_errors_triton(cw, 1e-2, 8, seed=42), repeat=20)
        add(f"inject_bit_errors (ber 1e-2, own stream) [{label}]", n, 2, lambda: inject_bit_errors_triton(cw, 1e-2, 8, seed=42),
            lambda: ck.inject_bit_errors_fast(cw, 1e-2, 8, seed=42), repeat=20)
        d_ref = hamming84_decode(noisy, return_error_types=True)
        d_our = ck.hamming84_decode(noisy, return_error_types=True)
        assert torch.equal(d_ref[0], d_our[0]) and torch.equal(d_ref[1], d_our[1]) and tuple(d_ref[2]) == tuple(d_our[2])
        add(f"hamming84_decode (+types, host counts) [{label}]", n, 3, lambda: hamming84_decode(noisy, return_error_types=True),
            lambda: ck.hamming84_decode(noisy, return_error_types=True))
        add(f"encode->inject->decode, 3 kernels [{label}]", n, 7,
            lambda: hamming84_decode(inject_bit_errors_triton(hamming84_encode(x), 1e-2, 8, seed=42)),
            lambda: ck.hamming84_decode(ck.inject_bit_errors_fast(ck.hamming84_encode(x), 1e-2, 8, seed=42)), repeat=20)
        ncw = n // 3
        trip = torch.randint(0, 16, (ncw, 3), dtype=torch.uint8, device=dev)
        gcw = golay_encode(trip)
        assert torch.equal(gcw, ck.golay_encode(trip))
        add(f"golay_encode [{label}]", 3 * ncw, 7 / 3, lambda: golay_encode(trip), lambda: ck.golay_encode(trip))
        gn = inject_bit_errors_triton(gcw, 1e-2, 24, seed=7)
        g_ref, g_our = golay_decode(gn), ck.golay_decode(gn)
        assert torch.equal(g_ref[0], g_our[0]) and tuple(g_ref[1]) == tuple(g_our[1])
        add(f"golay_decode (host counts) [{label}]", 3 * ncw, 7 / 3, lambda: golay_decode(gn), lambda: ck.golay_decode(gn))
        # temporal interpolation as ECCBackend.attend calls it: [ctx, Hkv, D] along the context axis (seq_dim=0) ...
        q3 = d_ref[0][: (n // 1024) * 1024].view(-1, 8, 128)
        e3 = d_ref[1][: (n // 1024) * 1024].view(-1, 8, 128)
        assert torch.equal(interpolate_double_errors(q3, e3, seq_dim=0), ck.interpolate_double_errors(q3, e3, seq_dim=0))
        add(f"interpolate_double_errors [ctx,8,128] dim 0 [{label}]", q3.numel(), 3, lambda: interpolate_double_errors(q3, e3, seq_dim=0),
            lambda: ck.interpolate_double_errors(q3, e3, seq_dim=0), repeat=20)
        # ... and along the contiguous axis of a 2-D tensor (the form the reference's unit tests use)
        q2, e2 = q3.view(-1, 1024), e3.view(-1, 1024)
        assert torch.equal(interpolate_double_errors(q2, e2), ck.interpolate_double_errors(q2, e2))
        add(f"interpolate_double_errors [rows,1024] last dim [{label}]", q2.numel(), 3, lambda: interpolate_double_errors(q2, e2),
            lambda: ck.interpolate_double_errors(q2, e2), repeat=20)
        del x, cw, noisy, d_ref, d_our, trip, gcw, gn, g_ref, g_our, q2, e2, q3, e3
        torch.cuda.empty_cache()

    if attention:
        # MHA cache with the C5 shard's bytes: 8 sequences x ctx tokens x 8 heads x 128, block 16
        Bq, H, D, BS = 8, 8, 128, 16
        nblk_seq = ctx // BS
        nblk = Bq * nblk_seq
        m = SimpleBlockManager(nblk, BS, 1, H, D, device=dev, codec="hamming84")
        m.block_table = torch.randperm(nblk, generator=torch.Generator().manual_seed(3)).to(torch.int32).view(Bq, nblk_seq).to(dev)
        P = _lib.ptr
        sp = _lib.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        step = min(4096, ctx)
        for s0 in range(0, ctx, step):
            kk = torch.randn(Bq, step, H * D, device=dev, dtype=torch.float16)
            start = torch.full((Bq,), s0, dtype=torch.int32, device=dev)
            _lib.check(_lib.load().eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), nblk_seq, P(start), Bq, step, 0,
                                                    _lib.INJECT_PHILOX_FAST, None, None, 1e-3, 1, s0, None, sp), "cache_write")
        q = torch.randn(Bq, H, D, device=dev, dtype=torch.float32)
        cl = torch.full((Bq,), ctx, dtype=torch.int32, device=dev)
        o_ref = ref_attention(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, v_scales=m.v_scales)
        o_our = paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, v_scales=m.v_scales)
        err = float((o_ref - o_our).abs().max())
        nbytes = Bq * ctx * (2 * H * D + 2 * H * 4)
        r = cuda_timer(torch, lambda: ref_attention(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, v_scales=m.v_scales), 1, 3)
        o = cuda_timer(torch, lambda: paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, v_scales=m.v_scales), 5, 50)
        rows.append({"op": f"paged_attention_ecc MHA 8x{ctx}x8x128 fp32 q", "values": Bq * ctx * H * D * 2, "reference_triton_us": r, "ours_us": o,
                     "speedup": r / o, "reference_gbs": nbytes / r / 1e3, "ours_gbs": nbytes / o / 1e3, "ours_frac_of_hbm": nbytes / o / 1e3 / peak,
                     "max_abs_diff": err})
        print(f"{rows[-1]['op']:44s} reference Triton {r:10.1f} us   ours {o:9.1f} us   x{r / o:7.2f}   ours {nbytes / o / 1e3:7.0f} GB/s   max|diff| {err:.2e}",
              flush=True)
    return {"gpu": torch.cuda.get_device_name(dev), "hbm_peak_gbs": peak, "timer": "cuda_timer (benchmark_harness.py:42-57)", "rows": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--ctx", type=int, default=32768)
    ap.add_argument("--no-attention", action="store_true")
    args = ap.parse_args()
    res = run(args.ctx, not args.no_attention)
    if args.json:
        Path(args.json).write_text(json.dumps(res, indent=1))
    if "unavailable" in res:
        print(res["unavailable"])

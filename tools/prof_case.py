#!/usr/bin/env python
"""Run ONE kernel variant of the hot path a few times (for ncu captures and quick timings).

    python tools/prof_case.py <case> [reps]
    cases: copy h84enc h84dec write0 write3 write3lo read0 read1 readint4 writeg readg attn attni attng attn4 prefill   (C3 / C5 shapes)
Prints the median CUDA-event time and algorithmic GB/s when not under a profiler."""
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from eccq_b200 import _lib  # noqa: E402
from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc  # noqa: E402

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
lib = _lib.load()
P = _lib.ptr
SP = _lib.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
ctr = _lib.new_counters(dev)
import os
H, D, BS = int(os.environ.get("PC_H", 8)), int(os.environ.get("PC_D", 128)), int(os.environ.get("PC_BS", 16))
torch.manual_seed(0)


def timed(fn, nbytes):
    import os
    import time
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if reps > 2 and "NV_COMPUTE_PROFILER_PERFWORKS_DIR" not in os.environ:   # let the clocks ramp up (idle: 120 MHz)
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            for _ in range(8):
                fn()
            torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    print(f"{case}: {ms:.4f} ms  {nbytes / ms / 1e6:.1f} GB/s", flush=True)


if case == "copy":      # box-to-box HBM reference: torch device copy, 1 R + 1 W
    x = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
    y = torch.empty_like(x)
    timed(lambda: y.copy_(x), 2 * x.numel())
elif case in ("h84enc", "h84dec"):
    from eccq_b200.ecc_codecs import cuda_kernels as ck
    N = 1 << 28
    x = torch.randint(0, 16, (N,), dtype=torch.uint8, device=dev)
    cw = ck.inject_bit_errors_fast(ck.hamming84_encode(x), 1e-2, 8, seed=1)
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    if case == "h84enc":
        timed(lambda: lib.eccq_h84_encode(P(x), P(o1), N, SP), 2 * N)
    else:
        timed(lambda: lib.eccq_h84_decode(P(cw), P(o1), P(o2), P(ctr), N, 0, SP), 3 * N)
elif case.startswith(("write", "read")):
    B, S = 16, 8192
    codec = "golay" if "g" in case else "int4" if "int4" in case else "hamming84"
    nblk = B * S // BS
    k = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    v = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    cl = torch.full((B,), S, dtype=torch.int32, device=dev)
    vals = 2 * B * S * H * D
    m = SimpleBlockManager(nblk, BS, 1, H, D, device=dev, codec=codec)
    m.block_table = torch.randperm(nblk).to(torch.int32).view(B, S // BS).to(dev)
    desc = m.desc()
    cwb = 4 * ((D + 2) // 3) / D if codec == "golay" else 1.0

    def write(inj, ber):
        return lib.eccq_cache_write(desc, P(k), P(v), _lib.ECCQ_F16, P(m.block_table), S // BS, None, B, S, 0, inj, None, None, ber, 42, 0, P(ctr), SP)

    if case.startswith("write"):
        inj, ber = (_lib.INJECT_NONE, 0.0) if "0" in case else (_lib.INJECT_PHILOX_FAST, 1e-4 if case.endswith("lo") else 1e-2)
        timed(lambda: write(inj, ber), vals * (2 + cwb + 4 / D))
    else:
        write(_lib.INJECT_PHILOX_FAST, 1e-2)
        ko = torch.empty(B, S, H, D, device=dev, dtype=torch.float16)
        vo = torch.empty_like(ko)
        interp = 1 if case == "read1" else 0
        timed(lambda: lib.eccq_cache_read(desc, P(m.block_table), S // BS, P(cl), B, S, 0, interp, _lib.ECCQ_F16, P(ko), P(vo), None, None, P(ctr), SP),
              vals * (cwb + 4 / D + 2))
elif case == "prefill":     # ECC-fused causal prefill, GPT-2 small geometry (config 2 with 1 024-token texts): 12 heads x 64, MHA
    from eccq_b200.kv_cache import paged_prefill_attention_ecc
    Hh, Dh, Sq = 12, 64, 1024
    nblk = Sq // BS
    m = SimpleBlockManager(nblk, BS, 1, Hh, Dh, device=dev, codec="hamming84")
    m.block_table = torch.randperm(nblk).to(torch.int32).view(1, nblk).to(dev)
    kk = torch.randn(1, Sq, Hh * Dh, device=dev, dtype=torch.float16)
    lib.eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), nblk, None, 1, Sq, 0, _lib.INJECT_PHILOX_FAST,
                         None, None, 1e-3, 1, 0, None, SP)
    q = torch.randn(1, Hh, Sq, Dh, device=dev, dtype=torch.float32)
    cl = torch.full((1,), Sq, dtype=torch.int32, device=dev)
    timed(lambda: paged_prefill_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, m.v_scales, 0, BS,
                                              use_interpolation=True), 2 * Sq * Hh * Dh)
else:
    Bq, ctx, HQ = 8, 32768, 32
    codec = "golay" if case == "attng" else "int4" if case == "attn4" else "hamming84"
    nblk = Bq * ctx // BS
    m = SimpleBlockManager(nblk, BS, 1, H, D, device=dev, codec=codec)
    m.block_table = torch.randperm(nblk).to(torch.int32).view(Bq, ctx // BS).to(dev)
    for s0 in range(0, ctx, 4096):
        kk = torch.randn(Bq, 4096, H * D, device=dev, dtype=torch.float16)
        st = torch.full((Bq,), s0, dtype=torch.int32, device=dev)
        lib.eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), ctx // BS, P(st), Bq, 4096, 0, _lib.INJECT_PHILOX_FAST,
                             None, None, 1e-3, 1, s0, None, SP)
    q = torch.randn(Bq, HQ, D, device=dev, dtype=torch.float16)
    cl = torch.full((Bq,), ctx, dtype=torch.int32, device=dev)
    cwb = 4 * ((D + 2) // 3) if codec == "golay" else D
    timed(lambda: paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, codec=codec, v_scales=m.v_scales,
                                      use_interpolation=(case == "attni")), Bq * ctx * (2 * H * cwb + 2 * H * 4))

#!/usr/bin/env python
"""Decode-attention A/B on the C5 per-GPU shard (8 sequences x 32 768 context, Hq 32 / Hkv 8 / D 128, block 16).
Usage: python tools/attn_bench.py [codec] [interp 0|1] [ctx] [q dtype f16|f32|bf16]
The kernel variant is chosen by the environment (read once per process): ECCQ_ATTN_TC=0 -> mma.sync kernel,
ECCQ_ATTN_TC_NDEC=8|16, ECCQ_ATTN_TC_CHUNK=<tokens>.  Prints one line: ms per call, GB/s of algorithmic bytes, fraction of
the measured HBM peak, max |err| against the fused-read + torch softmax route on sequence 0."""
import json
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from eccq_b200 import _lib  # noqa: E402
from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc, reference_attention_ecc  # noqa: E402

codec = sys.argv[1] if len(sys.argv) > 1 else "hamming84"
interp = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
qdt = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[4] if len(sys.argv) > 4 else "f16"]
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda", 0)
lib = _lib.load()
P = _lib.ptr
Bq, HQ, HKV, D, BS = 8, 32, 8, 128, 16
nblk_seq = ctx // BS
nblk = Bq * nblk_seq
m = SimpleBlockManager(nblk, BS, 1, HKV, D, device=dev, codec=codec)
m.block_table = torch.randperm(nblk, generator=torch.Generator().manual_seed(3)).to(torch.int32).view(Bq, nblk_seq).to(dev)
sp = _lib.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
chunk = min(4096, ctx)
torch.manual_seed(0)
for s0 in range(0, ctx, chunk):
    kk = torch.randn(Bq, chunk, HKV * D, device=dev, dtype=torch.float16)
    start = torch.full((Bq,), s0, dtype=torch.int32, device=dev)
    _lib.check(lib.eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), nblk_seq, P(start), Bq, chunk, 0,
                                    _lib.INJECT_PHILOX_FAST, None, None, float(os.environ.get("ATTN_BENCH_BER", "1e-3")), 1, s0, None, sp), "cache_write")
q = torch.randn(Bq, HQ, D, device=dev, dtype=qdt)
cl = torch.full((Bq,), ctx, dtype=torch.int32, device=dev)


def run():
    return paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, codec=codec, v_scales=m.v_scales,
                               use_interpolation=interp)


out = run()
torch.cuda.synchronize()
err = None
if not interp and codec != "golay":
    ks = m.k_scales
    # reference route needs one scale tensor for K and V: compare with v_scales = k_scales on both sides
    o2 = paged_attention_ecc(q[:1], m.k_cache, m.v_cache, m.block_table[:1], cl[:1], ks, 0, BS, codec=codec)
    ref = reference_attention_ecc(q[:1].float(), m.k_cache, m.v_cache, m.block_table[:1], cl[:1], ks, 0, BS, D, codec=codec)
    err = float((o2.float() - ref).abs().max())
ts = []
for _ in range(5):
    run()
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ms = statistics.median(ts)
cwb = 4 * ((D + 2) // 3) / D if codec == "golay" else 1.0
nbytes = Bq * ctx * (2 * HKV * D * cwb + 2 * HKV * 4)
env = {k: v for k, v in os.environ.items() if k.startswith("ECCQ_") and k != "ECCQ_LIB"}
lib_name = os.path.basename(os.environ.get("ECCQ_LIB", "main"))
print(f"[{lib_name}] {ms:.4f} ms (min {min(ts):.4f})  {nbytes / ms / 1e6:.0f} GB/s  {nbytes / ms / 1e6 / PEAK:.3f} of {PEAK:.0f}   attn {codec} interp={int(interp)} "
      f"ctx={ctx} q={qdt} env={env} max|err| vs fused-read route: {err}", flush=True)

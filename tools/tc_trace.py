#!/usr/bin/env python
"""Pipeline trace of the tcgen05 decode-attention kernel (library built with -DECCQ_TC_TRACE, selected with ECCQ_LIB):
runs one C5-shard call and prints, for CTA 0, the SM clock of every hand-off per tile relative to the tile's load issue."""
import ctypes
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from eccq_b200 import _lib  # noqa: E402
from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
P = _lib.ptr
Bq, HQ, HKV, D, BS, ctx = 8, 32, 8, 128, 16, 32768
nblk_seq = ctx // BS
nblk = Bq * nblk_seq
m = SimpleBlockManager(nblk, BS, 1, HKV, D, device=dev, codec="hamming84")
m.block_table = torch.randperm(nblk, generator=torch.Generator().manual_seed(3)).to(torch.int32).view(Bq, nblk_seq).to(dev)
INTERP = os.environ.get("TC_TRACE_INTERP", "0") == "1"
sp = _lib.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for s0 in range(0, ctx, 4096):   # valid codewords with flips at BER 1e-3, as in bench.py / attn_bench.py
    kk = torch.randn(Bq, 4096, HKV * D, device=dev, dtype=torch.float16)
    start = torch.full((Bq,), s0, dtype=torch.int32, device=dev)
    _lib.check(lib.eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), nblk_seq, P(start), Bq, 4096, 0,
                                    _lib.INJECT_PHILOX_FAST, None, None, 1e-3, 1, s0, None, sp), "cache_write")
q = torch.randn(Bq, HQ, D, device=dev, dtype=torch.float16)
cl = torch.full((Bq,), ctx, dtype=torch.int32, device=dev)
for _ in range(3):
    paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, BS, v_scales=m.v_scales, use_interpolation=INTERP)
torch.cuda.synchronize()
raw = ctypes.CDLL(str(_lib.LIB_PATH))
buf = np.zeros(512 * 16, dtype=np.int64)
rc = raw.eccq_tc_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
t = buf.reshape(512, 16)
names = ["loadK", "decstart", "decK", "decV", "Sissue", "smstart", "smW", "PVissue", "Oready", "Odone", "SwaitK", "loadV"]
n_tiles = int((t[:, 1] != 0).sum())
base = t[0, 0]
print("tiles traced:", n_tiles)
print("tile " + " ".join(f"{x:>9s}" for x in names) + "   (cycles since the first load issue)")
for n in range(min(n_tiles, int(sys.argv[1]) if len(sys.argv) > 1 else 48)):
    print(f"{n:4d} " + " ".join(f"{int(t[n, e] - base):9d}" for e in range(12)))
d = np.diff(t[8:n_tiles, 3])
print("steady-state tile period (decV to decV): mean %.0f cycles, min %d max %d" % (d.mean(), d.min(), d.max()))
for a, b, label in ((0, 1, "load issue -> decode start (load latency + queueing)"), (1, 2, "decode K"), (2, 3, "decode V"), (2, 4, "decK -> S issue"),
                    (4, 5, "S issue -> softmax start"), (5, 6, "softmax A"), (6, 7, "W ready -> PV issue"), (7, 8, "PV issue -> O ready"),
                    (8, 9, "O drain"), (3, 8, "decV -> O ready (stage hold after decode)")):
    x = t[8:n_tiles, b] - t[8:n_tiles, a]
    print(f"  {label:55s} mean {x.mean():8.0f}  min {x.min():7d}  max {x.max():7d}")

import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import eccq_oracle as oracle
from tests.test_cuda_attention_tc import build_cache, dev
from eccq_b200.kv_cache import paged_attention_ecc
D, bs = 128, 16
lens = [9000, 4133, 2048]
for Hq, Hkv in ((8, 2),):
    oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=Hq + Hkv, ber=0.01)
    for qs in (0.5, 1.0, 2.0):
        q = np.random.default_rng(2).standard_normal((len(lens), Hq, D)).astype(np.float32) * qs
        ref, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, precision=np.float64)
        ref32, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
        out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs, v_scales=dev(oc.v_scales)).cpu().numpy()
        e = np.abs(out - ref)
        print(os.environ.get("ECCQ_ATTN_TC", "1"), "qscale", qs, "err vs f64 per seq", [float(e[b].max()) for b in range(3)], "oracle32 vs f64", float(np.abs(ref32 - ref).max()), "|ref|max", float(np.abs(ref).max()))

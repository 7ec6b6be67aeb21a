"""The Blackwell-native decode-attention kernel (tcgen05 / TMEM / TMA gather4, csrc/eccq_attention_tc.cu) against the CPU
oracle at sizes and geometries the small tests do not reach: every block size it accepts, contexts long enough for many
128-token tiles / several splits / several items per CTA, holes in the block table, decode-error counters, and -- in fresh
processes -- the mma.sync kernel it replaced (ECCQ_ATTN_TC=0) and Golay decode attention as the first library call."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
TOL_F32 = 1e-4     # the reference's own bar for fp32 queries (tests/test_paged_attention_ecc.py:115-116)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_cache(oracle, codec, Hkv, D, bs, lens, seed, ber=0.02, holes=()):
    """Oracle cache with a random block placement, flip masks at `ber`; `holes` = (sequence, logical block) pairs unmapped after the write."""
    rng = np.random.default_rng(seed)
    nblk = sum(-(-max(n, 1) // bs) for n in lens) + 3
    mb = max(-(-max(n, 1) // bs) for n in lens) + 2
    oc = oracle.PagedCache(nblk, bs, 2, Hkv, D, codec)
    perm = rng.permutation(nblk)
    bt = np.full((len(lens), mb), -1, dtype=np.int32)
    p = 0
    nbits = oracle.CODEC_BITS[codec]
    for b, n in enumerate(lens):
        nb_ = -(-n // bs)
        bt[b, :nb_] = perm[p:p + nb_]
        p += nb_
        if n:
            k = rng.standard_normal((1, n, Hkv * D)).astype(np.float32) * rng.uniform(0.2, 3.0, (1, n, 1)).astype(np.float32)
            v = rng.standard_normal((1, n, Hkv * D)).astype(np.float32) * rng.uniform(0.2, 3.0, (1, n, 1)).astype(np.float32)
            fl = rng.random((1, n, Hkv, oc.cw, nbits)) < ber
            m1 = (fl * (1 << np.arange(nbits))).sum(-1).astype(oc.k_cache.dtype)
            oracle.cache_write(oc, k, v, 1, bt[b], k_mask=m1, v_mask=np.roll(m1, 5, axis=1))
    for b, lb in holes:
        bt[b, lb] = -1
    return oc, bt


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [16, 32, 64, 128])
@pytest.mark.parametrize("codec", ["hamming84", "hamming74", "int4"])
def test_block_sizes_and_counters(oracle, codec, bs):
    from eccq_b200.kv_cache import paged_attention_ecc
    Hq, Hkv, D = 8, 2, 128
    lens = [700, 129, 5, 0, 256]
    oc, bt = build_cache(oracle, codec, Hkv, D, bs, lens, seed=bs + len(codec))
    q = np.random.default_rng(1).standard_normal((len(lens), Hq, D)).astype(np.float32)
    ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                              codec=codec, v_scales=dev(oc.v_scales), counters=ctr)
    assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
    assert (int(ctr[0]), int(ctr[1])) == (corr, det)
    assert not out[3].any()


@pytest.mark.gpu
def test_long_ragged_contexts_many_tiles(oracle):
    """Contexts of 9 000 / 4 133 / 2 048 tokens: several 2 048-token splits per sequence, 70 tiles, more items than one CTA
    of a small grid would see in the small tests, a ragged last tile; GQA group 4 and group 8 (two passes)."""
    from eccq_b200.kv_cache import paged_attention_ecc
    D, bs = 128, 16
    lens = [9000, 4133, 2048]
    for Hq, Hkv in ((8, 2), (8, 1)):
        oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=Hq + Hkv, ber=0.01)
        q = np.random.default_rng(2).standard_normal((len(lens), Hq, D)).astype(np.float32) * 2.0
        # yardstick in float64: over 9 000 tokens the order of the reference's own fp32 sums is worth ~1e-4
        ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, precision=np.float64)
        ref32, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        args = (dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs)
        out = paged_attention_ecc(dev(q), *args, v_scales=dev(oc.v_scales), counters=ctr)
        err, err32 = np.abs(out.cpu().numpy() - ref).max(), np.abs(ref32 - ref).max()
        assert err < 1e-5 and err < err32, (err, err32)    # closer to the exact result than the fp32 formula itself
        assert np.abs(out.cpu().numpy() - ref32).max() < 2 * TOL_F32
        assert (int(ctr[0]), int(ctr[1])) == (corr, det)
        ref_k, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, v_use_k_scales=True, precision=np.float64)
        out_k = paged_attention_ecc(dev(q), *args)                      # v_scales defaults to k_scales (attention_ecc.py:666-668)
        assert np.abs(out_k.cpu().numpy() - ref_k).max() < TOL_F32
        qh = dev(q).half()
        ref_h, _, _ = oracle.paged_decode_attention(qh.float().cpu().numpy(), oc, bt, np.array(lens), 1, precision=np.float64)
        out_h = paged_attention_ecc(qh, *args, v_scales=dev(oc.v_scales))
        # fp16 output: one rounding to half (2^-11 relative) on top of the fp32 result
        assert np.abs(out_h.float().cpu().numpy() - ref_h).max() < 6e-4 * np.abs(ref_h).max() + 1e-4


@pytest.mark.gpu
def test_extreme_score_ranges(oracle):
    """Large queries: the running maximum moves by far more than the lazy-rescaling slack between tiles, single tokens dominate
    the softmax, and the fixed-point weights span their whole range."""
    from eccq_b200.kv_cache import paged_attention_ecc
    Hq, Hkv, D, bs = 4, 1, 128, 16
    lens = [3000, 640]
    oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=11, ber=0.0)
    rng = np.random.default_rng(3)
    for qscale in (0.01, 30.0, 300.0):
        q = rng.standard_normal((len(lens), Hq, D)).astype(np.float32) * qscale
        ref, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, precision=np.float64)
        out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                                  v_scales=dev(oc.v_scales))
        assert np.isfinite(out.cpu().numpy()).all()
        # |score| reaches ~1e3 at the largest scale: one fp32 rounding of the score is already ~1e-4 relative in the weights
        assert np.abs(out.cpu().numpy() - ref).max() < (1e-4 if qscale < 100 else 2e-3) * max(1.0, np.abs(ref).max()), qscale


@pytest.mark.gpu
def test_unmapped_blocks_inside_the_context(oracle):
    """A block-table hole inside a context: the reference's gather skips the block (ecc_shim.py:931-977, `continue`), i.e. its
    tokens take no part in the softmax; scale rows of the hole are NaN to prove they are never read into the result."""
    from eccq_b200.kv_cache import paged_attention_ecc
    Hq, Hkv, D, bs = 8, 2, 128, 16
    lens = [500, 300]
    oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=4, holes=((0, 3), (0, 17), (1, 0)))
    q = np.random.default_rng(5).standard_normal((2, Hq, D)).astype(np.float32)
    ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
    ks, vs = dev(oc.k_scales), dev(oc.v_scales)
    spare = [b for b in range(oc.k_cache.shape[0]) if b not in set(bt[bt >= 0].tolist())]
    ks[spare] = float("nan")
    vs[spare] = float("nan")
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), ks, 1, bs, v_scales=vs,
                              counters=ctr)
    assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
    assert (int(ctr[0]), int(ctr[1])) == (corr, det)


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [16, 32, 64, 128])
@pytest.mark.parametrize("ber", [0.02, 0.08])
def test_single_pass_interpolation(oracle, bs, ber):
    """Hamming(8,4) + temporal interpolation inside the tile pipeline (INTERP variant) against the oracle: double errors in
    adjacent rows (no cascade), in the first / last row of a tile (neighbour in the previous / next tile), at the first and the
    last token of a context (clamped neighbour), ragged last tiles, counters incl. `interpolated`.  BER 0.08 flags nearly
    every row of every tile (2 % of the values are double errors)."""
    from eccq_b200.kv_cache import paged_attention_ecc
    Hq, Hkv, D = 8, 2, 128
    lens = [700, 129, 128, 1, 0, 257, 2]
    oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=3 * bs + int(100 * ber), ber=ber)
    q = np.random.default_rng(1).standard_normal((len(lens), Hq, D)).astype(np.float32)
    ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, use_interpolation=True)
    ref0, _, _ = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, use_interpolation=False)
    assert np.abs(ref - ref0).max() > 1e-3       # interpolation matters on this input
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                              codec="hamming84", v_scales=dev(oc.v_scales), use_interpolation=True, counters=ctr)
    assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
    assert (int(ctr[0]), int(ctr[1]), int(ctr[5])) == (corr, det, det)
    out2 = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                               codec="hamming84", v_scales=dev(oc.v_scales), use_interpolation=True)      # variant without counters
    assert torch.equal(out, out2)
    # a caller that knows its BER is >= 5e-3 is routed to the two-pass path (faster there): same corrected nibbles, same result
    out3 = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                               codec="hamming84", v_scales=dev(oc.v_scales), use_interpolation=True, ber_hint=ber)
    assert np.abs(out3.cpu().numpy() - ref).max() < TOL_F32


@pytest.mark.gpu
def test_single_pass_interpolation_across_splits(oracle):
    """Contexts longer than one 2 048-token split: the neighbour of a split's first / last row lives in another work item
    (fetched from the cache by the issuer warps); several items per CTA; GQA group 8 = two passes over the same tiles."""
    from eccq_b200.kv_cache import paged_attention_ecc
    D, bs = 128, 16
    lens = [6200, 4096, 2049]
    for Hq, Hkv in ((8, 2), (8, 1)):
        oc, bt = build_cache(oracle, "hamming84", Hkv, D, bs, lens, seed=Hq + Hkv, ber=0.03)
        # double errors exactly at the split boundaries of sequence 0 (tokens 2047 | 2048 and 4095 | 4096), every dim of head 0
        for tok in (2047, 2048, 4095, 4096):
            blk, slot = bt[0, tok // bs], tok % bs
            oc.k_cache[blk, 1, 0, slot * D:(slot + 1) * D] ^= 0x11
            oc.v_cache[blk, 1, 0, slot * D:(slot + 1) * D] ^= 0x06
        q = np.random.default_rng(2).standard_normal((len(lens), Hq, D)).astype(np.float32)
        ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, use_interpolation=True, precision=np.float64)
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                                  v_scales=dev(oc.v_scales), use_interpolation=True, counters=ctr)
        assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
        assert (int(ctr[0]), int(ctr[1]), int(ctr[5])) == (corr, det, det)


@pytest.mark.gpu
def test_single_pass_interpolation_matches_two_pass_route():
    """The single-pass kernel against the round-1 route (streaming read into corrected INT4 pages, then the tile kernel) at the
    C5 geometry with 4 096-token contexts: same corrected nibbles -> outputs agree to fp32 rounding, counters identical."""
    code = r"""
import os, sys, torch
sys.path.insert(0, %r)
from eccq_b200 import _lib
from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc
two_pass = os.environ.get("ECCQ_ATTN_INTERP_TWO_PASS") == "1"
torch.manual_seed(0)
B, ctx, Hq, Hkv, D, bs = 4, 4096, 32, 8, 128, 16
m = SimpleBlockManager(B * ctx // bs, bs, 1, Hkv, D, device="cuda", codec="hamming84")
m.block_table = torch.randperm(B * ctx // bs, generator=torch.Generator().manual_seed(3)).to(torch.int32).view(B, ctx // bs).cuda()
lib, P = _lib.load(), _lib.ptr
sp = _lib.c_void_p(torch.cuda.current_stream().cuda_stream)
kk = torch.randn(B, ctx, Hkv * D, device="cuda", dtype=torch.float16)
_lib.check(lib.eccq_cache_write(m.desc(), P(kk), P(kk.flip(1)), _lib.ECCQ_F16, P(m.block_table), ctx // bs, None, B, ctx, 0,
                                _lib.INJECT_PHILOX_FAST, None, None, 2e-2, 7, 0, None, sp), "cache_write")
q = torch.randn(B, Hq, D, device="cuda")
cl = torch.tensor([4096, 4000, 2100, 77], dtype=torch.int32, device="cuda")
ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
out = paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, bs, v_scales=m.v_scales, use_interpolation=True, counters=ctr)
torch.save((out.cpu(), ctr.cpu()), sys.argv[1])
""" % str(ROOT)
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for two_pass in ("0", "1"):
            path = os.path.join(tmp, f"o{two_pass}.pt")
            env = dict(os.environ, ECCQ_ATTN_INTERP_TWO_PASS=two_pass)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            outs.append(torch.load(path))
    (o1, c1), (o2, c2) = outs
    assert (o1 - o2).abs().max() < 2e-5
    assert c1.tolist() == c2.tolist() and c1[1] > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [16, 32, 64, 128])
def test_golay_single_pass_block_sizes_and_counters(oracle, bs):
    """Golay(24,12) decode attention in ONE pass of the tcgen05 kernel (the decode warps read the 43-codeword rows from global
    memory, syndrome table in shared memory): every block size, ragged contexts, holes, corrected-bit / uncorrectable counters."""
    from eccq_b200.kv_cache import paged_attention_ecc
    Hq, Hkv, D = 8, 2, 128
    lens = [700, 129, 5, 0, 256]
    oc, bt = build_cache(oracle, "golay", Hkv, D, bs, lens, seed=bs + 1, ber=0.03, holes=((0, 1),))
    q = np.random.default_rng(1).standard_normal((len(lens), Hq, D)).astype(np.float32)
    ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
    assert det > 0                                  # BER 0.03: some words carry four or more flips
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                              codec="golay", v_scales=dev(oc.v_scales), counters=ctr)
    assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
    assert (int(ctr[0]), int(ctr[1])) == (corr, det)
    assert not out[3].any()


@pytest.mark.gpu
def test_golay_single_pass_long_contexts(oracle):
    """Several 2 048-token splits and items per CTA, GQA group 8 (two passes over the same tiles: counted once)."""
    from eccq_b200.kv_cache import paged_attention_ecc
    D, bs = 128, 16
    lens = [4300, 2049, 2048]
    for Hq, Hkv in ((8, 2), (8, 1)):
        oc, bt = build_cache(oracle, "golay", Hkv, D, bs, lens, seed=Hq + Hkv, ber=0.01)
        q = np.random.default_rng(2).standard_normal((len(lens), Hq, D)).astype(np.float32)
        ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1, precision=np.float64)
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs,
                                  codec="golay", v_scales=dev(oc.v_scales), counters=ctr)
        assert np.abs(out.cpu().numpy() - ref).max() < TOL_F32
        assert (int(ctr[0]), int(ctr[1])) == (corr, det)


_CHILD = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from oracle import eccq_oracle as oracle
from tests.test_cuda_attention_tc import build_cache, dev
from eccq_b200.kv_cache import paged_attention_ecc
codec = {codec!r}
D, bs, lens = 128, 16, [600, 77]
oc, bt = build_cache(oracle, codec, 2, D, bs, lens, seed=9, ber=0.01)
q = np.random.default_rng(6).standard_normal((2, 8, D)).astype(np.float32)
ref, corr, det = oracle.paged_decode_attention(q, oc, bt, np.array(lens), 1)
ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
out = paged_attention_ecc(dev(q), dev(oc.k_cache), dev(oc.v_cache), dev(bt), dev(np.array(lens, np.int32)), dev(oc.k_scales), 1, bs, codec=codec,
                          v_scales=dev(oc.v_scales), counters=ctr)          # the FIRST library call of this process
err = float(np.abs(out.cpu().numpy() - ref).max())
assert err < 1e-4, err
assert (int(ctr[0]), int(ctr[1])) == (corr, det), (ctr.tolist(), corr, det)
print("child ok", err)
"""


def _run_child(codec, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT), codec=codec)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_golay_decode_attention_as_first_library_call():
    """eccq_cache_correct_pages did not upload the Golay tables; as the first call of a process every codeword
    decoded as clean (zero syndrome table) and the counters stayed 0."""
    _run_child("golay", {})


@pytest.mark.gpu
def test_mma_sync_kernel_still_matches_the_oracle():
    """ECCQ_ATTN_TC=0 routes head_dim 128 through the mma.sync tile kernel (the round-1 path, still used for head_dim 64)."""
    _run_child("hamming84", {"ECCQ_ATTN_TC": "0"})


def test_library_contains_blackwell_tensor_core_and_tma_code():
    """SASS evidence that the decode-attention kernel is Blackwell-native: tcgen05.mma (UTCIMMA), tcgen05.ld (LDTM), TMA
    tile::gather4 (UTMALDG) -- checked on the built object without a GPU."""
    import shutil
    obj = ROOT / "quantized-kv-cache-ecc-protection_b200" / "lib" / "obj" / "eccq_attention_tc.o"
    if shutil.which("cuobjdump") is None or not obj.exists():
        pytest.skip("cuobjdump or the object file is not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCIMMA", "LDTM", "UTMALDG.2D.GATHER4", "UTCBAR", "SYNCS.PHASECHK"):
        assert mnemonic in sass, mnemonic

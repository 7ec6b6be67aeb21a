"""Fused cache write + fused read at production chunk sizes against the oracle: 2 sequences x 4 096 tokens x 8 heads x 128
with flip masks at BER 1e-2 and temporal interpolation -- a persistent warp of the streaming read crosses many chunk
boundaries with real double errors on both sides (the large cases were only counted, never value-checked)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("codec,interp", [("hamming84", True), ("hamming84", False), ("golay", False)])
def test_write_read_4k_tokens_vs_oracle(oracle, codec, interp):
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager, read_cache
    B, S, H, D, bs = 2, 4096, 8, 128, 16
    rng = np.random.default_rng(17)
    nblk = B * S // bs
    k = (rng.standard_normal((B, S, H * D)) * rng.uniform(0.1, 4.0, (B, S, 1))).astype(np.float32)
    v = (rng.standard_normal((B, S, H * D)) * rng.uniform(0.1, 4.0, (B, S, 1))).astype(np.float32)
    k[0, 100] = 0.0                                           # an all-zero row: scale 1.0
    kh, vh = torch.from_numpy(k).half(), torch.from_numpy(v).half()
    oc = oracle.PagedCache(nblk, bs, 1, H, D, codec)
    nbits = oracle.CODEC_BITS[codec]
    fl = rng.random((B, S, H, oc.cw, nbits)) < 1e-2
    km = (fl * (1 << np.arange(nbits, dtype=np.int64))).sum(-1).astype(oc.k_cache.dtype)
    vm = np.roll(km, 7, axis=1)
    perm = rng.permutation(nblk).astype(np.int32)
    bt = perm.reshape(B, S // bs)
    for b in range(B):
        oracle.cache_write(oc, kh[b:b + 1].float().numpy(), vh[b:b + 1].float().numpy(), 0, bt[b], k_mask=km[b:b + 1], v_mask=vm[b:b + 1])
    cfg = ECCShimConfig(codec=codec, block_size=bs, num_blocks=nblk, use_interpolation=interp, backend="cuda", append=True)
    mgr = SimpleBlockManager(nblk, bs, 1, H, D, device="cuda", codec=codec)
    mgr.free_blocks = [int(x) for x in perm]
    be = ECCBackend(mgr, cfg, H)
    be.set_flip_masks(torch.from_numpy(km).cuda(), torch.from_numpy(vm).cuda())
    be.write(kh.cuda(), vh.cuda(), 0)
    assert np.array_equal(mgr.block_table[:B, :S // bs].cpu().numpy(), bt)
    assert np.array_equal(mgr.k_cache.cpu().numpy(), oc.k_cache) and np.array_equal(mgr.v_cache.cpu().numpy(), oc.v_cache)
    assert np.array_equal(mgr.k_scales.cpu().numpy(), oc.k_scales) and np.array_equal(mgr.v_scales.cpu().numpy(), oc.v_scales)
    cl = torch.full((B,), S, dtype=torch.int32, device="cuda")
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    kf, vf, nk, nv = read_cache(mgr.k_cache, mgr.v_cache, mgr.k_scales, mgr.v_scales, mgr.block_table[:B], cl, 0, bs, D, codec,
                                use_interpolation=interp, counters=ctr, return_nibbles=True)
    corr = det = 0
    for b in range(B):
        okf, ovf, c, d = oracle.cache_read(oc, 0, bt[b], S, use_interpolation=interp)
        assert np.array_equal(kf[b].cpu().numpy(), okf) and np.array_equal(vf[b].cpu().numpy(), ovf), (codec, interp, b)
        kc, vc, _, _ = oracle.cache_gather(oc, 0, bt[b], S)
        okn, _, _ = oracle.decode_rows(kc, codec, D, interp)
        assert np.array_equal(nk[b].cpu().numpy(), okn)
        corr += c
        det += d
    assert (int(ctr[0]), int(ctr[1])) == (corr, det)
    assert det > 50 or codec == "golay"      # the case really contains double errors on many pages


def test_chunked_prefill_equals_one_shot_prefill():
    """append mode with a multi-token chunk after existing context needs a bottom-right aligned causal mask (and a
    key-padding mask for ragged batches); SDPA's is_causal is top-left aligned."""
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    Hq, Hkv, D, bs = 8, 2, 64, 16
    torch.manual_seed(3)
    B, S = 2, 96
    k = torch.randn(B, S, Hkv * D, device="cuda")
    v = torch.randn(B, S, Hkv * D, device="cuda")
    q = torch.randn(B, Hq, S, D, device="cuda")

    def backend():
        cfg = ECCShimConfig(codec="hamming84", block_size=bs, num_blocks=64, backend="cuda", append=True)
        mgr = SimpleBlockManager(64, bs, 1, Hkv, D, device="cuda", codec="hamming84")
        return ECCBackend(mgr, cfg, Hq)

    one = backend()
    one.write(k, v, 0)
    ref = one.attend(q, 0)
    chunked = backend()
    outs = []
    for lo, hi in ((0, 40), (40, 41), (41, 96)):
        chunked.write(k[:, lo:hi], v[:, lo:hi], 0)
        outs.append(chunked.attend(q[:, :, lo:hi], 0))
    got = torch.cat(outs, dim=2)
    assert (got - ref).abs().max() < 2e-5
    # ragged batch: the second sequence stops early; its queries must not see the zero padding of the longer one
    ragged = backend()
    ragged.write(k[:1, :80], v[:1, :80], 0, seq_id=0)
    ragged.write(k[1:, :48], v[1:, :48], 0, seq_id=1)
    # one more chunk for both (16 tokens): lens become 96 / 64
    ragged.write(torch.cat([k[:1, 80:96], k[1:, 48:64]]), torch.cat([v[:1, 80:96], v[1:, 48:64]]), 0)
    out = ragged.attend(torch.cat([q[:1, :, 80:96], q[1:, :, 48:64]]), 0)
    assert (out[0] - ref[0, :, 80:96]).abs().max() < 2e-5
    assert (out[1] - ref[1, :, 48:64]).abs().max() < 2e-5

"""Parity of the production-shape fast paths added in round 1 (tiled fused write, streaming fused read, Golay page read)
on the inputs that stress their index arithmetic: appends at unaligned positions, several sequences per call, 12 kv-heads
(GPT-2), 16- and 32-token blocks, ragged contexts.  The checker is the CPU oracle (oracle/eccq_oracle.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _mgr(codec, D, bs, H, nblk, L=2, append=True, interp=False):
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    cfg = ECCShimConfig(codec=codec, block_size=bs, num_blocks=nblk, use_interpolation=interp, append=append)
    mgr = SimpleBlockManager(nblk, bs, L, H, D, device="cuda", codec=codec)
    return ECCBackend(mgr, cfg, H), mgr


@pytest.mark.parametrize("codec", ["hamming84", "hamming74", "int4", "golay"])
@pytest.mark.parametrize("D,bs,H", [(128, 16, 8), (64, 16, 12), (64, 32, 3), (128, 8, 2)])
def test_piecewise_append_equals_one_shot(oracle, codec, D, bs, H):
    """Writing a sequence in pieces of odd lengths (tiles that start and end inside a cache block) gives the cache image
    of a single write, and that image equals the oracle's."""
    from eccq_b200.kv_cache import read_cache
    pieces = [5, 20, 3, 17, 1, 31]
    S = sum(pieces)
    nblk = -(-S // bs) + 2
    rng = np.random.default_rng(D + bs + H)
    k = rng.standard_normal((1, S, H * D)).astype(np.float32)
    v = rng.standard_normal((1, S, H * D)).astype(np.float32)
    k[0, 7] = 0.0                                   # an all-zero row: scale 1.0, code 8
    kh, vh = dev(k).half(), dev(v).half()
    be1, m1 = _mgr(codec, D, bs, H, nblk)
    be1.write(kh, vh, layer_idx=1, seq_id=0)
    be2, m2 = _mgr(codec, D, bs, H, nblk)
    pos = 0
    for n in pieces:
        be2.write(kh[:, pos:pos + n].contiguous(), vh[:, pos:pos + n].contiguous(), layer_idx=1, seq_id=0)
        pos += n
    assert torch.equal(m1.block_table[0], m2.block_table[0])
    assert torch.equal(m1.k_cache, m2.k_cache) and torch.equal(m1.v_cache, m2.v_cache)
    assert torch.equal(m1.k_scales, m2.k_scales) and torch.equal(m1.v_scales, m2.v_scales)
    oc = oracle.PagedCache(nblk, bs, 2, H, D, codec)
    oracle.cache_write(oc, kh.float().cpu().numpy(), vh.float().cpu().numpy(), 1, m1.block_table[0].cpu().numpy())
    assert np.array_equal(m1.k_cache.cpu().numpy(), oc.k_cache) and np.array_equal(m1.v_cache.cpu().numpy(), oc.v_cache)
    assert np.array_equal(m1.k_scales.cpu().numpy(), oc.k_scales)
    cl = torch.tensor([S], dtype=torch.int32, device="cuda")
    kf, vf = read_cache(m2.k_cache, m2.v_cache, m2.k_scales, m2.v_scales, m2.block_table[:1], cl, 1, bs, D, codec)
    okf, ovf, _, _ = oracle.cache_read(oc, 1, m1.block_table[0].cpu().numpy(), S)
    assert np.array_equal(kf[0].cpu().numpy(), okf) and np.array_equal(vf[0].cpu().numpy(), ovf)


@pytest.mark.parametrize("codec,interp", [("hamming84", True), ("hamming84", False), ("hamming74", False), ("golay", False)])
@pytest.mark.parametrize("D,bs,H", [(128, 16, 8), (64, 16, 12), (64, 32, 2)])
def test_multi_sequence_ragged_read(oracle, codec, interp, D, bs, H):
    """Several sequences appended in ONE call at different start positions, then read with ragged context lengths,
    injected errors and (hamming84) interpolation across page boundaries; dequantised values, nibbles and counters."""
    from eccq_b200.kv_cache import read_cache
    first = [3, 40, 0, 2 * bs]          # existing lengths before the batched append
    add = 37
    B = len(first)
    nblk = sum(-(-(f + add) // bs) for f in first) + 2
    be, mgr = _mgr(codec, D, bs, H, nblk, L=1, interp=interp)
    mgr.free_blocks = [int(b) for b in np.random.default_rng(5).permutation(nblk)]
    rng = np.random.default_rng(1000 + D + bs)
    cwn = oracle.golay_cw_per_row(D) if codec == "golay" else D
    nbits = oracle.CODEC_BITS[codec]
    mdt = np.int32 if codec == "golay" else np.uint8
    seqs_k, seqs_v, masks_k, masks_v = [[] for _ in range(B)], [[] for _ in range(B)], [[] for _ in range(B)], [[] for _ in range(B)]

    def mk(n):
        kk = rng.standard_normal((n, H * D)).astype(np.float32)
        vv = rng.standard_normal((n, H * D)).astype(np.float32)
        fl = rng.random((n, H, cwn, nbits)) < 0.03
        mkk = (fl * (1 << np.arange(nbits))).sum(-1).astype(mdt)
        return kk, vv, mkk, np.roll(mkk, 3, axis=0)

    for b in range(B):                  # sequence b gets `first[b]` tokens on its own
        if first[b]:
            kk, vv, mkk, mvv = mk(first[b])
            be.set_flip_masks(dev(mkk[None]), dev(mvv[None]))
            be.write(dev(kk[None]).half(), dev(vv[None]).half(), 0, seq_id=b)
            seqs_k[b].append(kk); seqs_v[b].append(vv); masks_k[b].append(mkk); masks_v[b].append(mvv)
    batch = [mk(add) for _ in range(B)]  # then all sequences append `add` tokens in one call
    be.set_flip_masks(dev(np.stack([x[2] for x in batch])), dev(np.stack([x[3] for x in batch])))
    be.write(dev(np.stack([x[0] for x in batch])).half(), dev(np.stack([x[1] for x in batch])).half(), 0, seq_id=0)
    for b in range(B):
        seqs_k[b].append(batch[b][0]); seqs_v[b].append(batch[b][1]); masks_k[b].append(batch[b][2]); masks_v[b].append(batch[b][3])

    oc = oracle.PagedCache(nblk, bs, 1, H, D, codec)
    lens = [first[b] + add for b in range(B)]
    half = lambda x: torch.from_numpy(x).half().float().numpy()
    for b in range(B):
        kk, vv = half(np.concatenate(seqs_k[b]))[None], half(np.concatenate(seqs_v[b]))[None]
        oracle.cache_write(oc, kk, vv, 0, mgr.block_table[b].cpu().numpy(), k_mask=np.concatenate(masks_k[b])[None],
                           v_mask=np.concatenate(masks_v[b])[None])
    assert np.array_equal(mgr.k_cache.cpu().numpy(), oc.k_cache) and np.array_equal(mgr.v_cache.cpu().numpy(), oc.v_cache)
    assert np.array_equal(mgr.k_scales.cpu().numpy(), oc.k_scales) and np.array_equal(mgr.v_scales.cpu().numpy(), oc.v_scales)
    # ragged read: context lengths end inside blocks, one sequence shorter than what was written
    ctx = [lens[0], lens[1] - 5, lens[2] - bs - 1, lens[3]]
    cl = torch.tensor(ctx, dtype=torch.int32, device="cuda")
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    for odt in (torch.float32, torch.float16):
        ctr.zero_()
        kf, vf = read_cache(mgr.k_cache, mgr.v_cache, mgr.k_scales, mgr.v_scales, mgr.block_table[:B], cl, 0, bs, D, codec,
                            use_interpolation=interp, out_dtype=odt, counters=ctr)
        tc = td = 0
        for b in range(B):
            okf, ovf, c, d = oracle.cache_read(oc, 0, mgr.block_table[b].cpu().numpy(), ctx[b], use_interpolation=interp)
            assert torch.equal(kf[b, :ctx[b]].cpu(), torch.from_numpy(okf).to(odt)), (b, odt)
            assert torch.equal(vf[b, :ctx[b]].cpu(), torch.from_numpy(ovf).to(odt)), (b, odt)
            assert not kf[b, ctx[b]:].any()
            tc, td = tc + c, td + d
        assert (int(ctr[0]), int(ctr[1])) == (tc, td)
        assert tc > 0


def test_fast_injector_window_kernel_formats():
    """The windowed own-stream injector on the four codeword formats of the path, sizes with and without a tail chunk."""
    from eccq_b200.ecc_codecs import cuda_kernels as ck
    for dtype, n_bits in ((torch.uint8, 8), (torch.uint8, 7), (torch.uint8, 4), (torch.int32, 24)):
        for n in (1 << 20, (1 << 20) + 5, 16, 7):
            x = torch.zeros(n, dtype=dtype, device="cuda")
            out, (flips, affected) = ck.inject_bit_errors_fast(x, 2e-2, n_bits, seed=7, return_stats=True)
            o = out.cpu().numpy()
            assert int(np.unpackbits(o.view(np.uint8)).sum()) == flips and int((o != 0).sum()) == affected
            assert not (o.astype(np.int64) >> n_bits).any()
            bits = n * n_bits
            assert abs(flips - bits * 2e-2) < 5 * (bits * 2e-2) ** 0.5 + 3
            assert torch.equal(out, ck.inject_bit_errors_fast(x, 2e-2, n_bits, seed=7))

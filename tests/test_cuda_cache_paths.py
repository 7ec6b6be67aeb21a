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


@pytest.mark.parametrize("codec", ["hamming84", "golay"])
def test_fused_roundtrip_properties_at_c3_size(oracle, codec):
    """Size-independent properties of the fused write -> read path at BASELINE config-3 size (16 x 8192 tokens x 8 kv-heads
    x 128 = 134 M values per tensor, random page placement): one flipped bit per codeword is corrected back to the clean
    read (corrected == number of codewords); two flipped bits are all detected and, with interpolation, all replaced;
    sampled rows agree with the oracle."""
    from eccq_b200 import _lib
    from eccq_b200.kv_cache import SimpleBlockManager, read_cache
    B, S, H, D, bs = 16, 8192, 8, 128, 16
    nblk = B * S // bs
    lib = _lib.load()
    P = _lib.ptr
    sp = _lib.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda").manual_seed(5)
    k = torch.randn(B, S, H * D, device="cuda", dtype=torch.float16, generator=gen)
    v = torch.randn(B, S, H * D, device="cuda", dtype=torch.float16, generator=gen)
    m = SimpleBlockManager(nblk, bs, 1, H, D, device="cuda", codec=codec)
    m.block_table = torch.randperm(nblk, generator=torch.Generator().manual_seed(3)).to(torch.int32).view(B, S // bs).cuda()
    cl = torch.full((B,), S, dtype=torch.int32, device="cuda")
    cw = (D + 2) // 3 if codec == "golay" else D
    ncw = 2 * B * S * H * cw

    def write(km=None, vm=None):
        inj = _lib.INJECT_MASK if km is not None else _lib.INJECT_NONE
        _lib.check(lib.eccq_cache_write(m.desc(), P(k), P(v), _lib.ECCQ_F16, P(m.block_table), S // bs, None, B, S, 0, inj, P(km), P(vm),
                                        0.0, 0, 0, None, sp), "cache_write")

    def read(interp=False):
        ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
        kf, vf = read_cache(m.k_cache, m.v_cache, m.k_scales, m.v_scales, m.block_table, cl, 0, bs, D, codec, use_interpolation=interp,
                            out_dtype=torch.float16, counters=ctr)
        return kf, vf, ctr.tolist()

    write()
    k0, v0, c0 = read()
    assert c0[0] == 0 and c0[1] == 0
    # sampled rows against the oracle (first and last sequence, a few tokens)
    for b, t in ((0, 0), (0, 4097), (B - 1, S - 1)):
        row = k[b, t].float().cpu().numpy().reshape(H, D)
        q, sc = oracle.quantize_int4(row)
        assert np.array_equal(k0[b, t].cpu().numpy(), oracle.dequantize_int4(q, sc).astype(np.float16))
    nbits = 24 if codec == "golay" else 8
    mdt = torch.int32 if codec == "golay" else torch.uint8
    bit = torch.randint(0, nbits if codec == "golay" else 7, (B, S, H, cw), device="cuda", dtype=torch.int32)
    one = torch.bitwise_left_shift(torch.ones_like(bit), bit).to(mdt)
    write(one, one)
    k1, v1, c1 = read()
    assert torch.equal(k1, k0) and torch.equal(v1, v0)
    assert c1[0] == ncw and c1[1] == 0
    if codec == "hamming84":
        two = (one.to(torch.int32) | torch.bitwise_left_shift(torch.ones_like(bit), (bit + 1) % 8)).to(mdt)
        write(two, two)
        _, _, c2 = read(interp=True)
        assert c2[0] == 0 and c2[1] == ncw and c2[5] == ncw
    else:
        three = torch.full_like(one, 0x800401)    # three flips per codeword: still corrected
        write(three, three)
        k3, v3, c3 = read()
        assert torch.equal(k3, k0) and torch.equal(v3, v0) and c3[0] == 3 * ncw and c3[1] == 0


def test_decode_attention_corrects_single_flips_at_32k_context():
    """ECC-fused paged decode attention at the C5 context length (32 768 tokens, GQA 32/8, random pages): one flipped bit in
    EVERY codeword of K and V leaves the output bit-identical to the clean cache's (table prologue), both prologues agree
    with each other within the fp16 tolerance, and the counted decode reports every codeword as corrected."""
    from eccq_b200 import _lib
    from eccq_b200.kv_cache import SimpleBlockManager, paged_attention_ecc
    B, ctx, Hq, H, D, bs = 2, 32768, 32, 8, 128, 16
    nblk = B * ctx // bs
    lib = _lib.load()
    P = _lib.ptr
    sp = _lib.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda").manual_seed(9)
    k = torch.randn(B, ctx, H * D, device="cuda", dtype=torch.float16, generator=gen)
    v = torch.randn(B, ctx, H * D, device="cuda", dtype=torch.float16, generator=gen)
    q = torch.randn(B, Hq, D, device="cuda", dtype=torch.float16, generator=gen)
    m = SimpleBlockManager(nblk, bs, 1, H, D, device="cuda", codec="hamming84")
    m.block_table = torch.randperm(nblk, generator=torch.Generator().manual_seed(4)).to(torch.int32).view(B, ctx // bs).cuda()
    cl = torch.tensor([ctx, ctx - 1000], dtype=torch.int32, device="cuda")

    def write(mask=None):
        inj = _lib.INJECT_MASK if mask is not None else _lib.INJECT_NONE
        _lib.check(lib.eccq_cache_write(m.desc(), P(k), P(v), _lib.ECCQ_F16, P(m.block_table), ctx // bs, None, B, ctx, 0, inj, P(mask), P(mask),
                                        0.0, 0, 0, None, sp), "cache_write")

    def attend(counters=None):
        return paged_attention_ecc(q, m.k_cache, m.v_cache, m.block_table, cl, m.k_scales, 0, bs, v_scales=m.v_scales, counters=counters)

    write()
    clean = attend()
    bit = torch.randint(0, 8, (B, ctx, H, D), device="cuda", dtype=torch.int32)
    one = torch.bitwise_left_shift(torch.ones_like(bit), bit).to(torch.uint8)
    write(one)
    noisy = attend()
    assert torch.equal(noisy, clean)
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    counted = attend(ctr)                       # SWAR prologue (flags needed for the counters)
    assert (counted.float() - clean.float()).abs().max().item() < 2e-3
    # every codeword of the attended context is a single error: data-bit and parity-bit flips alike have syndrome != 0
    # except flips of the overall parity bit (bit 7), which are "parity only" and not counted (hamming84_triton.py:185-187)
    attended = int(cl.sum().item()) * H * D * 2
    par_only = 0
    for b in range(B):
        n = int(cl[b].item())
        par_only += 2 * int((bit[b, :n] == 7).sum().item())
    assert int(ctr[0]) == attended - par_only and int(ctr[1]) == 0


@pytest.mark.parametrize("codec,interp", [("hamming84", True), ("hamming84", False), ("hamming74", False), ("golay", False)])
def test_correct_pages_matches_oracle(oracle, codec, interp):
    """eccq_cache_correct_pages: corrected (+ interpolated) INT4 pages and row scales of one layer, against the oracle's decode
    of the gathered context; geometries outside the streaming kernels answer ECCQ_EUNSUPPORTED."""
    from eccq_b200 import _lib
    from eccq_b200.kv_cache import ECCBackend, ECCShimConfig, SimpleBlockManager
    D, bs, H, L, S = 128, 16, 3, 2, 77
    nblk = -(-S // bs) + 3
    cfg = ECCShimConfig(codec=codec, ber=0.03, inject_errors=True, block_size=bs, num_blocks=nblk, use_interpolation=interp, seed=7)
    mgr = SimpleBlockManager(nblk, bs, L, H, D, device="cuda", codec=codec)
    mgr.free_blocks = [int(b) for b in np.random.default_rng(2).permutation(nblk)]
    be = ECCBackend(mgr, cfg, H)
    rng = np.random.default_rng(17)
    k = rng.standard_normal((1, S, H * D)).astype(np.float32)
    v = rng.standard_normal((1, S, H * D)).astype(np.float32)
    be.write(dev(k).half(), dev(v).half(), layer_idx=1, seq_id=0)
    lib = _lib.load()
    P = _lib.ptr
    kp = torch.zeros(nblk, 1, H, bs * D, dtype=torch.uint8, device="cuda")
    vp, ksp = torch.zeros_like(kp), torch.zeros(nblk, 1, H, bs, dtype=torch.float32, device="cuda")
    vsp = torch.zeros_like(ksp)
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    cl = torch.tensor([S - 5], dtype=torch.int32, device="cuda")
    bt = mgr.block_table[:1].contiguous()
    rc = lib.eccq_cache_correct_pages(mgr.desc(), P(bt), bt.shape[1], P(cl), 1, 1, int(interp), P(kp), P(vp), P(ksp), P(vsp), P(ctr),
                                      _lib.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.eccq_last_error()
    # oracle: gather + decode the same context
    oc = oracle.PagedCache(nblk, bs, L, H, D, codec)
    oc.k_cache[:] = mgr.k_cache.cpu().numpy(); oc.v_cache[:] = mgr.v_cache.cpu().numpy()
    oc.k_scales[:] = mgr.k_scales.cpu().numpy(); oc.v_scales[:] = mgr.v_scales.cpu().numpy()
    btn = bt[0].cpu().numpy()
    kc, vc, ks, vs = oracle.cache_gather(oc, 1, btn, S - 5)
    kd, c1, d1 = oracle.decode_rows(kc, codec, D, interp)
    vd, c2, d2 = oracle.decode_rows(vc, codec, D, interp)
    kpn, vpn, kspn = kp.cpu().numpy(), vp.cpu().numpy(), ksp.cpu().numpy()
    for t in range(S - 5):
        blk, slot = int(btn[t // bs]), t % bs
        assert np.array_equal(kpn[blk, 0, :, slot * D:(slot + 1) * D], kd[t]), t
        assert np.array_equal(vpn[blk, 0, :, slot * D:(slot + 1) * D], vd[t]), t
        assert np.array_equal(kspn[blk, 0, :, slot], ks[t])
    assert (int(ctr[0]), int(ctr[1])) == (c1 + c2, d1 + d2)
    # a geometry the streaming kernels do not cover
    m2 = SimpleBlockManager(4, 4, 1, 2, 32, device="cuda", codec=codec)
    z8 = torch.zeros(4, 1, 2, 4 * 32, dtype=torch.uint8, device="cuda")
    zf = torch.zeros(4, 1, 2, 4, dtype=torch.float32, device="cuda")
    bt2 = torch.zeros(1, 4, dtype=torch.int32, device="cuda")
    rc = lib.eccq_cache_correct_pages(m2.desc(), P(bt2), 4, P(cl), 1, 0, 0, P(z8), P(z8.clone()), P(zf), P(zf.clone()), None, None)
    assert rc == _lib.ECCQ_EUNSUPPORTED

#!/usr/bin/env python
"""
Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference, never the GPU box):

    TRITON_INTERPRET=1 python tests/golden/make_golden.py [core] [gpt2] [unprotected]

The reference's @triton.jit kernels execute on CPU tensors under the Triton interpreter.  Its
Python wrappers assert ``tensor.is_cuda``; we re-exec each wrapper's source with that one line
removed (inside its own module namespace, nothing is written to the reference tree) so that the
reference's real host logic (ECCBackend.write / attend, golay_decode's statistics, the
interpolation permutes ...) produces the vectors, not a restatement of it.

Outputs (*.npz, a few hundred KB in total) are committed; tests/test_oracle_golden.py pins
oracle/eccq_oracle.py against them and the -m gpu tests pin the CUDA path against both.
"""
import inspect
import os
import re
import sys
from pathlib import Path

os.environ.setdefault("TRITON_INTERPRET", "1")
REF = os.environ.get("ECCQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

OUT = Path(__file__).resolve().parent


def _strip_cuda_asserts():
    import ecc_codecs.triton_kernels.hamming74_triton as h74
    import ecc_codecs.triton_kernels.hamming84_triton as h84
    import ecc_codecs.triton_kernels.golay_triton as gol
    import ecc_codecs.triton_kernels.fault_injection_triton as fi
    import ecc_codecs.triton_kernels.interpolation_triton as ip
    import ecc_codecs.triton_kernels as pkg
    import kv_cache.attention_ecc as att
    import kv_cache.ecc_shim as shim
    import kv_cache.paged_cache_ecc as pc

    todo = [
        (h74, ["hamming74_encode", "hamming74_decode"]),
        (h84, ["hamming84_encode", "hamming84_decode"]),
        (gol, ["golay_encode", "golay_decode"]),
        (fi, ["inject_bit_errors_triton", "inject_bit_errors_triton_vectorized"]),
        (ip, ["interpolate_double_errors"]),
        (att, ["paged_attention_ecc"]),
    ]
    for mod, names in todo:
        for name in names:
            src = inspect.getsource(getattr(mod, name))
            src = re.sub(r"^\s*assert .*is_cuda.*$", "", src, flags=re.M)
            exec(compile(src, mod.__file__, "exec"), mod.__dict__)
            for user in (pkg, shim, att, pc):
                if hasattr(user, name):
                    setattr(user, name, getattr(mod, name))
    return h74, h84, gol, fi, ip, att, shim, pc


def core_sections(mods):
    h74, h84, gol, fi, ip, att, shim, pc = mods
    from ecc_codecs.triton_kernels.config import build_golay_syndrome_table, GOLAY_H_ROW_MASKS

    rng = np.random.default_rng(20251018)
    T = torch.from_numpy

    # ---------------------------------------------------------------- 1. exhaustive Hamming tables
    nib = torch.arange(16, dtype=torch.uint8)
    allb = torch.arange(256, dtype=torch.int32).to(torch.uint8)
    h74_enc = h74.hamming74_encode(nib)
    h84_enc = h84.hamming84_encode(nib)
    d74, f74, (n74,) = h74.hamming74_decode(allb, return_error_detected=True)
    d84, e84, (c84, dd84) = h84.hamming84_decode(allb, return_error_types=True)
    # encode must ignore the high nibble
    h84_enc_hi = h84.hamming84_encode(allb)
    np.savez_compressed(
        OUT / "hamming_tables.npz",
        h74_enc=h74_enc.numpy(), h84_enc=h84_enc.numpy(), h84_enc_all256=h84_enc_hi.numpy(),
        h74_dec_data=d74.numpy(), h74_dec_flag=f74.numpy(), h74_count=np.int64(n74),
        h84_dec_data=d84.numpy(), h84_dec_etype=e84.numpy(), h84_counts=np.array([c84, dd84], dtype=np.int64),
    )

    # ---------------------------------------------------------------- 2. Golay
    table = build_golay_syndrome_table().numpy()
    trip = rng.integers(0, 16, size=(3000, 3), dtype=np.uint8)
    trip[:5] = np.array([[0, 0, 0], [1, 2, 3], [15, 15, 15], [8, 8, 8], [5, 10, 7]], dtype=np.uint8)
    cw = gol.golay_encode(T(trip)).numpy()
    # error patterns of weight 0..7, random positions
    pat = np.zeros(3000, dtype=np.int32)
    for i in range(3000):
        w = i % 8
        pos = rng.choice(24, size=w, replace=False)
        for p in pos:
            pat[i] |= 1 << int(p)
    rx = (cw ^ pat).astype(np.int32)
    dec, cnt, (bits, unc) = gol.golay_decode(T(rx), return_error_counts=True)
    # all 24 single flips and all 276 double flips of one codeword (verification.py:407-471)
    base = int(cw[4])
    sd = [base ^ (1 << i) for i in range(24)] + [base ^ (1 << i) ^ (1 << j) for i in range(24) for j in range(i + 1, 24)]
    sd = np.array(sd, dtype=np.int32)
    dec_sd, cnt_sd, st_sd = gol.golay_decode(T(sd), return_error_counts=True)
    np.savez_compressed(
        OUT / "golay_vectors.npz",
        syndrome_table=table, h_masks=np.array(GOLAY_H_ROW_MASKS, dtype=np.int64),
        b_cols=np.array([getattr(gol, f"B_COL_{i}") for i in range(12)], dtype=np.int64),
        triplets=trip, codewords=cw, patterns=pat, received=rx, decoded=dec.numpy(), errcnt=cnt.numpy(),
        stats=np.array([bits, unc], dtype=np.int64),
        sd_received=sd, sd_decoded=dec_sd.numpy(), sd_errcnt=cnt_sd.numpy(), sd_stats=np.array(st_sd, dtype=np.int64),
    )

    # ---------------------------------------------------------------- 3. fault injection (Philox stream)
    inj = {}
    cases = [  # (tag, dtype, N, n_bits, seed, ber, vectorized)
        ("u8_4096_8_s42_b05", np.uint8, 4096, 8, 42, 0.05, False),
        ("u8_1000_7_s7_b01", np.uint8, 1000, 7, 7, 0.01, False),
        ("u8_64_8_s42_b1", np.uint8, 64, 8, 42, 0.1, False),
        ("u8_64_8_s43_b1", np.uint8, 64, 8, 43, 0.1, False),
        ("u8_128_4_s99_b2", np.uint8, 128, 4, 99, 0.2, False),
        ("u8_128_8_s123456789_b1", np.uint8, 128, 8, 123456789, 0.1, False),  # int32 key wrap
        ("i32_1000_24_s42_b02", np.int32, 1000, 24, 42, 0.02, False),
        ("i32_43_24_s45_b1", np.int32, 43, 24, 45, 0.1, False),
        ("u8_4096_8_s42_b05_vec", np.uint8, 4096, 8, 42, 0.05, True),
        ("i32_1000_24_s42_b02_vec", np.int32, 1000, 24, 42, 0.02, True),
    ]
    for tag, dt, n, nb, seed, ber, vec in cases:
        z = torch.zeros(n, dtype=torch.uint8 if dt == np.uint8 else torch.int32)
        f = fi.inject_bit_errors_triton_vectorized if vec else fi.inject_bit_errors_triton
        out, (tot, aff) = f(z, ber, nb, seed, return_stats=True)
        inj[tag + "_mask"] = out.numpy()
        inj[tag + "_meta"] = np.array([n, nb, seed, int(vec), tot, aff], dtype=np.int64)
        inj[tag + "_ber"] = np.float64(ber)
    np.savez_compressed(OUT / "inject_vectors.npz", **inj)

    # ---------------------------------------------------------------- 4. interpolation
    ipd = {}
    gold = [  # tests/test_triton_interpolation.py:19-125
        ([4, 8, 12, 8, 4], [0, 0, 2, 0, 0]),
        ([15, 4, 8, 12], [2, 0, 0, 0]),
        ([4, 8, 12, 15], [0, 0, 0, 2]),
        ([4, 0, 0, 0, 4], [0, 2, 2, 2, 0]),
        ([7], [2]),
        ([3, 200, 15, 16, 0], [0, 0, 2, 0, 1]),  # non-INT4 inputs saturate
    ]
    for i, (qv, ev) in enumerate(gold):
        o = ip.interpolate_double_errors(torch.tensor(qv, dtype=torch.uint8), torch.tensor(ev, dtype=torch.uint8))
        ipd[f"g{i}_q"], ipd[f"g{i}_e"], ipd[f"g{i}_o"] = np.array(qv, np.uint8), np.array(ev, np.uint8), o.numpy()
    q3 = rng.integers(0, 16, size=(37, 3, 8), dtype=np.uint8)
    e3 = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(37, 3, 8), p=[0.6, 0.1, 0.25, 0.05])
    o3 = ip.interpolate_double_errors(T(q3), T(e3), seq_dim=0)
    o3l = ip.interpolate_double_errors(T(q3), T(e3), seq_dim=-1)
    ipd.update(r_q=q3, r_e=e3, r_out_seq0=o3.contiguous().numpy(), r_out_last=o3l.contiguous().numpy())
    np.savez_compressed(OUT / "interp_vectors.npz", **ipd)

    # ---------------------------------------------------------------- 5. quantisation
    x32 = rng.standard_normal((64, 24)).astype(np.float32)
    x32[0] = 0.0                                   # zero row -> scale 1, nibble 8
    x32[1] = np.array([7.0, 3.5, 2.5, 1.5, 0.5, -0.5, -1.5, -2.5, -3.5, -7.0, 6.5, -6.5] * 2, dtype=np.float32)  # ties
    x32[2, :] = 1e-30
    x32[3, 0] = 6.0e4                               # large but finite in fp16
    x16 = T(x32).to(torch.float16)
    xb16 = T(x32).to(torch.bfloat16)
    qd = {}
    for tag, x in (("f32", T(x32)), ("f16", x16), ("bf16", xb16)):
        xf = x.float()
        s = pc.compute_quantization_scales(xf, dim=-1)
        qv = (torch.round(xf / s.unsqueeze(-1)).clamp(-8, 7) + 8).to(torch.uint8)      # ecc_shim.py:575-580
        qd[tag + "_x"] = x.float().numpy() if tag != "f32" else x32
        qd[tag + "_bits"] = x.view(torch.int16).numpy() if tag != "f32" else x32.view(np.int32)
        qd[tag + "_scales"] = s.numpy()
        qd[tag + "_q"] = qv.numpy()
    np.savez_compressed(OUT / "quant_vectors.npz", **qd)

    # ---------------------------------------------------------------- 6. the shim: ECCBackend.write/attend on CPU
    sh = {}
    B, S, Hq, Hkv, D, L, bs, nblk = 2, 21, 4, 2, 16, 2, 4, 12
    kin = rng.standard_normal((B, S, Hkv * D)).astype(np.float32)
    vin = rng.standard_normal((B, S, Hkv * D)).astype(np.float32)
    kin[1, 3, :D] = 0.0  # an all-zero row
    qpre = rng.standard_normal((1, Hq, S, D)).astype(np.float32)
    qdec = rng.standard_normal((1, Hq, 1, D)).astype(np.float32)
    sh.update(k=kin, v=vin, q_prefill=qpre, q_decode=qdec,
              dims=np.array([B, S, Hq, Hkv, D, L, bs, nblk], dtype=np.int64))
    modes = [("hamming74", False), ("hamming84", False), ("hamming84", True), ("golay", False)]
    for codec, interp in modes:
        for ber in (0.0, 0.05):
            tag = f"{codec}{'_interp' if interp else ''}_ber{ber}"
            cfg = shim.ECCShimConfig(codec=codec, ber=ber, block_size=bs, num_blocks=nblk,
                                     inject_errors=ber > 0, seed=42, use_interpolation=interp)
            mgr = shim.SimpleBlockManager(nblk, bs, L, Hkv, D, device="cpu", codec=codec)
            be = shim.ECCBackend(mgr, cfg, Hq)
            be.write(T(kin), T(vin), layer_idx=1, seq_id=0)
            sh[tag + "_k_cache"] = mgr.k_cache.numpy().copy()
            sh[tag + "_v_cache"] = mgr.v_cache.numpy().copy()
            sh[tag + "_k_scales"] = mgr.k_scales.numpy().copy()
            sh[tag + "_v_scales"] = mgr.v_scales.numpy().copy()
            sh[tag + "_block_table"] = mgr.block_table[0].numpy().copy()
            o_pre = be.attend(T(qpre), layer_idx=1, seq_id=0)
            st_pre = (be._errors_corrected, be._errors_detected)
            o_dec = be.attend(T(qdec), layer_idx=1, seq_id=0)  # hamming84/no-interp takes the Triton attention kernel
            # MHA twin of the decode step: the Triton fast path is only correct for Hq == Hkv
            be_mha = shim.ECCBackend(mgr, cfg, Hkv)
            sh[tag + "_out_decode_mha"] = be_mha.attend(T(qdec[:, :Hkv]), layer_idx=1, seq_id=0).numpy()
            sh[tag + "_out_prefill"] = o_pre.numpy()
            sh[tag + "_out_decode"] = o_dec.numpy()
            sh[tag + "_stats"] = np.array([be._injection_count, st_pre[0], st_pre[1], be._errors_corrected,
                                           be._errors_detected, be._total_values], dtype=np.int64)
    np.savez_compressed(OUT / "shim_vectors.npz", **sh)

    # ---------------------------------------------------------------- 7. paged decode attention (Triton kernel, MHA)
    at = {}
    Bq, H, D2, L2, bs2, nb2 = 3, 4, 32, 2, 8, 16
    mb = 5
    ctx = np.array([37, 8, 1], dtype=np.int32)
    kq = rng.integers(0, 16, size=(nb2, L2, H, bs2 * D2), dtype=np.uint8)
    vq = rng.integers(0, 16, size=(nb2, L2, H, bs2 * D2), dtype=np.uint8)
    kcw = h84.hamming84_encode(T(kq))
    vcw = h84.hamming84_encode(T(vq))
    kcw = fi.inject_bit_errors_triton(kcw, 0.02, 8, seed=5)
    vcw = fi.inject_bit_errors_triton(vcw, 0.02, 8, seed=6)
    ksc = (rng.random((nb2, L2, H, bs2)).astype(np.float32) * 0.2 + 0.01)
    vsc = (rng.random((nb2, L2, H, bs2)).astype(np.float32) * 0.2 + 0.01)
    bt = np.full((Bq, mb), -1, dtype=np.int32)
    perm = rng.permutation(nb2)
    bt[0, :5] = perm[:5]
    bt[1, :1] = perm[5:6]
    bt[2, :1] = perm[6:7]
    qa = rng.standard_normal((Bq, H, D2)).astype(np.float32)
    o_tri = att.paged_attention_ecc(T(qa), kcw, vcw, T(bt), T(ctx), T(ksc), 1, bs2, v_scales=T(vsc))
    o_ref_k = att.reference_attention_ecc(T(qa), kcw, vcw, T(bt), T(ctx), T(ksc), 1, bs2, D2)
    o_tri_k = att.paged_attention_ecc(T(qa), kcw, vcw, T(bt), T(ctx), T(ksc), 1, bs2)
    at.update(q=qa, k_cache=kcw.numpy(), v_cache=vcw.numpy(), k_scales=ksc, v_scales=vsc, block_table=bt,
              context_lens=ctx, dims=np.array([Bq, H, D2, L2, bs2, nb2, mb], dtype=np.int64),
              out_triton_kv=o_tri.numpy(), out_triton_kscale_only=o_tri_k.numpy(), out_pyref_kscale_only=o_ref_k.numpy())
    # golay python reference path (attention_ecc.py:759-775), D not a multiple of 3
    Dg = 10
    ncw = (Dg + 2) // 3
    gq = rng.integers(0, 16, size=(nb2, L2, H, bs2, ncw * 3), dtype=np.uint8)
    gq[..., Dg:] = 0
    gk = gol.golay_encode(T(gq.reshape(-1, 3))).view(nb2, L2, H, bs2 * ncw)
    gv = gol.golay_encode(T(np.ascontiguousarray(gq[::-1]).reshape(-1, 3))).view(nb2, L2, H, bs2 * ncw)
    gk = fi.inject_bit_errors_triton(gk, 0.03, 24, seed=11)
    gv = fi.inject_bit_errors_triton(gv, 0.03, 24, seed=12)
    qg = rng.standard_normal((Bq, H, Dg)).astype(np.float32)
    o_g = att.paged_attention_ecc(T(qg), gk, gv, T(bt), T(ctx), T(ksc), 0, bs2, codec="golay")
    at.update(g_q=qg, g_k_cache=gk.numpy(), g_v_cache=gv.numpy(), g_out=o_g.numpy(), g_dims=np.array([Dg, ncw], dtype=np.int64))
    np.savez_compressed(OUT / "attention_vectors.npz", **at)



def gpt2_section(mods):
    """Config 2 in miniature: a random-init 2-layer GPT-2 under patch_model_with_ecc_attention on CPU, every ECC mode,
    BER 0 and 1e-2 (evaluation/sweep.py:512-590 call pattern: reset_ecc_cache, forward with labels, get_ecc_stats)."""
    shim = mods[6]
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(1234)
    cfg = GPT2Config(n_layer=2, n_head=2, n_embd=32, vocab_size=100, n_positions=64, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    model = GPT2LMHeadModel(cfg).eval()
    ids = torch.randint(0, 100, (1, 24), generator=torch.Generator().manual_seed(5))
    out = {"input_ids": ids.numpy()}
    for k, v in model.state_dict().items():
        out["w::" + k] = v.numpy()
    with torch.no_grad():
        out["baseline_logits"] = model(ids).logits.numpy()
    for codec, interp in (("hamming74", False), ("hamming84", False), ("hamming84", True), ("golay", False)):
        for ber in (0.0, 1e-2):
            tag = f"{codec}{'_interp' if interp else ''}_ber{ber}"
            c = shim.ECCShimConfig(codec=codec, ber=ber, inject_errors=ber > 0, seed=42, use_interpolation=interp, block_size=16, num_blocks=8)
            with shim.patch_model_with_ecc_attention(model, c, num_blocks=8):
                shim.reset_ecc_cache(model)
                with torch.no_grad():
                    r = model(ids, labels=ids)
                st = shim.get_ecc_stats(model)
            out[tag + "_logits"] = r.logits.numpy()
            out[tag + "_loss"] = np.float64(r.loss.item())
            out[tag + "_stats"] = np.array([st["injection_count"], st["errors_corrected"], st["errors_detected"], st["total_values"],
                                            st["allocated_blocks"], st["free_blocks"], st["sequences"]], dtype=np.int64)
            print(tag, float(r.loss), st)
    np.savez_compressed(OUT / "gpt2_shim_vectors.npz", **out)


def unprotected_section(mods):
    """The unprotected comparison modes: ECCBackend with codec fp16 / fp8 / int4 (ecc_shim.py:511-625, :802-928) and
    kv_cache/unprotected_shim.py's UnprotectedBackend (INT4, bit errors injected after the cache read), BER 0 and 0.05."""
    shim = mods[6]
    import kv_cache.unprotected_shim as un
    import ecc_codecs.triton_kernels.fault_injection_triton as fi
    un.inject_bit_errors_triton = fi.inject_bit_errors_triton          # the assert-stripped wrapper
    T = torch.from_numpy
    rng = np.random.default_rng(77)
    B, S, Hq, Hkv, D, L, bs, nblk = 2, 21, 4, 2, 16, 2, 4, 12
    kin = rng.standard_normal((B, S, Hkv * D)).astype(np.float32)
    vin = rng.standard_normal((B, S, Hkv * D)).astype(np.float32)
    qpre = rng.standard_normal((1, Hq, S, D)).astype(np.float32)
    qdec = rng.standard_normal((1, Hq, 1, D)).astype(np.float32)
    out = dict(k=kin, v=vin, q_prefill=qpre, q_decode=qdec, dims=np.array([B, S, Hq, Hkv, D, L, bs, nblk], dtype=np.int64))
    for codec in ("fp16", "fp8", "int4"):
        for ber in (0.0, 0.05):
            tag = f"{codec}_ber{ber}"
            cfg = shim.ECCShimConfig(codec=codec, ber=ber, block_size=bs, num_blocks=nblk, inject_errors=ber > 0, seed=42)
            mgr = shim.SimpleBlockManager(nblk, bs, L, Hkv, D, device="cpu", codec=codec)
            be = shim.ECCBackend(mgr, cfg, Hq)
            be.write(T(kin), T(vin), layer_idx=1, seq_id=0)
            raw = (lambda t: t.view(torch.uint8).numpy().copy()) if codec == "fp8" else (lambda t: t.numpy().copy())
            out[tag + "_k_cache"], out[tag + "_v_cache"] = raw(mgr.k_cache), raw(mgr.v_cache)
            out[tag + "_k_scales"], out[tag + "_v_scales"] = mgr.k_scales.numpy().copy(), mgr.v_scales.numpy().copy()
            out[tag + "_out_prefill"] = be.attend(T(qpre), layer_idx=1, seq_id=0).numpy()
            out[tag + "_out_decode"] = be.attend(T(qdec), layer_idx=1, seq_id=0).numpy()
            out[tag + "_stats"] = np.array([be._injection_count, be._errors_corrected, be._errors_detected, be._total_values], dtype=np.int64)
    for ber in (0.0, 0.05):
        tag = f"unprotected_ber{ber}"
        cfg = un.UnprotectedShimConfig(ber=ber, block_size=bs, num_blocks=nblk, inject_errors=ber > 0, seed=42)
        mgr = shim.SimpleBlockManager(nblk, bs, L, Hkv, D, device="cpu", codec="int4")
        be = un.UnprotectedBackend(mgr, cfg, Hq)
        be.write(T(kin), T(vin), layer_idx=1, seq_id=0)
        out[tag + "_k_cache"], out[tag + "_v_cache"] = mgr.k_cache.numpy().copy(), mgr.v_cache.numpy().copy()
        out[tag + "_k_scales"], out[tag + "_v_scales"] = mgr.k_scales.numpy().copy(), mgr.v_scales.numpy().copy()
        out[tag + "_out_prefill"] = be.attend(T(qpre), layer_idx=1, seq_id=0).numpy()
        out[tag + "_out_prefill2"] = be.attend(T(qpre), layer_idx=1, seq_id=0).numpy()       # a second call: the next seed
        out[tag + "_out_decode"] = be.attend(T(qdec), layer_idx=1, seq_id=0).numpy()
        out[tag + "_stats"] = np.array([be._injection_count, be._bits_flipped, be._total_bits], dtype=np.int64)
    np.savez_compressed(OUT / "unprotected_vectors.npz", **out)


def main():
    sel = set(sys.argv[1:])
    mods = _strip_cuda_asserts()
    if not sel or "core" in sel:
        core_sections(mods)
    if not sel or "gpt2" in sel:
        gpt2_section(mods)
    if not sel or "unprotected" in sel:
        unprotected_section(mods)
    print("golden fixtures written to", OUT)
    for f in sorted(OUT.glob("*.npz")):
        print(f"  {f.name:28s} {f.stat().st_size:8d} B")


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz by running the UNMODIFIED reference: the compiled cbench.ans coder
(oracle/_ref, built by oracle/Makefile) and the reference's Python y-path imported from /root/reference
through tests/golden/ref_shim.py.  Run in the build container only (the reference does not travel):

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no golden vectors (all its tests are unseeded round trips), so
these seeded vectors are what pins oracle/ and, through it, the CUDA path.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from oracle import ref_loader  # noqa: E402


def u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8)


def coder_vectors():
    R = ref_loader.load("ans")
    rng = np.random.default_rng(20240917)
    out = {}
    # (a) tests/ans_test.py:17-43 shapes, reduced batch: 8 tables x 512 random freqs, 6 % escapes
    T, M = 8, 512
    freqs = rng.integers(1, 1024, (T, M)).astype(np.int32)
    nsym, offs = np.full(T, M, np.int32), np.zeros(T, np.int32)
    data = rng.integers(0, M + 32, (6, 3, 32, 32)).astype(np.int32)
    idx = rng.integers(0, T, data.shape).astype(np.int32)
    enc, dec = R.Rans64Encoder(bypass_coding=True), R.Rans64Decoder(bypass_coding=True)
    enc.init_params(freqs, nsym, offs)
    dec.init_params(freqs, nsym, offs)
    bs = enc.encode_with_indexes(data, idx)
    assert np.array_equal(dec.decode_with_indexes(bs, idx), data)
    cdfs = enc.get_cdfs()[:, :M + 2].copy()
    out.update(a_freqs=freqs, a_nsym=nsym, a_offsets=offs, a_data=data, a_idx=idx, a_rans=u8(bs), a_cdfs=cdfs)
    # (b) same tables, offsets != 0, negative and very large escapes, bypass precision 4
    offs_b = rng.integers(-300, -200, T).astype(np.int32)
    data_b = (rng.integers(0, M, (4000,)) + offs_b[idx.reshape(-1)[:4000]]).astype(np.int32)
    data_b[::7] -= 100000
    data_b[::11] += 2000000
    data_b[5] = -(2 ** 26)   # raw must stay < 2^28: the reference's digit-count loop (rans64.cpp:299) never ends beyond that
    idx_b = idx.reshape(-1)[:4000].copy()
    enc.init_params(freqs, nsym, offs_b)
    dec.init_params(freqs, nsym, offs_b)
    bs = enc.encode_with_indexes(data_b, idx_b)
    assert np.array_equal(dec.decode_with_indexes(bs, idx_b), data_b)
    out.update(b_offsets=offs_b, b_data=data_b, b_idx=idx_b, b_rans=u8(bs))
    # (c) bypass disabled, precision 12 (non-default freq_precision), in-range symbols only
    enc12, dec12 = R.Rans64Encoder(12, False, 4), R.Rans64Decoder(12, False, 4)
    f12 = rng.integers(1, 64, (4, 40)).astype(np.int32)
    n12 = np.array([40, 17, 3, 29], np.int32)
    o12 = np.array([0, -8, -1, 5], np.int32)
    enc12.init_params(f12, n12, o12)
    dec12.init_params(f12, n12, o12)
    idx_c = rng.integers(0, 4, (3000,)).astype(np.int32)
    data_c = (rng.integers(0, 1 << 20, (3000,)) % (n12[idx_c] + 1) + o12[idx_c]).astype(np.int32)
    bs = enc12.encode_with_indexes(data_c, idx_c)
    assert np.array_equal(dec12.decode_with_indexes(bs, idx_c), data_c)
    c12 = enc12.get_cdfs()
    for t in range(4):
        c12[t, n12[t] + 2:] = 0
    out.update(c_freqs=f12, c_nsym=n12, c_offsets=o12, c_data=data_c, c_idx=idx_c, c_rans=u8(bs), c_cdfs=c12)
    # (d) tANS, tests/ans_test.py:112-137 shapes (table_log 11, max_symbol_value 511, bypass)
    tenc = R.TansEncoder(max_symbol_value=M - 1, bypass_coding=True)
    tdec = R.TansDecoder(max_symbol_value=M - 1, bypass_coding=True)
    tenc.init_params(freqs, nsym, offs)
    tdec.init_params(freqs, nsym, offs)
    bs = tenc.encode_with_indexes(data, idx)
    assert np.array_equal(tdec.decode_with_indexes(bs, idx), data)
    out.update(d_tans=u8(bs))
    # (e) tANS, skewed tables (exercise the -1 / low-probability and non-fast paths), table_log 10, no bypass
    fe = np.maximum((4000 * np.exp(-0.08 * np.arange(200))[None, :] * rng.uniform(0.5, 1.5, (3, 200))), 1).astype(np.int32)
    fe[1, 0] = 200000
    ne = np.array([200, 150, 90], np.int32)
    oe = np.array([0, -3, 7], np.int32)
    tenc = R.TansEncoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    tdec = R.TansDecoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    tenc.init_params(fe, ne, oe)
    tdec.init_params(fe, ne, oe)
    idx_e = rng.integers(0, 3, (5000,)).astype(np.int32)
    data_e = (np.minimum(rng.geometric(0.08, 5000) - 1, ne[idx_e] - 2) + oe[idx_e]).astype(np.int32)
    bs = tenc.encode_with_indexes(data_e, idx_e)
    assert np.array_equal(tdec.decode_with_indexes(bs, idx_e), data_e)
    out.update(e_freqs=fe, e_nsym=ne, e_offsets=oe, e_data=data_e, e_idx=idx_e, e_tans=u8(bs))
    # (f) pmf_to_quantized_cdf on its own (tests/ans_test.py:96-102 usage)
    p = (freqs[0].astype(np.float32) / freqs[0].sum()).tolist() + [1e-8]
    out.update(f_pmf=np.array(p, np.float32), f_cdf=np.array(R.pmf_to_quantized_cdf(p, 16), np.int32))
    np.savez_compressed(os.path.join(HERE, "coder_vectors.npz"), **out)
    print("coder_vectors.npz", {k: v.shape for k, v in out.items()})


def ypath_vectors():
    Coder, Ctx = ref_shim.load()
    cases = [  # name, C, G, method, B, H, W, seed, pgm kind, context model?
        ("ckbd", 24, 1, "checkerboard", 2, 6, 8, 0, None, True),
        ("meanscale", 24, 1, "none", 2, 6, 8, 1, None, False),
        ("cwckbd", 24, 4, "channelwise-checkerboard", 2, 5, 7, 2, None, True),
        ("scanline", 12, 1, "scanline", 1, 4, 5, 3, None, True),
        ("learned_int", 24, 4, "none", 3, 5, 7, 4, "int6", True),
        ("raster", 24, 2, "raster2x2", 1, 6, 6, 5, None, True),
        ("learned_logits", 24, 4, "none", 2, 6, 8, 6, "logits8", True),
    ]
    out = {}
    tables_done = False
    for name, C, G, method, B, H, W, seed, pk, ctx in cases:
        torch.manual_seed(seed)
        kw = dict(in_channels=C, channel_groups=G, default_topo_group_method=method)
        if ctx:
            kw["topo_group_context_model"] = Ctx(in_channels=C, out_channels=2 * C)
        else:
            kw["use_param_merger"] = False
        coder = Coder(**kw)
        coder.eval()
        coder.update_state()
        y, p = 3 * torch.randn(B, C, H, W), torch.randn(B, 2 * C, H, W)
        pgm = None
        if pk == "int6":
            pgm = torch.randint(0, 6, (1, G, 2, 2))
        elif pk == "logits8":
            pgm = torch.randn(1, G * 8, 2, 2)
        rec = {}
        real = coder.ans_encoder

        class Spy:
            def encode_with_indexes(self, data, indexes, **k):
                rec["sym"], rec["idx"] = data.copy(), indexes.copy()
                return real.encode_with_indexes(data, indexes, **k)
        coder.ans_encoder = Spy()
        with torch.no_grad():
            bs = coder.encode(y, prior=p, pgm=pgm)
            coder.ans_encoder = real
            yh = coder.decode(bs, prior=p, pgm=pgm)
            tg = coder._get_pgm(y, input_shape=y.shape, pgm=pgm, fast_mode=True)
            params0 = coder._pgm_inference_group_mask(torch.zeros_like(y), None, pgm=tg, prior=p)
            params_full = coder._pgm_inference_group_mask(yh, None, pgm=tg, prior=p)
        sd = {k: v.detach().cpu().numpy() for k, v in coder.state_dict().items()
              if "context" in k or "param_merger" in k}
        for k, v in sd.items():
            out[f"{name}.sd.{k}"] = v
        out[f"{name}.meta"] = np.array([C, G, B, H, W, int(ctx)], np.int32)
        out[f"{name}.y"], out[f"{name}.prior"] = y.numpy(), p.numpy()
        if pgm is not None:
            out[f"{name}.pgm"] = pgm.numpy()
        out[f"{name}.tg"] = tg.cpu().numpy().astype(np.int32)
        out[f"{name}.sym"], out[f"{name}.idx"] = rec["sym"], rec["idx"]
        out[f"{name}.bytes"], out[f"{name}.yhat"] = u8(bs), yh.numpy()
        out[f"{name}.params0"], out[f"{name}.params_full"] = params0.numpy(), params_full.numpy()
        print(name, "bytes", len(bs), "max|yhat-y|", float((yh - y).abs().max()))
        if not tables_done:
            freqs, nsym, offs = coder._get_ans_params()
            cd = real.get_cdfs()
            flat = np.concatenate([cd[t, :nsym[t] + 2] for t in range(len(nsym))]).astype(np.int32)
            tab = dict(scale_table=coder.scale_table.numpy(), freqs=freqs, nsym=nsym, offsets=offs, cdf_flat=flat,
                       sha256=np.frombuffer(hashlib.sha256(flat.tobytes()).digest(), dtype=np.uint8))
            np.savez_compressed(os.path.join(HERE, "gaussian_tables.npz"), **tab)
            print("gaussian_tables.npz nsym", nsym.min(), nsym.max(), int(nsym.sum()), "cdf0", flat[:5])
            tables_done = True
    np.savez_compressed(os.path.join(HERE, "ypath_vectors.npz"), **out)


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference and oracle/_ref (make -C oracle ref)"
    coder_vectors()
    ypath_vectors()

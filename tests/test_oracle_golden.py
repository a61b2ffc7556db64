"""The CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; the CUDA path is then checked against the
oracle in the -m gpu tests."""
import hashlib
import os

import struct

import numpy as np
import pytest
import torch

from oracle import ans_oracle as O
from oracle import ypath_oracle as Y


@pytest.fixture(scope="module")
def cv(golden_dir):
    return np.load(os.path.join(golden_dir, "coder_vectors.npz"))


def _pair(cv, offsets_key="a_offsets", **kw):
    enc, dec = O.Rans64Encoder(**kw), O.Rans64Decoder(**kw)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv[offsets_key])
    return enc, dec


def test_rans64_tables_match_reference(cv):
    enc, _ = _pair(cv, bypass_coding=True)
    assert np.array_equal(enc.get_cdfs(), cv["a_cdfs"])


def test_rans64_stream_byte_exact_ans_test_shapes(cv):
    enc, dec = _pair(cv, bypass_coding=True)
    bs = enc.encode_with_indexes(cv["a_data"], cv["a_idx"])
    assert bs == cv["a_rans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["a_idx"]), cv["a_data"])


def test_rans64_escapes_negative_and_large(cv):
    enc, dec = _pair(cv, "b_offsets", bypass_coding=True)
    bs = enc.encode_with_indexes(cv["b_data"], cv["b_idx"])
    assert bs == cv["b_rans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["b_idx"]), cv["b_data"])


def test_rans64_decode_stream_in_pieces(cv):
    _, dec = _pair(cv, "b_offsets", bypass_coding=True)
    dec.set_stream(cv["b_rans"].tobytes())
    idx = cv["b_idx"]
    parts = [dec.decode_stream(idx[:1]), dec.decode_stream(idx[1:1234]), dec.decode_stream(idx[1234:])]
    assert np.array_equal(np.concatenate(parts), cv["b_data"])


def test_rans64_precision12_no_bypass(cv):
    enc, dec = O.Rans64Encoder(12, False, 4), O.Rans64Decoder(12, False, 4)
    for c in (enc, dec):
        c.init_params(cv["c_freqs"], cv["c_nsym"], cv["c_offsets"])
    assert np.array_equal(enc.get_cdfs(), cv["c_cdfs"])
    bs = enc.encode_with_indexes(cv["c_data"], cv["c_idx"])
    assert bs == cv["c_rans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["c_idx"]), cv["c_data"])


def test_out_of_range_without_bypass_is_an_error(cv):
    enc = O.Rans64Encoder(12, False, 4)
    enc.init_params(cv["c_freqs"], cv["c_nsym"], cv["c_offsets"])
    with pytest.raises(ValueError):
        enc.encode_with_indexes(np.array([1000], np.int32), np.array([0], np.int32))


def test_not_initialised_raises(cv):
    with pytest.raises(ValueError):
        O.Rans64Encoder().encode_with_indexes(np.zeros(4, np.int32), np.zeros(4, np.int32))
    with pytest.raises(ValueError):
        O.Rans64Encoder().init_params(np.zeros((2, 4), np.int32), np.zeros(3, np.int32), np.zeros(3, np.int32))


def test_pmf_to_quantized_cdf(cv):
    assert O.pmf_to_quantized_cdf(cv["f_pmf"], 16) == cv["f_cdf"].tolist()


def test_rcp_form_equals_division():
    """rans64.h:167-278: the exact-reciprocal update used by the CUDA lanes=1 encoder."""
    import ctypes as C
    lib = O.lib()
    rng = np.random.default_rng(1)
    for prec in (12, 16):
        for _ in range(4000):
            freq = int(rng.integers(1, (1 << prec)))
            start = int(rng.integers(0, (1 << prec) - freq + 1))
            lo, hi = freq << (31 - prec), (freq << (31 - prec)) << 32   # states valid before C(s, x)
            x = int(rng.integers(lo, hi)) if rng.random() < 0.9 else int(rng.choice([lo, hi - 1]))
            rcp, sh, bias, cmpl = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint32()
            lib.orc_rans64_rcp(C.c_uint32(start), C.c_uint32(freq), C.c_uint32(prec), C.byref(rcp), C.byref(sh),
                               C.byref(bias), C.byref(cmpl))
            got = lib.orc_rans64_put_rcp(x, rcp.value, sh.value, bias.value, cmpl.value)
            assert got == ((x // freq) << prec) + (x % freq) + start


def test_tans_byte_exact(cv):
    enc = O.TansEncoder(max_symbol_value=511, bypass_coding=True)
    dec = O.TansDecoder(max_symbol_value=511, bypass_coding=True)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv["a_offsets"])
    bs = enc.encode_with_indexes(cv["a_data"], cv["a_idx"])
    assert bs == cv["d_tans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["a_idx"]), cv["a_data"])


def test_tans_skewed_tables_log10(cv):
    enc = O.TansEncoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    dec = O.TansDecoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    for c in (enc, dec):
        c.init_params(cv["e_freqs"], cv["e_nsym"], cv["e_offsets"])
    bs = enc.encode_with_indexes(cv["e_data"], cv["e_idx"])
    assert bs == cv["e_tans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["e_idx"]), cv["e_data"])


def test_tans_limits_mirror_reference():
    """SURVEY hard part 8: table_log > 12 is refused by the decoder, nsym > 2^(table_log-1) by both."""
    f = np.ones((1, 600), np.int32) * 5
    with pytest.raises(ValueError, match="tableLog"):
        O.TansDecoder(table_log=16, max_symbol_value=255).init_params(f[:, :100], np.array([100]), np.array([0]))
    with pytest.raises(ValueError, match="generic"):
        O.TansEncoder(table_log=10, max_symbol_value=255).init_params(f, np.array([600]), np.array([0]))


# ---------------------------------------------------------------------------- Gaussian tables / y path
def test_gaussian_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "gaussian_tables.npz"))
    tab = Y.get_scale_table()
    assert np.array_equal(tab.numpy(), g["scale_table"])
    freqs, nsym, offs = Y.gaussian_ans_params(tab)
    assert np.array_equal(freqs, g["freqs"]) and np.array_equal(nsym, g["nsym"]) and np.array_equal(offs, g["offsets"])
    enc = O.Rans64Encoder()
    enc.init_params(freqs, nsym, offs)
    cd = enc.get_cdfs()
    flat = np.concatenate([cd[t, :nsym[t] + 2] for t in range(64)]).astype(np.int32)
    assert np.array_equal(flat, g["cdf_flat"])
    assert hashlib.sha256(flat.tobytes()).digest() == g["sha256"].tobytes()
    assert flat[:5].tolist() == [0, 1, 65534, 65535, 65536]


YCASES = ["ckbd", "meanscale", "cwckbd", "scanline", "learned_int", "raster", "learned_logits"]


def load_ycase(yv, name):
    C, G, B, H, W, ctx = [int(v) for v in yv[name + ".meta"]]
    sd = {k[len(name) + 4:]: torch.from_numpy(yv[k]) for k in yv.files if k.startswith(name + ".sd.")}
    if ctx:
        w = Y.weights_from_state_dict(sd)
    else:
        w = {"ctx_w": sd["context_prediction.weight"], "ctx_b": sd["context_prediction.bias"]}
    get = lambda k: yv[f"{name}.{k}"]
    return dict(C=C, G=G, B=B, H=H, W=W, w=w, y=torch.from_numpy(get("y")), prior=torch.from_numpy(get("prior")),
                tg=torch.from_numpy(get("tg")).long(), sym=get("sym"), idx=get("idx"), bytes=get("bytes").tobytes(),
                yhat=torch.from_numpy(get("yhat")), pgm=torch.from_numpy(get("pgm")) if f"{name}.pgm" in yv.files else None,
                params0=torch.from_numpy(get("params0")), params_full=torch.from_numpy(get("params_full")))


@pytest.fixture(scope="module")
def yv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_vectors.npz"))


@pytest.mark.parametrize("name", YCASES)
def test_ypath_matches_reference(yv, name):
    c = load_ycase(yv, name)
    with torch.no_grad():
        sym, idx, yhat = Y.encode_symbols(c["y"], c["prior"], c["tg"], c["w"], Y.get_scale_table())
    assert np.array_equal(sym, c["sym"]) and np.array_equal(idx, c["idx"])
    o = Y.YPathOracle(c["C"], c["G"], c["w"])
    o.update_state()
    with torch.no_grad():
        assert o.encode(c["y"], c["prior"], c["tg"]) == c["bytes"]
        assert torch.equal(o.decode(c["bytes"], c["prior"], c["tg"]), c["yhat"])
        assert torch.equal(Y.params_for(c["yhat"], c["tg"], c["prior"], c["w"]), c["params_full"])


ICASES = ["int_ckbd", "int_cwckbd", "int_raster"]


def load_icase(iv, name):
    C, G, B, H, W = [int(v) for v in iv[name + ".meta"]]
    sd = {k[len(name) + 4:]: torch.from_numpy(iv[k]) for k in iv.files if k.startswith(name + ".sd.")}
    get = lambda k: iv[f"{name}.{k}"]
    return dict(C=C, G=G, B=B, H=H, W=W, sd=sd, w=Y.weights_from_state_dict(sd, prefix=""), method=str(get("method")),
                y=torch.from_numpy(get("y")), prior=torch.from_numpy(get("prior")), tg=torch.from_numpy(get("tg")).long(),
                bytes=get("bytes").tobytes(), yhat=torch.from_numpy(get("yhat")), params_full=torch.from_numpy(get("params_full")))


@pytest.fixture(scope="module")
def iv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_internal_vectors.npz"))


@pytest.mark.parametrize("name", ICASES)
def test_internal_merger_matches_reference(iv, name):
    """the coder's own context_prediction + param_merger over 2G channel groups (pgm_coder.py:1177-1239,
    :1606-1638), golden vectors from the unmodified reference (tests/golden/make_internal_golden.py)."""
    c = load_icase(iv, name)
    o = Y.YPathOracle(c["C"], c["G"], c["w"])
    o.update_state()
    with torch.no_grad():
        assert torch.equal(Y.params_for(c["yhat"], c["tg"], c["prior"], c["w"]), c["params_full"])
        assert o.encode(c["y"], c["prior"], c["tg"]) == c["bytes"]
        assert torch.equal(o.decode(c["bytes"], c["prior"], c["tg"]), c["yhat"])


JCASES = ["jar_a", "jar_b"]


def load_jcase(jv, name):
    C, B, H, W = [int(v) for v in jv[name + ".meta"]]
    sd = {k[len(name) + 4:]: torch.from_numpy(jv[k]) for k in jv.files if k.startswith(name + ".sd.")}
    get = lambda k: jv[f"{name}.{k}"]
    return dict(C=C, B=B, H=H, W=W, sd=sd, w=Y.joint_ar_weights_from_state_dict(sd), y=torch.from_numpy(get("y")),
                prior=torch.from_numpy(get("prior")), bytes=get("bytes").tobytes(), yhat=torch.from_numpy(get("yhat")))


@pytest.fixture(scope="module")
def jv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_jointar_vectors.npz"))


@pytest.mark.parametrize("name", JCASES)
def test_joint_ar_matches_reference(jv, name):
    """use_joint_ar_model_impl=True (pixel-by-pixel CompressAI-style coder), golden vectors from the
    unmodified reference (tests/golden/make_jointar_golden.py)."""
    c = load_jcase(jv, name)
    o = Y.JointAROracle(c["C"], 1, c["w"])
    o.update_state()
    with torch.no_grad():
        assert o.encode(c["y"], c["prior"]) == c["bytes"]
        assert torch.equal(o.decode(c["bytes"], c["prior"]), c["yhat"])


def test_internal_merger_expanded_bottleneck_matches_reference(golden_dir):
    """param_merger_expand_bottleneck=True (4C -> 8C -> 8C -> 4C, pgm_coder.py:1215): the oracle against the reference."""
    ev = np.load(os.path.join(golden_dir, "ypath_internal_expand_vectors.npz"))
    c = load_icase(ev, "int_expand")
    assert tuple(c["w"]["pm2_w"].shape[:2]) == (8 * c["C"], 8 * c["C"])
    o = Y.YPathOracle(c["C"], c["G"], c["w"])
    o.update_state()
    with torch.no_grad():
        assert torch.equal(Y.params_for(c["yhat"], c["tg"], c["prior"], c["w"]), c["params_full"])
        assert o.encode(c["y"], c["prior"], c["tg"]) == c["bytes"]
        assert torch.equal(o.decode(c["bytes"], c["prior"], c["tg"]), c["yhat"])


def test_group_maps(yv):
    for name, method in [("ckbd", "checkerboard"), ("cwckbd", "channelwise-checkerboard"), ("scanline", "scanline"),
                         ("raster", "raster2x2"), ("meanscale", "none")]:
        c = load_ycase(yv, name)
        assert torch.equal(Y.default_pgm(method, c["G"], c["H"], c["W"]), c["tg"])
    for name in ("learned_int", "learned_logits"):   # 2x2 patches tiled with fold; odd sizes leave group 0
        c = load_ycase(yv, name)
        assert torch.equal(Y.tile_pgm(c["pgm"], c["G"], c["H"], c["W"]), c["tg"])


# ------------------------------------------------------------------------ multi-lane format (CPU spec)
def test_lanes_slices_round_trip(cv):
    """Multi-slice segments (one slice per coding group, lane states carried across slices): lossless, a single
    slice is the plain segment, and the flush overhead is paid once per chunk -- not once per slice."""
    enc, dec = O.Rans64Encoder(bypass_coding=True), O.Rans64Decoder(bypass_coding=True)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv["b_offsets"])
    data, idx = cv["b_data"].reshape(-1), cv["b_idx"].reshape(-1)
    n = data.size
    for slices in ([n], [n // 3, n - n // 3], [100, 0, n - 1100, 1000], [1] * 5 + [n - 5]):
        for nc in (1, 3, 16):
            bs = enc.encode_lanes_slices(data, idx, slices, nc)
            out, used = dec.decode_lanes_slices(bs, idx, slices)
            assert used == len(bs) and np.array_equal(out, data)
    one = enc.encode_lanes_slices(data, idx, [n], 4)
    chunk = struct.unpack_from("<I", one, 8)[0]
    assert one == enc.encode_lanes(data, idx, chunk)
    ref = len(enc.encode_with_indexes(data, idx))
    many = enc.encode_lanes_slices(data, idx, [n // 8] * 7 + [n - 7 * (n // 8)], 4)
    assert len(many) <= ref + 4 * 136 + 8 * 4 + 16 + 8 * 4 * 64     # + idle tail blocks cost nothing but word rounding


@pytest.mark.parametrize("bp", [1, 2, 3, 4, 5, 8])
def test_lanes_escape_units_round_trip(cv, bp):
    """Escape units of the multi-lane format for every bypass precision: lossless, and the same number of escape bits
    as the token-by-token code of the lanes=1 stream (size within the flush overhead)."""
    rng = np.random.default_rng(bp)
    n = 5000
    idx = rng.integers(0, 8, n).astype(np.int32)
    sym = rng.integers(-600, 600, n).astype(np.int32)
    sym[::5] = rng.integers(-(2 ** 30), 2 ** 30, sym[::5].size)
    enc, dec = O.Rans64Encoder(bypass_coding=True, bypass_precision=bp), O.Rans64Decoder(bypass_coding=True, bypass_precision=bp)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv["b_offsets"])
    bs = enc.encode_lanes(sym, idx, 512)
    out, used = dec.decode_lanes(bs, idx, 512)
    assert used == len(bs) and np.array_equal(out, sym)
    ref = len(enc.encode_with_indexes(sym, idx))
    assert len(bs) <= ref + (-(-n // 512)) * 136 + 16


@pytest.mark.parametrize("chunk", [128, 512, 4096])
def test_lanes_round_trip_and_size(cv, chunk):
    enc, dec = _pair(cv, "b_offsets", bypass_coding=True)
    for data, idx in ((cv["b_data"], cv["b_idx"]), (cv["a_data"].reshape(-1)[:5001], cv["a_idx"].reshape(-1)[:5001])):
        if data is cv["a_data"] or data.base is cv["a_data"]:
            enc, dec = _pair(cv, bypass_coding=True)
        bs = enc.encode_lanes(data, idx, chunk)
        out, used = dec.decode_lanes(bs, idx, chunk)
        assert np.array_equal(out, data) and used == len(bs)
        ref = len(enc.encode_with_indexes(data, idx))
        nchunks = -(-data.size // chunk)
        assert len(bs) <= ref + nchunks * 136 + 16      # per chunk: 128 B states + 4 B count (+ word rounding)

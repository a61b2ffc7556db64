"""-m gpu: the CUDA coder (through the C ABI, via the cbench.ans-shaped shim) against the CPU oracle and the
golden vectors of the unmodified reference.  Bar: bit-exact (integer / byte work)."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import ans_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from cbench_basic_b200 import ans
    return ans


@pytest.fixture(scope="module")
def cv(golden_dir):
    return np.load(os.path.join(golden_dir, "coder_vectors.npz"))


@pytest.fixture(scope="module")
def gauss(golden_dir):
    return np.load(os.path.join(golden_dir, "gaussian_tables.npz"))


def _mk(A, cv, offsets="a_offsets", lanes=1, **kw):
    enc, dec = A.Rans64Encoder(lanes=lanes, **kw), A.Rans64Decoder(lanes=lanes, **kw)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv[offsets])
    return enc, dec


def _gauss_pair(A, gauss, lanes):
    enc, dec = A.Rans64Encoder(lanes=lanes), A.Rans64Decoder(lanes=lanes)
    oenc, odec = O.Rans64Encoder(), O.Rans64Decoder()
    for c in (enc, dec, oenc, odec):
        c.init_params(gauss["freqs"], gauss["nsym"], gauss["offsets"])
    return enc, dec, oenc, odec


def _gauss_data(gauss, n, seed, geometric=False):
    rng = np.random.default_rng(seed)
    idx = (np.minimum(rng.geometric(0.12, n) - 1, 63) if geometric else rng.integers(0, 64, n)).astype(np.int32)
    sym = np.rint(rng.standard_normal(n) * gauss["scale_table"][idx]).astype(np.int32)
    return sym, idx


# ----------------------------------------------------------------------------------------------- tables
def test_tables_built_on_device_match_reference(A, cv, gauss):
    enc, _ = _mk(A, cv)
    assert np.array_equal(enc.get_cdfs(), cv["a_cdfs"])
    g = A.Rans64Encoder()
    g.init_params(gauss["freqs"], gauss["nsym"], gauss["offsets"])
    cd = g.get_cdfs()
    assert cd.shape == (64, 2219)
    flat = np.concatenate([cd[t, :gauss["nsym"][t] + 2] for t in range(64)])
    assert np.array_equal(flat, gauss["cdf_flat"])
    for t in range(64):
        assert not cd[t, gauss["nsym"][t] + 2:].any()
    e12 = A.Rans64Encoder(12, False, 4)
    e12.init_params(cv["c_freqs"], cv["c_nsym"], cv["c_offsets"])
    assert np.array_equal(e12.get_cdfs(), cv["c_cdfs"])


def test_pmf_to_quantized_cdf(A, cv):
    assert A.pmf_to_quantized_cdf(cv["f_pmf"], 16) == cv["f_cdf"].tolist()


def test_init_cdf_params_round_trip(A, cv):
    enc, dec = A.Rans64Encoder(), A.Rans64Decoder()
    sizes = cv["a_nsym"] + 2
    for c in (enc, dec):
        c.init_cdf_params(cv["a_cdfs"], sizes, cv["a_offsets"])
    assert np.array_equal(enc.get_cdfs(), cv["a_cdfs"])
    bs = enc.encode_with_indexes(cv["a_data"], cv["a_idx"])
    assert bs == cv["a_rans"].tobytes()


def test_errors_mirror_reference(A, cv):
    with pytest.raises(ValueError, match="not initialized"):
        A.Rans64Encoder().encode_with_indexes(np.zeros(4, np.int32), np.zeros(4, np.int32))
    with pytest.raises(ValueError, match="freqs should be 2-dimensional"):
        A.Rans64Encoder().init_params(np.zeros((2, 4), np.int32), np.zeros(3, np.int32), np.zeros(3, np.int32))
    with pytest.raises(ValueError, match="cdfs should be 2-dimensional"):
        A.Rans64Encoder().init_cdf_params(np.zeros((2, 4), np.int32), np.zeros(3, np.int32), np.zeros(3, np.int32))
    enc = A.Rans64Encoder(12, False, 4)
    enc.init_params(cv["c_freqs"], cv["c_nsym"], cv["c_offsets"])
    with pytest.raises(ValueError):   # out of range with bypass off: the reference has UB, we refuse
        enc.encode_with_indexes(np.array([1000], np.int32), np.array([0], np.int32))
    with pytest.raises(ValueError):   # index beyond the tables
        enc.encode_with_indexes(np.array([0], np.int32), np.array([99], np.int32))


# -------------------------------------------------------------------------------- lanes = 1: byte exact
def test_lanes1_byte_exact_ans_test_shapes(A, cv):
    enc, dec = _mk(A, cv)
    bs = enc.encode_with_indexes(cv["a_data"], cv["a_idx"])
    assert bs == cv["a_rans"].tobytes()
    out = dec.decode_with_indexes(cv["a_rans"].tobytes(), cv["a_idx"])
    assert out.shape == cv["a_idx"].shape and out.dtype == np.int32 and np.array_equal(out, cv["a_data"])


def test_lanes1_escapes_and_stream_api(A, cv):
    enc, dec = _mk(A, cv, "b_offsets")
    assert enc.encode_with_indexes(cv["b_data"], cv["b_idx"]) == cv["b_rans"].tobytes()
    dec.set_stream(cv["b_rans"].tobytes())
    idx = cv["b_idx"]
    parts = [dec.decode_stream(idx[:1]), dec.decode_stream(idx[1:1234]), dec.decode_stream(idx[1234:])]
    assert np.array_equal(np.concatenate(parts), cv["b_data"])


def test_lanes1_precision12_no_bypass(A, cv):
    enc, dec = A.Rans64Encoder(12, False, 4), A.Rans64Decoder(12, False, 4)
    for c in (enc, dec):
        c.init_params(cv["c_freqs"], cv["c_nsym"], cv["c_offsets"])
    bs = enc.encode_with_indexes(cv["c_data"], cv["c_idx"])
    assert bs == cv["c_rans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["c_idx"]), cv["c_data"])


def test_lanes1_gaussian_tables_vs_oracle_and_cuda_tensors(A, gauss):
    enc, dec, oenc, _ = _gauss_pair(A, gauss, 1)
    sym, idx = _gauss_data(gauss, 200_000, 3)
    ref = oenc.encode_with_indexes(sym, idx)
    assert enc.encode_with_indexes(sym, idx) == ref
    ts, ti = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
    assert enc.encode_with_indexes(ts, ti) == ref                       # device-resident operands
    out = dec.decode_with_indexes(ref, ti)
    assert out.is_cuda and torch.equal(out.cpu(), torch.from_numpy(sym))


def test_lanes1_empty_and_single(A, cv):
    enc, dec = _mk(A, cv)
    e = np.zeros(0, np.int32)
    oenc = O.Rans64Encoder(bypass_coding=True)
    oenc.init_params(cv["a_freqs"], cv["a_nsym"], cv["a_offsets"])
    assert enc.encode_with_indexes(e, e) == oenc.encode_with_indexes(e, e)
    one = np.array([7], np.int32), np.array([3], np.int32)
    bs = enc.encode_with_indexes(*one)
    assert bs == oenc.encode_with_indexes(*one)
    assert dec.decode_with_indexes(bs, one[1]).tolist() == [7]


def test_cache_and_flush_lanes1(A, cv):
    enc, _ = _mk(A, cv, "b_offsets")
    d, i = cv["b_data"], cv["b_idx"]
    assert enc.encode_with_indexes(d[:1500], i[:1500], cache=True) == b""
    assert enc.encode_with_indexes(d[1500:], i[1500:], cache=True) == b""
    assert enc.flush() == cv["b_rans"].tobytes()      # == one stream over the concatenation (what _encode_with_pgm builds)


def test_lanes1_batch_of_streams(A, cv, gauss):
    """encode_batch / decode_batch: B independent reference streams in one launch (one CTA per stream) = what B single calls
    give, byte for byte; numpy and CUDA-tensor operands; escapes; an empty batch."""
    enc, dec = _mk(A, cv, "b_offsets")
    data, idx = cv["a_data"].reshape(6, -1)[:, :3000].copy(), cv["a_idx"].reshape(6, -1)[:, :3000].copy()
    data[2, ::50] = 100_000
    data[4, 1::70] = -70_000
    singles = [enc.encode_with_indexes(data[b], idx[b]) for b in range(6)]
    assert enc.encode_batch(data, idx) == singles
    assert np.array_equal(dec.decode_batch(singles, idx), data)
    g_enc, g_dec, oenc, _ = _gauss_pair(A, gauss, 1)
    sym, ix = _gauss_data(gauss, 24 * 18432, 3, geometric=True)
    sym, ix = sym.reshape(24, -1), ix.reshape(24, -1)
    bs = g_enc.encode_batch(torch.from_numpy(sym).cuda(), torch.from_numpy(ix).cuda())
    assert bs == [oenc.encode_with_indexes(sym[b], ix[b]) for b in range(24)]
    out = g_dec.decode_batch(bs, torch.from_numpy(ix).cuda())
    assert out.is_cuda and torch.equal(out.cpu(), torch.from_numpy(sym))
    assert enc.encode_batch(np.zeros((0, 5), np.int32), np.zeros((0, 5), np.int32)) == []
    with pytest.raises(ValueError):
        dec.decode_batch([singles[0][:-4]], idx[:1])               # a word short: truncated stream


# ------------------------------------------------------------------------ multi-lane: format + lossless
@pytest.mark.parametrize("lanes", [32, 64, 1000, 4096])
def test_multilane_matches_cpu_spec_and_is_lossless(A, cv, lanes):
    for off, data, idx in (("b_offsets", cv["b_data"], cv["b_idx"]),
                           ("a_offsets", cv["a_data"].reshape(-1)[:18001], cv["a_idx"].reshape(-1)[:18001])):
        enc, dec = _mk(A, cv, off, lanes=lanes)
        oenc = O.Rans64Encoder(bypass_coding=True)
        oenc.init_params(cv["a_freqs"], cv["a_nsym"], cv[off])
        bs = enc.encode_with_indexes(data, idx)
        magic, n_chunks, n_slices, chunk = struct.unpack_from("<IIII", bs, 0)
        assert magic == 0x31534C42 and n_slices == 1 and chunk % 128 == 0 and n_chunks == -(-data.size // chunk)
        assert n_chunks <= -(-lanes // 32)
        assert bs[4:] == oenc.encode_lanes(data, idx, chunk)       # byte-identical to the CPU specification
        assert np.array_equal(dec.decode_with_indexes(bs, idx), data)


def test_multilane_auto_bpp_within_half_percent(A, gauss):
    for geometric, n in ((False, 7_077_888), (True, 7_077_888), (False, 294_912)):
        enc, dec, oenc, _ = _gauss_pair(A, gauss, 0)
        sym, idx = _gauss_data(gauss, n, 5, geometric)
        ref = oenc.encode_with_indexes(sym, idx)
        ts, ti = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
        bs = enc.encode_with_indexes(ts, ti)
        assert len(bs) <= len(ref) * 1.005 + 160, (len(bs), len(ref))
        assert torch.equal(dec.decode_with_indexes(bs, ti).cpu(), torch.from_numpy(sym))


def test_multilane_segments_cache_flush_decode_stream(A, gauss):
    enc, dec, _, _ = _gauss_pair(A, gauss, 256)
    sym, idx = _gauss_data(gauss, 50_000, 9)
    cuts = [0, 1, 12_345, 12_345, 50_000]          # includes an empty segment and a 1-symbol segment
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert enc.encode_with_indexes(sym[a:b], idx[a:b], cache=True) == b""
    bs = enc.flush()
    dec.set_stream(bs)
    parts = [dec.decode_stream(idx[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), sym)


def test_multilane_escape_heavy_and_ragged(A, cv):
    rng = np.random.default_rng(11)
    for n in (1, 31, 127, 128, 129, 4097):
        enc, dec = _mk(A, cv, "b_offsets", lanes=64)
        idx = rng.integers(0, 8, n).astype(np.int32)
        sym = rng.integers(-5000, 5000, n).astype(np.int32)        # mostly escapes
        sym[::3] = rng.integers(-(2 ** 26), 2 ** 26, sym[::3].size)
        bs = enc.encode_with_indexes(sym, idx)
        assert np.array_equal(dec.decode_with_indexes(bs, idx), sym)


@pytest.mark.parametrize("bp", [1, 2, 3, 4, 5, 8])
def test_multilane_escape_units_any_bypass_precision(A, cv, bp):
    """Escape tokens travel in units of floor(16 / bypass_precision) tokens (oracle/ans_oracle.c section 4): every unit
    width, one-unit and many-unit escapes (|value| up to 2^30), byte-identical to the CPU specification and lossless."""
    rng = np.random.default_rng(100 + bp)
    n = 6000
    idx = rng.integers(0, 8, n).astype(np.int32)
    sym = rng.integers(-600, 600, n).astype(np.int32)
    sym[::5] = rng.integers(-(2 ** 30), 2 ** 30, sym[::5].size)
    sym[1::7] = rng.integers(-40000, 40000, sym[1::7].size)
    enc, dec = _mk(A, cv, "b_offsets", lanes=96, bypass_precision=bp)
    oenc = O.Rans64Encoder(bypass_coding=True, bypass_precision=bp)
    oenc.init_params(cv["a_freqs"], cv["a_nsym"], cv["b_offsets"])
    bs = enc.encode_with_indexes(sym, idx)
    chunk = struct.unpack_from("<I", bs, 12)[0]
    assert bs[4:] == oenc.encode_lanes(sym, idx, chunk)
    assert np.array_equal(dec.decode_with_indexes(bs, idx), sym)


@pytest.mark.parametrize("bp", [4, 2])
def test_multilane_tables_larger_than_shared_memory(A, bp):
    """400 tables x 512 symbols = 411 KB of CDFs: the table image does not fit an SM's shared memory, the kernels read it
    through L2 instead (rans_lanes.cu Tab<false>; both escape instantiations).  Same bytes as the CPU specification."""
    rng = np.random.default_rng(77 + bp)
    T, M, n = 400, 512, 40_000
    freqs = rng.integers(1, 1024, (T, M)).astype(np.int32)
    nsym, offs = np.full(T, M, np.int32), rng.integers(-300, 1, T).astype(np.int32)
    idx = rng.integers(0, T, n).astype(np.int32)
    sym = (rng.integers(0, M + 40, n) + offs[idx] - 20).astype(np.int32)       # a few percent escapes on both sides
    sym[::97] = rng.integers(-(2 ** 28), 2 ** 28, sym[::97].size)
    enc, dec = A.Rans64Encoder(lanes=640, bypass_precision=bp), A.Rans64Decoder(lanes=640, bypass_precision=bp)
    oenc = O.Rans64Encoder(bypass_coding=True, bypass_precision=bp)
    for c in (enc, dec, oenc):
        c.init_params(freqs, nsym, offs)
    bs = enc.encode_with_indexes(sym, idx)
    chunk = struct.unpack_from("<I", bs, 12)[0]
    assert bs[4:] == oenc.encode_lanes(sym, idx, chunk)
    assert np.array_equal(dec.decode_with_indexes(bs, idx), sym)
    e1, d1 = A.Rans64Encoder(lanes=1, bypass_precision=bp), A.Rans64Decoder(lanes=1, bypass_precision=bp)
    for c in (e1, d1):
        c.init_params(freqs, nsym, offs)
    ref = oenc.encode_with_indexes(sym, idx)
    assert e1.encode_with_indexes(sym, idx) == ref and np.array_equal(d1.decode_with_indexes(ref, idx), sym)


def test_multilane_rejects_garbage(A, gauss):
    from cbench_basic_b200 import _native
    _, dec, _, _ = _gauss_pair(A, gauss, 64)
    idx = np.zeros(1000, np.int32)
    with pytest.raises(ValueError):
        dec.decode_with_indexes(b"\x00" * 64, idx)
    enc, _, _, _ = _gauss_pair(A, gauss, 64)
    bs = enc.encode_with_indexes(np.zeros(1000, np.int32), idx)
    with pytest.raises(ValueError):
        dec.decode_with_indexes(bs[:len(bs) // 2], idx)            # truncated
    with pytest.raises(ValueError):
        dec.decode_with_indexes(bs, idx[:500])                     # wrong symbol count


# --------------------------------------------------------------------------------------------------- tANS
def test_tans_byte_exact(A, cv):
    enc = A.TansEncoder(max_symbol_value=511, bypass_coding=True)
    dec = A.TansDecoder(max_symbol_value=511, bypass_coding=True)
    for c in (enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv["a_offsets"])
    bs = enc.encode_with_indexes(cv["a_data"], cv["a_idx"])
    assert bs == cv["d_tans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["a_idx"]), cv["a_data"])


def test_tans_skewed_log10(A, cv):
    enc = A.TansEncoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    dec = A.TansDecoder(table_log=10, max_symbol_value=255, bypass_coding=False)
    for c in (enc, dec):
        c.init_params(cv["e_freqs"], cv["e_nsym"], cv["e_offsets"])
    bs = enc.encode_with_indexes(cv["e_data"], cv["e_idx"])
    assert bs == cv["e_tans"].tobytes()
    assert np.array_equal(dec.decode_with_indexes(bs, cv["e_idx"]), cv["e_data"])


def test_tans_limits_mirror_reference(A):
    f = np.ones((1, 600), np.int32) * 5
    with pytest.raises(ValueError, match="tableLog"):
        A.TansDecoder(table_log=16, max_symbol_value=255).init_params(f[:, :100], np.array([100]), np.array([0]))
    with pytest.raises(ValueError, match="generic"):
        A.TansEncoder(table_log=10, max_symbol_value=255).init_params(f, np.array([600]), np.array([0]))


# ----------------------------------------------------------------------------------------------- several GPUs, one process
def test_coders_on_two_devices_in_one_process(A, gauss):
    """The opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: coders created on cuda:0 and
    cuda:1 by the same process must both launch (tables image 155 KB) and agree byte for byte."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sym, idx = _gauss_data(gauss, 50_000, seed=5, geometric=True)
    out = []
    for dev in (0, 1):
        for lanes in (1, 0):
            enc, dec = A.Rans64Encoder(lanes=lanes, device=dev), A.Rans64Decoder(lanes=lanes, device=dev)
            for c in (enc, dec):
                c.init_params(gauss["freqs"], gauss["nsym"], gauss["offsets"])
            bs = enc.encode_with_indexes(sym, idx)
            assert np.array_equal(dec.decode_with_indexes(bs, idx), sym)
            out.append(bs)
    assert out[0] == out[2] and out[1] == out[3]


# ----------------------------------------------------------------------------------------------- in-coder AR lookup (row f3)
@pytest.mark.parametrize("order", [1, 2])
def test_ar_table_lookup_matches_reference(A, cv, order):
    """ans_interface.hpp:58-105 (table branch): the table of element i is ar_tables[ar_index][index][v0]([v1]) with
    v_k = off_k[i] > 0 ? symbol[i - off_k[i]] + 1 : 0.  Stream byte-identical to the unmodified reference coder's (oracle/_ref),
    both directions, and each side decodes the other's stream."""
    from oracle import ref_loader
    R = ref_loader.load("ans")
    if R is None:
        pytest.skip("oracle/_ref/ans not built")
    rng = np.random.default_rng(11 + order)
    n, T, hi = 20000, 8, 40
    sym = rng.integers(0, hi, n).astype(np.int32)
    idx = rng.integers(0, T, n).astype(np.int32)
    ar_idx = rng.integers(0, 2, n).astype(np.int32)
    shape = (2, T) + (hi + 1,) * order
    ar_tables = rng.integers(0, T, shape).astype(np.int32)
    off = np.stack([np.minimum(rng.choice([0, 1, 2, 7], n), np.arange(n)) for _ in range(order)]).astype(np.int32)
    ar_offsets_init = np.zeros((2, order, 2), dtype=np.int32)     # stored, never read by the reference's coding calls
    renc, rdec = R.Rans64Encoder(16, True, 4), R.Rans64Decoder(16, True, 4)
    enc, dec = A.Rans64Encoder(lanes=1), A.Rans64Decoder(lanes=1)
    for c in (renc, rdec, enc, dec):
        c.init_params(cv["a_freqs"], cv["a_nsym"], cv["a_offsets"])
        c.init_ar_params(ar_tables, ar_offsets_init)
    ref_bytes = renc.encode_with_indexes(sym, idx, ar_idx, off, False)
    got = enc.encode_with_indexes(sym, idx, ar_idx, off)
    assert got == ref_bytes
    assert np.array_equal(dec.decode_with_indexes(ref_bytes, idx, ar_idx, off), sym)
    assert np.array_equal(rdec.decode_with_indexes(got, idx, ar_idx, off), sym)
    # device-resident operands, no ar_indexes (= table set 0)
    ts, ti, to = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(off).cuda()
    assert enc.encode_with_indexes(ts, ti, None, to) == renc.encode_with_indexes(sym, idx, None, off, False)
    with pytest.raises(ValueError):
        enc.encode_with_indexes(sym, idx, ar_idx, None)           # "ar_offsets is required for ar coding!"
    bad = ar_tables.copy()
    bad[0, 0, 0] = T + 3                                          # a table entry past the coder's tables: an error, not UB
    enc.init_ar_params(bad, ar_offsets_init)
    with pytest.raises(ValueError):
        enc.encode_with_indexes(sym, idx, np.zeros(n, np.int32), off)

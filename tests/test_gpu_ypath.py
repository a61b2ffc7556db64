"""-m gpu: Gaussian conditional, context model and the whole y path (CUDA, through the C ABI) against the CPU
oracle and the golden vectors of the unmodified reference.

Bars (BASELINE.json north_star): symbols, scale indexes, CDF tables and decoded latents bit-exact; means and
scales within 1e-5 relative error before quantisation; lanes=1 streams byte-for-byte; multi-lane lossless and
within 0.5 % of the lanes=1 size."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import ypath_oracle as Y
from tests.test_oracle_golden import YCASES, load_ycase

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5      # north_star: floating-point means / scales, relative error before quantisation


@pytest.fixture(scope="module")
def yv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_vectors.npz"))


def make_coder(c, lanes, method="none", ctx_precision="fp32"):
    """ctx_precision="fp32": the parity mode (exact FP32 context model) -- what the oracle comparisons below pin; the
    tensor-core mode has its own tests (tests/test_gpu_ctx_tc.py)."""
    from cbench_basic_b200.prior_coder import (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder,
                                               TopoGroupDynamicMaskConv2dContextModel as Ctx)
    w, C_, G = c["w"], c["C"], c["G"]
    if "m1_w" in w:
        cm = Ctx(in_channels=C_, out_channels=2 * C_)
        sd = {"context_prediction.weight": w["ctx_w"], "context_prediction.bias": w["ctx_b"],
              "param_merger_in.weight": w["m1_w"], "param_merger_in.bias": w["m1_b"],
              "param_merger_out.1.weight": w["m2_w"], "param_merger_out.1.bias": w["m2_b"],
              "param_merger_out.3.weight": w["m3_w"], "param_merger_out.3.bias": w["m3_b"]}
        cm.load_state_dict(sd)
        coder = Coder(in_channels=C_, channel_groups=G, default_topo_group_method=method, topo_group_context_model=cm,
                      lanes=lanes, ans_params_device="cpu", ctx_precision=ctx_precision)
    else:
        coder = Coder(in_channels=C_, channel_groups=G, default_topo_group_method=method, use_param_merger=False,
                      lanes=lanes, ans_params_device="cpu", ctx_precision=ctx_precision)
        coder.load_state_dict({"context_prediction.weight": w["ctx_w"], "context_prediction.bias": w["ctx_b"]})
    coder = coder.cuda().eval()
    coder.update_state()
    return coder


def rel_err(a, b):
    """|a - b| relative to max(1, |b|): relative for large values, absolute 1e-5 below 1."""
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


def assert_latents_match(got, ref):
    """Decoded latents are symbol + mean: the integer symbol must be the reference's, the float mean may differ by
    the rounding of a different (deterministic) summation order -- north_star's 1e-5 relative bar."""
    assert rel_err(got, ref) <= REL_TOL, rel_err(got, ref)
    assert torch.equal(torch.round(got - ref), torch.zeros_like(ref))


# ------------------------------------------------------------------------------------- update_state
def test_update_state_tables_match_reference(golden_dir):
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder
    g = np.load(os.path.join(golden_dir, "gaussian_tables.npz"))
    coder = Coder(in_channels=8, use_param_merger=False, ans_params_device="cpu").cuda()
    assert np.array_equal(coder.scale_table.numpy(), g["scale_table"])
    freqs, nsym, offs = coder._get_ans_params()
    assert np.array_equal(freqs, g["freqs"]) and np.array_equal(nsym, g["nsym"]) and np.array_equal(offs, g["offsets"])
    coder.update_state()
    cd = coder.ans_encoder.get_cdfs()
    flat = np.concatenate([cd[t, :nsym[t] + 2] for t in range(64)])
    assert np.array_equal(flat, g["cdf_flat"])
    # the same float32 evaluation on the GPU (what a GPU-resident reference would do) may differ from the CPU
    # libm in the last ulp of a few truncated counts; alphabet sizes and offsets must not
    coder2 = Coder(in_channels=8, use_param_merger=False).cuda()
    f2, n2, o2 = coder2._get_ans_params()
    assert np.array_equal(n2, nsym) and np.array_equal(o2, offs)
    assert np.abs(f2.astype(np.int64) - freqs).max() <= 1


# --------------------------------------------------------------------------- quantise + scale index
def test_quantize_index_bit_exact():
    from cbench_basic_b200 import _native as N, ans
    torch.manual_seed(0)
    B, C_, H, W = 3, 16, 9, 11
    tab = Y.get_scale_table()
    y = 3 * torch.randn(B, C_, H, W)
    params = torch.randn(B, 2 * C_, H, W) * 2
    sc = params[:, 1::2]
    # stress the argmin: exact table entries, exact midpoints (ties -> lower index), negatives, huge, tiny
    flat = sc.reshape(-1)
    mids = ((tab[:-1].double() + tab[1:].double()) / 2).float()
    flat[:64] = tab
    flat[64:127] = mids
    flat[127:190] = torch.nextafter(mids, torch.tensor(float("inf")))
    flat[190:253] = torch.nextafter(mids, torch.tensor(0.0))
    flat[253:258] = torch.tensor([-1.0, 0.0, 1e9, 0.1099999, 256.0])
    ym = y - params[:, 0::2]
    y.reshape(-1)[:400] = (params[:, 0::2].reshape(-1)[:400] + torch.arange(400).float() % 7 - 3 + 0.5)  # exact .5 ties
    enc = ans.Rans64Encoder()
    N.check(N.lib().basic_coder_set_scale_table(enc.handle, np.ascontiguousarray(tab.numpy()).ctypes.data, 64))
    pos = torch.randperm(C_ * H * W)[:1000].sort().values.int()
    dy, dp, dpos = y.cuda(), params.cuda(), pos.cuda()
    n = B * pos.numel()
    sym, idx = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    buf = torch.zeros_like(dy)
    N.check(N.lib().basic_gauss_quantize_index(enc.handle, dy.data_ptr(), dp.data_ptr(), dpos.data_ptr(), pos.numel(), B, C_,
                                               H * W, sym.data_ptr(), idx.data_ptr(), buf.data_ptr(), 0))
    torch.cuda.synchronize()
    mean, scale = Y.split_mean_scale(params)
    p = pos.long()
    mean_s, scale_s, y_s = (t.reshape(B, -1)[:, p] for t in (mean, scale, y))
    ref_idx = Y.select_indexes(scale_s, tab)
    ref_sym = torch.round(y_s - mean_s)
    assert torch.equal(idx.cpu().reshape(B, -1).long(), ref_idx)
    assert torch.equal(sym.cpu().reshape(B, -1).float(), ref_sym)
    ref_buf = torch.zeros(B, C_ * H * W)
    ref_buf[:, p] = ref_sym + mean_s
    assert torch.equal(buf.cpu().reshape(B, -1), ref_buf)
    # decoder side
    out = torch.zeros_like(dy)
    N.check(N.lib().basic_gauss_dequantize(enc.handle, sym.data_ptr(), dp.data_ptr(), dpos.data_ptr(), pos.numel(), B, C_, H * W,
                                           out.data_ptr(), 0))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().reshape(B, -1), ref_buf * 1.0 + 0.0)


# ----------------------------------------------------------------------------------- context model
@pytest.mark.parametrize("name", [n for n in YCASES])
def test_context_model_params_within_tolerance(yv, name):
    """Every stage's distribution parameters, computed cell by cell on the GPU, against the reference's
    full-tensor evaluation on the final y_hat (golden `params_full`): groups < g are all a cell may see, so the
    two agree wherever the cell is coded."""
    from cbench_basic_b200 import _native as N
    c = load_ycase(yv, name)
    coder = make_coder(c, lanes=1)
    coder._set_map(c["tg"])
    B, C_, H, W = c["B"], c["C"], c["H"], c["W"]
    buf, prior = c["yhat"].cuda().contiguous(), c["prior"].cuda().contiguous()
    params = torch.full((B, 2 * C_, H, W), float("nan"), device="cuda")
    S = N.lib().basic_ctx_num_stages(coder._ctx)
    assert S == int(c["tg"].max()) + 1
    for g in range(S):
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, buf.data_ptr(), prior.data_ptr(), B, params.data_ptr(), 0))
    torch.cuda.synchronize()
    got = params.cpu()
    assert not torch.isnan(got).any()            # every cell belongs to exactly one stage
    assert rel_err(got, c["params_full"]) <= REL_TOL, rel_err(got, c["params_full"])


@pytest.mark.parametrize("name", YCASES)
@pytest.mark.parametrize("lanes", [1, 0, 64])
def test_ypath_golden(yv, name, lanes):
    c = load_ycase(yv, name)
    coder = make_coder(c, lanes=lanes)
    kw = dict(prior=c["prior"].cuda(), pgm=c["tg"])
    bs, yhat_enc = coder.encode(c["y"].cuda(), return_yhat=True, **kw)
    yhat = coder.decode(bs, **kw)
    assert torch.equal(yhat, yhat_enc * 1.0 + 0.0)                 # lossless: decoder reproduces the encoder bit for bit
    assert_latents_match(yhat.cpu(), c["yhat"])
    if c["w"].get("m1_w") is None:
        assert torch.equal(yhat.cpu(), c["yhat"])                  # no float reduction involved: bit-exact vs the reference
    if lanes == 1:
        assert bs == c["bytes"]                                    # reference bitstream byte for byte (=> symbols and
        assert_latents_match(coder.decode(c["bytes"], **kw).cpu(), c["yhat"])   # scale indexes identical everywhere)
    else:
        assert len(bs) <= len(c["bytes"]) + 140 * (int(c["tg"].max()) + 1) * (1 if lanes == 0 else 2) + 8


@pytest.mark.parametrize("name", ["int_ckbd", "int_cwckbd", "int_raster"])
@pytest.mark.parametrize("lanes", [1, 0])
def test_internal_merger_golden(iv, name, lanes):
    """the coder's own context_prediction + param_merger (2G channel groups, no topo_group_context_model),
    a reference state_dict loaded as it is.  lanes = 1: the reference's bytes; parameters within 1e-5; multi-lane lossless."""
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder
    from tests.test_oracle_golden import load_icase
    c = load_icase(iv, name)
    coder = Coder(in_channels=c["C"], channel_groups=c["G"], default_topo_group_method=c["method"], lanes=lanes,
                  ans_params_device="cpu")
    sd = dict(c["sd"])
    sd["conv_kernel_weight"] = torch.zeros_like(sd["context_prediction.weight"])     # a reference checkpoint has these too
    sd["lower_bound_scale.bound"] = torch.tensor([0.11])
    coder.load_state_dict(sd)
    coder = coder.cuda().eval()
    coder.update_state()
    y, prior = c["y"].cuda(), c["prior"].cuda()
    from cbench_basic_b200 import _native as N
    params = torch.full((c["B"], 2 * c["C"], c["H"], c["W"]), float("nan"), device="cuda")
    coder._set_map(c["tg"])
    yhat_ref = c["yhat"].cuda().contiguous()
    for g in range(int(c["tg"].max()) + 1):   # every cell at its own stage, from the full reconstruction
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, yhat_ref.data_ptr(), prior.data_ptr(), c["B"], params.data_ptr(), 0))
    torch.cuda.synchronize()
    assert not torch.isnan(params).any()
    assert rel_err(params.cpu(), c["params_full"]) <= REL_TOL, rel_err(params.cpu(), c["params_full"])
    bs = coder.encode(y, prior=prior)
    yhat = coder.decode(bs, prior=prior)
    if lanes == 1:
        assert bs == c["bytes"]
    assert_latents_match(yhat.cpu(), c["yhat"])


@pytest.mark.parametrize("lanes", [1, 0])
def test_internal_merger_expand_bottleneck_golden(golden_dir, lanes):
    """param_merger_expand_bottleneck=True (pgm_coder.py:1216-1222: the merger's hidden width is 8C instead of 4C) on the CUDA
    path against the unmodified reference's vectors (tests/golden/make_internal_golden.py): bytes at lanes = 1, parameters
    within 1e-5, multi-lane lossless."""
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder
    from cbench_basic_b200 import _native as N
    from tests.test_oracle_golden import load_icase
    ev = np.load(os.path.join(golden_dir, "ypath_internal_expand_vectors.npz"))
    c = load_icase(ev, "int_expand")
    coder = Coder(in_channels=c["C"], channel_groups=c["G"], default_topo_group_method=c["method"], lanes=lanes,
                  ans_params_device="cpu", param_merger_expand_bottleneck=True)
    coder.load_state_dict(dict(c["sd"]))
    coder = coder.cuda().eval()
    coder.update_state()
    y, prior = c["y"].cuda(), c["prior"].cuda()
    params = torch.full((c["B"], 2 * c["C"], c["H"], c["W"]), float("nan"), device="cuda")
    coder._set_map(c["tg"])
    yhat_ref = c["yhat"].cuda().contiguous()
    for g in range(int(c["tg"].max()) + 1):
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, yhat_ref.data_ptr(), prior.data_ptr(), c["B"], params.data_ptr(), 0))
    torch.cuda.synchronize()
    assert not torch.isnan(params).any()
    assert rel_err(params.cpu(), c["params_full"]) <= REL_TOL, rel_err(params.cpu(), c["params_full"])
    bs = coder.encode(y, prior=prior)
    yhat = coder.decode(bs, prior=prior)
    if lanes == 1:
        assert bs == c["bytes"]
    assert_latents_match(yhat.cpu(), c["yhat"])


@pytest.mark.parametrize("name", ["jar_a", "jar_b"])
@pytest.mark.parametrize("lanes", [1, 0])
def test_joint_ar_serial_coder_golden(jv, name, lanes):
    """use_joint_ar_model_impl=True -- the reference codes pixel by pixel with a causally masked 5x5
    convolution and entropy_parameters(cat(prior, ctx)); here it is the scanline map with remapped merger matrices.  A
    reference state_dict loads as it is; lanes = 1 reproduces the reference's bytes."""
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder
    from tests.test_oracle_golden import load_jcase
    c = load_jcase(jv, name)
    coder = Coder(in_channels=c["C"], use_joint_ar_model_impl=True, lanes=lanes, ans_params_device="cpu")
    coder.load_state_dict(dict(c["sd"], conv_kernel_bias=torch.zeros(2 * c["C"])))
    coder = coder.cuda().eval()
    coder.update_state()
    y, prior = c["y"].cuda(), c["prior"].cuda()
    bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
    if lanes == 1:
        assert bs == c["bytes"]
    yhat = coder.decode(bs, prior=prior)
    assert torch.equal(yhat, yhat_enc * 1.0 + 0.0)
    assert_latents_match(yhat.cpu(), c["yhat"])


@pytest.fixture(scope="module")
def jv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_jointar_vectors.npz"))


@pytest.fixture(scope="module")
def iv(golden_dir):
    return np.load(os.path.join(golden_dir, "ypath_internal_vectors.npz"))


@pytest.mark.parametrize("name,method", [("ckbd", "checkerboard"), ("cwckbd", "channelwise-checkerboard"),
                                         ("scanline", "scanline"), ("raster", "raster2x2"), ("meanscale", "none")])
def test_default_maps_through_constructor(yv, name, method):
    c = load_ycase(yv, name)
    coder = make_coder(c, lanes=1, method=method)
    bs = coder.encode(c["y"].cuda(), prior=c["prior"].cuda())
    assert bs == c["bytes"]
    assert_latents_match(coder.decode(bs, prior=c["prior"].cuda()).cpu(), c["yhat"])


# ------------------------------------------------------- larger, seeded: oracle parity + properties
def _random_case(C_, G, B, H, W, seed, method="checkerboard", ctx=True):
    torch.manual_seed(seed)
    w = Y.random_weights(C_, seed) if ctx else None
    y, prior = 3 * torch.randn(B, C_, H, W), torch.randn(B, 2 * C_, H, W)
    tg = Y.default_pgm(method, G, H, W)
    if w is None:
        w = {"ctx_w": torch.zeros(2 * C_, C_, 5, 5), "ctx_b": torch.zeros(2 * C_)}
    return dict(C=C_, G=G, B=B, H=H, W=W, w=w, y=y, prior=prior, tg=tg)


def _stagewise_symbols(coder, c):
    """Our symbols / scale indexes in stream order, through the C-ABI building blocks (the same calls
    basic_ypath_encode makes): per stage, context model -> quantise + index -> write-back."""
    from cbench_basic_b200 import _native as N
    B, C_, H, W = c["B"], c["C"], c["H"], c["W"]
    coder._set_map(c["tg"])
    y, prior = c["y"].cuda().contiguous(), c["prior"].cuda().contiguous()
    buf = torch.zeros_like(y)
    params = torch.zeros(B, 2 * C_, H, W, device="cuda")
    syms, idxs = [], []
    h = coder.ans_encoder.handle
    for g in range(N.lib().basic_ctx_num_stages(coder._ctx)):
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, buf.data_ptr(), prior.data_ptr(), B, params.data_ptr(), 0))
        pos, n_pos = C.c_void_p(), C.c_int64()
        N.check(N.lib().basic_ctx_stage_positions(coder._ctx, g, C.byref(pos), C.byref(n_pos)))
        n = B * n_pos.value
        sym, idx = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
        N.check(N.lib().basic_gauss_quantize_index(h, y.data_ptr(), params.data_ptr(), pos, n_pos.value, B, C_, H * W,
                                                   sym.data_ptr(), idx.data_ptr(), buf.data_ptr(), 0))
        syms.append(sym)
        idxs.append(idx)
    torch.cuda.synchronize()
    return torch.cat(syms).cpu().numpy(), torch.cat(idxs).cpu().numpy(), params.cpu()


def test_ypath_vs_oracle_c192_checkerboard():
    """BASELINE configs[1] geometry (C = 192, checkerboard) on one Kodak-shape image, against the CPU oracle:
    295k symbols through a 4800-deep FP32 reduction.  Symbols must agree everywhere.  A scale index may differ
    only where the ORACLE's own scale sits within the float tolerance of the midpoint between two table entries
    (a different, equally valid FP32 summation order decides such a tie) -- any other disagreement fails."""
    c = _random_case(192, 1, 1, 32, 48, 7)
    tab = Y.get_scale_table()
    with torch.no_grad():
        sym, idx, yhat_ref = Y.encode_symbols(c["y"], c["prior"], c["tg"], c["w"], tab)
        params_ref = Y.params_for(yhat_ref, c["tg"], c["prior"], c["w"])
    o = Y.YPathOracle(192, 1, c["w"])
    o.update_state()
    ref_bytes = o.enc.encode_with_indexes(sym, idx)
    coder = make_coder(c, lanes=1, method="checkerboard")
    gsym, gidx, gparams = _stagewise_symbols(coder, c)
    assert rel_err(gparams, params_ref) <= REL_TOL
    assert np.array_equal(gsym, sym), f"{int((gsym != sym).sum())} symbols differ from the oracle"
    bad = np.nonzero(gidx != idx)[0]
    assert bad.size <= 1, f"{bad.size} scale indexes differ"   # observed on B200: 0 (profiles/r2_tie_counts.txt); a tie is checked below
    if bad.size:
        # stream order -> element: stage-major, then (c, h, w) of the stage's mask
        gmap = Y.group_of_elements(c["tg"], 1, 192).reshape(-1)
        order = torch.cat([torch.nonzero(gmap == g).reshape(-1) for g in range(2)])
        sc = Y.split_mean_scale(params_ref)[1].reshape(-1)[order][bad].double()
        lo = np.minimum(gidx[bad], idx[bad])
        assert np.all(np.abs(gidx[bad] - idx[bad]) == 1)
        mid = (tab.double()[lo] + tab.double()[lo + 1]) / 2
        assert float(((sc - mid).abs() / mid).max()) <= REL_TOL, "scale-index disagreement away from a tie"
    for lanes in (1, 0):
        coder = make_coder(c, lanes=lanes, method="checkerboard")
        bs, yhat_enc = coder.encode(c["y"].cuda(), prior=c["prior"].cuda(), return_yhat=True)
        yhat = coder.decode(bs, prior=c["prior"].cuda())
        assert torch.equal(yhat.cpu(), yhat_enc.cpu() * 1.0 + 0.0)          # lossless w.r.t. the encoder's own y_hat
        assert_latents_match(yhat.cpu(), yhat_ref)                          # same integer symbols as the oracle
        if lanes == 1:
            assert bs == o.enc.encode_with_indexes(gsym, gidx)               # the reference coder on OUR symbols/indexes
            if bad.size == 0:
                assert bs == ref_bytes
        else:
            assert len(bs) <= len(ref_bytes) * 1.005 + 300


def test_ypath_multilane_container_matches_cpu_spec():
    """The y path's multi-lane container = magic + ONE segment with one slice per coding group (lane states carried
    across groups, one flush per lane): byte-identical to the CPU specification run on our symbols / indexes."""
    import struct
    from oracle import ans_oracle as O
    c = _random_case(24, 4, 3, 9, 14, 5, method="channelwise-checkerboard")
    coder = make_coder(c, lanes=96, method="channelwise-checkerboard")
    gsym, gidx, _ = _stagewise_symbols(coder, c)
    bs = coder.encode(c["y"].cuda(), prior=c["prior"].cuda())
    magic, n_chunks, n_slices = struct.unpack_from("<III", bs, 0)
    gmap = Y.group_of_elements(c["tg"], c["B"], c["C"]).reshape(-1)
    slice_n = [int((gmap == g).sum()) for g in range(n_slices)]
    # "BLS0": this geometry (6 channels per group) runs the exact FP32 kernels, and the container says so
    assert magic == 0x30534C42 and n_slices == int(c["tg"].max()) + 1 and sum(slice_n) == gsym.size
    o = Y.YPathOracle(24, 4, c["w"])
    o.update_state()
    assert bs[4:] == o.enc.encode_lanes_slices(gsym, gidx, slice_n, 3)
    out, used = o.dec.decode_lanes_slices(bs[4:], gidx, slice_n)
    assert used == len(bs) - 4 and np.array_equal(out, gsym)
    yhat = coder.decode(bs, prior=c["prior"].cuda())
    assert float((yhat.cpu() - c["y"]).abs().max()) <= 0.5 + 1e-5


def test_ypath_round_trip_properties_large():
    """Size-independent properties at a bigger batch: lossless round trip, |y_hat - y| <= 0.5, stream size close to
    the lanes=1 stream."""
    c = _random_case(192, 1, 6, 32, 48, 21)
    coder1, coder0 = make_coder(c, 1, "checkerboard"), make_coder(c, 0, "checkerboard")
    y, p = c["y"].cuda(), c["prior"].cuda()
    b1, yh1 = coder1.encode(y, prior=p, return_yhat=True)
    b0, yh0 = coder0.encode(y, prior=p, return_yhat=True)
    assert torch.equal(yh0, yh1)
    d1, d0 = coder1.decode(b1, prior=p), coder0.decode(b0, prior=p)
    assert torch.equal(d0, d1) and torch.equal(d0, yh0 * 1.0 + 0.0)
    assert float((d0 - y).abs().max()) <= 0.5 + 1e-5
    assert len(b0) <= len(b1) * 1.005 + 300


@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_host_tensor_inputs_equal_device_inputs(precision):
    """Pinned or pageable CPU tensors go to the C ABI as host pointers (uploaded on the coder's copy stream, overlapped
    with the first kernels): same stream, same reconstruction as device inputs; repeated calls reuse the staging buffers."""
    c = _random_case(96, 1, 3, 16, 24, 33)
    coder = make_coder(c, 0, "checkerboard", ctx_precision=precision)
    y, p = c["y"].contiguous(), c["prior"].contiguous()
    ref = coder.encode(y.cuda(), prior=p.cuda())
    for yy, pp in ((y.pin_memory(), p.pin_memory()), (y, p), (y.pin_memory(), p.cuda())):
        for _ in range(2):
            bs, yh = coder.encode(yy, prior=pp, return_yhat=True)
            assert bs == ref
            out = coder.decode(bs, prior=pp)
            assert out.is_cuda and torch.equal(out, yh * 1.0 + 0.0)


def test_mean_scale_4k_320ch_tiles():
    """BASELINE configs[4] geometry: 135 x 240 x 320 latent, mean-scale coder (tiling is exact there), coded as
    row-band tiles; every tile decodes on its own and the union equals the untiled result."""
    from cbench_basic_b200 import sharding
    c = _random_case(320, 1, 1, 135, 240, 5, method="none", ctx=False)
    coder = make_coder(c, 0, "none")
    y, p = c["y"].cuda(), c["prior"].cuda()
    whole, yh = coder.encode(y, prior=p, return_yhat=True)
    parts, total = [], 0
    for band in sharding.row_band_tiles(135, 8):
        ys, ps = y[:, :, band.start:band.stop].contiguous(), p[:, :, band.start:band.stop].contiguous()
        bs = coder.encode(ys, prior=ps)
        total += len(bs)
        parts.append(coder.decode(bs, prior=ps))
    assert torch.equal(torch.cat(parts, dim=2), yh * 1.0 + 0.0)
    assert total <= len(whole) * 1.01


def test_combined_coder_dispatch(yv):
    from cbench_basic_b200.prior_coder import CombinedNNTrainablePGMPriorCoder
    c = load_ycase(yv, "ckbd")
    a, b = make_coder(c, 1, "checkerboard"), make_coder(c, 1, "none")
    comb = CombinedNNTrainablePGMPriorCoder([a, b])
    w = torch.tensor([0.1, 0.9])
    bs = comb.encode(c["y"].cuda(), prior=c["prior"].cuda(), blend_weight=w)
    assert bs == b.encode(c["y"].cuda(), prior=c["prior"].cuda())
    assert torch.equal(comb.decode(bs, prior=c["prior"].cuda(), blend_weight=w), b.decode(bs, prior=c["prior"].cuda()))


def test_cfg2_batch_against_oracle():
    """BASELINE configs[1] at its full size -- 24 Kodak-shape images, C = 192, the TIMED mode (multi-lane, 3xFP16 context model):
    images 0 and 23 of the batch against the CPU oracle (images are independent, so the oracle runs them one at a time).
    Symbols and scale indexes identical, except where the oracle's own value sits on a rounding tie; parameters within 1e-5."""
    import bench
    dev = torch.device("cuda", 0)
    y, prior, w = bench.make_inputs("cfg2", 0)
    coder = bench.build_coder("cfg2", w, 0, dev)
    yd, pd = y.to(dev), prior.to(dev)
    bs, yhat_enc = coder.encode(yd, prior=pd, return_yhat=True)
    out = coder.decode(bs, prior=pd)
    assert torch.equal(out, yhat_enc * 1.0 + 0.0)
    tg, tab = Y.default_pgm("checkerboard", 1, 32, 48), Y.get_scale_table()
    for b in (0, 23):
        with torch.no_grad():
            _, _, yhat_o = Y.encode_symbols(y[b:b + 1], prior[b:b + 1], tg, w, tab)
            mean_o, _ = Y.split_mean_scale(Y.params_for(yhat_o, tg, prior[b:b + 1], w))
        sym_o = torch.round(yhat_o - mean_o)
        sym_g = torch.round(out[b:b + 1].cpu() - mean_o)
        bad = sym_g != sym_o
        frac = (y[b:b + 1] - mean_o)[bad].double()
        assert int(bad.sum()) <= 2 and bool((((frac - torch.floor(frac)) - 0.5).abs() <= 2e-5 * (1 + frac.abs())).all()), int(bad.sum())
        ok = ~bad
        assert float(((out[b:b + 1].cpu() - yhat_o)[ok].abs() / yhat_o[ok].abs().clamp_min(1.0)).max()) <= REL_TOL
    r = bench.oracle_mismatches("cfg2", coder, y, prior, w, yd, pd)
    assert r["off_tie"] == 0 and r["symbol_mismatches"] <= 2 and r["index_mismatches"] <= 2 and r["params_max_rel_err"] <= REL_TOL, r


@pytest.mark.parametrize("B", [1, 3, 6, 9, 20])
def test_scanline_stage_kernel_rows_against_oracle(B):
    """The persistent stage kernels (ctx_scan.cu) with 1 .. 20 rows per stage (<= 4: k_scan_stages, weights resident; 6, 9, 20:
    k_scan_blocks with one, two and three row blocks): y_hat of encoder and both decoders (lanes = 0: chunk warps inside the one
    launch; lanes = 1: a launch per stage that first dequantises the previous stage) against the CPU oracle -- same symbols,
    means within 1e-5."""
    c = _random_case(12, 1, B, 6, 7, 40 + B, method="scanline")
    tab = Y.get_scale_table()
    with torch.no_grad():
        _, _, yhat_o = Y.encode_symbols(c["y"], c["prior"], c["tg"], c["w"], tab)
    y, prior = c["y"].cuda(), c["prior"].cuda()
    sizes = {}
    for lanes in (0, 1):
        coder = make_coder(c, lanes, method="scanline")
        bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
        out = coder.decode(bs, prior=prior)
        assert torch.equal(out, yhat_enc * 1.0 + 0.0)          # lossless, bit for bit
        assert_latents_match(out.cpu(), yhat_o)
        sizes[lanes] = len(bs)
    assert sizes[0] >= sizes[1]


def test_scanline_one_launch_decoder_on_damaged_streams():
    """The in-kernel chunk decoder (lanes = 0 on the stage kernel) must come back on damaged input: a truncated container is
    refused before anything is launched, flipped stream words decode to different latents or raise -- neither hangs."""
    from cbench_basic_b200._native import StreamError
    c = _random_case(12, 1, 2, 6, 7, 77, method="scanline")
    coder = make_coder(c, 0, method="scanline")
    y, prior = c["y"].cuda(), c["prior"].cuda()
    bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
    assert torch.equal(coder.decode(bs, prior=prior), yhat_enc * 1.0 + 0.0)
    with pytest.raises((StreamError, ValueError)):
        coder.decode(bs[:len(bs) // 2], prior=prior)
    bad = bytearray(bs)
    for at in range(len(bad) - 40, len(bad) - 8):   # the tail of the word region
        bad[at] ^= 0x5A
    try:
        out = coder.decode(bytes(bad), prior=prior)
    except (StreamError, ValueError):
        out = None
    assert out is None or not torch.equal(out, yhat_enc * 1.0 + 0.0)
    # the coder is still usable afterwards
    assert torch.equal(coder.decode(bs, prior=prior), yhat_enc * 1.0 + 0.0)


def test_zero_copy_view_equals_bytes():
    """encode(zero_copy=True): a read-only view of the coder's page-locked buffer with the bytes of the plain call; decode() takes
    the view (uploading straight out of page-locked memory) and a bytes copy of it alike."""
    c = _random_case(12, 1, 2, 8, 8, 5)
    coder = make_coder(c, 0, method="checkerboard")
    y, prior = c["y"].cuda(), c["prior"].cuda()
    bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
    view = coder.encode(y, prior=prior, zero_copy=True)
    assert isinstance(view, memoryview) and view.readonly and bytes(view) == bs
    out_v = coder.decode(view, prior=prior)
    out_b = coder.decode(bs, prior=prior)
    assert torch.equal(out_v, out_b) and torch.equal(out_v, yhat_enc * 1.0 + 0.0)
    coder_h = make_coder(c, 0, method="checkerboard")
    coder_h.force_input_prior_shape_aligned = False          # a framing header cannot be prepended without a copy
    with pytest.raises(ValueError):
        coder_h.encode(y, prior=prior, zero_copy=True)


@pytest.mark.parametrize("C_,G,method", [(12, 1, "checkerboard"), (12, 2, "channelwise-checkerboard"), (192, 1, "checkerboard")])
def test_host_inputs_in_sub_batches(C_, G, method):
    """Host tensors of a batch of >= 8 images are uploaded in image sub-batches (capi.cu plan_subs): same bytes and the same
    latents as device inputs, also with two channel groups (later stages read activations cached by earlier ones) and at
    C = 192 (the 3xFP16 tensor-core kernels on sub-batches)."""
    c = _random_case(C_, G, 16, 6, 8, 91 + G, method=method)
    coder = make_coder(c, 0, method=method, ctx_precision="auto")
    y, prior = c["y"], c["prior"]
    bs_d, yhat_d = coder.encode(y.cuda(), prior=prior.cuda(), return_yhat=True)
    bs_h = coder.encode(y.pin_memory(), prior=prior.pin_memory())
    assert bs_h == bs_d
    out_h = coder.decode(bs_h, prior=prior.pin_memory())
    out_d = coder.decode(bs_d, prior=prior.cuda())
    assert torch.equal(out_h, out_d) and torch.equal(out_d, yhat_d * 1.0 + 0.0)


@pytest.mark.parametrize("B,H,W", [(1, 5, 6), (3, 6, 7), (7, 4, 9), (13, 9, 8)])
def test_stage_kernels_on_zigzag_maps(B, H, W):
    """Anti-diagonal stages: several cells per stage and a different number in every stage (rows = B x cells: 1 .. 104 here, so
    both stage kernels and several row blocks are exercised, with partly filled blocks).  Encoder and both decoders against the
    CPU oracle."""
    c = _random_case(12, 1, B, H, W, 300 + B, method="zigzag")
    tab = Y.get_scale_table()
    with torch.no_grad():
        _, _, yhat_o = Y.encode_symbols(c["y"], c["prior"], c["tg"], c["w"], tab)
    y, prior = c["y"].cuda(), c["prior"].cuda()
    for lanes in (0, 1):
        coder = make_coder(c, lanes, method="zigzag")
        bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
        out = coder.decode(bs, prior=prior)
        assert torch.equal(out, yhat_enc * 1.0 + 0.0)
        assert_latents_match(out.cpu(), yhat_o)


@pytest.mark.parametrize("B", [1, 6])
def test_stage_kernels_c192_against_oracle(B):
    """The stage kernels at the channel count of the BASELINE configurations (C = 192: 148 CTAs with two or three channel pairs
    each, the dedicated decoder CTAs; B = 6: row blocks with streamed weights) on a small scanline image, against the CPU oracle."""
    c = _random_case(192, 1, B, 4, 6, 500 + B, method="scanline")
    tab = Y.get_scale_table()
    with torch.no_grad():
        _, _, yhat_o = Y.encode_symbols(c["y"], c["prior"], c["tg"], c["w"], tab)
        mean_o, _ = Y.split_mean_scale(Y.params_for(yhat_o, c["tg"], c["prior"], c["w"]))
    y, prior = c["y"].cuda(), c["prior"].cuda()
    for lanes in (0, 1):
        coder = make_coder(c, lanes, method="scanline")
        bs, yhat_enc = coder.encode(y, prior=prior, return_yhat=True)
        out = coder.decode(bs, prior=prior)
        assert torch.equal(out, yhat_enc * 1.0 + 0.0)
        # same symbols except where the oracle's own value sits on a rounding tie (4800-deep sums), means within 1e-5
        sym_o, sym_g = torch.round(yhat_o - mean_o), torch.round(out.cpu() - mean_o)
        bad = sym_g != sym_o
        frac = (c["y"] - mean_o)[bad].double()
        assert int(bad.sum()) <= 2 and bool((((frac - torch.floor(frac)) - 0.5).abs() <= 2e-5 * (1 + frac.abs())).all()), int(bad.sum())
        ok = ~bad
        assert float(((out.cpu() - yhat_o)[ok].abs() / yhat_o[ok].abs().clamp_min(1.0)).max()) <= REL_TOL

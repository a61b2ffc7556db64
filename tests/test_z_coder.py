"""z-node factorized-prior coder (cbench_basic_b200/z_coder.py).

Pinned by tests/golden/z_vectors.npz (made with the reference's own cbench.rans coder and write_body framing,
tests/golden/make_z_golden.py): stream bytes, framing, tables -> streams.  The density network is pinned to the reference's in-tree copy of it
(z_pmf_vectors.npz, make_z_pmf_golden.py); the three lines that turn its logits into a pmf restate compressai 1.2.3."""
import os

import numpy as np
import pytest
import torch

from cbench_basic_b200 import z_coder
from oracle import ref_loader, z_oracle as Z

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "z_vectors.npz")
CASES = ["small", "c192"]


@pytest.fixture(scope="module")
def zv():
    return np.load(GOLDEN)


def _sd(zv, name):
    return {k.split("/sd/")[1]: torch.from_numpy(zv[k]) for k in zv.files if k.startswith(name + "/sd/")}


def _tables(zv, name):
    return zv[name + "/cdf"], zv[name + "/cdf_length"], zv[name + "/offset"]


# ------------------------------------------------------------------------------------------------ CPU: oracle + host logic
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(zv, name):
    if ref_loader.load("rans") is None:
        pytest.skip("oracle/_ref/rans not built")
    sd, x = _sd(zv, name), torch.from_numpy(zv[name + "/x"])
    tables = Z.build_tables(sd)
    for a, b in zip(tables, _tables(zv, name)):
        assert np.array_equal(a, b)
    strings = Z.compress(sd, tables, x)
    body = Z.write_body(x.shape[-2:], strings)
    assert body == zv[name + "/body"].tobytes()
    back, shape = Z.read_body(body)
    assert torch.equal(Z.decompress(sd, tables, back, shape), torch.from_numpy(zv[name + "/y_hat"]))


@pytest.mark.parametrize("name", CASES)
def test_cumulative_logits_match_reference_network(name):
    """The density network of the z node (the step whose original lives in compressai): the product's and the oracle's
    evaluation against the reference's own in-tree copy of that network (cbench/nn/layers/param_generator.py:158-199), run
    by tests/golden/make_z_pmf_golden.py on seeded parameters at the sample points update() uses.  Bit-exact on the CPU."""
    pv = np.load(os.path.join(os.path.dirname(GOLDEN), "z_pmf_vectors.npz"))
    sd = _sd(pv, name)
    samples = torch.from_numpy(pv[name + "/samples"])
    C = samples.shape[0]
    eb = z_coder.EntropyBottleneck(C)
    eb.load_state_dict(sd, strict=False)
    for shift, key in ((-0.5, "lower"), (0.5, "upper")):
        ref = torch.from_numpy(pv[f"{name}/{key}"])
        assert torch.equal(eb._logits_cumulative(samples + shift), ref)
        assert torch.equal(Z.logits_cumulative(sd, samples + shift), ref)


@pytest.mark.parametrize("name", CASES)
def test_framing_matches_reference(zv, name):
    """read_body / write_body of the product against bytes written by the reference's write_body (big-endian u32s)."""
    body = zv[name + "/body"].tobytes()
    strings, shape = z_coder.read_body(body)
    assert tuple(shape) == tuple(zv[name + "/x"].shape[-2:]) and len(strings) == zv[name + "/x"].shape[0]
    assert z_coder.write_body(shape, strings) == body
    with pytest.raises(ValueError):
        z_coder.read_body(body[:-1])
    with pytest.raises(ValueError):
        z_coder.read_body(body[:7])


def test_state_dict_layout(zv):
    """Parameter names / shapes of compressai's EntropyBottleneck (a reference checkpoint must load), incl. the later
    ParameterList spelling and table buffers of any size."""
    sd = _sd(zv, "small")
    eb = z_coder.EntropyBottleneck(8)
    keys = set(eb.state_dict().keys())
    assert {"_matrix0", "_bias0", "_factor0", "_matrix4", "_bias4", "quantiles", "target", "_offset", "_quantized_cdf",
            "_cdf_length"} <= keys and "_factor4" not in keys
    assert tuple(eb._matrix0.shape) == (8, 3, 1) and tuple(eb._matrix4.shape) == (8, 1, 3) and tuple(eb.quantiles.shape) == (8, 1, 3)
    full = dict(sd, target=eb.target.clone(), _offset=torch.from_numpy(zv["small/offset"]),
                _quantized_cdf=torch.from_numpy(zv["small/cdf"]), _cdf_length=torch.from_numpy(zv["small/cdf_length"]))
    eb.load_state_dict(full)
    assert torch.equal(eb._quantized_cdf, torch.from_numpy(zv["small/cdf"]))
    renamed = {k.replace("_matrix", "matrices.").replace("_bias", "biases.").replace("_factor", "factors."): v for k, v in full.items()}
    eb2 = z_coder.EntropyBottleneck(8)
    eb2.load_state_dict(renamed)
    assert torch.equal(eb2._matrix2, sd["_matrix2"])
    coder = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=8)
    assert any(k.startswith("entropy_bottleneck._matrix0") for k in coder.state_dict())


# ------------------------------------------------------------------------------------------------ GPU: the product path
def _coder_from_golden(zv, name, with_tables):
    C_ = zv[name + "/x"].shape[1]
    coder = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=C_)
    eb = coder.entropy_bottleneck
    sd = _sd(zv, name)
    full = dict(sd, target=eb.target.clone(), _offset=torch.from_numpy(zv[name + "/offset"]),
                _quantized_cdf=torch.from_numpy(zv[name + "/cdf"]), _cdf_length=torch.from_numpy(zv[name + "/cdf_length"]))
    if not with_tables:
        full.update(_offset=torch.IntTensor(), _quantized_cdf=torch.IntTensor(), _cdf_length=torch.IntTensor())
    eb.load_state_dict(full)
    return coder


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_streams_byte_exact_from_reference_tables(zv, name):
    """Tables as the reference's pmf_to_quantized_cdf made them -> our CUDA coder writes the reference's bytes (every
    image's lanes=1 stream and the framing) and decodes them to the reference's y_hat."""
    coder = _coder_from_golden(zv, name, with_tables=True)
    x = torch.from_numpy(zv[name + "/x"])
    body = coder.encode(x.cuda())
    assert body == zv[name + "/body"].tobytes()
    y_hat = coder.decode(body)
    assert y_hat.is_cuda and torch.equal(y_hat.cpu(), torch.from_numpy(zv[name + "/y_hat"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_update_state_builds_the_tables(zv, name):
    """update_state(): pmf out of the parameters (restated, unpinned) + pmf_to_quantized_cdf on the device (pinned)."""
    coder = _coder_from_golden(zv, name, with_tables=False)
    coder.update_state()
    eb = coder.entropy_bottleneck
    assert np.array_equal(eb._quantized_cdf.numpy(), zv[name + "/cdf"])
    assert np.array_equal(eb._cdf_length.numpy(), zv[name + "/cdf_length"])
    assert np.array_equal(eb._offset.numpy(), zv[name + "/offset"])
    assert coder.encode(torch.from_numpy(zv[name + "/x"])) == zv[name + "/body"].tobytes()


@pytest.mark.gpu
def test_gpu_round_trip_with_gains_and_empty_batch():
    torch.manual_seed(0)
    C_ = 16
    coder = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=C_)
    coder.entropy_bottleneck.load_state_dict(dict(Z.init_params(C_, seed=7), target=coder.entropy_bottleneck.target.clone(),
                                                  _offset=torch.IntTensor(), _quantized_cdf=torch.IntTensor(),
                                                  _cdf_length=torch.IntTensor()))
    coder.update_state()
    x = 6 * torch.randn(5, C_, 3, 7)
    gains, inv = torch.rand(C_) + 0.5, None
    inv = 1.0 / gains
    body = coder.encode(x.cuda(), channel_gains=gains)
    y_hat = coder.decode(body, channel_gains_inv=inv)
    assert float((y_hat.cpu() * gains.view(1, -1, 1, 1) - x * gains.view(1, -1, 1, 1)).abs().max()) <= 0.5 + 1e-4
    empty = coder.encode(torch.zeros(0, C_, 3, 7).cuda())
    assert coder.decode(empty).shape == (0, C_, 3, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_multilane_z_container(zv, name):
    """lanes = 0: the whole batch in one multi-lane container behind the same framing; same symbols as the reference
    streams (y_hat identical), lossless, and not larger than the per-image streams plus the lane flush."""
    C_ = zv[name + "/x"].shape[1]
    coder = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=C_, lanes=0)
    eb = coder.entropy_bottleneck
    eb.load_state_dict(dict(_sd(zv, name), target=eb.target.clone(), _offset=torch.from_numpy(zv[name + "/offset"]),
                            _quantized_cdf=torch.from_numpy(zv[name + "/cdf"]), _cdf_length=torch.from_numpy(zv[name + "/cdf_length"])))
    x = torch.from_numpy(zv[name + "/x"])
    body = coder.encode(x.cuda())
    strings, shape = z_coder.read_body(body)
    assert len(strings) == 1 and strings[0][4:8] == b"BLS1" and tuple(shape) == tuple(x.shape[-2:])
    assert torch.equal(coder.decode(body).cpu(), torch.from_numpy(zv[name + "/y_hat"]))
    assert len(body) <= len(zv[name + "/body"]) + 200


@pytest.mark.gpu
def test_gpu_hyperprior_latent_codec_round_trip():
    """z -> h_s -> y on the device: both node coders, a random hyper-analysis / hyper-synthesis pair, the
    reference's container (z segment with its u32 length, then y).  The decoder reproduces the encoder's y_hat."""
    import struct
    import torch.nn as nn
    from cbench_basic_b200.latent_codec import HyperpriorLatentCodec
    from cbench_basic_b200.prior_coder import (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as YCoder,
                                               TopoGroupDynamicMaskConv2dContextModel as Ctx)
    torch.manual_seed(1)
    C_, Cz, B, H, W = 24, 16, 3, 8, 12
    zc = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=Cz)
    zc.entropy_bottleneck.load_state_dict(dict(Z.init_params(Cz, seed=5), target=zc.entropy_bottleneck.target.clone(),
                                               _offset=torch.IntTensor(), _quantized_cdf=torch.IntTensor(),
                                               _cdf_length=torch.IntTensor()))
    yc = YCoder(in_channels=C_, default_topo_group_method="checkerboard",
                topo_group_context_model=Ctx(in_channels=C_, out_channels=2 * C_), lanes=0)
    h_a = nn.Conv2d(C_, Cz, 3, stride=2, padding=1)
    h_s = nn.Sequential(nn.ConvTranspose2d(Cz, 2 * C_, 4, stride=2, padding=1))
    codec = HyperpriorLatentCodec(zc, yc, hyper_synthesis=h_s, hyper_analysis=h_a).cuda().eval()
    codec.update_state()
    y = (3 * torch.randn(B, C_, H, W)).cuda()
    data = codec.encode(y)
    (zlen,) = struct.unpack_from("I", data, 0)
    strings, shape = z_coder.read_body(data[4:4 + zlen])
    assert len(strings) == B and tuple(shape) == (H // 2, W // 2) and data[4 + zlen:8 + zlen] in (b"BLS0", b"BLS1", b"BLS2")
    y_hat = codec.decode(data)
    assert y_hat.shape == y.shape and float((y_hat - y).abs().max()) <= 0.5 + 1e-4
    assert torch.equal(codec.decode(data), y_hat)

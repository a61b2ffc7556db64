"""The upper boundary (SURVEY 8 row b): the reference's own caller drives the drop-in coders.

CPU part (skipped without /root/reference): the reference's REAL ``LatentGraphicalANSEntropyCoder`` is built through
``tests/golden/ref_shim`` around the two drop-in classes bound by ``cbench_basic_b200.reference_integration``; the native layer
(the C ABI) is replaced by a recording fake, so what is checked is construction, ``update_state`` wiring, the module protocol
(cache, profiler, isinstance) and the exact call order of latent_graph.py:826-841 / :1232-1301.

GPU part: the same call transcript replayed on the CUDA coders (forward -> encode -> merge_bytes -> split -> decode with
``stream=None``), the way ``_generative_process`` issues it.
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))


# ---------------------------------------------------------------------------------------------- module protocol (CPU)
def test_module_protocol_standalone():
    from cbench_basic_b200.prior_coder import (CombinedNNTrainablePGMPriorCoder,
                                               GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder)
    from cbench_basic_b200.z_coder import CompressAIEntropyBottleneckPriorCoder as ZCoder
    y = Coder(in_channels=8, use_param_merger=False).eval()
    z = ZCoder(entropy_bottleneck_channels=4).eval()
    comb = CombinedNNTrainablePGMPriorCoder([y]).eval()
    for m in (y, z, comb):
        assert set(m.cache_names) >= {"common", "loss_dict", "metric_dict", "moniter_dict", "hist_dict"}
        m.update_cache("metric_dict", a=1.0)
        assert m.get_raw_cache("metric_dict") == {"a": 1.0}
        assert m.get_cache("metric_dict") == {"metric_dict/a": 1.0}
        m.reset_all_cache()
        assert m.get_raw_cache("metric_dict") == {}
        with m.profiler.start_time_profile("scope"):
            pass
        assert "scope (ms)" in m.collect_profiler_results()
        assert m.device.type == "cpu"
    comb.coders[0].update_cache("loss_dict", r=2.0)
    assert comb.get_cache("loss_dict") == {"loss_dict/coders/0/r": 2.0}   # nn.ModuleList recursion, cbench/nn/base.py:310-314
    # eval forward = the dequantised input (pgm_coder.py:391-398, :539; compressai_coder.py:203-227)
    t = torch.tensor([[[[0.5, 1.5, -0.5, 2.4]]] * 8])
    assert torch.equal(y(t, prior=torch.zeros(1, 16, 1, 4)), torch.round(t))
    assert torch.equal(comb(t, prior=torch.zeros(1, 16, 1, 4)), torch.round(t))
    q = Coder(in_channels=8, use_param_merger=False, quantizer_params=[0.25, 128, 0.5]).eval()
    assert torch.allclose(q(t), torch.round((t - 0.25) / 0.5) * 0.5 + 0.25)
    tz = torch.randn(2, 4, 3, 3) * 4
    med = z.entropy_bottleneck.quantiles[:, 0, 1].detach().view(1, 4, 1, 1)
    assert torch.equal(z(tz), torch.round(tz - med) + med)
    y.train()
    with pytest.raises(NotImplementedError):
        y(t)


# ---------------------------------------------------------------------------------------------- the reference's own caller
class _FakeLib:
    """Stands in for libbasic_b200.so: records every C-ABI call; the y path "codes" by stashing the quantised tensor."""

    def __init__(self):
        self.calls, self.stash, self.pending = [], {}, b""

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append(name)
            impl = type(self).__dict__.get("_" + name)
            return impl(self, *args) if impl else 0
        return fn

    def _basic_device_count(self):
        return 1

    def _basic_coder_create(self, *args):
        args[-1]._obj.value = 0x1000 + len(self.calls)
        return 0

    def _basic_ctx_create(self, *args):
        args[-1]._obj.value = 0x2000 + len(self.calls)
        return 0

    def _basic_ypath_encode(self, h, ctx, y, prior, B, C, H, W, lanes, out, cap, out_len, yhat, stream):
        n = B * C * H * W
        arr = np.ctypeslib.as_array(ctypes.cast(y, ctypes.POINTER(ctypes.c_float)), shape=(n,)).copy()
        key = b"Y" + len(self.stash).to_bytes(3, "little")
        self.stash[key] = np.round(arr)
        self.pending = key
        return 0

    def _basic_coder_output_size(self, h):
        return len(self.pending)

    def _basic_coder_take_output(self, h, dst, n):
        ctypes.memmove(dst, self.pending, n)
        return 0

    def _basic_ypath_decode(self, h, ctx, enc, n_enc, prior, B, C, H, W, lanes, yhat, stream):
        key = ctypes.string_at(enc, n_enc)
        arr = self.stash[key].astype(np.float32)
        ctypes.memmove(yhat, arr.ctypes.data, arr.nbytes)
        return 0


def _need_reference():
    import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference (or oracle/_ref) is not available")
    ref_shim.load()


def test_reference_latent_graph_drives_the_dropins(monkeypatch):
    _need_reference()
    from cbench.modules.entropy_coder.latent_graph import LatentGraphicalANSEntropyCoder
    from cbench.nn.base import NNCacheImpl, NNTrainableModule
    from cbench_basic_b200 import _native, reference_integration, z_coder
    cls = reference_integration.bind()
    fake = _FakeLib()
    monkeypatch.setattr(_native, "_lib", fake)
    YCoder, ZCoder = cls["GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder"], cls["CompressAIEntropyBottleneckPriorCoder"]
    CtxModel = cls["TopoGroupDynamicMaskConv2dContextModel"]
    monkeypatch.setattr(YCoder, "_dev_index", lambda self: 0)
    monkeypatch.setattr(YCoder, "_stream", lambda self: 0)
    monkeypatch.setattr(YCoder, "_sync", lambda self: None)
    # the z coder's stream plumbing needs CUDA tensors: its two arithmetic entry points are replaced, the framing is real
    zstash = {}

    def z_compress(self, x):
        fake.calls.append("z.compress")
        med = self._get_medians().view(1, -1, 1, 1)
        out = []
        for b in range(x.shape[0]):
            key = b"Z" + len(zstash).to_bytes(3, "little")
            zstash[key] = (torch.round(x[b:b + 1] - med) + med).clone()
            out.append(key)
        return out

    def z_decompress(self, strings, size):
        fake.calls.append("z.decompress")
        return torch.cat([zstash[bytes(s)] for s in strings])

    monkeypatch.setattr(z_coder.EntropyBottleneck, "compress", z_compress)
    monkeypatch.setattr(z_coder.EntropyBottleneck, "decompress", z_decompress)
    monkeypatch.setattr(z_coder.EntropyBottleneck, "update", lambda self, force=False: fake.calls.append("z.update") or True)

    C, N = 12, 6
    torch.manual_seed(0)
    y_coder = YCoder(in_channels=C, default_topo_group_method="checkerboard",
                     topo_group_context_model=CtxModel(in_channels=C, out_channels=2 * C), ans_params_device="cpu")
    z_coder_ = ZCoder(entropy_bottleneck_channels=N)
    assert isinstance(y_coder, NNTrainableModule) and isinstance(z_coder_, NNCacheImpl)
    h_a = nn.Conv2d(C, N, 3, stride=2, padding=1)
    h_s = nn.Sequential(nn.ConvTranspose2d(N, 2 * C, 3, stride=2, padding=1, output_padding=1))
    # x -(g_a)-> y -(h_a)-> z ; z -(h_s)-> y -(g_s)-> x, the hyperprior graph of configs/lossy_latent_graph_topogroup.py:203-240
    # with identity backbones around y (the x node gets the reference's LossyDummyEntropyCoder)
    graph = LatentGraphicalANSEntropyCoder(
        use_lossy_compression=True,
        latent_node_entropy_coder_dict={"y": y_coder, "z": z_coder_},
        latent_inference_dict={"x_y": nn.Identity(), "y_z": h_a},
        latent_generative_dict={"z_y": h_s, "y_x": nn.Identity()},
        latent_node_inference_topo_order=["x", "y", "z"],
        latent_node_generative_topo_order=["z", "y", "x"],
    ).eval()
    assert graph.latent_node_entropy_coders["y"] is y_coder          # stored in the reference's nn.ModuleDict
    graph.update_state()                                              # latent_graph.py:1297-1301
    assert "z.update" in fake.calls and "basic_coder_init_params" in fake.calls and "basic_ctx_set_weights" in fake.calls
    fake.calls.clear()

    y = torch.randn(2, C, 8, 6) * 3
    data = graph.encode(y)                                            # latent_graph.py:1232-1263
    # _generative_process, do_encode: for z then y -- forward (no native call), then encode
    assert fake.calls == ["z.compress", "basic_ctx_set_map", "basic_ypath_encode", "basic_coder_output_size",
                          "basic_coder_take_output"], fake.calls
    from cbench_basic_b200.bytes_ops import split_merged_bytes
    z_bytes, y_bytes = split_merged_bytes(data, num_segments=2)
    assert y_bytes.startswith(b"Y") and z_bytes[:12] == (4).to_bytes(4, "big") + (3).to_bytes(4, "big") + (2).to_bytes(4, "big")
    fake.calls.clear()
    out = graph.decode(data)                                          # latent_graph.py:1265-1295
    assert fake.calls == ["z.decompress", "basic_ypath_decode"], fake.calls      # same map: not uploaded again
    assert torch.equal(out, torch.round(y))
    # the prior the y coder received is h_s(z_hat) with z_hat from the z coder's forward (encode) / decode: identical
    res = graph.collect_profiler_results(recursive=True)
    assert any("time_ans_encode" in k for k in res) and any("pgm_generate_coding" in k for k in res), sorted(res)
    # cache API reaches the drop-ins through the reference's recursion (cbench/nn/base.py:306-321)
    y_coder.update_cache("metric_dict", probe=1.0)
    assert any(k.endswith("latent_node_entropy_coders/y/probe") for k in graph.get_cache("metric_dict"))
    graph.reset_all_cache()
    assert y_coder.get_raw_cache("metric_dict") == {}


def test_install_patches_reference_modules():
    _need_reference()
    from cbench_basic_b200 import ans, reference_integration
    import cbench.modules.prior_model.prior_coder.pgm_coder as ref_pgm
    saved = {name: getattr(__import__(mod, fromlist=[name]), name) for name, (_, mod) in reference_integration.TARGETS.items()}
    saved_ans = sys.modules.get("cbench.ans")
    try:
        classes = reference_integration.install()
        assert ref_pgm.GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder is classes["GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder"]
        assert sys.modules["cbench.ans"] is ans
    finally:
        for name, (_, mod) in reference_integration.TARGETS.items():
            setattr(__import__(mod, fromlist=[name]), name, saved[name])
        import cbench
        if saved_ans is not None:
            sys.modules["cbench.ans"] = saved_ans
            cbench.ans = saved_ans


# ---------------------------------------------------------------------------------------------- the transcript on the GPU
@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 0])
def test_gpu_generative_process_transcript(lanes):
    """latent_graph.py:826-841 with do_encode and stream=None, then :1265-1295: per node in generative order
    ``node_data = coder(data, **prior_kwargs)``; ``bytes = coder.encode(data, **prior_kwargs)``; edges run on node_data;
    container = merge_bytes; decode: ``coder.decode(bytes, stream=None, **prior_kwargs)``."""
    from cbench_basic_b200.bytes_ops import merge_bytes, split_merged_bytes
    from cbench_basic_b200.prior_coder import (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as YCoder,
                                               TopoGroupDynamicMaskConv2dContextModel as CtxModel)
    from cbench_basic_b200.z_coder import CompressAIEntropyBottleneckPriorCoder as ZCoder
    torch.manual_seed(1)
    C, N = 24, 8
    coders = nn.ModuleDict({
        "z": ZCoder(entropy_bottleneck_channels=N),
        "y": YCoder(in_channels=C, default_topo_group_method="checkerboard", lanes=lanes,
                    topo_group_context_model=CtxModel(in_channels=C, out_channels=2 * C)),
    }).cuda().eval()
    h_a = nn.Conv2d(C, N, 3, stride=2, padding=1).cuda()
    # (a transposed convolution may pick an atomics-based cuDNN kernel whose sums differ run to run; encoder and decoder
    # must see the same prior bit for bit -- the same requirement the reference's codec has on its backbone)
    h_s = nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(N, 2 * C, 1)).cuda()
    for c in coders.values():
        c.update_state()
    with torch.no_grad():
        y = torch.randn(3, C, 16, 12, device="cuda") * 3
        latent = {"y": y, "z": h_a(y) * 4}
        data, prior, node_out = {}, {}, {}
        for node in ("z", "y"):                                       # encode side
            kw = {"prior": prior[node]} if node in prior else {}
            node_out[node] = coders[node](latent[node], **kw)        # forward BEFORE encode
            data[node] = coders[node].encode(latent[node], **kw)
            assert isinstance(data[node], bytes)
            if node == "z":
                prior["y"] = h_s(node_out["z"])
        blob = merge_bytes([data["z"], data["y"]], num_segments=2)
        zb, yb = split_merged_bytes(blob, num_segments=2)
        z_hat = coders["z"].decode(zb, stream=None)                   # decode side
        assert torch.equal(z_hat, node_out["z"])                      # forward == what the decoder reconstructs
        prior_dec = h_s(z_hat)
        assert torch.equal(prior_dec, prior["y"])
        y_hat = coders["y"].decode(yb, stream=None, prior=prior_dec)
    assert y_hat.shape == y.shape and float((y_hat - y).abs().max()) <= 0.5 + 1e-5
    assert float((node_out["y"] - torch.round(y)).abs().max()) == 0.0
    prof = coders["y"].collect_profiler_results()
    assert "time_ans_encode (ms)" in prof and "pgm_generate_coding (ms)" in prof

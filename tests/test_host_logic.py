"""CPU tests of the host-side logic: C-ABI surface, group maps, byte framing, sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cbench_basic_b200 import _native, build
    build.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    header = open(os.path.join(REPO, "include", "basic_b200.h")).read()
    declared = set(re.findall(r"\b(basic_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu():
    from cbench_basic_b200 import _native, ans
    if _native.lib().basic_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.CudaError):
        ans.Rans64Encoder()
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder
    coder = Coder(in_channels=8, use_param_merger=False)
    with pytest.raises(_native.CudaError):
        coder.update_state()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "cbench_basic_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/ans_oracle.c section 4", ""), f


def test_group_maps_match_reference_golden(golden_dir):
    from cbench_basic_b200 import topo_groups
    yv = np.load(os.path.join(golden_dir, "ypath_vectors.npz"))
    for name, method in [("ckbd", "checkerboard"), ("cwckbd", "channelwise-checkerboard"), ("scanline", "scanline"),
                         ("raster", "raster2x2"), ("meanscale", "none")]:
        C, G, B, H, W, _ = [int(v) for v in yv[name + ".meta"]]
        assert np.array_equal(topo_groups.default_map(method, G, H, W).numpy(), yv[name + ".tg"])
    for name in ("learned_int", "learned_logits"):
        C, G, B, H, W, _ = [int(v) for v in yv[name + ".meta"]]
        got = topo_groups.tile_map(torch.from_numpy(yv[name + ".pgm"]), G, H, W)
        assert np.array_equal(got.numpy(), yv[name + ".tg"])
    # odd sizes: only whole 2x2 patches are tiled, the leftover row / column stays group 0 (cfg 5: H = 135)
    tg = topo_groups.tile_map(torch.tensor([[[[1, 2], [3, 1]]]]), 1, 5, 5)
    assert tg[0, 0, 4].tolist() == [0] * 5 and tg[0, 0, :, 4].tolist() == [0] * 5 and tg[0, 0, 0, :4].tolist() == [1, 2, 1, 2]


def test_bytes_ops_framing():
    from cbench_basic_b200.bytes_ops import merge_bytes, split_merged_bytes
    parts = [b"abc", b"", b"\x00" * 7]
    assert split_merged_bytes(merge_bytes(parts)) == parts
    m = merge_bytes(parts[:2], num_segments=2)
    assert m == b"\x03\x00\x00\x00abc" and split_merged_bytes(m, num_segments=2) == parts[:2]


def test_partition_and_container():
    from cbench_basic_b200 import sharding
    for n, w in [(24, 8), (5, 4), (3, 8), (512, 8), (0, 2)]:
        got = [i for r in range(w) for i in sharding.partition(n, w, r)]
        assert got == list(range(n))
        sizes = [len(sharding.partition(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    assert [len(t) for t in sharding.row_band_tiles(135, 8)] == [17] * 7 + [16]
    streams = [b"", b"xyz", b"\x01" * 100]
    assert sharding.split_container(sharding.assemble_container(streams)) == streams


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from cbench_basic_b200 import sharding
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
enc = lambda u: bytes([u]) * (u + 1)          # stand-in for the per-image GPU encoder: unit u -> u+1 bytes
streams, sizes, off = sharding.encode_sharded(enc, 5, rank, 2)
flat, offs, total = sharding.unit_offsets(sizes)
assert sizes == [[1, 2, 3], [4, 5]], sizes
assert total == 15 and off == (0 if rank == 0 else 6), (total, off)
gathered = [None, None]
dist.all_gather_object(gathered, streams)
if rank == 0:
    container = sharding.assemble_container([s for r in gathered for s in r])
    assert sharding.split_container(container) == [enc(u) for u in range(5)]
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_sharded_encode_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), REPO, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_numa_binding_helper_is_safe_without_topology():
    """sharding.bind_to_gpu_numa: cpulist parsing, and a no-op (never an exception, affinity untouched) where there is no
    GPU / no sysfs topology -- the pool's boxes report numa_node = -1."""
    import os
    from cbench_basic_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = sharding.bind_to_gpu_numa(0)
    assert isinstance(info, dict) and info["bound"] is False
    assert os.sched_getaffinity(0) == before


def test_latent_codec_container_layout():
    """HyperpriorLatentCodec: segments in generative order (z, y), u32 length before all but the last
    (latent_graph.py:1260-1263 + bytes_ops.merge_bytes) -- with stub node coders, no GPU."""
    import struct
    import torch
    from cbench_basic_b200.latent_codec import HyperpriorLatentCodec

    class Stub(torch.nn.Module):
        def __init__(self, tag):
            super().__init__()
            self.tag, self.seen = tag, []

        def update_state(self):
            self.seen.append("update")

        def forward(self, x, prior=None):                    # latent_graph.py:836: forward before encode
            self.seen.append("fwd")
            return torch.full((1, 2), float(x.numel()))

        def encode(self, x, prior=None):
            self.seen.append(("enc", None if prior is None else float(prior.sum())))
            return self.tag * int(x.numel())

        def decode(self, b, prior=None):
            self.seen.append(("dec", bytes(b), None if prior is None else float(prior.sum())))
            return torch.full((1, 2), float(len(b)))

    z, y = Stub(b"z"), Stub(b"y")
    codec = HyperpriorLatentCodec(z, y, hyper_synthesis=lambda zh: zh * 2, hyper_analysis=lambda t: t[:, :3])
    codec.update_state()
    data = codec.encode(torch.ones(1, 5))
    assert data == struct.pack("I", 3) + b"zzz" + b"yyyyy"
    assert y.seen[-1] == ("enc", 12.0)                      # prior = h_s(z_hat) = 2 * [3, 3]
    assert z.seen == ["update", "fwd", ("enc", None)]
    out = codec.decode(data)
    assert z.seen[-1][:2] == ("dec", b"zzz") and y.seen[-1] == ("dec", b"yyyyy", 12.0) and float(out[0, 0]) == 5.0
    assert list(codec.state_dict().keys()) == [] and set(codec.latent_node_entropy_coders.keys()) == {"z", "y"}


def test_serial_coder_is_the_scanline_map_with_remapped_matrices(golden_dir):
    """prior_coder.joint_ar_remap (row f4): the reference's pixel-by-pixel loop (oracle restatement, byte-identical to the
    reference) and the masked-merger formulation the kernels run -- scanline map, [ctx | prior] columns, interleaved
    (mean, scale) rows -- produce the same symbols, scale indexes and reconstruction."""
    import numpy as np
    import torch
    from cbench_basic_b200.prior_coder import joint_ar_remap
    from oracle import ypath_oracle as Y
    jv = np.load(os.path.join(golden_dir, "ypath_jointar_vectors.npz"))
    for name in ("jar_a", "jar_b"):
        C_, B, H, W = [int(v) for v in jv[name + ".meta"]]
        sd = {k[len(name) + 4:]: torch.from_numpy(jv[k]) for k in jv.files if k.startswith(name + ".sd.")}
        w = Y.joint_ar_weights_from_state_dict(sd)
        y, prior = torch.from_numpy(jv[name + ".y"]), torch.from_numpy(jv[name + ".prior"])
        w0, w4, b4 = joint_ar_remap(w["e0_w"], w["e4_w"], w["e4_b"], C_)
        wm = {"ctx_w": w["ctx_w"], "ctx_b": w["ctx_b"], "m1_w": w0.reshape(*w0.shape, 1, 1), "m1_b": w["e0_b"],
              "m2_w": w["e2_w"], "m2_b": w["e2_b"], "m3_w": w4.reshape(*w4.shape, 1, 1), "m3_b": b4}
        tg = Y.default_pgm("scanline", 1, H, W)
        with torch.no_grad():
            sym_a, idx_a, yhat_a = Y.joint_ar_encode_symbols(y, prior, w, Y.get_scale_table())
            sym_b, idx_b, yhat_b = Y.encode_symbols(y, prior, tg, wm, Y.get_scale_table())
        assert np.array_equal(sym_a, sym_b) and np.array_equal(idx_a, idx_b)
        assert float((yhat_a - yhat_b).abs().max()) <= 1e-5


def test_reference_state_dicts_load_into_every_coder_variant(golden_dir):
    """Drop-in at the checkpoint level: the state_dict keys of the reference coder -- its internal context model with and
    without the param merger, the serial coder's entropy_parameters, plus the entries that carry nothing for coding
    (conv_kernel_*, lower_bound_scale.bound) -- load strictly (no GPU needed for that)."""
    import numpy as np
    import torch
    from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder

    def ref_sd(npz, name, C_):
        sd = {k[len(name) + 4:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(name + ".sd.")}
        sd.update({"conv_kernel_weight": torch.zeros(2 * C_, C_, 5, 5), "conv_kernel_bias": torch.zeros(2 * C_),
                   "lower_bound_scale.bound": torch.tensor([0.11])})
        return sd
    iv = np.load(os.path.join(golden_dir, "ypath_internal_vectors.npz"))
    C_, G = int(iv["int_cwckbd.meta"][0]), int(iv["int_cwckbd.meta"][1])
    coder = Coder(in_channels=C_, channel_groups=G, default_topo_group_method="channelwise-checkerboard")
    coder.load_state_dict(ref_sd(iv, "int_cwckbd", C_))
    assert set(coder.state_dict()) == {"context_prediction.weight", "context_prediction.bias"} | \
        {f"param_merger.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")}
    assert tuple(coder.param_merger[0].weight.shape) == (4 * C_, 4 * C_, 1, 1)
    wide = Coder(in_channels=C_, channel_groups=G, param_merger_expand_bottleneck=True)
    assert tuple(wide.param_merger[2].weight.shape) == (8 * C_, 8 * C_, 1, 1)
    jv = np.load(os.path.join(golden_dir, "ypath_jointar_vectors.npz"))
    Cj = int(jv["jar_a.meta"][0])
    serial = Coder(in_channels=Cj, use_joint_ar_model_impl=True)
    serial.load_state_dict(ref_sd(jv, "jar_a", Cj))
    assert tuple(serial.entropy_parameters[0].weight.shape) == (2 * Cj * 5 // 3, 4 * Cj, 1, 1)
    assert serial._get_pgm((1, Cj, 3, 4)).reshape(-1).tolist() == list(range(12))      # pixel = coding group
    plain = Coder(in_channels=Cj, use_param_merger=False)
    assert set(plain.state_dict()) == {"context_prediction.weight", "context_prediction.bias"}
    with pytest.raises(ValueError):
        Coder(in_channels=Cj, channel_groups=2, use_joint_ar_model_impl=True)

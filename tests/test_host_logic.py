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
    out = codec.decode(data)
    assert z.seen[-1][:2] == ("dec", b"zzz") and y.seen[-1] == ("dec", b"yyyyy", 12.0) and float(out[0, 0]) == 5.0
    assert list(codec.state_dict().keys()) == [] and set(codec.latent_node_entropy_coders.keys()) == {"z", "y"}

"""configs[3] (the combined coder): encode / decode time of every sub-coder on 64 crops of 256x256."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cbench_basic_b200 import _native as N
y, prior, w = bench.make_inputs("cfg4", 0)
dev = torch.device("cuda", 0)
coder = bench.build_coder("cfg4", w, 0, dev)
yd, pd = y.to(dev), prior.to(dev)
for level, (name, G, S) in enumerate(bench.CFG4_LEVELS):
    kw = {"blend_weight": torch.eye(len(bench.CFG4_LEVELS))[level]}
    if name != "scanline":
        kw["pgm"] = bench.learned_style_map(G, S, 100 + level)
    for rep in range(2):
        N.launch_count(reset=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bs = coder.encode(yd, prior=pd, **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = coder.decode(bs, prior=pd, **kw)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:9s} G={G} S={S:3d}: encode {1e3 * (t1 - t0):7.2f} ms  decode {1e3 * (t2 - t1):7.2f} ms  launches {N.launch_count()}  "
          f"bytes {len(bs)}  max err {float((out - yd).abs().max()):.3f}")

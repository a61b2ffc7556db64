"""One scanline (configs[2]) encode + decode of a Kodak-shape image: the target of ncu / timing runs on the many-stage path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cbench_basic_b200 import _native as N
y, prior, w = bench.make_inputs("cfg3", 0)
dev = torch.device("cuda", 0)
coder = bench.build_coder("cfg3", w, int(os.environ.get("LANES", "0")), dev)
yd, pd = y.to(dev), prior.to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    N.launch_count(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bs = coder.encode(yd, prior=pd)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    n_enc = N.launch_count()
    out = coder.decode(bs, prior=pd)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"encode {1e3 * (t1 - t0):.1f} ms  decode {1e3 * (t2 - t1):.1f} ms  launches {n_enc} / {N.launch_count() - n_enc}  bytes {len(bs)}  max err {float((out - yd).abs().max()):.3f}")

"""Two steps (encode + decode) of a bench workload, nothing else: the command ncu wraps (profiles/README.md)."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
y, prior, w = bench.make_inputs(wl, 0)
coder = bench.build_coder(wl, w, lanes, dev)
yd, pd = y.to(dev), prior.to(dev)
for _ in range(steps):
    bs = coder.encode(yd, prior=pd)
    out = coder.decode(bs, prior=pd)
torch.cuda.synchronize()
print("ok", len(bs), float((out - yd).abs().max()))

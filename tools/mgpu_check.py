"""torchrun --nproc-per-node N tools/mgpu_check.py: round trips of the bench workload on every rank's own GPU and inputs."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
rank, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
wl = "cfg2"
y, prior, w = bench.make_inputs(wl, rank)
coder = bench.build_coder(wl, w, 0, dev)
yd, pd = y.to(dev), prior.to(dev)
yp, pp = y.pin_memory(), prior.pin_memory()
ref = None
for it in range(int(os.environ.get("ITERS", 4))):
    host = it % 2 == 1
    bs, yhat_enc = coder.encode(yp if host else yd, prior=pp if host else pd, return_yhat=True)
    out = coder.decode(bs, prior=pp if host else pd)
    bad = out != yhat_enc
    nbad = int(bad.sum())
    err = float((out - yd).abs().max())
    eerr = float((yhat_enc - yd).abs().max())
    if ref is None:
        ref = bs
    msg = f"[rank {rank} it {it} host={host}] bytes {len(bs)} same-as-first {bs == ref} nbad {nbad} max|out-y| {err:.3f} max|enc-y| {eerr:.3f}"
    if nbad:
        idx = bad.nonzero()
        par = ((idx[:, 2] + idx[:, 3]) % 2)
        msg += f" | images {sorted(set(idx[:, 0].tolist()))[:8]} parity0 {int((par == 0).sum())} parity1 {int((par == 1).sum())} first {idx[0].tolist()}"
    print(msg, flush=True)

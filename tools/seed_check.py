"""Round trip of the bench workload for the inputs of ranks 0..7 on one GPU (each rank draws its own seed)."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
for rank in range(8):
    y, prior, w = bench.make_inputs(wl, rank)
    for lanes in (0, 1):
        coder = bench.build_coder(wl, w, lanes, dev)
        yd, pd = y.to(dev), prior.to(dev)
        bs, yhat_enc = coder.encode(yd, prior=pd, return_yhat=True)
        out = coder.decode(bs, prior=pd)
        same = torch.equal(out, yhat_enc * 1.0 + 0.0)
        err = float((out - yd).abs().max())
        nbad = int((out != yhat_enc).sum())
        print(f"rank {rank} lanes {lanes}: decoder==encoder {same} (differing {nbad}), max|yhat-y| {err:.4f}, enc max|yhat-y| {float((yhat_enc - yd).abs().max()):.4f}, bytes {len(bs)}", flush=True)

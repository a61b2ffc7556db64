"""BASELINE configs[2]: the scanline (fixed-AR) map on one Kodak-shape image -- 1536 coding groups, one launch wave each."""
import os, sys, time, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from cbench_basic_b200.prior_coder import (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder,
                                           TopoGroupDynamicMaskConv2dContextModel as Ctx)
from oracle import ypath_oracle as Y
from cbench_basic_b200 import _native as N
dev = torch.device("cuda", 0)
C_, H, W = 192, 32, 48
w = Y.random_weights(C_, 1234)
cm = Ctx(in_channels=C_, out_channels=2 * C_)
cm.load_state_dict({"context_prediction.weight": w["ctx_w"], "context_prediction.bias": w["ctx_b"],
                    "param_merger_in.weight": w["m1_w"], "param_merger_in.bias": w["m1_b"],
                    "param_merger_out.1.weight": w["m2_w"], "param_merger_out.1.bias": w["m2_b"],
                    "param_merger_out.3.weight": w["m3_w"], "param_merger_out.3.bias": w["m3_b"]})
for lanes in (0, 1):
    coder = Coder(in_channels=C_, default_topo_group_method="scanline", topo_group_context_model=cm, lanes=lanes).to(dev).eval()
    coder.update_state()
    g = torch.Generator().manual_seed(0)
    y, prior = 3 * torch.randn(1, C_, H, W, generator=g), torch.randn(1, 2 * C_, H, W, generator=g)
    yd, pd = y.to(dev), prior.to(dev)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bs, yh_enc = coder.encode(yd, prior=pd, return_yhat=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = coder.decode(bs, prior=pd)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    N.profile(True); N.profile_read()
    bs = coder.encode(yd, prior=pd); torch.cuda.synchronize(); ph_e = N.profile_read()
    out = coder.decode(bs, prior=pd); torch.cuda.synchronize(); ph_d = N.profile_read(); N.profile(False)
    print('  encode phases (ms):', {k: round(v[0], 1) for k, v in ph_e.items() if v[0]}, '\n  decode phases (ms):', {k: round(v[0], 1) for k, v in ph_d.items() if v[0]})
    ok = torch.equal(out, yh_enc * 1.0 + 0.0) and float((out - yd).abs().max()) <= 0.5 + 1e-5
    print(f"scanline lanes={lanes}: {len(bs)} B, encode {1e3 * (t1 - t0):.1f} ms, decode {1e3 * (t2 - t1):.1f} ms, round trip {'ok' if ok else 'FAILED'}")

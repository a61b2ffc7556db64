import os, sys, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tests.test_oracle_golden import load_ycase
from tests.test_gpu_ypath import make_coder
from cbench_basic_b200 import _native as N
yv = np.load(os.path.join(REPO, "tests/golden/ypath_vectors.npz"))
for name in ("ckbd", "raster"):
    c = load_ycase(yv, name)
    coder = make_coder(c, lanes=1)
    kw = dict(prior=c["prior"].cuda(), pgm=c["tg"])
    bs, yhat_enc = coder.encode(c["y"].cuda(), return_yhat=True, **kw)
    print(name, "bytes equal", bs == c["bytes"], len(bs), len(c["bytes"]))
    ye = yhat_enc.cpu()
    print(" enc yhat mismatches", int((ye != c["yhat"]).sum()))
    yd = coder.decode(bs, **kw).cpu()
    bad = (yd != c["yhat"])
    print(" dec mismatches", int(bad.sum()), "of", bad.numel())
    if bad.any():
        ii = bad.nonzero()[:10]
        tg = c["tg"]
        for b, ch, h, w in ii.tolist():
            print("   ", (b, ch, h, w), "grp", int(tg[0, ch // (c["C"] // c["G"]), h, w]), float(yd[b, ch, h, w]), float(c["yhat"][b, ch, h, w]))
        gmap = tg[0].repeat_interleave(c["C"] // c["G"], 0).unsqueeze(0).expand_as(bad)
        for g in range(int(tg.max()) + 1):
            print("   group", g, "bad", int((bad & (gmap == g)).sum()), "of", int((gmap == g).sum()))
    # params abs error per stage
    B, C_, H, W = c["B"], c["C"], c["H"], c["W"]
    buf, prior = c["yhat"].cuda().contiguous(), c["prior"].cuda().contiguous()
    params = torch.full((B, 2 * C_, H, W), float("nan"), device="cuda")
    for g in range(N.lib().basic_ctx_num_stages(coder._ctx)):
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, buf.data_ptr(), prior.data_ptr(), B, params.data_ptr(), 0))
    torch.cuda.synchronize()
    d = (params.cpu() - c["params_full"]).abs()
    print(" params max abs err", float(d.max()), "max |ref|", float(c["params_full"].abs().max()), "rel-to-max(1,|ref|)", float((d / c["params_full"].abs().clamp_min(1)).max()))

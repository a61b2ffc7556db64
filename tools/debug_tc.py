"""Debug helper: tensor-core context model vs the FP32 kernel, error localisation."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tests.test_gpu_ctx_tc import build, stage_params, CASES

for (C_, G, method, B, H, W, merger) in CASES:
    for nacc in (1,):
        torch.manual_seed(1)
        buf = (3 * torch.randn(B, C_, H, W)).round().cuda() + torch.randn(B, C_, H, W).cuda()
        prior = torch.randn(B, 2 * C_, H, W).cuda()
        ref, S = stage_params(build(C_, G, method, "fp32", 1, merger=merger), buf, prior)
        got, _ = stage_params(build(C_, G, method, "tf32x3", nacc, merger=merger), buf, prior)
        d = (got - ref).abs() / ref.abs().clamp_min(1.0)
        print(f"C={C_} G={G} {method} merger={merger} S={S} nacc={nacc}: max rel err {float(d.max()):.3e} mean {float(d.mean()):.3e}"
              f" nan={int(torch.isnan(got).sum())}")
        # where
        per_ch = d.amax(dim=(0, 2, 3))
        per_b = d.amax(dim=(1, 2, 3))
        per_hw = d.amax(dim=(0, 1))
        print("  per-batch", [f"{float(v):.1e}" for v in per_b])
        print("  per-channel (first 16)", [f"{float(v):.1e}" for v in per_ch[:16]], " max at ch", int(per_ch.argmax()))
        bad = (per_ch > 1e-4).nonzero().flatten().tolist()
        print("  bad channels:", len(bad), bad[:20], "...", bad[-5:])
        print("  per-hw row0", [f"{float(v):.0e}" for v in per_hw[0][:12]])
        print("  per-hw row1", [f"{float(v):.0e}" for v in per_hw[1][:12]])
        i = int(d.flatten().argmax()); print("  worst: got", float(got.flatten()[i]), "ref", float(ref.flatten()[i]))

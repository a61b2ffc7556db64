"""Coder alone at trained-model rates: the lane sweep of bench.py without the rest (A/B runs while tuning the kernels)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cbench_basic_b200.prior_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder  # noqa: E402

coder = Coder(in_channels=8, use_param_merger=False).cuda().eval()
r = bench.coder_lane_sweep(coder, 0, bench.peaks()[0], n_sym=int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24)
if "rows" not in r:
    print(r)
for row in r.get("rows", []):
    print(f"lanes {row['lanes']:7d} chunks {row['chunks']:5d}  dec {row['decode_ms']:.3f} ms ({row['decode_frac_of_hbm']:.4f})  "
          f"enc {row['encode_ms']:.3f} ms ({row['encode_frac_of_hbm']:.4f})  size {row['size_vs_lanes1']:+.4f}")

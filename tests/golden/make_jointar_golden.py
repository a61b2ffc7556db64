"""Generates tests/golden/ypath_jointar_vectors.npz: the reference coder with use_joint_ar_model_impl=True (the CompressAI-style
serial coder, pgm_coder.py:1975-2066 -- SURVEY 8 row f4), run unmodified through tests/golden/ref_shim.py.  Build container:

    make -C oracle ref && python tests/golden/make_jointar_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402


def main():
    Coder, _ = ref_shim.load()
    out = {}
    for name, C, B, H, W, seed in [("jar_a", 24, 2, 5, 7, 20), ("jar_b", 12, 1, 6, 6, 21)]:
        torch.manual_seed(seed)
        coder = Coder(in_channels=C, use_joint_ar_model_impl=True)
        with torch.no_grad():
            for prm in coder.parameters():
                prm.add_(0.05 * torch.randn_like(prm))
        coder.eval()
        coder.update_state()
        y, p = 3 * torch.randn(B, C, H, W), torch.randn(B, 2 * C, H, W)
        with torch.no_grad():
            bs = coder.encode(y, prior=p)
            yh = coder.decode(bs, prior=p)
        for k, v in coder.state_dict().items():
            if k.startswith("context_prediction") or k.startswith("entropy_parameters"):
                out[f"{name}.sd.{k}"] = v.detach().cpu().numpy()
        out[f"{name}.meta"] = np.array([C, B, H, W], np.int32)
        out[f"{name}.y"], out[f"{name}.prior"] = y.numpy(), p.numpy()
        out[f"{name}.bytes"], out[f"{name}.yhat"] = np.frombuffer(bs, dtype=np.uint8), yh.numpy()
        print(name, "bytes", len(bs), "max|yhat-y|", float((yh - y).abs().max()))
    np.savez_compressed(os.path.join(HERE, "ypath_jointar_vectors.npz"), **out)


if __name__ == "__main__":
    main()

"""Generates tests/golden/ypath_internal_vectors.npz: the y path of the reference coder's INTERNAL context model
(context_prediction + param_merger over 2G channel groups, no topo_group_context_model; pgm_coder.py:1177-1239,
:1606-1638 -- SURVEY 8 row a14), run unmodified through tests/golden/ref_shim.py.  Build container only:

    make -C oracle ref && python tests/golden/make_internal_golden.py
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
    cases = [("int_ckbd", 24, 1, "checkerboard", 2, 6, 8, 10, False), ("int_cwckbd", 24, 2, "channelwise-checkerboard", 2, 5, 7, 11, False),
             ("int_raster", 12, 1, "raster2x2", 1, 6, 6, 12, False)]
    generate(Coder, cases, "ypath_internal_vectors.npz")
    # the expanded bottleneck (param_merger_expand_bottleneck=True: 4C -> 8C -> 8C -> 4C), kept in its own file
    generate(Coder, [("int_expand", 12, 2, "channelwise-checkerboard", 2, 4, 6, 13, True)], "ypath_internal_expand_vectors.npz")


def generate(Coder, cases, filename):
    out = {}
    for name, C, G, method, B, H, W, seed, expand in cases:
        torch.manual_seed(seed)
        coder = Coder(in_channels=C, channel_groups=G, default_topo_group_method=method, param_merger_expand_bottleneck=expand)
        with torch.no_grad():   # the masked convolutions start from a structured init: perturb so every weight matters
            for prm in coder.parameters():
                prm.add_(0.05 * torch.randn_like(prm))
        coder.eval()
        coder.update_state()
        y, p = 3 * torch.randn(B, C, H, W), torch.randn(B, 2 * C, H, W)
        with torch.no_grad():
            bs = coder.encode(y, prior=p)
            yh = coder.decode(bs, prior=p)
            tg = coder._get_pgm(y, input_shape=y.shape, pgm=None, fast_mode=True)
            params_full = coder._pgm_inference_group_mask(yh, None, pgm=tg, prior=p)
        for k, v in coder.state_dict().items():
            if k.startswith("context_prediction") or k.startswith("param_merger"):
                out[f"{name}.sd.{k}"] = v.detach().cpu().numpy()
        out[f"{name}.meta"] = np.array([C, G, B, H, W], np.int32)
        out[f"{name}.method"] = np.array(method)
        out[f"{name}.y"], out[f"{name}.prior"] = y.numpy(), p.numpy()
        out[f"{name}.tg"] = tg.cpu().numpy().astype(np.int32)
        out[f"{name}.bytes"], out[f"{name}.yhat"] = np.frombuffer(bs, dtype=np.uint8), yh.numpy()
        out[f"{name}.params_full"] = params_full.numpy()
        print(name, "bytes", len(bs), "max|yhat-y|", float((yh - y).abs().max()))
    np.savez_compressed(os.path.join(HERE, filename), **out)


if __name__ == "__main__":
    main()

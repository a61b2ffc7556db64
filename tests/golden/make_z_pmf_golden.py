"""Generates tests/golden/z_pmf_vectors.npz: the cumulative-logits network of the z node's factorized density evaluated by
the reference's OWN in-tree copy of it (cbench/nn/layers/param_generator.py:158-199, class
DifferentiableIncreasingVectorGenerator._cumulative -- same parameter names and arithmetic as compressai's
EntropyBottleneck._logits_cumulative) on seeded, perturbed parameters at the sample points `update()` uses.  Build container only:

    make -C oracle ref && python tests/golden/make_z_pmf_golden.py

This pins the network evaluation of row f1 to reference code; what stays restated from compressai 1.2.3's published
algorithm is the three lines after it (sign trick, sigmoid difference, tail mass)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from oracle import z_oracle as Z  # noqa: E402


def main():
    ref_shim.load()
    from cbench.nn.layers.param_generator import DifferentiableIncreasingVectorGenerator as Net   # the reference's code
    out = {}
    for name, (C, seed) in {"small": (8, 1), "c192": (192, 2)}.items():
        sd = Z.init_params(C, seed=seed)
        net = Net(C)
        with torch.no_grad():
            for k, v in sd.items():
                if k != "quantiles":
                    getattr(net, k).copy_(v)
        q = sd["quantiles"].float()
        medians = q[:, 0, 1]
        minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
        maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
        n = int((maxima + minima + 1).max())
        samples = torch.arange(n, dtype=torch.float32)[None, :] + (medians - minima)[:, None, None]
        with torch.no_grad():
            lower = net._cumulative(samples - 0.5, True)       # `inputs += bias` is in place: hand over fresh tensors
            upper = net._cumulative(samples + 0.5, True)
        for k, v in sd.items():
            out[f"{name}/sd/{k}"] = v.numpy()
        out[f"{name}/samples"] = samples.numpy()
        out[f"{name}/lower"] = lower.numpy()
        out[f"{name}/upper"] = upper.numpy()
        print(name, "C", C, "samples", tuple(samples.shape), "logit range", float(lower.min()), float(upper.max()))
    np.savez_compressed(os.path.join(HERE, "z_pmf_vectors.npz"), **out)


if __name__ == "__main__":
    main()

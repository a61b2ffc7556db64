"""Generates tests/golden/z_vectors.npz (z-node factorized-prior coder).  Build container only:

    make -C oracle ref && python tests/golden/make_z_golden.py

What comes from the UNMODIFIED reference: the coder and pmf_to_quantized_cdf (cbench.rans compiled into oracle/_ref) and the
framing (write_body / read_body imported from cbench/modules/prior_model/prior_coder/compressai_coder.py through
ref_shim).  What does not: the pmf evaluation (oracle/z_oracle.py restates compressai 1.2.3, which is not installed --
parity unpinned for that step; the vectors store the float pmf inputs' results, i.e. the tables, so the coder and framing
are pinned independently of it)."""
import io
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
    from cbench.modules.prior_model.prior_coder import compressai_coder as cc   # the reference's own framing
    out = {}
    for name, (C, B, H, W, seed) in {"small": (8, 3, 4, 6, 1), "c192": (192, 2, 8, 12, 2)}.items():
        sd = Z.init_params(C, seed=seed)
        tables = Z.build_tables(sd)
        g = torch.Generator().manual_seed(100 + seed)
        q = sd["quantiles"][:, 0, :]
        x = q[:, 1].view(1, C, 1, 1) + (q[:, 2] - q[:, 0]).view(1, C, 1, 1) * 0.45 * torch.randn(B, C, H, W, generator=g)
        x[0, 0, 0, 0] = 300.0     # escapes on both sides
        x[0, 1, 0, 1] = -250.0
        strings = Z.compress(sd, tables, x)
        with io.BytesIO() as bio:
            cc.write_body(bio, x.size()[-2:], [[s] for s in strings])
            body = bio.getvalue()
        with io.BytesIO(body) as bio:
            back, shape = cc.read_body(bio)
        assert [b[0] for b in back] == strings and tuple(shape) == (H, W)
        y_hat = Z.decompress(sd, tables, strings, (H, W))
        assert float((y_hat - x).abs().max()) <= 0.5 + 1e-4
        for k, v in sd.items():
            out[f"{name}/sd/{k}"] = v.numpy()
        out[f"{name}/cdf"], out[f"{name}/cdf_length"], out[f"{name}/offset"] = tables
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/y_hat"] = y_hat.numpy()
        out[f"{name}/body"] = np.frombuffer(body, dtype=np.uint8)
        print(name, "C", C, "table widths", tables[1].min(), "..", tables[1].max(), "body bytes", len(body))
    np.savez_compressed(os.path.join(HERE, "z_vectors.npz"), **out)


if __name__ == "__main__":
    main()

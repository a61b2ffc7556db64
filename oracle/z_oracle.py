"""CPU restatement of the z-node factorized-prior coder -- TEST INFRASTRUCTURE ONLY (tests/ and the
golden generator import it; the product path never does).

The reference (compressai_coder.py:87-248) calls compressai==1.2.3 (requirements.txt:15), which is NOT in /root/reference
and not installed here.  Restated from compressai's published algorithm, anchored on the reference tree:
  * _logits_cumulative           <- cbench/nn/layers/param_generator.py:182-199 (in-tree copy of the same network)
  * update / _pmf_to_cdf         <- compressai entropy_models.py (EntropyBottleneck.update, EntropyModel._pmf_to_cdf)
  * compress / decompress        <- EntropyModel.compress / decompress: round(x - median), one stream per image
  * coder + pmf_to_quantized_cdf <- the UNMODIFIED cbench.rans (cbench/csrc/rans/rans_interface.cpp, in-tree clone of
                                    compressai.ans) compiled into oracle/_ref
  * framing                      <- compressai_coder.py:63-84
PARITY: coder, CDF quantisation and framing are the reference's own code (pinned); the pmf evaluation out of the network
parameters is **parity unpinned** (nothing here can run compressai's EntropyBottleneck).
"""
import struct

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_loader


def init_params(channels, filters=(3, 3, 3, 3), init_scale=10.0, seed=0, spread=0.3):
    """Seeded parameters in the state_dict layout of the module: compressai's initialisation plus a perturbation, so that
    the channels get different widths, medians and shapes (an untrained model has identical channels)."""
    g = torch.Generator().manual_seed(seed)
    f = (1,) + tuple(filters) + (1,)
    scale = init_scale ** (1 / (len(filters) + 1))
    sd = {}
    for i in range(len(filters) + 1):
        init = float(np.log(np.expm1(1 / scale / f[i + 1])))
        sd[f"_matrix{i}"] = torch.full((channels, f[i + 1], f[i]), init) + spread * torch.randn(channels, f[i + 1], f[i], generator=g)
        sd[f"_bias{i}"] = torch.rand(channels, f[i + 1], 1, generator=g) - 0.5
        if i < len(filters):
            sd[f"_factor{i}"] = spread * torch.randn(channels, f[i + 1], 1, generator=g)
    med = 2.0 * torch.randn(channels, generator=g)
    half = 2.0 + 12.0 * torch.rand(channels, generator=g)
    sd["quantiles"] = torch.stack([med - half, med, med + 0.7 * half], dim=1).unsqueeze(1)
    return sd


def logits_cumulative(sd, x, n_filters=4):
    for i in range(n_filters + 1):
        x = torch.matmul(F.softplus(sd[f"_matrix{i}"]), x)
        x = x + sd[f"_bias{i}"]
        if i < n_filters:
            x = x + torch.tanh(sd[f"_factor{i}"]) * torch.tanh(x)
    return x


def build_tables(sd, precision=16):
    """-> (quantized_cdf int32 [C, max_len + 2], cdf_length [C], offset [C]) with the reference's pmf_to_quantized_cdf."""
    R = ref_loader.load("rans")
    q = sd["quantiles"].float()
    medians = q[:, 0, 1]
    minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
    maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
    offset = -minima
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length, dtype=torch.float32)[None, :] + pmf_start[:, None, None]
    lower, upper = logits_cumulative(sd, samples - 0.5), logits_cumulative(sd, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    C = q.shape[0]
    cdf = np.zeros((C, max_length + 2), dtype=np.int32)
    for c in range(C):
        prob = torch.cat((pmf[c, :int(pmf_length[c])], tail[c]), dim=0)
        qc = R.pmf_to_quantized_cdf(prob.tolist(), precision)
        cdf[c, :len(qc)] = qc
    return cdf, (pmf_length + 2).numpy().astype(np.int32), offset.numpy().astype(np.int32)


def quantize(sd, x):
    med = sd["quantiles"][:, 0, 1].float().view(1, -1, 1, 1)
    return torch.round(x.float() - med).to(torch.int32)


def compress(sd, tables, x):
    R = ref_loader.load("rans")
    cdf, lens, offs = tables
    sym = quantize(sd, x)
    B, C, H, W = x.shape
    idx = np.broadcast_to(np.arange(C, dtype=np.int32)[:, None, None], (C, H, W)).reshape(-1)
    enc = R.RansEncoder()
    return [enc.encode_with_indexes(sym[i].reshape(-1).tolist(), idx.tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
            for i in range(B)]


def decompress(sd, tables, strings, size):
    R = ref_loader.load("rans")
    cdf, lens, offs = tables
    C = cdf.shape[0]
    H, W = size
    idx = np.broadcast_to(np.arange(C, dtype=np.int32)[:, None, None], (C, H, W)).reshape(-1)
    med = sd["quantiles"][:, 0, 1].float().view(-1, 1, 1)
    dec = R.RansDecoder()
    out = torch.empty(len(strings), C, H, W)
    for i, s in enumerate(strings):
        v = dec.decode_with_indexes(s, idx.tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
        out[i] = torch.tensor(v, dtype=torch.float32).view(C, H, W) + med
    return out


def write_body(shape, strings):
    out = struct.pack(">3I", int(shape[0]), int(shape[1]), len(strings))
    for s in strings:
        out += struct.pack(">I", len(s)) + s
    return out


def read_body(data):
    h, w, n = struct.unpack_from(">3I", data, 0)
    at, strings = 12, []
    for _ in range(n):
        (ln,) = struct.unpack_from(">I", data, at)
        strings.append(data[at + 4:at + 4 + ln])
        at += 4 + ln
    return strings, (h, w)

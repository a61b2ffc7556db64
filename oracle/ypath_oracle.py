"""CPU restatement of the reference's y-node coding path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with plain torch-CPU / numpy ops, what the reference's
GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder does in encode()/decode()/update_state()
(cbench/modules/prior_model/prior_coder/pgm_coder.py, torch_ans.py; cbench/nn/layers/masked_conv.py).
Each function cites the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Parity status: PINNED -- checked against the unmodified reference Python (tests/golden/ref_shim.py, in the
build container where /root/reference exists) and against the committed golden vectors it produced
(tests/golden/*.npz, tests/test_oracle_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ans_oracle

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS):
    """compressai_coder.py:23-30."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def gaussian_ans_params(scale_table, freq_precision=16, lower_bound_scale=0.11, device="cpu"):
    """torch_ans.py:284-310 _get_ans_params with pgm_coder.py:757-799 (_params_to_dist/_init_dist_params):
    per scale, integer counts of a zero-mean Normal.  Returns (freqs [T, max] int32, nsym [T], offsets [T]).
    The float32 torch calls are kept verbatim: their last-ulp behaviour decides truncated counts."""
    freq_cnt = 1 << freq_precision
    tail_mass = torch.tensor([0.5 / freq_cnt], device=device)
    bound = torch.tensor([float(lower_bound_scale)], device=device)
    cnts, nsym, offs = [], [], []
    for s in scale_table.to(device=device):
        mean = torch.zeros(1, 1, device=device)
        scale = torch.max(s.reshape(1, 1), bound)           # LowerBound(0.11) forward
        dist = torch.distributions.Normal(mean, scale)
        dmin = int(dist.icdf(tail_mass).floor().item())
        dmax = int(dist.icdf(1 - tail_mass).ceil().item())
        offs.append(dmin)
        nsym.append(dmax - dmin + 1)
        pts = torch.arange(dmin - 1, dmax + 1).type_as(dist.mean) + 0.5
        logp = (dist.cdf(pts[1:]) - dist.cdf(pts[:-1])).log()[0]
        pmf = torch.softmax(logp, dim=-1)
        cnt = (pmf * freq_cnt).clamp_min(1)
        cnts.append(cnt.detach().cpu().contiguous().numpy().astype(np.int32))
    freqs = np.zeros((len(cnts), max(len(c) for c in cnts)), dtype=np.int32)
    for i, c in enumerate(cnts):
        freqs[i, :len(c)] = c
    return freqs, np.array(nsym, dtype=np.int32), np.array(offs, dtype=np.int32)


def select_indexes(scales, scale_table):
    """pgm_coder.py:802-821: argmin_t |sigma - table[t]|, linear domain, no clamp, first minimum."""
    idx = (scales.reshape(-1).unsqueeze(-1) - scale_table.type_as(scales).unsqueeze(0)).abs().argmin(-1)
    return idx.reshape_as(scales)


def split_mean_scale(params):
    """pgm_coder.py:735-755 with mean_scale_split_method="split_interleave": ch 2c = mean, 2c+1 = scale."""
    return params[:, 0::2], params[:, 1::2]


# ------------------------------------------------------------------------------------------- group maps
def default_pgm(method, G, H, W):
    """pgm_coder.py:1416-1491 _get_default_pgm -> int64 (1, G, H, W)."""
    tg = torch.zeros(1, G, H, W, dtype=torch.long)
    if method == "none":
        pass
    elif method == "scanline":
        tg = torch.arange(H * W).reshape(1, 1, H, W).repeat(1, G, 1, 1)
    elif method == "zigzag":
        tg = (torch.arange(H).reshape(H, 1) + torch.arange(W).reshape(1, W)).reshape(1, 1, H, W).repeat(1, G, 1, 1)
    elif method == "checkerboard":
        tg[..., 0::2, 1::2] = 1
        tg[..., 1::2, 0::2] = 1
    elif method == "raster2x2":
        tg[..., 0::2, 1::2] = 1
        tg[..., 1::2, 0::2] = 2
        tg[..., 1::2, 1::2] = 3
    elif method == "channelwise":
        for i in range(G):
            tg[:, i] = i
    elif method == "channelwise-checkerboard":
        for i in range(G):
            tg[:, i] = i * 2
            tg[:, i, 1::2, 0::2] = i * 2 + 1
            tg[:, i, 0::2, 1::2] = i * 2 + 1
    else:
        raise NotImplementedError(method)
    return tg


def tile_pgm(pgm, G, H, W):
    """pgm_coder.py:1342-1385 _preprocess_pgm (integer map, or logits with fast_mode argmax): trim, then
    tile whole patches with F.fold -- leftover rows/cols keep fold's zero fill (group 0)."""
    if torch.is_floating_point(pgm):
        pgm = pgm.reshape(pgm.shape[0], G, pgm.shape[1] // G, *pgm.shape[2:]).movedim(2, -1).argmax(-1)
    assert pgm.shape[1] == G
    tg = pgm[:, :, :H, :W]
    ph, pw = tg.shape[2:4]
    if ph < H or pw < W:
        reps = (H // ph) * (W // pw)
        cols = tg.reshape(tg.shape[0], -1, 1).repeat(1, 1, reps)
        tg = F.fold(cols.float(), (H, W), (ph, pw), stride=(ph, pw)).type_as(pgm)
    return tg


def group_of_elements(tg, B, C):
    """pgm_coder.py:1763-1771: element (b,c,h,w) belongs to group tg[0, c // (C/G), h, w]."""
    G = tg.shape[1]
    return tg.unsqueeze(2).repeat(B // tg.shape[0], 1, C // G, 1, 1).reshape(B, C, *tg.shape[2:])


# --------------------------------------------------------------------------------------- context model
def masked_conv(x, tg, weight, bias, allow_same, out_group_mask=None):
    """masked_conv.py:102-228 TopoGroupDynamicMaskConv2d.forward (inference branch): im2col, visibility
    mask unfold(tg) < / <= centre, one masked matmul per output channel group."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    pad = (kh // 2, kw // 2)
    cols = F.unfold(x, (kh, kw), padding=pad).unsqueeze(1)                       # B,1,Cin*k2,L
    tgf = tg.type_as(x)
    tgo = tgf - tgf.max().ceil() - 1                                            # padding (0) = "future"
    Gin = tg.shape[1]
    centre = tgo.reshape(tg.shape[0], Gin, 1, -1)
    nb = F.unfold(tgo, (kh, kw), padding=pad).unsqueeze(1)                       # 1,1,Gin*k2,L
    vis = (nb <= centre) if allow_same else (nb < centre)                       # 1,Gin(out),Gin*k2,L
    vis = vis.reshape(tg.shape[0], Gin, Gin, kh * kw, -1).repeat(1, 1, 1, Cin // Gin, 1) \
        .reshape(tg.shape[0], Gin, Cin * kh * kw, -1)
    if out_group_mask is not None:
        vis = vis[:, out_group_mask]
    Gout = vis.shape[1]
    masked = cols * vis
    out = weight.reshape(1, Gout, Cout // Gout, Cin * kh * kw).matmul(masked)
    out = out + bias.reshape(1, Gout, Cout // Gout, 1)
    return out.reshape(B, Cout, H, W)


def context_model(buf, tg, prior, w):
    """masked_conv.py:287-305 TopoGroupDynamicMaskConv2dContextModel.forward.  `w` is a dict with
    ctx_w/ctx_b, m1_w/m1_b, m2_w/m2_b, m3_w/m3_b (the state_dict tensors)."""
    G = tg.shape[1]
    ctx = masked_conv(buf, tg, w["ctx_w"], w["ctx_b"], allow_same=False)
    cat = torch.cat([ctx, prior], dim=1)
    cat_tg = torch.cat([tg, torch.zeros_like(tg) - 1], dim=1)
    m = masked_conv(cat, cat_tg, w["m1_w"], w["m1_b"], allow_same=True, out_group_mask=[True] * G + [False] * G)
    m = masked_conv(F.leaky_relu(m), tg, w["m2_w"], w["m2_b"], allow_same=True)
    return masked_conv(F.leaky_relu(m), tg, w["m3_w"], w["m3_b"], allow_same=True)


def internal_merger(buf, tg, prior, w):
    """The coder's own context_prediction + param_merger (no topo_group_context_model): pgm_coder.py:1177-1239 and
    _merge_prior_params :1606-1638 -- three masked 1x1 convolutions over 2G channel groups [ctx groups (ids = map),
    prior groups (id -1)], all with the <= rule, every out-group computed in every layer, the first G kept at the end."""
    G = tg.shape[1]
    ctx = masked_conv(buf, tg, w["ctx_w"], w["ctx_b"], allow_same=False)
    cat_tg = torch.cat([tg, torch.zeros_like(tg) - 1], dim=1)
    m = masked_conv(torch.cat([ctx, prior], dim=1), cat_tg, w["pm0_w"], w["pm0_b"], allow_same=True)
    m = masked_conv(F.leaky_relu(m), cat_tg, w["pm2_w"], w["pm2_b"], allow_same=True)
    m = masked_conv(F.leaky_relu(m), cat_tg, w["pm4_w"], w["pm4_b"], allow_same=True)
    B, out = m.shape[0], ctx.shape[1]
    return m.reshape(B, 2 * G, out // G, *m.shape[2:])[:, :G].reshape(B, out, *m.shape[2:])


def weights_from_state_dict(sd, prefix="topo_group_context_model."):
    g = lambda k: sd[prefix + k].detach().float().cpu()
    if prefix + "param_merger.0.weight" in sd:   # the internal variant
        return {"ctx_w": g("context_prediction.weight"), "ctx_b": g("context_prediction.bias"),
                "pm0_w": g("param_merger.0.weight"), "pm0_b": g("param_merger.0.bias"),
                "pm2_w": g("param_merger.2.weight"), "pm2_b": g("param_merger.2.bias"),
                "pm4_w": g("param_merger.4.weight"), "pm4_b": g("param_merger.4.bias")}
    return {"ctx_w": g("context_prediction.weight"), "ctx_b": g("context_prediction.bias"),
            "m1_w": g("param_merger_in.weight"), "m1_b": g("param_merger_in.bias"),
            "m2_w": g("param_merger_out.1.weight"), "m2_b": g("param_merger_out.1.bias"),
            "m3_w": g("param_merger_out.3.weight"), "m3_b": g("param_merger_out.3.bias")}


def random_weights(C, seed, scale=1.0):
    """nn.Conv2d default init (kaiming-uniform a=sqrt(5)) of the reference context model, seeded."""
    gen = torch.Generator().manual_seed(seed)

    def conv(cout, cin, k):
        bound = 1.0 / math.sqrt(cin * k * k)
        wt = (torch.rand(cout, cin, k, k, generator=gen) * 2 - 1) * bound * scale
        b = (torch.rand(cout, generator=gen) * 2 - 1) * bound
        return wt, b
    o = 2 * C
    w = {}
    w["ctx_w"], w["ctx_b"] = conv(o, C, 5)
    w["m1_w"], w["m1_b"] = conv(o * 5 // 3, 2 * o, 1)
    w["m2_w"], w["m2_b"] = conv(o * 4 // 3, o * 5 // 3, 1)
    w["m3_w"], w["m3_b"] = conv(o, o * 4 // 3, 1)
    return w


def params_for(buf, tg, prior, w):
    """Distribution parameters of one coding step: context model, or prior alone when there is none
    (pgm_coder.py:1606-1638 use_param_merger=False with map "none": ctx == bias == 0 at init; cfg 1)."""
    if w is None:
        return prior
    if "pm0_w" in w:
        return internal_merger(buf, tg, prior, w)
    if "m1_w" not in w:   # internal variant, use_param_merger=False: params = ctx + prior (pgm_coder.py:1634-1635)
        return masked_conv(buf, tg, w["ctx_w"], w["ctx_b"], allow_same=False) + prior
    return context_model(buf, tg, prior, w)


# ---------------------------------------------------------------- CompressAI-style serial coder
def joint_ar_weights_from_state_dict(sd, prefix=""):
    g = lambda k: sd[prefix + k].detach().float().cpu()
    return {"ctx_w": g("context_prediction.weight"), "ctx_b": g("context_prediction.bias"),
            "e0_w": g("entropy_parameters.0.weight"), "e0_b": g("entropy_parameters.0.bias"),
            "e2_w": g("entropy_parameters.2.weight"), "e2_b": g("entropy_parameters.2.bias"),
            "e4_w": g("entropy_parameters.4.weight"), "e4_b": g("entropy_parameters.4.bias")}


def _joint_ar_params(w, y_crop, p, conv_w):
    """pgm_coder.py:1996-2008 / :2049-2059: 5x5 convolution on the crop, then the 1x1 entropy_parameters network on
    cat(prior, ctx); scales = first half of the channels, means = second (mean_scale_split_method "chunk", inverse)."""
    ctx_p = F.conv2d(y_crop, conv_w, bias=w["ctx_b"])
    x = torch.cat((p, ctx_p), dim=1)
    x = F.leaky_relu(F.conv2d(x, w["e0_w"], w["e0_b"]))
    x = F.leaky_relu(F.conv2d(x, w["e2_w"], w["e2_b"]))
    params = F.conv2d(x, w["e4_w"], w["e4_b"]).squeeze(-1).squeeze(-1)
    scales, means = params.chunk(2, 1)
    return means, scales


def joint_ar_encode_symbols(y, prior, w, scale_table, k=5):
    """use_joint_ar_model_impl, _encode_with_pgm (pgm_coder.py:1975-2027): pixel by pixel in raster order with the
    causally masked kernel; stream order = pixel-major, then (b, c)."""
    B, C, H, W = y.shape
    pad = k // 2
    y_hat = F.pad(y, (pad, pad, pad, pad))
    mask = torch.ones_like(w["ctx_w"])
    mask[:, :, k // 2, k // 2:] = 0
    mask[:, :, k // 2 + 1:] = 0
    mw = w["ctx_w"] * mask
    data, idx = [], []
    for h in range(H):
        for x in range(W):
            crop = y_hat[:, :, h:h + k, x:x + k]
            means, scales = _joint_ar_params(w, crop, prior[:, :, h:h + 1, x:x + 1], mw)
            sym = torch.round(crop[:, :, pad, pad] - means)
            y_hat[:, :, h + pad, x + pad] = sym + means
            data.append(sym)
            idx.append(select_indexes(scales, scale_table))
    return (torch.cat(data).numpy().astype(np.int32).reshape(-1), torch.cat(idx).numpy().astype(np.int32).reshape(-1),
            y_hat[:, :, pad:pad + H, pad:pad + W].clone())


def joint_ar_decode(decode_group, prior, w, scale_table, C, k=5):
    """_pgm_generate (pgm_coder.py:2031-2066): the unmasked kernel on a buffer whose future is still zero."""
    B, _, H, W = prior.shape
    pad = k // 2
    y_hat = torch.zeros(B, C, H + 2 * pad, W + 2 * pad)
    for h in range(H):
        for x in range(W):
            means, scales = _joint_ar_params(w, y_hat[:, :, h:h + k, x:x + k], prior[:, :, h:h + 1, x:x + 1], w["ctx_w"])
            sym = decode_group(select_indexes(scales, scale_table).numpy().astype(np.int32))
            y_hat[:, :, h + pad, x + pad] = torch.from_numpy(np.asarray(sym)).float().reshape(B, C) + means
    return y_hat[:, :, pad:pad + H, pad:pad + W].clone()


# ------------------------------------------------------------------------------------------- the path
def encode_symbols(y, prior, tg, w, scale_table):
    """pgm_coder.py:912-947 _encode_with_pgm: returns (symbols int32 [N], indexes int32 [N], y_hat) in the
    reference's stream order (group-major, then boolean-mask order b,c,h,w)."""
    B, C, H, W = y.shape
    gmap = group_of_elements(tg, B, C)
    buf = torch.zeros(B, C, H, W)
    syms, idxs = [], []
    for g in range(int(tg.max()) + 1):
        m = gmap == g
        params = params_for(buf, tg, prior, w)
        mean, scale = split_mean_scale(params)
        idx = select_indexes(scale, scale_table)[m]
        mu = mean[m]
        s = torch.round(y[m] - mu)
        buf[m] = s + mu
        syms.append(s)
        idxs.append(idx)
    return (torch.cat(syms).numpy().astype(np.int32), torch.cat(idxs).numpy().astype(np.int32), buf)


def decode_symbols(decode_group, prior, tg, w, scale_table, C):
    """pgm_coder.py:949-981 _pgm_generate; decode_group(indexes int32) -> symbols int32 consumes the stream."""
    B, _, H, W = prior.shape
    gmap = group_of_elements(tg, B, C)
    buf = torch.zeros(B, C, H, W)
    for g in range(int(tg.max()) + 1):
        m = gmap == g
        params = params_for(buf, tg, prior, w)
        mean, scale = split_mean_scale(params)
        idx = select_indexes(scale, scale_table)[m].contiguous().numpy().astype(np.int32)
        sym = decode_group(idx)
        buf[m] = torch.as_tensor(sym).float() + mean[m]
    return buf * 1.0 + 0.0      # torch_ans.py:161-180 _data_postprocess, "uniform" quantiser [0, 128, 1]


class YPathOracle:
    """encode()/decode()/update_state() of the reference y coder on the CPU, lanes = 1 stream."""

    def __init__(self, C, G=1, weights=None, freq_precision=16, bypass_precision=4):
        self.C, self.G, self.w = C, G, weights
        self.scale_table = get_scale_table()
        self.freq_precision, self.bypass_precision = freq_precision, bypass_precision

    def update_state(self):
        """torch_ans.py:237-251."""
        freqs, nsym, offs = gaussian_ans_params(self.scale_table, self.freq_precision)
        self.enc = ans_oracle.Rans64Encoder(self.freq_precision, True, self.bypass_precision)
        self.dec = ans_oracle.Rans64Decoder(self.freq_precision, True, self.bypass_precision)
        self.enc.init_params(freqs, nsym, offs)
        self.dec.init_params(freqs, nsym, offs)
        self.ans_params = (freqs, nsym, offs)

    def encode(self, y, prior, tg):
        sym, idx, _ = encode_symbols(y, prior, tg, self.w, self.scale_table)
        return self.enc.encode_with_indexes(sym, idx)

    def decode(self, data, prior, tg):
        self.dec.set_stream(data)
        return decode_symbols(self.dec.decode_stream, prior, tg, self.w, self.scale_table, self.C)


class JointAROracle(YPathOracle):
    """The CompressAI-style serial coder (use_joint_ar_model_impl=True): same tables and coder, the
    pixel-by-pixel loops of pgm_coder.py:1975-2066."""

    def encode(self, y, prior, tg=None):
        sym, idx, _ = joint_ar_encode_symbols(y, prior, self.w, self.scale_table)
        return self.enc.encode_with_indexes(sym, idx)

    def decode(self, data, prior, tg=None):
        self.dec.set_stream(data)
        return joint_ar_decode(self.dec.decode_stream, prior, self.w, self.scale_table, self.C)

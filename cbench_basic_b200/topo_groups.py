"""Host-side construction of the topological group map (reference: pgm_coder.py:1416-1491 _get_default_pgm,
:1299-1414 _preprocess_pgm).  The map is G*H*W small integers; it is built on the host (torch CPU ops) and
handed to the CUDA library, which derives the per-stage cell / position lists from it."""
import torch
import torch.nn.functional as F

# number of stages implied by the default methods (pgm_coder.py:1129-1173)
METHODS = ("none", "scanline", "zigzag", "checkerboard", "half-checkerboard", "halfinv-checkerboard",
           "quarter-checkerboard", "interlace-checkerboard", "raster2x2", "channelwise", "channelwise-checkerboard",
           "channelwise-scanline", "channelwise-g10", "elic")


def default_map(method, G, H, W):
    """int64 tensor (1, G, H, W) of group ids."""
    tg = torch.zeros(1, G, H, W, dtype=torch.long)
    if method == "none":
        return tg
    if method == "scanline":
        return torch.arange(H * W).reshape(1, 1, H, W).repeat(1, G, 1, 1)
    if method == "zigzag":
        z = torch.arange(H).reshape(H, 1) + torch.arange(W).reshape(1, W)
        return z.reshape(1, 1, H, W).repeat(1, G, 1, 1)
    if method == "checkerboard":
        tg[..., 0::2, 1::2] = 1
        tg[..., 1::2, 0::2] = 1
    elif method == "half-checkerboard":
        tg.fill_(1)
        tg[..., 1::2, 1::2] = 0
    elif method == "halfinv-checkerboard":
        tg[..., 1::2, 1::2] = 1
    elif method == "quarter-checkerboard":
        tg.fill_(1)
        tg[..., 1::4, 3::4] = 0
        tg[..., 3::4, 1::4] = 0
    elif method == "interlace-checkerboard":
        for i in range(G):
            if i % 2 == 0:
                tg[..., i, 0::2, 0::2] = 1
                tg[..., i, 1::2, 1::2] = 1
            else:
                tg[..., i, 0::2, 1::2] = 1
                tg[..., i, 1::2, 0::2] = 1
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
    elif method == "channelwise-scanline":
        for i in range(G):
            tg[:, i] = torch.arange(H * W).reshape(1, H, W) + i * H * W
    elif method == "channelwise-g10":
        splits, at = [1] * 9 + [G - 9], 0
        for i, n in enumerate(splits):
            tg[:, at:at + n] = i
            at += n
    elif method == "elic":
        splits, at = [1, 1, 2, 4, G - 8], 0
        for i, n in enumerate(splits):
            tg[:, at:at + n] = i * 2
            tg[:, at:at + n, 1::2, 0::2] = i * 2 + 1
            tg[:, at:at + n, 0::2, 1::2] = i * 2 + 1
            at += n
    else:
        raise NotImplementedError(f"Unknown default_topo_group_method {method}")
    return tg


def tile_map(pgm, G, H, W):
    """Explicit map (ints, or logits (1, G*S, h, w) resolved with argmax as in fast_mode) trimmed and tiled to
    H x W.  Only whole patches are laid down: with H or W not a multiple of the patch, the leftover rows /
    columns stay group 0 (F.fold's zero fill), exactly like the reference."""
    pgm = pgm.detach().cpu()
    if torch.is_floating_point(pgm):
        if pgm.ndim != 4 or pgm.shape[1] % G:
            raise ValueError("logits must have shape (1, channel_groups * num_groups, h, w)")
        pgm = pgm.reshape(pgm.shape[0], G, pgm.shape[1] // G, *pgm.shape[2:]).movedim(2, -1).argmax(-1)
    if pgm.ndim != 4 or pgm.shape[1] != G:
        raise ValueError("pgm must have shape (1, channel_groups, h, w)")
    tg = pgm[:, :, :H, :W].long()
    ph, pw = tg.shape[2:4]
    if ph < H or pw < W:
        reps = (H // ph) * (W // pw)
        cols = tg.reshape(tg.shape[0], -1, 1).repeat(1, 1, reps)
        tg = F.fold(cols.float(), (H, W), (ph, pw), stride=(ph, pw)).long()
    return tg[:1]

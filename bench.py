#!/usr/bin/env python
"""bench.py -- entropy encode+decode throughput of the BaSIC y-node hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg1|cfg3|cfg4|cfg5|cfg5t] [--rates synthetic|trained] [--lanes L]

One "step" = one full pass of the hot path over one batch of synthetic latents: encode(y, prior) -> bytes, then
decode(bytes, prior) -> y_hat (context model + scale index + quantisation + multi-lane rANS, both directions).
Metric (BASELINE.json): image Mpixel/s, pixels = 256 x latent positions x batch; whole job over all N GPUs
(weak scaling: every rank codes its own block of images, no data-path collective; one all_gather of stream sizes;
cfg5t is ONE 4K image cut into row-band tiles over the N GPUs = strong scaling).
Prints ONE JSON line on rank 0.  `--impl reference` times the reference's CPU path on the SAME batch and step count.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (B per GPU, C, H, W, group-map method, context model?, description)
    "cfg2": (24, 192, 32, 48, "checkerboard", True,
             "configs[1]: joint AR hyperprior, checkerboard 2-group context model, 24 Kodak-shape (768x512) images per GPU"),
    "cfg1": (1, 192, 32, 48, "none", False, "configs[0]: mean-scale hyperprior, one Kodak-shape image"),
    "cfg3": (1, 192, 32, 48, "scanline", True,
             "configs[2]: BaSIC fixed-AR preset (scanline map, 1536 coding groups), one Kodak-shape image; the slimmable "
             "levels only change backbone widths, the coder is the same at every level"),
    "cfg4": (64, 192, 16, 16, "combined", True,
             "configs[3]: BaSIC dynamic entropy coder = CombinedNNTrainablePGMPriorCoder of 5 sub-coders (scanline; learned "
             "2x2xG maps with 8 / 6 stages at G = 4, 4 stages at G = 1, 2 stages at G = 2), 64 crops of 256x256 per GPU, every "
             "level coded in turn"),
    "cfg5": (1, 320, 135, 240, "none", False, "configs[4]: 4K 3840x2160, 320-channel latent, mean-scale coder, one image per GPU"),
    "cfg5t": (1, 320, 135, 240, "none", False,
              "configs[4] tile-partitioned: ONE 4K 3840x2160 image, 320-channel latent, cut into row-band tiles over the N GPUs "
              "(strong scaling), mean-scale coder"),
}
# configs[3] sub-coders: (name, channel groups G, stages S); learned-style maps = seeded ints in [0, S) of shape (1, G, 2, 2)
CFG4_LEVELS = [("scanline", 1, 256), ("8-stage", 4, 8), ("6-stage", 4, 6), ("4-stage", 1, 4), ("2-stage", 2, 2)]
METRIC = "entropy encode+decode Mpixel/s (Kodak 768x512 shape)"


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        # NVML in-process (a fork + exec of nvidia-smi every 200 ms from a process holding pinned buffers perturbs the
        # host side of the timed region, more so with one sampler per rank); nvidia-smi only if NVML is not importable
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[n] else "Not Active"
                                                      for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                self.stop.wait(0.01)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def learned_style_map(G, S, seed):
    """What _preprocess_pgm produces from a trained topo_group_predictor (pgm_coder.py:1342-1385): an int map (1, G, 2, 2)
    with values in [0, S); here a seeded assignment of the G*4 cells to stages that uses every stage."""
    rng = np.random.default_rng(seed)
    cells = G * 4
    stages = np.concatenate([np.arange(S) % S, rng.integers(0, S, max(0, cells - S))])[:cells]
    if cells < S:
        stages = np.arange(cells)
    rng.shuffle(stages)
    return torch.from_numpy(stages.reshape(1, G, 2, 2).astype(np.int64))


def make_inputs(workload, seed):
    from oracle import ypath_oracle as Y  # only for seeded random-init weights + the CPU baseline (checker side)
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    g = torch.Generator().manual_seed(seed)
    y = 3 * torch.randn(B, C_, H, W, generator=g)
    prior = torch.randn(B, 2 * C_, H, W, generator=g)
    w = Y.random_weights(C_, 1234) if ctx else None
    return y, prior, w


def _ctx_module(C_, w):
    from cbench_basic_b200.prior_coder import TopoGroupDynamicMaskConv2dContextModel as Ctx
    cm = Ctx(in_channels=C_, out_channels=2 * C_)
    cm.load_state_dict({"context_prediction.weight": w["ctx_w"], "context_prediction.bias": w["ctx_b"],
                        "param_merger_in.weight": w["m1_w"], "param_merger_in.bias": w["m1_b"],
                        "param_merger_out.1.weight": w["m2_w"], "param_merger_out.1.bias": w["m2_b"],
                        "param_merger_out.3.weight": w["m3_w"], "param_merger_out.3.bias": w["m3_b"]})
    return cm


def build_coder(workload, w, lanes, device, ctx_precision=None):
    from cbench_basic_b200.prior_coder import (CombinedNNTrainablePGMPriorCoder as Combined,
                                               GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder)
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    prec = ctx_precision or os.environ.get("BASIC_CTX_PRECISION", "auto")
    if method == "combined":   # configs[3]: ...ar_models.py:256-348; every sub-coder has its own context model
        subs = []
        for name, G, S in CFG4_LEVELS:
            kw = dict(default_topo_group_method="scanline") if name == "scanline" else dict(channel_groups=G)
            subs.append(Coder(in_channels=C_, topo_group_context_model=_ctx_module(C_, w), lanes=lanes, ctx_precision=prec, **kw))
        coder = Combined(subs)
    elif ctx:
        coder = Coder(in_channels=C_, default_topo_group_method=method, topo_group_context_model=_ctx_module(C_, w), lanes=lanes,
                      ctx_precision=prec)
    else:
        coder = Coder(in_channels=C_, default_topo_group_method=method, use_param_merger=False, lanes=lanes)
        with torch.no_grad():
            coder.context_prediction.weight.zero_()
            coder.context_prediction.bias.zero_()
    coder = coder.to(device).eval()
    coder.update_state()
    return coder


def calibrated_inputs(coder, prior_d, seed):
    """Trained-rate-like inputs: y is SAMPLED from the coder's own model, group by group (mean + sigma[idx] * N(0, 1) with the
    parameters the context model predicts from the groups already drawn), so that the model is calibrated on its input the
    way a trained one is on natural images -- 0.3 - 1 bpp instead of the 13.7 bpp of y = 3 * randn.  Teacher-forced
    parameters come from the library (basic_ctx_stage_params); everything stays on the device."""
    import ctypes as C
    from cbench_basic_b200 import _native as N
    B, C2, H, W = prior_d.shape
    Cc = C2 // 2
    dev = prior_d.device
    g = torch.Generator(device=dev).manual_seed(seed)
    tg = coder._get_pgm((B, Cc, H, W))
    coder._set_map(tg)
    G = tg.shape[1]
    gmap = tg.to(dev).unsqueeze(2).repeat(B, 1, Cc // G, 1, 1).reshape(B, Cc, H, W)
    tab = coder.scale_table.to(dev)
    mids = (tab[1:] + tab[:-1]) / 2
    buf = torch.zeros(B, Cc, H, W, device=dev)
    y = torch.zeros_like(buf)
    params = torch.zeros(B, C2, H, W, device=dev)
    for s in range(int(tg.max()) + 1):
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, s, buf.data_ptr(), prior_d.data_ptr(), B, params.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
        m = gmap == s
        mean, scale = params[:, 0::2][m], params[:, 1::2][m]
        sigma = tab[torch.bucketize(scale, mids)]
        v = mean + sigma * torch.randn(mean.shape, generator=g, device=dev)
        y[m] = v
        buf[m] = torch.round(v - mean) + mean
    return y


def flush_l2(scratch):
    scratch.add_(1.0)  # 256 MB > 126 MB L2


# ----------------------------------------------------------------------------------------------- CPU side
def _cpu_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would time the CPU legs on one
    core at N > 1)."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return n


def cpu_reference_step(workload, y, prior, w, n_images, coder_alone=True):
    """The reference path on host cores for `n_images` images of the workload: the oracle's restatement of
    _encode_with_pgm / _pgm_generate (torch CPU ops, all host threads; the reference's work pattern: the full-tensor
    context model once per group) with the UNMODIFIED reference coder (oracle/_ref) when it loads on this box, else the C
    port.  Returns timings, the coder used, and the coder alone on one thread (the reference coder has no threading)."""
    from oracle import ans_oracle, ref_loader, ypath_oracle as Y
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    n_images = min(n_images, B)
    ys, ps = y[:n_images], prior[:n_images]
    tg = Y.default_pgm(method, 1, H, W)
    tab = Y.get_scale_table()
    freqs, nsym, offs = Y.gaussian_ans_params(tab)
    R = ref_loader.load("ans")
    mod, kind = (R, "reference") if R is not None else (ans_oracle, "port")
    enc, dec = mod.Rans64Encoder(16, True, 4), mod.Rans64Decoder(16, True, 4)
    enc.init_params(freqs, nsym, offs)
    dec.init_params(freqs, nsym, offs)
    with torch.no_grad():
        t0 = time.perf_counter()
        sym, idx, _ = Y.encode_symbols(ys, ps, tg, w, tab)
        bs = enc.encode_with_indexes(sym, idx)
        t1 = time.perf_counter()
        dec.set_stream(bs)
        Y.decode_symbols(lambda i: dec.decode_stream(i), ps, tg, w, tab, C_)
        t2 = time.perf_counter()
    out = dict(t_enc=t1 - t0, t_dec=t2 - t1, kind=kind, n_images=n_images, n_symbols=int(sym.size), bytes=len(bs))
    if coder_alone:   # single thread, as the reference coder is
        t3 = time.perf_counter()
        enc.encode_with_indexes(sym, idx)
        t4 = time.perf_counter()
        dec.decode_with_indexes(bs, idx)
        t5 = time.perf_counter()
        out.update(coder_enc=t4 - t3, coder_dec=t5 - t4)
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload in ("cfg3", "cfg4"):
        print(json.dumps({"impl": "reference", "unavailable": f"{args.workload}: the reference evaluates the full context model "
                          "once per coding group (256 - 1536 groups): minutes per step on host cores; see DESIGN.md"}), flush=True)
        return
    cores = _cpu_threads()
    y, prior, w = make_inputs(args.workload, 0)
    B, C_, H, W, method, ctx, desc = WORKLOADS[args.workload]
    n_img = B                     # the SAME batch as the GPU arm, every step
    times, r = [], None
    for it in range(args.warmup + args.steps):
        r = cpu_reference_step(args.workload, y, prior, w, n_img, coder_alone=(it == 0))
        if it == 0:
            first = r
        if it >= args.warmup:
            times.append((r["t_enc"], r["t_dec"]))
    t_enc = sum(t[0] for t in times) / len(times)
    t_dec = sum(t[1] for t in times) / len(times)
    ms = 1e3 * (t_enc + t_dec)
    pix = n_img * 256 * H * W
    val = pix / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 rANS state / int32 symbols (reference CPU coder); context model f32 on torch CPU", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "images_per_gpu": B, "latent": [C_, H, W], "step": "encode + decode",
                       "sample": f"all {n_img} images of the batch, every step"},
            "encode_mpixel_s": pix / t_enc / 1e6, "decode_mpixel_s": pix / t_dec / 1e6,
            "cpu_coder_enc_mpixel_s": pix / first["coder_enc"] / 1e6, "cpu_coder_dec_mpixel_s": pix / first["coder_dec"] / 1e6,
            "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                             "kind_detail": {"python_loop": "port: oracle/ypath_oracle.py restates _encode_with_pgm / _pgm_generate "
                                                            "(the reference module itself needs /root/reference, absent on the GPU box)",
                                             "coder": first["kind"] + (": the unmodified cbench.ans compiled into oracle/_ref"
                                                                       if first["kind"] == "reference" else ": oracle/ans_oracle.c")},
                             "sample": f"{n_img} of {B} images per step, encode+decode, {cores} torch threads; coder alone on 1 thread: "
                                       f"enc {first['n_symbols'] / first['coder_enc'] / 1e6:.1f} Msym/s, "
                                       f"dec {first['n_symbols'] / first['coder_dec'] / 1e6:.1f} Msym/s"},
            "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU side
def coder_lane_sweep(coder, device_index, hbm_peak, n_sym=1 << 24):
    """The multi-lane coder alone (standalone encode_with_indexes / decode_with_indexes API, operands resident) on
    symbols at trained-model rates (idx = min(Geom(0.12) - 1, 63), sym = rint(N(0,1) * sigma[idx])), for
    a range of lane counts: achieved algorithmic GB/s ((8 + c) bytes per symbol / kernel time, CUDA events inside the
    library) against the HBM peak, and the container overhead each lane count costs -- the trade-off the 0.5 % bpp bar
    decides.  lanes = 0 is the auto mode the y path uses."""
    from cbench_basic_b200 import _native as N, ans
    try:
        dev = torch.device("cuda", device_index)
        tab = coder.scale_table.detach().cpu().numpy()
        rng = np.random.default_rng(0)
        idx = np.minimum(rng.geometric(0.12, n_sym) - 1, 63).astype(np.int32)
        sym = np.rint(rng.standard_normal(n_sym) * tab[idx]).astype(np.int32)
        freqs, nsym, offs = coder._get_ans_params()
        ts, ti = torch.from_numpy(sym).to(dev), torch.from_numpy(idx).to(dev)
        rows, base_bytes = [], None
        ref_enc = ans.Rans64Encoder(lanes=1, device=device_index)
        ref_enc.init_params(freqs, nsym, offs)
        lanes1_bytes = len(ref_enc.encode_with_indexes(ts, ti))
        for lanes in (0, 148 * 8 * 32, 148 * 16 * 32, 148 * 32 * 32, 4 * 148 * 16 * 32):
            enc = ans.Rans64Encoder(lanes=lanes, device=device_index)
            dec = ans.Rans64Decoder(lanes=lanes, device=device_index)
            for c in (enc, dec):
                c.init_params(freqs, nsym, offs)
            bs = enc.encode_with_indexes(ts, ti)
            out = dec.decode_with_indexes(bs, ti)
            assert torch.equal(torch.as_tensor(out).to(dev).reshape(-1), ts), "lane sweep: lossless round trip failed"
            N.profile(True)
            N.profile_read()
            reps = 3
            for _ in range(reps):
                bs = enc.encode_with_indexes(ts, ti)
                dec.decode_with_indexes(bs, ti)
            torch.cuda.synchronize(dev)
            ph = N.profile_read()
            N.profile(False)
            enc_ms, dec_ms = ph["coder_encode"][0] / reps, ph["coder_decode"][0] / reps
            c_b = len(bs) / n_sym
            if base_bytes is None:
                base_bytes = len(bs)
            n_chunks = int(np.frombuffer(bs[4:8], dtype=np.uint32)[0])
            rows.append({"lanes": lanes, "chunks": n_chunks, "stream_bytes": len(bs), "bytes_per_symbol": c_b,
                         "size_vs_lanes1": len(bs) / lanes1_bytes - 1.0,
                         "decode_ms": dec_ms, "encode_ms": enc_ms,
                         "decode_gbs": (8 + c_b) * n_sym / (dec_ms * 1e-3) / 1e9, "encode_gbs": (8 + c_b) * n_sym / (enc_ms * 1e-3) / 1e9,
                         "decode_frac_of_hbm": (8 + c_b) * n_sym / (dec_ms * 1e-3) / 1e9 / hbm_peak,
                         "encode_frac_of_hbm": (8 + c_b) * n_sym / (enc_ms * 1e-3) / 1e9 / hbm_peak})
        return {"n_symbols": n_sym, "data": "idx = min(Geom(0.12) - 1, 63), sym = rint(N(0,1) * sigma[idx])",
                "lanes1_bytes": lanes1_bytes,
                "bytes_per_symbol_algorithmic": "8 + c (int32 symbol + int32 index + c stream bytes)", "rows": rows}
    except Exception as e:
        return {"error": repr(e)}


def oracle_mismatches(workload, coder, y, prior, w, yd, pd):
    """Image 0 of this rank's batch against the CPU oracle (the reference's arithmetic): quantised symbols and scale indexes
    of the MODE THAT IS TIMED.  north_star: any disagreement is a failure unless it sits on a float rounding tie -- the
    counts are reported, `off_tie` must be 0."""
    import ctypes as C
    from cbench_basic_b200 import _native as N
    from oracle import ypath_oracle as Y
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    dev = yd.device
    tg = Y.default_pgm(method, 1, H, W)
    tab = Y.get_scale_table()
    with torch.no_grad():
        sym_o, idx_o, yhat_o = Y.encode_symbols(y[:1], prior[:1], tg, w, tab)
        params_o = Y.params_for(yhat_o, tg, prior[:1], w)
    bs, yhat = coder.encode(yd[:1], prior=pd[:1], return_yhat=True)
    out = coder.decode(bs, prior=pd[:1])
    assert torch.equal(out, yhat * 1.0 + 0.0)
    # our parameters, teacher-forced on our own reconstruction, cell by cell at its stage
    params = torch.zeros(1, 2 * C_, H, W, device=dev)
    if ctx:
        for g in range(int(tg.max()) + 1):
            N.check(N.lib().basic_ctx_stage_params(coder._ctx, g, out.data_ptr(), pd[:1].contiguous().data_ptr(), 1, params.data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream))
    else:
        params = pd[:1].clone()
    params = params.cpu()
    mean, scale = params[:, 0::2], params[:, 1::2]
    mean_o, scale_o = Y.split_mean_scale(params_o)
    idx_g = Y.select_indexes(scale, tab)
    idx_ref = Y.select_indexes(scale_o, tab)
    sym_g = torch.round(out.cpu() - mean)
    sym_ref = torch.round(yhat_o - mean_o)
    bad_s, bad_i = (sym_g != sym_ref), (idx_g != idx_ref)
    # a disagreement is a tie when the ORACLE's own value sits within 2e-5 (relative) of the decision boundary
    tabd = tab.double()
    lo = torch.minimum(idx_g, idx_ref)[bad_i].clamp(max=tabd.numel() - 2)
    mid = (tabd[lo] + tabd[lo + 1]) / 2
    off_i = int((((scale_o[bad_i].double() - mid).abs() / mid) > 2e-5).sum())
    frac = (y[:1] - mean_o)[bad_s].double()
    off_s = int((((frac - torch.floor(frac)) - 0.5).abs() > 2e-5 * (1 + frac.abs())).sum())
    rel = float(((params - params_o).abs() / params_o.abs().clamp_min(1.0)).max())
    return {"image": 0, "symbols": int(sym_ref.numel()), "symbol_mismatches": int(bad_s.sum()), "index_mismatches": int(bad_i.sum()),
            "off_tie": off_s + off_i, "params_max_rel_err": rel, "tolerance": 1e-5,
            "mode": f"lanes={coder.lanes}, ctx_precision={coder.ctx_precision}"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from cbench_basic_b200 import _native as N, sharding
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # before any pinned allocation or library thread exists: stay on the CPUs next to this rank's GPU
    numa = sharding.bind_to_gpu_numa(local_rank, local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world))) \
        if not os.environ.get("BASIC_NO_BIND") else {"bound": False, "disabled": True}
    B, C_, H, W, method, ctx, desc = WORKLOADS[args.workload]
    tiled = args.workload == "cfg5t"
    y, prior, w = make_inputs(args.workload, 0 if tiled else rank)
    tiles = None
    if tiled:   # ONE image: this rank's block of row bands, coded as independent images (zero context across band borders)
        n_tiles = max(8, world)
        bands = sharding.row_band_tiles(H, n_tiles)
        mine = [bands[t] for t in sharding.partition(n_tiles, world, rank)]
        tiles = [(b.start, b.stop) for b in mine]
    coder = build_coder(args.workload, w, args.lanes, dev)
    yd, pd = y.to(dev), prior.to(dev)
    if args.rates == "trained":
        if method == "combined" or not ctx:
            raise SystemExit("--rates trained needs a single coder with a context model")
        # A random-init network predicts scales around zero (everything lands in the narrowest table: 0.001 bpp when y is drawn
        # from the model, 13.7 bpp when it is not).  A trained model's scales spread over the lower third of the table: the last
        # layer's scale rows are rescaled so that the predicted scales are ~ N(0.25, 0.35^2), then y is drawn from the model.
        import ctypes as C
        params = torch.zeros(B, 2 * C_, H, W, device=dev)
        coder._set_map(coder._get_pgm((B, C_, H, W)))
        N.check(N.lib().basic_ctx_stage_params(coder._ctx, 0, torch.zeros_like(yd).data_ptr(), pd.data_ptr(), B, params.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
        raw = params[:, 1::2][:, :, 0::2, 0::2]      # (checkerboard stage 0 owns the even-even positions among others)
        gain = 0.35 / float(raw.std())
        w = dict(w)
        w["m3_w"] = w["m3_w"].clone()
        w["m3_b"] = w["m3_b"].clone()
        w["m3_w"][1::2] *= gain
        w["m3_b"][1::2] = w["m3_b"][1::2] * gain + (0.25 - gain * float(raw.mean()))
        coder = build_coder(args.workload, w, args.lanes, dev)
        yd = calibrated_inputs(coder, pd, 77 + rank)
        y = yd.cpu()
    yp, pp = y.pin_memory(), prior.pin_memory()
    scratch = torch.zeros(64 * 1024 * 1024, device=dev)
    stream = torch.cuda.current_stream(dev)
    levels = list(range(len(CFG4_LEVELS))) if method == "combined" else [None]

    def one_hot(level):
        return None if level is None else torch.eye(len(CFG4_LEVELS))[level]

    def kw(level):
        if level is None:
            return {}
        name, G, S = CFG4_LEVELS[level]
        k = {"blend_weight": one_hot(level)}
        if name != "scanline":
            k["pgm"] = learned_style_map(G, S, 100 + level)
        return k

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def units(ty, tp):
        """The coding units of a step: the whole batch, or this rank's row bands of the one image."""
        if tiles is None:
            return [(ty, tp)]
        return [(ty[:, :, a:b].contiguous(), tp[:, :, a:b].contiguous()) for a, b in tiles]

    res_units, host_units = units(yd, pd), None
    mid_events = []

    def step_resident():
        outs = []
        for level in levels:
            for uy, up in res_units:
                bs = coder.encode(uy, prior=up, **kw(level))
                if mid_events and mid_events[-1][1] is None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(stream)
                    mid_events[-1][1] = e
                outs.append((bs, coder.decode(bs, prior=up, **kw(level))))
        return outs

    def step_zero_copy():   # the opt-in hand-over: encode returns a view of its page-locked buffer, decode uploads out of it
        outs = []
        for level in levels:
            for uy, up in res_units:
                view = coder.encode(uy, prior=up, zero_copy=True, **kw(level))
                outs.append(coder.decode(view, prior=up, **kw(level)))
        return outs

    out_host = torch.empty(B, C_, H, W, dtype=torch.float32).pin_memory()
    if tiles is not None:
        host_units = [(yp[:, :, a:b].contiguous().pin_memory(), pp[:, :, a:b].contiguous().pin_memory()) for a, b in tiles]
    else:
        host_units = [(yp, pp)]

    def step_e2e():   # pinned host buffers in, host bytes / pinned host tensor out: H2D / D2H inside the timed region
        outs = []
        for level in levels:
            for k, (uy, up) in enumerate(host_units):
                bs = coder.encode(uy, prior=up, **kw(level))       # the library uploads pinned host tensors itself (prior first,
                out = coder.decode(bs, prior=up, **kw(level))      # y beside the first group's context model)
                if tiles is None:
                    out_host.copy_(out, non_blocking=False)
                else:
                    out_host[:, :, tiles[k][0]:tiles[k][1]].copy_(out, non_blocking=False)
                outs.append(bs)
        return outs

    # correctness of what is timed: lossless + matches encoder-side reconstruction (|rint(y - m) + m - y| reaches 0.5 plus a
    # float rounding of the sum: rank 2's inputs hit 0.50000012)
    stream_bytes = 0
    for level in levels:
        for uy, up in res_units:
            bs, yhat_enc = coder.encode(uy, prior=up, return_yhat=True, **kw(level))
            out = coder.decode(bs, prior=up, **kw(level))
            if not (torch.equal(out, yhat_enc * 1.0 + 0.0) and float((out - uy).abs().max()) <= 0.5 + 1e-5):
                bad = (out != yhat_enc).nonzero()
                raise AssertionError(f"round trip failed on rank {rank} (level {level}): {bad.shape[0]} elements differ from the "
                                     f"encoder's reconstruction, max|out - y| {float((out - uy).abs().max())}")
            stream_bytes += len(bs)

    def timed(fn, steps, warmup, split=False):
        for _ in range(warmup):
            flush_l2(scratch)
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush_l2(scratch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if split:
                mid_events.append([e0, None])
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        total = sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps
        enc_ms = None
        if split and mid_events:
            enc_ms = sum(e0.elapsed_time(em) for e0, em in mid_events) / len(mid_events)
            mid_events.clear()
        return total, enc_ms

    single = levels == [None] and len(res_units) == 1
    with ClockSampler(local_rank) as clk:
        N.launch_count(reset=True)
        ms, enc_ms = timed(step_resident, args.steps, args.warmup, split=single)
        launches = N.launch_count() // (args.steps + args.warmup)
    ms_e2e, _ = timed(step_e2e, max(2, args.steps // 2), 2)
    ms_zc = None
    if method != "combined":
        try:
            ms_zc, _ = timed(step_zero_copy, max(2, args.steps // 2), 2)
        except (ValueError, TypeError):
            ms_zc = None

    # sizes gather: the one collective of the path
    sizes = sharding.gather_sizes([stream_bytes], device=dev)
    t = torch.tensor([ms, ms_e2e, enc_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, enc_ms = float(t[0]), float(t[1]), float(t[2]) or None
    if rank != 0:
        return
    n_levels = len(levels)
    pix_total = (1 if tiled else world) * B * 256 * H * W * n_levels
    n_sym = B * C_ * H * W
    hbm_peak, tf_peak, peak_src = peaks()

    # --- per-phase device time, live: CUDA events recorded by the library on the launching stream around the phases
    # of the same step that was timed above (basic_profile_enable / basic_profile_read, include/basic_b200.h)
    roofline, coder_roof, phases = None, None, None
    try:
        prof_steps = 5
        N.profile(True)
        N.profile_read()
        N.launch_count(reset=True)
        for _ in range(prof_steps):
            flush_l2(scratch)
            step_resident()
        torch.cuda.synchronize(dev)
        n_launch = N.launch_count() // prof_steps
        ph = N.profile_read()
        N.profile(False)
        phases = {k: {"ms_per_step": v[0] / prof_steps, "spans_per_step": v[1] / prof_steps} for k, v in ph.items()}
        n_coded = sum(hi - lo for lo, hi in tiles) * W * C_ if tiled else n_sym * n_levels
        cbytes = stream_bytes / max(n_coded, 1)
        dec_ms, enc_ph_ms = phases["coder_decode"]["ms_per_step"], phases["coder_encode"]["ms_per_step"]
        if dec_ms > 0:
            ach = (8 + cbytes) * n_coded / (dec_ms * 1e-3) / 1e9  # 4 B index + 4 B symbol + c stream bytes per symbol
            coder_roof = {"kernel": "k_bls_decode (multi-lane rANS decode, one launch per coding group)", "bound": "hbm",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                          "traffic": None,
                          "decode_ms_per_step": dec_ms, "launches_per_step": phases["coder_decode"]["spans_per_step"],
                          "encode_phase_ms_per_step": enc_ph_ms, "bytes_per_symbol": 8 + cbytes,
                          "decode_msym_s": n_coded / dec_ms / 1e3, "encode_msym_s": n_coded / enc_ph_ms / 1e3 if enc_ph_ms > 0 else None,
                          "peak_source": peak_src,
                          "note": "lane count is capped by the 0.5 % bpp bar (one flush per lane): latency-bound, not "
                                  "bandwidth-bound, at this stream size (DESIGN.md section 6)"}
        if ctx and method != "combined":
            ctx_ms = phases["context_model"]["ms_per_step"]            # both passes (encoder + decoder side)
            flops = 2 * 2 * 77.56 * C_ * C_ * B * H * W                # dense-equivalent, each position once, x 2 passes
            ach = flops / (ctx_ms * 1e-3) / 1e12
            mode = coder.ctx_precision if coder.ctx_precision != "auto" else ("fp32" if coder.lanes == 1 else "fp16x3")
            kname = {"fp16x3": "k_layer_tc<1> (context conv + 1x1 merger; tcgen05 kind::f16, 3 MMAs per product = error-compensated FP16, "
                               "operands pre-scaled by powers of two, FP32 accumulate)",
                     "tf32x3": "k_layer_tc<0> (context conv + 1x1 merger; tcgen05 kind::tf32, 3 MMAs per product = error-compensated TF32)",
                     "fp32": "k_layer (context conv + 1x1 merger, FP32 SIMT exact path)"}[mode]
            ceil = {"fp16x3": "3 FP16 MMAs per product: ceiling = bf16/fp16 peak / 3 (frac 0.333)",
                    "tf32x3": "3 TF32 MMAs per product: ceiling = tf32 peak / 3 = bf16 peak / 6 (frac 0.167)",
                    "fp32": "FP32 FMA pipe, not the tensor pipe"}[mode]
            roofline = {"kernel": kname, "bound": "tensor", "achieved": ach,
                        "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                        "ms_per_step": ctx_ms, "share_of_step": ctx_ms / ms, "peak_source": peak_src,
                        "note": "algorithmic FLOPs = 2*77.56*C^2 per latent position and pass (dense-equivalent, each position once; "
                                "masked taps are skipped, so executed FLOPs are lower: 0.69 of dense for this checkerboard). The 1e-5 "
                                "parity bar needs ~22 significant bits per product: " + ceil}
        else:
            roofline = coder_roof
    except Exception as e:  # never lose the headline line to a side measurement
        roofline = roofline or {"error": repr(e)}
        n_launch = None

    extra = {}
    simple = single and not tiled
    # --- Delta bpp against the lanes=1 reference stream on the same symbols
    dbpp = None
    if simple:
        try:
            ref_coder = build_coder(args.workload, w, 1, dev)
            nb = min(B, 2)
            b1 = ref_coder.encode(yd[:nb], prior=pd[:nb])
            b0 = coder.encode(yd[:nb], prior=pd[:nb])
            dbpp = {"images": nb, "lanes1_bytes": len(b1), "multilane_bytes": len(b0), "delta_frac": len(b0) / len(b1) - 1.0,
                    "full_batch_bytes": stream_bytes, "full_batch_bpp": stream_bytes * 8 / (B * 256 * H * W)}
            # parity mode: the byte-exact configuration (lanes = 1, exact FP32 context model), same batch
            t0 = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ref_coder.decode(ref_coder.encode(yd, prior=pd), prior=pd)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                t0.append(e0.elapsed_time(e1))
            extra["parity_mode"] = {"mode": "lanes=1 (reference bitstream byte for byte), ctx_precision=fp32", "ms_per_step": min(t0[1:]),
                                    "mpixel_s": B * 256 * H * W / (min(t0[1:]) * 1e-3) / 1e6}
            del ref_coder
        except Exception as e:
            dbpp = {"error": repr(e)}
        if C_ * H * W <= 400_000:   # (the oracle's argmin materialises 64 floats per symbol)
            try:
                extra["oracle_check"] = oracle_mismatches(args.workload, coder, y, prior, w, yd, pd)
            except Exception as e:
                extra["oracle_check"] = {"error": repr(e)}
    if tiled:   # what tiling costs in size: the same image coded as one unit
        try:
            whole = coder.encode(yd, prior=pd)
            extra["tiling"] = {"tiles_total": max(8, world), "tiles_this_rank": len(tiles), "untiled_bytes": len(whole),
                               "note": "mean-scale coder: symbols and indexes are identical to the untiled run, only the per-stream "
                                       "flush differs; delta vs untiled is computed over all ranks' streams"}
            tot = sum(s[0] for s in sizes)
            extra["tiling"]["tiled_bytes_all_ranks"] = tot
            extra["tiling"]["delta_frac"] = tot / len(whole) - 1.0
        except Exception as e:
            extra["tiling"] = {"error": repr(e)}

    sweep = coder_lane_sweep(coder if method != "combined" else coder.coders[1], local_rank, hbm_peak) if args.workload == "cfg2" else None

    cpu = None
    if args.workload not in ("cfg3", "cfg4"):
        try:
            cores = _cpu_threads()
            n_cpu = min(B, 4)
            r = cpu_reference_step(args.workload, y, prior, w, n_cpu)
            pix_cpu = n_cpu * 256 * H * W
            v = pix_cpu / (r["t_enc"] + r["t_dec"]) / 1e6
            cpu = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                   "kind_detail": {"python_loop": "port (oracle/ypath_oracle.py)", "coder": r["kind"]},
                   "sample": f"{n_cpu} of {B} images, encode+decode once; torch-CPU restatement of the y path with the {r['kind']} "
                             f"rANS coder"}
            extra["cpu_coder_enc_mpixel_s"] = pix_cpu / r["coder_enc"] / 1e6       # 1 core: the reference coder has no threading
            extra["cpu_coder_dec_mpixel_s"] = pix_cpu / r["coder_dec"] / 1e6
            extra["cpu_coder_cores"] = 1
        except Exception as e:
            cpu = {"error": repr(e)}

    if enc_ms:
        extra["encode_mpixel_s"] = pix_total / (enc_ms * 1e-3) / 1e6
        extra["decode_mpixel_s"] = pix_total / ((ms - enc_ms) * 1e-3) / 1e6
        extra["encode_ms"], extra["decode_ms"] = enc_ms, ms - enc_ms
        if extra.get("cpu_coder_dec_mpixel_s"):
            extra["x_cpu_coder_decode"] = extra["decode_mpixel_s"] / extra["cpu_coder_dec_mpixel_s"]   # target: >= 100 at 8 GPUs

    n_elem = sum(u[0].numel() for u in res_units) * n_levels
    h2d = n_elem * 4 * 4 + stream_bytes            # y once, the prior (2x) for encode and for decode, the stream for decode
    d2h = stream_bytes + n_elem * 4
    line = {"metric": METRIC, "value": pix_total / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if tiled else "weak",
            "vs_baseline": None,
            "dtype": "u32 rANS states / int32 symbols; context model f32 in / f32 accumulate, products as 3xFP16 on tcgen05", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "images_per_gpu": B, "latent": [C_, H, W], "lanes": args.lanes,
                       "rates": args.rates, "l2": "256 MB buffer rewritten between timed iterations", "step": "encode + decode",
                       "value_definition": "device-resident y / prior tensors in, host `bytes` out of encode and into decode (the "
                                           "reference API), y_hat left on the device"},
            "e2e": {"value": pix_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "zero_copy": None if ms_zc is None else {
                "ms_per_step": ms_zc, "value": pix_total / (ms_zc * 1e-3) / 1e6, "unit": "Mpixel/s", "rank": 0,
                "what": "same step through the opt-in hand-over: encode(zero_copy=True) returns a memoryview of the coder's page-locked "
                        "buffer instead of a bytes object, decode() uploads straight out of it (no host copy of the stream in either call)"},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "roofline_coder": coder_roof, "phases": phases,
            "coder_lane_sweep": sweep, "cpu_baseline": cpu, "delta_bpp": dbpp, "stream_bytes_per_rank": [s[0] for s in sizes],
            "bpp": sum(s[0] for s in sizes) * 8 / pix_total,
            "cpu_binding_rank0": numa}
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--rates", default="synthetic", choices=["synthetic", "trained"])
    ap.add_argument("--lanes", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)   # rank 0 alone; same batch, same step count as the GPU arm
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

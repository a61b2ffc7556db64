#!/usr/bin/env python
"""bench.py -- entropy encode+decode throughput of the BaSIC y-node hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4|cfg5]

One "step" = one full pass of the hot path over one batch of synthetic latents: encode(y, prior) -> bytes, then
decode(bytes, prior) -> y_hat (context model + scale index + quantisation + multi-lane rANS, both directions).
Metric (BASELINE.json): image Mpixel/s, pixels = 256 x latent positions x batch; whole job over all N GPUs
(weak scaling: every rank codes its own block of images, no data-path collective; one all_gather of stream sizes).
Prints ONE JSON line on rank 0.
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
    "cfg4": (64, 192, 16, 16, "raster2x2", True, "configs[3]-like: 64 crops of 256x256 per GPU, 4-stage 2x2 map"),
    "cfg5": (1, 320, 135, 240, "none", False, "configs[4]: 4K 3840x2160, 320-channel latent, mean-scale coder"),
}
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


def make_inputs(workload, seed):
    from oracle import ypath_oracle as Y  # only for seeded random-init weights + the CPU baseline (checker side)
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    g = torch.Generator().manual_seed(seed)
    y = 3 * torch.randn(B, C_, H, W, generator=g)
    prior = torch.randn(B, 2 * C_, H, W, generator=g)
    w = Y.random_weights(C_, 1234) if ctx else None
    return y, prior, w


def build_coder(workload, w, lanes, device):
    from cbench_basic_b200.prior_coder import (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder as Coder,
                                               TopoGroupDynamicMaskConv2dContextModel as Ctx)
    B, C_, H, W, method, ctx, _ = WORKLOADS[workload]
    if ctx:
        cm = Ctx(in_channels=C_, out_channels=2 * C_)
        cm.load_state_dict({"context_prediction.weight": w["ctx_w"], "context_prediction.bias": w["ctx_b"],
                            "param_merger_in.weight": w["m1_w"], "param_merger_in.bias": w["m1_b"],
                            "param_merger_out.1.weight": w["m2_w"], "param_merger_out.1.bias": w["m2_b"],
                            "param_merger_out.3.weight": w["m3_w"], "param_merger_out.3.bias": w["m3_b"]})
        coder = Coder(in_channels=C_, default_topo_group_method=method, topo_group_context_model=cm, lanes=lanes,
                      ctx_precision=os.environ.get("BASIC_CTX_PRECISION", "auto"))
    else:
        coder = Coder(in_channels=C_, default_topo_group_method=method, use_param_merger=False, lanes=lanes)
        with torch.no_grad():
            coder.context_prediction.weight.zero_()
            coder.context_prediction.bias.zero_()
    coder = coder.to(device).eval()
    coder.update_state()
    return coder


def flush_l2(scratch):
    scratch.add_(1.0)  # 256 MB > 126 MB L2


# ----------------------------------------------------------------------------------------------- CPU side
def cpu_reference_step(workload, y, prior, w, n_images):
    """The reference path on host cores for `n_images` images of the workload: the oracle's restatement of
    _encode_with_pgm / _pgm_generate (torch CPU ops, all host threads) with the UNMODIFIED reference coder
    (oracle/_ref) when it loads on this box, else the C port.  Returns (seconds encode, seconds decode, kind)."""
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
    # coder alone (single thread, as the reference coder is)
    t3 = time.perf_counter()
    enc.encode_with_indexes(sym, idx)
    t4 = time.perf_counter()
    dec.decode_with_indexes(bs, idx)
    t5 = time.perf_counter()
    return dict(t_enc=t1 - t0, t_dec=t2 - t1, kind=kind, n_images=n_images, coder_enc=t4 - t3, coder_dec=t5 - t4,
                n_symbols=int(sym.size), bytes=len(bs))


def run_reference(args, rank, world):
    if rank != 0:
        return
    y, prior, w = make_inputs(args.workload, 0)
    B, C_, H, W, method, ctx, desc = WORKLOADS[args.workload]
    n_img = 1
    times = []
    for it in range(args.warmup + args.steps):
        r = cpu_reference_step(args.workload, y, prior, w, n_img)
        if it >= args.warmup:
            times.append(r["t_enc"] + r["t_dec"])
    ms = 1e3 * sum(times) / len(times)
    pix = n_img * 256 * H * W
    val = pix / (ms * 1e-3) / 1e6
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 rANS state / int32 symbols (reference CPU coder); context model f32 on torch CPU", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "sample": f"{n_img} image per step"},
            "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                             "sample": f"{n_img} of {B} images per step, encode+decode, torch CPU context model + reference coder "
                                       f"({r['kind']}); coder alone single-thread: enc {r['n_symbols'] / r['coder_enc'] / 1e6:.1f} "
                                       f"Msym/s, dec {r['n_symbols'] / r['coder_dec'] / 1e6:.1f} Msym/s"},
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
        for lanes in (0, 148 * 8 * 32, 148 * 16 * 32, 4 * 148 * 16 * 32):
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
                         "flush_overhead_frac": n_chunks * 132 / len(bs),
                         "decode_ms": dec_ms, "encode_ms": enc_ms,
                         "decode_gbs": (8 + c_b) * n_sym / (dec_ms * 1e-3) / 1e9, "encode_gbs": (8 + c_b) * n_sym / (enc_ms * 1e-3) / 1e9,
                         "decode_frac_of_hbm": (8 + c_b) * n_sym / (dec_ms * 1e-3) / 1e9 / hbm_peak,
                         "encode_frac_of_hbm": (8 + c_b) * n_sym / (enc_ms * 1e-3) / 1e9 / hbm_peak})
        return {"n_symbols": n_sym, "data": "idx = min(Geom(0.12) - 1, 63), sym = rint(N(0,1) * sigma[idx])",
                "bytes_per_symbol_algorithmic": "8 + c (int32 symbol + int32 index + c stream bytes)", "rows": rows}
    except Exception as e:
        return {"error": repr(e)}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from cbench_basic_b200 import _native as N, sharding
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # before any pinned allocation or library thread exists: stay on the CPUs next to this rank's GPU
    numa = sharding.bind_to_gpu_numa(local_rank, local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world))) \
        if not os.environ.get("BASIC_NO_BIND") else {"bound": False, "disabled": True}
    B, C_, H, W, method, ctx, desc = WORKLOADS[args.workload]
    y, prior, w = make_inputs(args.workload, rank)
    coder = build_coder(args.workload, w, args.lanes, dev)
    yd, pd = y.to(dev), prior.to(dev)
    yp, pp = y.pin_memory(), prior.pin_memory()
    scratch = torch.zeros(64 * 1024 * 1024, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        bs = coder.encode(yd, prior=pd)
        out = coder.decode(bs, prior=pd)
        return bs, out

    out_host = torch.empty(B, C_, H, W, dtype=torch.float32).pin_memory()

    def step_e2e():   # pinned host buffers in, host bytes / pinned host tensor out: H2D / D2H inside the timed region
        bs = coder.encode(yp, prior=pp)        # pinned host tensors: the library uploads them (prior first, y beside the
        out = coder.decode(bs, prior=pp)       # first group's context model; the stream's staging beside the prior)
        out_host.copy_(out, non_blocking=False)
        return bs, out_host

    # correctness of what is timed: lossless + matches encoder-side reconstruction (|rint(y - m) + m - y| reaches 0.5 plus a
    # float rounding of the sum: rank 2's inputs hit 0.50000012)
    bs, yhat_enc = coder.encode(yd, prior=pd, return_yhat=True)
    out = coder.decode(bs, prior=pd)
    if not (torch.equal(out, yhat_enc * 1.0 + 0.0) and float((out - yd).abs().max()) <= 0.5 + 1e-5):
        bad = (out != yhat_enc).nonzero()
        raise AssertionError(f"round trip failed on rank {rank}: {bad.shape[0]} elements differ from the encoder's reconstruction "
                             f"(first {bad[0].tolist() if bad.shape[0] else None}, images {sorted(set(bad[:, 0].tolist()))[:8]}), "
                             f"max|out - y| {float((out - yd).abs().max())}, max|yhat_enc - y| {float((yhat_enc - yd).abs().max())}")
    stream_bytes = len(bs)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            flush_l2(scratch)
            fn()
        barrier()
        total_ms, evs = 0.0, []
        for _ in range(steps):
            flush_l2(scratch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        for e0, e1 in evs:
            total_ms += e0.elapsed_time(e1)
        return total_ms / steps

    with ClockSampler(local_rank) as clk:
        N.launch_count(reset=True)
        ms = timed(step_resident, args.steps, args.warmup)
        launches = N.launch_count() // (args.steps + args.warmup)
    ms_e2e = timed(step_e2e, max(2, args.steps // 2), 2)

    # sizes gather: the one collective of the path
    sizes = sharding.gather_sizes([stream_bytes], device=dev)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    pix_total = world * B * 256 * H * W
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
        cbytes = stream_bytes / n_sym
        dec_ms, enc_ms = phases["coder_decode"]["ms_per_step"], phases["coder_encode"]["ms_per_step"]
        if dec_ms > 0:
            ach = (8 + cbytes) * n_sym / (dec_ms * 1e-3) / 1e9  # 4 B index + 4 B symbol + c stream bytes per symbol
            coder_roof = {"kernel": "k_bls_decode (multi-lane rANS decode, one launch per coding group)", "bound": "hbm",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                          # dram__bytes of the two launches of a step (profiles/r1e_ncu_full_k_bls_coder.csv): indexes + stream
                          # are read once; the decoded symbols (28 MB) stay in L2 for the dequantiser
                          "traffic": 45.6e6 if (args.workload == "cfg2" and args.lanes == 0) else None,
                          "decode_ms_per_step": dec_ms, "launches_per_step": phases["coder_decode"]["spans_per_step"],
                          "encode_phase_ms_per_step": enc_ms, "bytes_per_symbol": 8 + cbytes,
                          "decode_msym_s": n_sym / dec_ms / 1e3, "encode_msym_s": n_sym / enc_ms / 1e3 if enc_ms > 0 else None,
                          "peak_source": peak_src,
                          "note": "lane count is capped by the 0.5 % bpp bar (one 132 B flush per 32 lanes): "
                                  "latency-bound, not bandwidth-bound, at this stream size (DESIGN.md section 6)"}
        if ctx:
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
            # dram__bytes_read + write of the 8 launches of one pass (profiles/r1e_ncu_full_k_layer_tc_fp16x3.csv, cold
            # caches under ncu) x 2 passes per step, for the cfg2 geometry only; algorithmic bytes = activations once
            traffic = 2 * 342.7e6 if (args.workload == "cfg2" and mode == "fp16x3") else None
            roofline = {"kernel": kname, "bound": "tensor", "achieved": ach,
                        "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": traffic,
                        "traffic_note": "DRAM bytes per step (16 launches), ncu --set full, cold L2; algorithmic = 2 x 0.59 GB "
                                        "(every activation read and written once per pass; most of the writes and "
                                        "re-reads stay in the 126 MB L2)",
                        "ms_per_step": ctx_ms, "launches_per_step": phases["context_model"]["spans_per_step"] * 4,
                        "share_of_step": ctx_ms / ms, "peak_source": peak_src,
                        "note": "algorithmic FLOPs = 2*77.56*C^2 per latent position and pass (dense-equivalent, each position once; "
                                "masked taps are skipped, so executed FLOPs are lower: 0.69 of dense for this checkerboard). The 1e-5 "
                                "parity bar needs ~22 significant bits per product: " + ceil}
        else:
            roofline = coder_roof
    except Exception as e:  # never lose the headline line to a side measurement
        roofline = roofline or {"error": repr(e)}
        n_launch = None

    # --- Delta bpp against the lanes=1 reference stream on the same symbols
    dbpp = None
    try:
        ref_coder = build_coder(args.workload, w, 1, dev)
        nb = min(B, 2)
        b1 = ref_coder.encode(yd[:nb], prior=pd[:nb])
        b0 = coder.encode(yd[:nb], prior=pd[:nb])
        dbpp = {"images": nb, "lanes1_bytes": len(b1), "multilane_bytes": len(b0), "delta_frac": len(b0) / len(b1) - 1.0,
                "full_batch_bytes": stream_bytes, "full_batch_bpp": stream_bytes * 8 / (B * 256 * H * W)}
    except Exception as e:
        dbpp = {"error": repr(e)}

    sweep = coder_lane_sweep(coder, local_rank, hbm_peak) if rank == 0 else None

    cpu = None
    if world == 1 or rank == 0:
        try:
            r = cpu_reference_step(args.workload, y, prior, w, 1)
            v = 256 * H * W / (r["t_enc"] + r["t_dec"]) / 1e6
            cpu = {"value": v, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"1 of {B} images, encode+decode; torch-CPU restatement of the y path (oracle/ypath_oracle.py) with the "
                             f"{r['kind']} rANS coder; coder alone, 1 thread: enc {r['n_symbols'] / r['coder_enc'] / 1e6:.1f} Msym/s, "
                             f"dec {r['n_symbols'] / r['coder_dec'] / 1e6:.1f} Msym/s "
                             f"(= {256 * H * W / r['coder_dec'] / 1e6:.1f} Mpixel/s decode)"}
        except Exception as e:
            cpu = {"error": repr(e)}

    h2d = (y.numel() + 2 * prior.numel()) * 4 + stream_bytes
    d2h = stream_bytes + y.numel() * 4
    line = {"metric": METRIC, "value": pix_total / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 rANS states / int32 symbols; context model f32 in / f32 accumulate, products as 3xFP16 on tcgen05", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "images_per_gpu": B, "latent": [C_, H, W], "lanes": args.lanes,
                       "l2": "256 MB buffer rewritten between timed iterations", "step": "encode + decode"},
            "e2e": {"value": pix_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "roofline_coder": coder_roof, "phases": phases,
            "coder_lane_sweep": sweep, "cpu_baseline": cpu, "delta_bpp": dbpp, "stream_bytes_per_rank": [s[0] for s in sizes],
            "cpu_binding_rank0": numa}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5
        # all the host threads the box gives this process: torchrun exports OMP_NUM_THREADS=1, which would time the
        # CPU arm on one core at N > 1 (rank 0 alone runs it, the other ranks exit)
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except Exception:
            pass
        run_reference(args, rank, world)
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

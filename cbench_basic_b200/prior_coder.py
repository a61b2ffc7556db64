"""Drop-in for the reference's y-node prior coder on the encode / decode / update_state path:

    cbench.modules.prior_model.prior_coder.pgm_coder.GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder
    cbench.nn.layers.masked_conv.TopoGroupDynamicMaskConv2dContextModel

Same class names, constructor keywords (the ones that matter for coding), state_dict keys, and
``encode(input, prior=, pgm=) -> bytes`` / ``decode(bytes, prior=, pgm=) -> Tensor`` / ``update_state()``
semantics (pgm_coder.py:544-618, :912-981, torch_ans.py:237-251).  The hot loop -- per group: context model,
scale index, quantisation, ANS coding, write-back -- runs entirely in the CUDA library
(basic_ypath_encode / basic_ypath_decode, include/basic_b200.h); Python only builds the group map and
frames the bytes.  ``forward()`` in eval mode returns what the reference returns there -- the dequantised input
(pgm_coder.py:391-398, :539), which is what ``LatentGraphicalANSEntropyCoder._generative_process`` feeds to the next edge
before it calls ``encode`` (latent_graph.py:836-841); the training likelihood stays with the reference module.  The module
protocol the reference's harness uses on a coder (cache dicts, ``profiler.start_time_profile`` scopes, ``device``) comes from
``module_api.CoderModuleBase``; ``reference_integration.bind()`` additionally derives the classes from the reference's own
``NNTrainableModule``.
"""
import ctypes as C
import math
import struct
import time

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from . import ans, topo_groups
from .module_api import CoderModuleBase

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):  # noqa: A002  (reference signature)
    """compressai_coder.py:23-30."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def joint_ar_remap(e0_weight, e4_weight, e4_bias, in_channels):
    """The serial coder's entropy_parameters network (pgm_coder.py:1204-1213, :2000-2003) in the layout of the masked-merger
    kernels: its first matrix multiplies cat(prior, ctx) -- the kernels read [ctx | prior]; its last one emits scales in the
    first C channels and means in the second ("chunk" split, inverse_mean_scale) -- the kernels emit (mean_c, scale_c) pairs.
    Returns (first weight [N, 4C] with swapped column halves, last weight and bias with interleaved rows)."""
    o, Cc = 2 * in_channels, in_channels
    w0 = e0_weight.reshape(e0_weight.shape[0], 2 * o)
    w0 = torch.cat([w0[:, o:], w0[:, :o]], dim=1)
    rows = torch.stack([torch.arange(Cc, 2 * Cc), torch.arange(0, Cc)], dim=1).reshape(-1)
    return w0, e4_weight.reshape(o, -1)[rows], e4_bias[rows]


class _CtxHandle:
    """Owns a basic_ctx* (kept out of nn.Module attribute machinery so it can be freed at interpreter exit)."""

    def __init__(self):
        self.h = None

    def reset(self):
        h, self.h = self.h, None
        if h:
            try:
                N.lib().basic_ctx_destroy(h)
            except Exception:
                pass

    __del__ = reset


class TopoGroupDynamicMaskConv2dContextModel(CoderModuleBase):
    """Weight container with the reference's parameter names (masked_conv.py:231-305): a 5x5 context
    convolution C -> 2C and the three 1x1 "param merger" convolutions 4C -> 10C/3 -> 8C/3 -> 2C, all with
    nn.Conv2d's default initialisation, so a reference state_dict loads unchanged.  Inference happens in the
    CUDA library; this module has no forward()."""

    def __init__(self, in_channels=192, out_channels=384, kernel_size=5, use_param_merger=True,
                 param_merger_in_channels=None, param_merger_mid_channels_list=None, param_merger_kernel_size=1, **kwargs):
        super().__init__()
        if param_merger_kernel_size != 1:
            raise NotImplementedError("param_merger_kernel_size != 1")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.use_param_merger = use_param_merger
        self.context_prediction = nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2)
        if use_param_merger:
            cin = out_channels * 2 if param_merger_in_channels is None else param_merger_in_channels
            mids = [out_channels * 5 // 3, out_channels * 4 // 3] if param_merger_mid_channels_list is None \
                else list(param_merger_mid_channels_list)
            if cin != out_channels * 2 or len(mids) != 2:
                raise NotImplementedError("only the default merger geometry (4C -> 10C/3 -> 8C/3 -> 2C) is accelerated")
            self.param_merger_in = nn.Conv2d(cin, mids[0], 1)
            self.param_merger_out = nn.Sequential(nn.LeakyReLU(inplace=True), nn.Conv2d(mids[0], mids[1], 1),
                                                  nn.LeakyReLU(inplace=True), nn.Conv2d(mids[1], out_channels, 1))

    def forward(self, *a, **k):
        raise NotImplementedError("inference runs in the CUDA library (prior_coder.encode / decode)")


class GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder(CoderModuleBase):
    """encode / decode / update_state of the reference coder of the same name (pgm_coder.py:983-2070).

    Extra keywords: ``lanes`` (1 = reference bitstream byte for byte; 0 = multi-lane container within 0.5 % of it,
    the default; N = ceil(N/32) chunks per group segment), ``ans_params_device`` (device of the float32 Gaussian
    pmf evaluation in update_state; None = the module's device, like the reference; "cpu" reproduces a
    CPU-run reference table bit for bit), ``ctx_precision`` ("fp32" = exact FP32 FMA kernel, "tf32x3" / "fp16x3" = tcgen05
    tensor cores with error-compensated TF32 / FP16 products (both within 1e-5 of fp32; fp16x3 falls back to tf32x3 when an
    activation reaches 4000 in magnitude and records that in the stream), "auto" = fp32 in the lanes=1 compatibility mode and
    fp16x3 otherwise; a multi-lane stream tells the decoder which tensor mode wrote it, fp32 and the tensor modes do not mix) and ``ctx_accumulators`` (k-blocks of 32 accumulated in tensor memory before a
    partial sum is drained into the FP32 register accumulator)."""

    def __init__(self, *args, in_channels=256, channel_groups=1, default_topo_group_method="none",
                 default_num_topo_groups=-1, topo_group_context_model=None, kernel_size=5, use_param_merger=True,
                 use_joint_ar_model_impl=False, coder_type="rans64", freq_precision=16, use_bypass_coding=True,
                 bypass_precision=4, data_precision=8, quantizer_type="uniform", quantizer_params=None,
                 fixed_input_shape=None, force_input_prior_shape_aligned=True, use_autoregressive_encode=True,
                 lower_bound_scale=0.11, scale_table=None, topo_group_predictor=None, lanes=0, ans_params_device=None,
                 ctx_precision="auto", ctx_accumulators=4, param_merger_expand_bottleneck=False, **kwargs):
        super().__init__()
        if use_joint_ar_model_impl and (channel_groups != 1 or topo_group_context_model is not None):
            raise ValueError("use_joint_ar_model_impl needs channel_groups == 1 and the coder's own context model")
        if coder_type not in ("rans", "rans64"):
            # torch_ans.py:241-243 builds Tans* with table_log = freq_precision = 16, which the reference decoder
            # itself refuses (TANS_MAX_TABLELOG = 12); tANS is available through cbench_basic_b200.ans only.
            raise NotImplementedError(f"Unknown coder type {coder_type}")
        # torch_ans.py:32-56: "uniform" [zero point, levels, step] and "uniform_scale" [step]; the affine transform is an
        # elementwise torch op around the fused call (identity, bit for bit, for the default [0, 128, 1])
        if quantizer_type == "uniform":
            quantizer_params = [0.0, 1 << data_precision - 1, 1.0] if quantizer_params is None else quantizer_params
            assert len(quantizer_params) == 3
        elif quantizer_type == "uniform_scale":
            quantizer_params = [1.0] if quantizer_params is None else quantizer_params
            assert len(quantizer_params) == 1
        else:
            raise NotImplementedError(f"Unknown quantizer_type {quantizer_type}")   # the reference's own transforms raise for these
        if not use_autoregressive_encode:
            # pgm_coder.py:310-335 + :975: the encoder then truncates round(y) - mean while the decoder adds round(mean): the
            # reference's own round trip is lossy there and no BaSIC configuration uses it
            raise NotImplementedError("use_autoregressive_encode=False")
        if default_topo_group_method not in topo_groups.METHODS:
            raise NotImplementedError(f"Unknown default_topo_group_method {default_topo_group_method}")
        self.in_channels = in_channels
        self.channel_groups = channel_groups
        if default_topo_group_method in ("channelwise-g10", "elic"):
            self.channel_groups = in_channels // 16
        self.default_topo_group_method = default_topo_group_method
        self.kernel_size = kernel_size
        self.use_param_merger = use_param_merger
        self.use_joint_ar_model_impl = use_joint_ar_model_impl
        self.coder_type, self.freq_precision = coder_type, freq_precision
        self.quantizer_type, self.data_precision = quantizer_type, data_precision
        self.register_buffer("quantizer_params", torch.as_tensor(quantizer_params, dtype=torch.float32), persistent=False)
        self.use_bypass_coding, self.bypass_precision = use_bypass_coding, bypass_precision
        self.fixed_input_shape = fixed_input_shape
        self.force_input_prior_shape_aligned = force_input_prior_shape_aligned
        self.lower_bound_scale = lower_bound_scale
        self.lanes = lanes
        self.ans_params_device = ans_params_device
        if ctx_precision not in ("auto", "fp32", "tf32x3", "fp16x3"):
            raise ValueError(f"Unknown ctx_precision {ctx_precision}")
        self.ctx_precision, self.ctx_accumulators = ctx_precision, ctx_accumulators
        self.scale_table = get_scale_table() if scale_table is None else torch.as_tensor(scale_table, dtype=torch.float32)
        self.topo_group_predictor = topo_group_predictor
        if topo_group_predictor is not None:
            self.register_buffer("topo_group_predictor_cache", topo_group_predictor())   # pgm_coder.py:1097-1103
        self.topo_group_context_model = topo_group_context_model
        if topo_group_context_model is None:
            # the coder's own context model (pgm_coder.py:1177-1239), parameter names of the reference
            out = 2 * in_channels
            self.context_prediction = nn.Conv2d(in_channels, out, kernel_size, padding=kernel_size // 2)
            if use_joint_ar_model_impl:
                # CompressAI-style serial coder (pgm_coder.py:1204-1213, :1975-2066): pixel by pixel in raster
                # order, causally masked 5x5 convolution, entropy_parameters(cat(prior, ctx)) with scales in the first half
                # of its output and means in the second.  On the device that is the scanline map with the merger's first
                # matrix read as [ctx | prior] columns and its last one as interleaved (mean, scale) rows (_upload_weights).
                self.entropy_parameters = nn.Sequential(nn.Conv2d(out * 2, out * 5 // 3, 1), nn.LeakyReLU(inplace=True),
                                                        nn.Conv2d(out * 5 // 3, out * 4 // 3, 1), nn.LeakyReLU(inplace=True),
                                                        nn.Conv2d(out * 4 // 3, out, 1))
            elif use_param_merger:
                # three masked 1x1 convolutions over 2G channel groups [context | prior], 4C -> bottleneck -> bottleneck -> 4C,
                # the context half of the result kept (_merge_prior_params, :1606-1638); runs on the exact FP32 kernels
                bott = out * 4 if param_merger_expand_bottleneck else out * 2
                if (bott // 2) % self.channel_groups or in_channels % self.channel_groups:
                    raise ValueError("channel counts must be divisible by channel_groups")
                self.param_merger = nn.Sequential(nn.Conv2d(out * 2, bott, 1), nn.LeakyReLU(inplace=True),
                                                  nn.Conv2d(bott, bott, 1), nn.LeakyReLU(inplace=True),
                                                  nn.Conv2d(bott, out * 2, 1))
        self.out_channels = 2 * in_channels
        self._ctx_holder = _CtxHandle()
        self._map_key = None       # (H, W, bytes of the group map) the library currently holds

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        # entries of a reference checkpoint that carry nothing for coding: the unused compatibility copy of the conv kernel
        # (pgm_coder.py:1179-1181) and the LowerBound buffer (its value is the constructor's lower_bound_scale)
        for key in ("conv_kernel_weight", "conv_kernel_bias", "lower_bound_scale.bound"):
            state_dict.pop(prefix + key, None)
        super()._load_from_state_dict(state_dict, prefix, *a, **k)

    # ------------------------------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self._device_indicator.device

    def _dev_index(self):
        if self.device.type != "cuda":
            raise N.CudaError("cbench_basic_b200 coders run on CUDA devices only: move the module with .cuda() "
                              "(there is no CPU fallback)")
        return self.device.index if self.device.index is not None else torch.cuda.current_device()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _sync(self):
        torch.cuda.synchronize(self.device)

    def _get_ans_params(self):
        """torch_ans.py:284-310, the float32 torch calls verbatim (their rounding decides truncated counts)."""
        device = self.device if self.ans_params_device is None else torch.device(self.ans_params_device)
        freq_cnt = 1 << self.freq_precision
        tail_mass = torch.tensor([0.5 / freq_cnt], device=device)
        bound = torch.tensor([float(self.lower_bound_scale)], device=device)
        cnts, nsym, offs = [], [], []
        for s in self.scale_table.to(device=device):
            scale = torch.max(s.reshape(1, 1), bound)
            dist = torch.distributions.Normal(torch.zeros(1, 1, device=device), scale)
            dmin = int(dist.icdf(tail_mass).floor().item())
            dmax = int(dist.icdf(1 - tail_mass).ceil().item())
            offs.append(dmin)
            nsym.append(dmax - dmin + 1)
            pts = torch.arange(dmin - 1, dmax + 1, device=device).type_as(dist.mean) + 0.5
            logp = (dist.cdf(pts[1:]) - dist.cdf(pts[:-1])).log()[0]
            cnt = (torch.softmax(logp, dim=-1) * freq_cnt).clamp_min(1)
            cnts.append(cnt.detach().cpu().contiguous().numpy().astype(np.int32))
        freqs = np.zeros((len(cnts), max(len(c) for c in cnts)), dtype=np.int32)
        for i, c in enumerate(cnts):
            freqs[i, :len(c)] = c
        return freqs, np.array(nsym, dtype=np.int32), np.array(offs, dtype=np.int32)

    def update_state(self, *args, **kwargs) -> None:
        """torch_ans.py:237-251: builds the coder objects and their tables -- here on the GPU, plus the
        context-model weights in the layout the kernels want."""
        dev = self._dev_index()
        kw = dict(freq_precision=self.freq_precision, bypass_coding=self.use_bypass_coding,
                  bypass_precision=self.bypass_precision, lanes=self.lanes, device=dev)
        encoder, decoder = ans.Rans64Encoder(**kw), ans.Rans64Decoder(**kw)
        freqs, nfreqs, offsets = self._get_ans_params()
        tab = np.ascontiguousarray(self.scale_table.detach().cpu().numpy(), dtype=np.float32)
        for coder in (encoder, decoder):
            coder.init_params(freqs, nfreqs, offsets)
            N.check(N.lib().basic_coder_set_scale_table(coder.handle, tab.ctypes.data, tab.size))
        self.ans_encoder, self.ans_decoder = encoder, decoder
        self._upload_weights(dev)

    def _upload_weights(self, dev):
        self._ctx_holder.reset()
        self._map_key = None
        h = C.c_void_p()
        N.check(N.lib().basic_ctx_create(self.in_channels, self.channel_groups, self.kernel_size, dev, C.byref(h)))
        self._ctx_holder.h = h
        keep = []

        def ptr(t):
            t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()
        cm = self.topo_group_context_model
        if cm is not None:
            ptrs = [ptr(cm.context_prediction.weight), ptr(cm.context_prediction.bias)]
            if getattr(cm, "use_param_merger", True):
                ptrs += [ptr(cm.param_merger_in.weight), ptr(cm.param_merger_in.bias),
                         ptr(cm.param_merger_out[1].weight), ptr(cm.param_merger_out[1].bias),
                         ptr(cm.param_merger_out[3].weight), ptr(cm.param_merger_out[3].bias)]
            else:
                ptrs += [None] * 6
        elif self.use_joint_ar_model_impl:
            e0, e2, e4 = self.entropy_parameters[0], self.entropy_parameters[2], self.entropy_parameters[4]
            w0, w4, b4 = joint_ar_remap(e0.weight, e4.weight, e4.bias, self.in_channels)
            ptrs = [ptr(self.context_prediction.weight), ptr(self.context_prediction.bias), ptr(w0), ptr(e0.bias),
                    ptr(e2.weight), ptr(e2.bias), ptr(w4), ptr(b4)]
        elif self.use_param_merger:
            o = 2 * self.in_channels
            pm0, pm2, pm4 = self.param_merger[0], self.param_merger[2], self.param_merger[4]
            half = pm0.weight.shape[0] // 2
            w0, w2, w4 = pm0.weight.reshape(2 * half, 2 * o), pm2.weight.reshape(2 * half, 2 * half), pm4.weight.reshape(2 * o, 2 * half)
            ptrs = [ptr(self.context_prediction.weight), ptr(self.context_prediction.bias),
                    ptr(w0[:half]), ptr(pm0.bias[:half]), ptr(w2[:half]), ptr(pm2.bias[:half]), ptr(w4[:o]), ptr(pm4.bias[:o]),
                    ptr(w0[half:, o:]), ptr(pm0.bias[half:]), ptr(w2[half:, half:]), ptr(pm2.bias[half:])]
        else:
            ptrs = [ptr(self.context_prediction.weight), ptr(self.context_prediction.bias)] + [None] * 6
        self._sync()
        if len(ptrs) == 12:
            N.check(N.lib().basic_ctx_set_weights_internal(self._ctx, *ptrs, int(half)))
        else:
            N.check(N.lib().basic_ctx_set_weights(self._ctx, *ptrs))
        prec = self.ctx_precision
        if prec == "auto":
            prec = "fp32" if self.lanes == 1 else "fp16x3"
        N.check(N.lib().basic_ctx_set_precision(self._ctx, {"fp32": N.CTX_FP32, "tf32x3": N.CTX_TF32X3, "fp16x3": N.CTX_FP16X3}[prec],
                                                int(self.ctx_accumulators)))

    @property
    def _ctx(self):
        return self._ctx_holder.h

    # ------------------------------------------------------------------------------------------ group map
    def _get_pgm(self, input_shape, pgm=None):
        """pgm_coder.py:1498-1604 (eval branch): explicit map > cached predictor output > default method."""
        H, W = int(input_shape[-2]), int(input_shape[-1])
        if self.use_joint_ar_model_impl:   # the serial coder ignores the map arguments (pgm_coder.py:1978-1985)
            return topo_groups.default_map("scanline", 1, H, W)
        if pgm is None and self.topo_group_predictor is not None:
            pgm = self.topo_group_predictor_cache
        if pgm is None:   # (the default map of a shape is built once: ~30 us of host time per call otherwise)
            cache = self.__dict__.setdefault("_default_maps", {})
            key = (self.default_topo_group_method, self.channel_groups, H, W)
            if key not in cache:
                if len(cache) > 64:
                    cache.clear()
                cache[key] = topo_groups.default_map(self.default_topo_group_method, self.channel_groups, H, W)
            return cache[key]
        return topo_groups.tile_map(pgm, self.channel_groups, H, W)

    def _set_map(self, tg):
        """Hands the group map to the library -- once per (H, W, map): the cell lists, tap masks and position lists it derives
        stay on the device between calls."""
        if self._map_key is not None and tg is self.__dict__.get("_map_obj"):
            return   # the very tensor of the last call (a cached default map)
        tg32 = np.ascontiguousarray(tg[0].numpy(), dtype=np.int32)
        key = (tg32.shape, tg32.tobytes())
        if key != self._map_key:
            self._map_key = None
            N.check(N.lib().basic_ctx_set_map(self._ctx, tg32.ctypes.data, tg32.shape[1], tg32.shape[2]))
            self._map_key = key
        self.__dict__["_map_obj"] = tg

    def _operand(self, t):
        """float32, contiguous; a CPU tensor stays where it is: the C ABI takes host pointers and uploads them on its own
        copy stream, overlapped with the first kernels (pinned memory makes that asynchronous)."""
        t = t.detach()
        if t.device.type == "cpu":
            return t.to(dtype=torch.float32).contiguous()
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------------------------------ quantiser (torch_ans.py:105-180)
    def _qparams(self, quantizer_params):
        if quantizer_params is None:
            # the module's own buffer lives on the device: read it once per (storage, version), not with a device
            # synchronisation in every call
            q = self.quantizer_params
            key = (q.data_ptr(), q._version, str(q.device))
            cached = self.__dict__.get("_q_host")
            if cached is None or cached[0] != key:
                cached = (key, self._zero_step(torch.as_tensor(q.detach().cpu().tolist(), dtype=torch.float32)))
                self.__dict__["_q_host"] = cached
            return cached[1]
        return self._zero_step(torch.as_tensor(quantizer_params, dtype=torch.float32))

    def _zero_step(self, q):
        if self.quantizer_type == "uniform":
            return float(q[0]), float(q[2])
        return 0.0, float(q.reshape(-1)[0])

    def _transform(self, x, quantizer_params=None):
        zero, step = self._qparams(quantizer_params)
        return x if (zero == 0.0 and step == 1.0) else (x - zero) / step

    def _inverse_transform(self, x, quantizer_params=None):
        zero, step = self._qparams(quantizer_params)
        return x if (zero == 0.0 and step == 1.0) else x * step + zero

    def _align_prior(self, prior, spatial):
        """pgm_coder.py:571-577, :611-616: same spatial size, or the prior narrowed to the input's."""
        if self.force_input_prior_shape_aligned:
            assert tuple(spatial) == tuple(prior.shape[2:]), "Input and prior shape not aligned! Consider setting force_input_prior_shape_aligned = False, which may add a little overhead to the bitstream to save the input shape!"
        else:
            for dim, size in enumerate(spatial, 2):
                prior = prior.narrow(dim, 0, size)
        return prior

    # ------------------------------------------------------------------------------------------ coding
    def forward(self, input, prior=None, pgm=None, quantizer_params=None, **kwargs):
        """Eval mode of pgm_coder.py:391-539: returns ``input_dequant`` = dequantise(round(transform(input))) -- the tensor the
        reference hands to the generative edges while it encodes (latent_graph.py:836-841).  The likelihood / rate terms of
        that method feed training and the ``prior_entropy`` monitor only; they are not evaluated here."""
        if self.training:
            raise NotImplementedError("training likelihood (pgm_coder.py:391-539) is outside the accelerated hot path; "
                                      "use the reference module for training and load its state_dict here for coding")
        if prior is not None:
            self._align_prior(prior, input.shape[2:])
        with self.profiler.start_time_profile("time_data_preprocess_encode"):
            return self._inverse_transform(torch.round(self._transform(input, quantizer_params)), quantizer_params)

    def encode(self, input, *args, prior=None, pgm=None, quantizer_params=None, return_yhat=False, zero_copy=False, **kwargs) -> bytes:
        """pgm_coder.py:912-951 + the framing of :581-597.  zero_copy=True (ours, opt-in; needs a configuration that writes no
        framing header): instead of `bytes` a read-only memoryview over the coder's page-locked output buffer, valid until the
        next encode() of this object -- decode() takes it as it is and uploads straight from it."""
        assert hasattr(self, "ans_encoder"), "Not Initialized! Should call self.update_state() before coding!"
        if prior is None:
            raise ValueError("prior should not be None!")
        B, Cc, H, W = input.shape
        prior = self._align_prior(prior, (H, W))
        if Cc != self.in_channels or tuple(prior.shape) != (B, 2 * Cc, H, W):
            raise ValueError(f"expected input [B, {self.in_channels}, H, W] and prior [B, {2 * self.in_channels}, H, W], got "
                             f"{tuple(input.shape)} and {tuple(prior.shape)}")
        with self.profiler.start_time_profile("time_prior_preprocess_encode"):
            y, p = self._operand(self._transform(input, quantizer_params)), self._operand(prior)
            self._set_map(self._get_pgm(input.shape, pgm))
        head = b""
        if self.fixed_input_shape is not None:
            assert B == self.fixed_input_shape[0] and tuple(input.shape[2:]) == tuple(self.fixed_input_shape[1:])
        elif not self.force_input_prior_shape_aligned:
            head = struct.pack("B", 3) + struct.pack("<H", B) + struct.pack("<H", H) + struct.pack("<H", W)  # :581-597
        with self.profiler.start_time_profile("time_ans_encode"):
            h = self.ans_encoder.handle
            out_len = C.c_int64(0)
            yhat = torch.empty(y.shape, dtype=torch.float32, device=self.device) if return_yhat else None
            # out = NULL: the stream lands in the coder's pinned host buffer; one copy makes the bytes object -- header and
            # stream in ONE object filled in place (prepending seven bytes to a 16 MB stream would copy it again)
            N.check(N.lib().basic_ypath_encode(h, self._ctx, y.data_ptr(), p.data_ptr(), B, Cc, H, W, self.lanes, None, 0,
                                               C.byref(out_len), yhat.data_ptr() if return_yhat else None,
                                               self._stream()))
            if zero_copy and head:
                raise ValueError("zero_copy needs force_input_prior_shape_aligned=True or fixed_input_shape (no framing header)")
            byte_string = N.last_output_view(h) if zero_copy else N.last_output(h, prefix=head)
        if return_yhat:
            return byte_string, self._inverse_transform(yhat, quantizer_params)
        return byte_string

    def decode(self, byte_string: bytes, *args, prior=None, pgm=None, quantizer_params=None, **kwargs) -> torch.Tensor:
        assert hasattr(self, "ans_decoder"), "Not Initialized! Should call self.update_state() before coding!"
        assert prior is not None
        if self.fixed_input_shape is not None:
            ptr, B, spatial = 0, self.fixed_input_shape[0], tuple(self.fixed_input_shape[1:])
        elif self.force_input_prior_shape_aligned:
            ptr, B, spatial = 0, prior.shape[0], tuple(prior.shape[2:])
        else:
            nd = struct.unpack("B", byte_string[:1])[0]
            flat = [struct.unpack("<H", byte_string[1 + 2 * i:3 + 2 * i])[0] for i in range(nd)]
            ptr, B, spatial = 1 + 2 * nd, flat[0], tuple(flat[1:])
        prior = self._align_prior(prior, spatial)
        H, W = spatial
        Cc = self.in_channels
        if tuple(prior.shape) != (B, 2 * Cc, H, W):
            raise ValueError(f"expected prior [{B}, {2 * Cc}, {H}, {W}], got {tuple(prior.shape)}")
        with self.profiler.start_time_profile("time_prior_preprocess_decode"):
            p = self._operand(prior)
            self._set_map(self._get_pgm((B, Cc, H, W), pgm))
        with self.profiler.start_time_profile("pgm_generate_coding"):
            enc = np.frombuffer(byte_string, dtype=np.uint8, offset=ptr)
            yhat = torch.empty(B, Cc, H, W, dtype=torch.float32, device=self.device)
            N.check(N.lib().basic_ypath_decode(self.ans_decoder.handle, self._ctx, enc.ctypes.data if enc.size else None, enc.size,
                                               p.data_ptr(), B, Cc, H, W, self.lanes, yhat.data_ptr(),
                                               self._stream()))
        return self._inverse_transform(yhat, quantizer_params)


class CombinedNNTrainablePGMPriorCoder(CoderModuleBase):
    """BaSIC "dynamic entropy coder" dispatch (pgm_coder.py:632-715): picks one sub-coder by blend_weight.argmax()."""

    def __init__(self, coders, *args, fix_weight=False, **kwargs):
        super().__init__()
        self.coders = nn.ModuleList(coders)
        if fix_weight:
            self.register_buffer("default_blend_weight", torch.zeros(len(coders)), persistent=False)
        else:
            self.default_blend_weight = nn.Parameter(torch.zeros(len(coders)))

    def _pick(self, blend_weight):
        if blend_weight is None:
            blend_weight = torch.softmax(self.default_blend_weight, dim=0)
        return self.coders[int(blend_weight.argmax().item())]

    def forward(self, input, prior=None, blend_weight=None, **kwargs):
        """Eval branch of pgm_coder.py:651-696: the selected sub-coder's forward."""
        if self.training:
            raise NotImplementedError("training forward stays with the reference module")
        return self._pick(blend_weight)(input, prior=prior, **kwargs)

    def encode(self, input, *args, prior=None, blend_weight=None, **kwargs) -> bytes:
        return self._pick(blend_weight).encode(input, prior=prior, **kwargs)

    def decode(self, byte_string, *args, prior=None, blend_weight=None, **kwargs):
        return self._pick(blend_weight).decode(byte_string, prior=prior, **kwargs)

    def update_state(self, *args, **kwargs) -> None:
        for coder in self.coders:
            coder.update_state(*args, **kwargs)

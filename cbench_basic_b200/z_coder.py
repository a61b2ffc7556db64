"""z-node factorized-prior coder: drop-in for the reference's
``CompressAIEntropyBottleneckPriorCoder`` (cbench/modules/prior_model/prior_coder/compressai_coder.py:87-248).

The reference delegates the arithmetic to a third-party dependency that is NOT vendored:
``compressai==1.2.3`` (requirements.txt:15) -- ``EntropyBottleneck.update / compress / decompress`` and its
``compressai.ans`` coder.  What is restated here is compressai's published algorithm (Balle et al. 2018, appendix 6.1;
compressai/entropy_models/entropy_models.py), anchored on what the reference tree itself holds:

* the cumulative-logits network: an in-tree copy, cbench/nn/layers/param_generator.py:158-199 (same parameter names
  ``_matrix{i}``, ``_bias{i}``, ``_factor{i}``);
* the coder: ``cbench.rans`` (cbench/csrc/rans/rans_interface.cpp, an in-tree clone of ``compressai.ans``) produces the
  same bytes as ``cbench.ans`` Rans64 for the same CDFs (SURVEY 8c probe; tests/test_oracle_golden.py pins it), i.e. the
  lanes=1 stream our CUDA coder reproduces byte for byte;
* ``pmf_to_quantized_cdf``: rans_interface.cpp / rans64.cpp:69-126, built bit-exactly on the device (tables.cu);
* the framing: ``write_body`` / ``read_body``, compressai_coder.py:63-84 (BIG-endian u32 h, w, n_strings, then per image a
  big-endian u32 length + bytes).

Parity status: stream format and framing PINNED (golden vectors made with the reference's own write_body and cbench.rans,
tests/golden/make_z_golden.py); the CDF construction out of the network parameters is **parity unpinned** -- compressai is
absent from this image, so nothing can run the original.

Every image is one lanes=1 stream (the reference's format); all images of a call are coded in ONE launch, one CTA per
stream, through the C ABI (`basic_coder_encode_batch / _decode_batch`, tables resident on the device, staged into shared
memory by every CTA).  Quantisation is two elementwise torch ops on the
device (N_z = N / 16: not a hot spot).  No CPU fallback: without the library or a GPU the coder raises.
"""
import io
import struct

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ans
from .module_api import CoderModuleBase


# ---------------------------------------------------------------------------------------------------- framing
def write_body(shape, out_strings) -> bytes:
    """compressai_coder.py:76-84 with segments = 1."""
    with io.BytesIO() as fd:
        fd.write(struct.pack(">3I", int(shape[0]), int(shape[1]), len(out_strings)))
        for s in out_strings:
            fd.write(struct.pack(">I", len(s)))
            fd.write(s)
        return fd.getvalue()


def read_body(data: bytes):
    """compressai_coder.py:63-73 with segments = 1; raises ValueError on a truncated body."""
    if len(data) < 12:
        raise ValueError("z stream shorter than its header")
    h, w, n = struct.unpack_from(">3I", data, 0)
    at, strings = 12, []
    for _ in range(n):
        if at + 4 > len(data):
            raise ValueError("z stream truncated")
        (ln,) = struct.unpack_from(">I", data, at)
        at += 4
        if at + ln > len(data):
            raise ValueError("z stream truncated")
        strings.append(data[at:at + ln])
        at += ln
    return strings, (h, w)


# ---------------------------------------------------------------------------------------------------- model
class EntropyBottleneck(nn.Module):
    """Factorized density model of compressai 1.2.3 (constructor keywords, parameter / buffer names and state_dict keys
    of the original, so a trained reference checkpoint loads)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), likelihood_bound=1e-9,
                 entropy_coder_precision=16, device_index=0, lanes=1, **kwargs):
        super().__init__()
        self.lanes = int(lanes)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self._device_index = int(device_index)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = float(np.log(np.expm1(1 / scale / filters[i + 1])))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((self.channels, filters[i + 1], filters[i]), init)))
            bias = torch.empty(self.channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(self.channels, filters[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(self.channels, 1, 1))
        target = float(np.log(2 / self.tail_mass - 1))
        self.register_buffer("target", torch.tensor([-target, 0.0, target]))
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._enc = self._dec = None

    @property
    def device_index(self):
        """The GPU the coder objects live on: the module's own device once it has been moved to CUDA (what the reference's
        harness does with the whole codec), else the constructor's ``device_index``."""
        d = self.quantiles.device
        if d.type == "cuda":
            return d.index if d.index is not None else torch.cuda.current_device()
        return self._device_index

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        # later compressai releases keep the same tensors in ParameterLists (matrices.0, biases.0, factors.0)
        for new, old in (("matrices.", "_matrix"), ("biases.", "_bias"), ("factors.", "_factor")):
            for key in [x for x in state_dict if x.startswith(prefix + new)]:
                state_dict[prefix + old + key[len(prefix + new):]] = state_dict.pop(key)
        for name in ("_offset", "_quantized_cdf", "_cdf_length"):   # table buffers take the checkpoint's size
            if prefix + name in state_dict:
                setattr(self, name, state_dict[prefix + name].to(torch.int32).clone())
        super()._load_from_state_dict(state_dict, prefix, *a, **k)

    def _logits_cumulative(self, inputs):
        """param_generator.py:182-199 (the reference's in-tree copy) with stop_gradient = True; evaluated where `inputs`
        live -- update() passes CPU tensors, so the tables do not depend on the GPU's libm."""
        par = lambda name: getattr(self, name).detach().to(device=inputs.device, dtype=torch.float32)  # noqa: E731
        logits = inputs
        for i in range(len(self.filters) + 1):
            logits = torch.matmul(F.softplus(par(f"_matrix{i:d}")), logits)
            logits = logits + par(f"_bias{i:d}")
            if i < len(self.filters):
                logits = logits + torch.tanh(par(f"_factor{i:d}")) * torch.tanh(logits)
        return logits

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    @torch.no_grad()
    def update(self, force=False):
        """EntropyBottleneck.update + EntropyModel._pmf_to_cdf: per channel the pmf on [median - minima, median + maxima]
        plus the tail mass, quantised to 16-bit CDFs on the device; then the tables go to the CUDA coder."""
        if self._offset.numel() > 0 and not force and self._enc is not None:
            return False
        q = self.quantiles.detach().float().cpu()
        medians = q[:, 0, 1]
        minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
        maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
        self._offset = (-minima).to(torch.int32)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length, dtype=torch.float32)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5)
        upper = self._logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        cdf = torch.zeros(self.channels, max_length + 2, dtype=torch.int32)
        for c in range(self.channels):
            prob = torch.cat((pmf[c, :int(pmf_length[c])], tail[c]), dim=0)
            qc = ans.pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision, device=self.device_index)
            cdf[c, :len(qc)] = torch.tensor(qc, dtype=torch.int32)
        self._quantized_cdf = cdf
        self._cdf_length = (pmf_length + 2).to(torch.int32)
        self._push_tables()
        return True

    def _coders(self):
        if self._enc is None or self._enc.device != self.device_index:
            self._push_tables()
        return self._enc, self._dec

    def _push_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        self._enc = ans.Rans64Encoder(freq_precision=self.entropy_coder_precision, lanes=self.lanes, device=self.device_index)
        self._dec = ans.Rans64Decoder(freq_precision=self.entropy_coder_precision, lanes=self.lanes, device=self.device_index)
        for c in (self._enc, self._dec):
            c.init_cdf_params(self._quantized_cdf.cpu().numpy(), self._cdf_length.cpu().numpy(), self._offset.cpu().numpy())

    def _indexes(self, B, spatial, device):
        idx = torch.arange(self.channels, dtype=torch.int32, device=device).view(1, -1, *([1] * len(spatial)))
        return idx.expand(B, self.channels, *spatial).contiguous()

    @torch.no_grad()
    def compress(self, x):
        """One lanes=1 stream per image: symbols = round(x - median), table = channel (EntropyModel.compress)."""
        enc, _ = self._coders()
        dev = torch.device("cuda", self.device_index)
        x = x.detach().to(device=dev, dtype=torch.float32)
        med = self._get_medians().to(dev).view(1, -1, *([1] * (x.dim() - 2)))
        sym = torch.round(x - med).to(torch.int32)
        idx = self._indexes(x.shape[0], x.shape[2:], dev)
        if x.shape[0] == 0:
            return []
        if self.lanes != 1:  # our multi-lane container over the whole batch: u32 batch size | container
            return [struct.pack("<I", x.shape[0]) + enc.encode_with_indexes(sym.reshape(-1), idx.reshape(-1))]
        return enc.encode_batch(sym.reshape(x.shape[0], -1), idx.reshape(x.shape[0], -1))  # one CTA per image

    @torch.no_grad()
    def decompress(self, strings, size):
        _, dec = self._coders()
        dev = torch.device("cuda", self.device_index)
        spatial = tuple(int(s) for s in size)
        if self.lanes != 1 and len(strings) == 1:
            if len(strings[0]) < 4:
                raise ValueError("z stream truncated")
            (B,) = struct.unpack_from("<I", strings[0], 0)
            idx = self._indexes(B, spatial, dev)
            med = self._get_medians().to(dev).view(1, -1, *([1] * len(spatial)))
            sym = dec.decode_with_indexes(strings[0][4:], idx.reshape(-1))
            return sym.view(B, self.channels, *spatial).to(torch.float32) + med
        idx = self._indexes(len(strings), spatial, dev)
        med = self._get_medians().to(dev).view(1, -1, *([1] * len(spatial)))
        if len(strings) == 0:
            return torch.empty(0, self.channels, *spatial, dtype=torch.float32, device=dev)
        sym = dec.decode_batch(strings, idx.reshape(len(strings), -1))            # one CTA per image
        return sym.view(len(strings), self.channels, *spatial).to(torch.float32) + med


class CompressAIEntropyBottleneckPriorCoder(CoderModuleBase):
    """compressai_coder.py:87-248: encode(input, channel_gains=) -> bytes, decode(bytes, channel_gains_inv=) -> Tensor,
    update_state(), and the eval-mode forward() -- the dequantised latent the reference's latent graph hands to h_s while it
    encodes (latent_graph.py:836-841).  Training forward() (likelihoods, aux loss) stays with the reference module."""

    def __init__(self, entropy_bottleneck_channels=256, eps=1e-7, use_inner_aux_opt=False, use_bit_rate_loss=True,
                 freeze_params=False, training_output_straight_through=False, device_index=0, lanes=1, **kwargs):
        super().__init__()
        # lanes = 1: the reference's streams (one per image).  lanes = 0 / N: ONE multi-lane container
        # over the whole batch inside the same write_body framing -- not readable by compressai, ~30x faster
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels, device_index=device_index, lanes=lanes)
        self.eps = eps
        if freeze_params:
            for p in self.parameters():
                p.requires_grad = False

    @staticmethod
    def _channelwise_mul(x, gain):
        return (x.view(x.shape[0], x.shape[1], -1) * gain.to(x.device).unsqueeze(0).unsqueeze(-1)).view_as(x)

    def update_state(self, *args, **kwargs) -> None:
        self.entropy_bottleneck.update(force=True)

    def encode(self, input, *args, channel_gains=None, channel_gains_inv=None, **kwargs) -> bytes:
        if channel_gains is not None:
            input = self._channelwise_mul(input, channel_gains)
        return write_body(input.shape[-2:], self.entropy_bottleneck.compress(input))

    def decode(self, byte_string, *args, channel_gains=None, channel_gains_inv=None, **kwargs):
        strings, shape = read_body(bytes(byte_string))
        y_hat = self.entropy_bottleneck.decompress(strings, shape)
        if channel_gains_inv is not None:
            y_hat = self._channelwise_mul(y_hat, channel_gains_inv)
        return y_hat

    def forward(self, input, *args, channel_gains=None, channel_gains_inv=None, **kwargs):
        """Eval mode of compressai_coder.py:203-227: ``entropy_bottleneck(input)[0]`` outside training is
        quantize(x, "dequantize", medians) = round(x - median) + median -- exactly what decode() returns for encode(input).
        The likelihoods only feed the rate loss / the ``prior_entropy`` monitor and are not evaluated."""
        if self.training:
            raise NotImplementedError("training likelihoods (compressai_coder.py:203-227) are outside the accelerated path")
        if channel_gains is not None:
            input = self._channelwise_mul(input, channel_gains)
        med = self.entropy_bottleneck._get_medians().to(input.device).view(1, -1, *([1] * (input.dim() - 2)))
        y_hat = torch.round(input - med) + med
        if channel_gains_inv is not None:
            y_hat = self._channelwise_mul(y_hat, channel_gains_inv)
        return y_hat

"""The two-node latent graph of a hyperprior codec (z -> y) on top of the two CUDA coders: the byte container and the
device-resident hand-over of the prior.

In the reference this wiring is `LatentGraphicalANSEntropyCoder` (cbench/modules/entropy_coder/latent_graph.py:1232-1301):
the node coders are called in generative topological order (z, then y with `prior = h_s(z_hat)`), and their byte strings are
joined with `merge_bytes(..., num_segments=len(nodes))` (bytes_ops.py:19-33: native-endian u32 length before every segment
but the last).  The general Bayesian-network machinery of that class (arbitrary node graphs, training losses, complexity
search) is the reference's control plane and stays there; this class is the fixed z -> y instance of its encode / decode, so
that the z stream is decoded, the hyper-synthesis runs, and the y coder reads its prior without leaving the GPU.

Parity: the container layout is the reference's (bytes_ops, pinned by tests); the node streams are the coders' (pinned in
their own tests).  The wiring itself has no reference vector (the latent graph needs the reference's config system to run).
"""
import torch
import torch.nn as nn

from .bytes_ops import merge_bytes, split_merged_bytes


class HyperpriorLatentCodec(nn.Module):
    """z_coder: `z_coder.CompressAIEntropyBottleneckPriorCoder`; y_coder: `prior_coder.Gaussian...PGMPriorCoder`;
    hyper_synthesis: any module z_hat -> prior [B, 2C, H, W] (the reference's h_s backbone, untouched PyTorch);
    hyper_analysis (optional): y -> z for callers that hand in y only."""

    NODES = ("z", "y")   # generative topological order = order of the segments in the container

    def __init__(self, z_coder, y_coder, hyper_synthesis, hyper_analysis=None):
        super().__init__()
        self.latent_node_entropy_coders = nn.ModuleDict({"z": z_coder, "y": y_coder})   # the reference's attribute name
        self.hyper_synthesis = hyper_synthesis
        self.hyper_analysis = hyper_analysis

    def update_state(self, *args, **kwargs) -> None:
        for coder in self.latent_node_entropy_coders.values():   # latent_graph.py:1297-1301
            coder.update_state(*args, **kwargs)

    @torch.no_grad()
    def encode(self, y, z=None, **y_kwargs) -> bytes:
        coders = self.latent_node_entropy_coders
        if z is None:
            if self.hyper_analysis is None:
                raise ValueError("z not given and no hyper_analysis module")
            z = self.hyper_analysis(y)
        z_hat = coders["z"](z)                       # eval forward = the dequantised z the decoder will see (latent_graph.py:836)
        z_bytes = coders["z"].encode(z)
        prior = self.hyper_synthesis(z_hat)
        y_bytes = coders["y"].encode(y, prior=prior, **y_kwargs)
        return merge_bytes([z_bytes, y_bytes], num_segments=len(self.NODES))

    @torch.no_grad()
    def decode(self, data: bytes, **y_kwargs):
        coders = self.latent_node_entropy_coders
        z_bytes, y_bytes = split_merged_bytes(data, num_segments=len(self.NODES))
        prior = self.hyper_synthesis(coders["z"].decode(z_bytes))
        return coders["y"].decode(y_bytes, prior=prior, **y_kwargs)

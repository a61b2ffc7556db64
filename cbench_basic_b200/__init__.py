"""cbench_basic_b200 -- B200-native (sm_100a) implementation of BaSIC's entropy-coding hot path.

Drop-in replacements, all backed by hand-written CUDA behind the C ABI in include/basic_b200.h:

* ``cbench_basic_b200.ans``         -- the native coder module ``cbench.ans`` (Rans64Encoder/Decoder,
                                      TansEncoder/Decoder, pmf_to_quantized_cdf)
* ``cbench_basic_b200.prior_coder`` -- the y-node prior coder (encode / decode / update_state of
                                      GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder)
* ``cbench_basic_b200.z_coder``     -- the z-node factorized-prior coder (CompressAIEntropyBottleneckPriorCoder: per-image
                                      lanes=1 streams + write_body / read_body framing), SURVEY 8 row f1
* ``cbench_basic_b200.latent_codec``-- the z -> h_s -> y wiring and byte container of the two-node latent graph (row f2)
* ``cbench_basic_b200.sharding``    -- per-image / per-tile partitioning over the GPUs of one node

There is no CPU fallback: importing works anywhere, but every call needs the CUDA library and a GPU.
"""
__version__ = "0.1.0"

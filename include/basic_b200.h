/*
 * basic_b200.h -- C ABI of the B200-native BaSIC entropy-coding hot path.
 *
 * This is the drop-in boundary: the entry points below are what a binding for the reference's native
 * coder module `cbench.ans` (pybind11, cbench/csrc/ans/lib.cpp:10-33) and for the y-node prior coder
 * hot loop (cbench/modules/prior_model/prior_coder/pgm_coder.py:912-981) would call.  Plain pointers and
 * sizes only; every data pointer may be HOST or DEVICE memory (detected with cudaPointerGetAttributes),
 * tables and all computation live on the GPU.  There is no CPU fallback: every call fails with
 * BASIC_ERR_CUDA when no sm_100 device is usable.
 *
 * Return value: 0 on success, negative BASIC_ERR_* otherwise; basic_last_error() gives the message of the
 * last failure on the calling thread.  The Python shim (cbench_basic_b200/ans.py) raises ValueError for
 * BASIC_ERR_VALUE, exactly where the reference raises py::value_error.
 *
 * `stream` arguments are cudaStream_t passed as void* (NULL = default stream).  Calls that return host
 * results (lengths, bytes into host memory) synchronise that stream before returning.
 */
#ifndef BASIC_B200_H
#define BASIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BASIC_OK 0
#define BASIC_ERR_VALUE (-1)    /* the reference's py::value_error categories (rans64.cpp:132,166,209; tans.cpp:416) */
#define BASIC_ERR_CUDA (-2)     /* CUDA runtime failure / no device */
#define BASIC_ERR_CAPACITY (-3) /* caller's output buffer too small */
#define BASIC_ERR_STREAM (-4)   /* malformed / truncated multi-lane container */

#define BASIC_KIND_RANS64 0 /* cbench.ans.Rans64Encoder/Decoder, rans64.hpp:127-148 */
#define BASIC_KIND_TANS 1   /* cbench.ans.TansEncoder/Decoder,   tans.hpp:147-157 */

/* `lanes` argument of the coding calls:
 *   1  -> the reference bitstream, byte for byte (single 64-bit rANS state / single FSE tANS state);
 *   0  -> multi-lane container, chunk count chosen so the container is within `BASIC_AUTO_BPP_BUDGET`
 *         (0.5 %) of the lanes=1 size;
 *  >1  -> multi-lane container with ceil(lanes / 32) chunks of 32 interleaved lanes (rounded so every chunk
 *         holds a multiple of 128 symbols). */
#define BASIC_LANES_REFERENCE 1
#define BASIC_LANES_AUTO 0

typedef struct basic_coder basic_coder; /* one encoder-or-decoder object; replaces a pybind Rans64.../Tans... instance */
typedef struct basic_ctx basic_ctx;     /* context-model weights resident in HBM */

const char *basic_last_error(void);
int basic_device_count(void);

/* ---- coder object: constructor args of rans64.hpp:130,139 / tans.hpp:149,154 ------------------------- */
int basic_coder_create(int kind, unsigned precision /* freq_precision | table_log */, unsigned max_symbol_value,
                       int bypass_coding, unsigned bypass_precision, int device, basic_coder **out);
void basic_coder_destroy(basic_coder *c);

/* init_params (rans64.cpp:128-159, tans.cpp:370-386): freqs int32 [T, M] row-major, num_symbols/offsets [T].
 * The quantised CDFs (or tANS tables) are built ON THE DEVICE and stay resident there. */
int basic_coder_init_params(basic_coder *c, const int32_t *freqs, int T, int M, const int32_t *num_symbols,
                            const int32_t *offsets);
/* init_cdf_params (rans64.cpp:162-182) -- rANS only. */
int basic_coder_init_cdf_params(basic_coder *c, const int32_t *cdfs, int T, int M, const int32_t *cdf_sizes,
                                const int32_t *offsets);
/* get_cdfs (rans64.hpp:40-51): shape, then the table itself ([T, M] int32, zero padded) into host memory. */
int basic_coder_cdfs_shape(basic_coder *c, int *T, int *M);
int basic_coder_get_cdfs(basic_coder *c, int32_t *out);
/* pmf_to_quantized_cdf (rans64.cpp:69-126): module-level helper, pmf host float [n] -> cdf host int32 [n+1]. */
int basic_pmf_to_quantized_cdf(const float *pmf, int n, int precision, int device, int32_t *cdf_out);

/* ---- encode_with_indexes / flush (rans64.cpp:203-386, tans.cpp:534-720) ------------------------------- */
/* Upper bound of the encoded size for n symbols (use it to size `out`). */
int64_t basic_coder_encode_bound(basic_coder *c, int64_t n, int lanes);
/* cache != 0: keep the segment inside the object and return out_len = 0; flush() then returns everything
 * cached so far as one stream (reference: one lanes=1 stream over the concatenated symbols; multi-lane: one
 * container with one segment per cached call, decodable segment by segment with decode_stream). */
int basic_coder_encode(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, int lanes,
                       int cache, uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream);
int basic_coder_flush(basic_coder *c, int lanes, uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream);
/* Every encoding call (basic_coder_encode / _flush / basic_ypath_encode) accepts out == NULL: the stream then stays
 * in the coder's own pinned host buffer and this call returns its address and length (valid until the next
 * encoding call on the same coder).  Saves sizing and page-faulting a worst-case caller buffer. */
int basic_coder_last_output(basic_coder *c, const uint8_t **ptr, int64_t *len);
/* Same delivery, copied into the caller's buffer (a freshly allocated `bytes` object in the Python binding): with
 * out == NULL the encoding call returns while the device-to-host copy is still running in chunks, and this call
 * moves every chunk to `dst` as it lands (a few host threads), so the bus transfer and the host copy overlap.
 * basic_coder_output_size = length of the pending delivery. */
int64_t basic_coder_output_size(basic_coder *c);
int basic_coder_take_output(basic_coder *c, uint8_t *dst, int64_t cap);

/* ---- decode_with_indexes / set_stream / decode_stream (rans64.cpp:389-598, rans64.hpp:104-124) ------- */
int basic_coder_decode(basic_coder *c, const uint8_t *encoded, int64_t len, const int32_t *indexes, int64_t n,
                       int lanes, int32_t *out, void *stream);
int basic_coder_set_stream(basic_coder *c, const uint8_t *encoded, int64_t len, int lanes, void *stream);
int basic_coder_decode_stream(basic_coder *c, const int32_t *indexes, int64_t n, int32_t *out, void *stream);

/* ---- in-coder autoregressive table lookup (ans_interface.hpp:58-105 table branch, ans_interface.cpp:75-135 init_ar_params;
 * used by the reference's lossless coders, entropy_coder/ans.py:78-158) -------------------------------------------------
 * basic_coder_init_ar_params: ar_tables int32 [A, I, D1] (one neighbour) or [A, I, D1, D2] (two; D2 = 0 otherwise), host or
 * device.  With AR tables set, an element is coded with table  ar_tables[ar_index][index][v0]([v1]),
 * v_k = ar_offsets[k][i] > 0 ? symbol[i - ar_offsets[k][i]] + 1 : 0  (rans64.cpp:259-263, :439-443).
 * basic_coder_encode_ar / _decode_ar: ar_indexes int32 [n] or NULL (= 0), ar_offsets int32 [order, n] (order = 1 or 2 as
 * initialised); the reference (lanes = 1) stream only -- the lookup makes every symbol depend on earlier ones.  Out-of-range
 * lookups, undefined behaviour in the reference, are BASIC_ERR_VALUE here. */
int basic_coder_init_ar_params(basic_coder *c, const int32_t *ar_tables, int A, int I, int D1, int D2);
int basic_coder_encode_ar(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *ar_indexes,
                          const int32_t *ar_offsets, int order, uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream);
int basic_coder_decode_ar(basic_coder *c, const uint8_t *encoded, int64_t len, const int32_t *indexes, int64_t n,
                          const int32_t *ar_indexes, const int32_t *ar_offsets, int order, int32_t *out, void *stream);

/* ---- batches of independent reference streams (the z node: compressai_coder.py:233,242 code one stream per image) --
 * n_streams runs of n symbols each ([n_streams, n] row-major, host or device) become n_streams lanes=1 rANS64 streams in
 * ONE launch (one CTA per stream), delivered back to back into `out` (or, with out == NULL, into the coder's pinned
 * buffer: basic_coder_last_output / _take_output); out_lens: host int64 [n_streams], the byte length of every stream.
 * basic_coder_decode_batch is the mirror: `encoded` = the streams back to back, lens = their lengths (host). */
int basic_coder_encode_batch(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, int n_streams,
                             uint8_t *out, int64_t out_cap, int64_t *out_lens, void *stream);
int basic_coder_decode_batch(basic_coder *c, const uint8_t *encoded, const int64_t *lens, int n_streams,
                             const int32_t *indexes, int64_t n, int32_t *out, void *stream);

/* ---- Gaussian conditional: quantise + scale index (pgm_coder.py:802-821,927-941; torch_ans.py:105-159) -- */
/* scale_table: host float [n_scales] (compressai_coder.py:23-30). */
int basic_coder_set_scale_table(basic_coder *c, const float *scale_table, int n_scales);
/* For the n elements selected by `positions` (int32 element offsets inside one image's [C,H,W] volume, the
 * same list for every image, n_pos per image; NULL = all C*H*W in order) of each of B images:
 *   idx = argmin_t |scale - table[t]| (first minimum), sym = rint(y - mean), yhat = sym + mean.
 * params is the [B, 2C, H, W] fp32 tensor with channel 2c = mean, 2c+1 = scale of latent channel c.
 * Outputs in stream order (b-major, then positions order): symbols/indexes int32 [B * n_pos]; yhat is
 * scattered into the [B, C, H, W] buffer `yhat_buf` (may be NULL).  y == NULL selects the DECODER variant:
 * only indexes are produced. */
int basic_gauss_quantize_index(basic_coder *c, const float *y, const float *params, const int32_t *positions,
                               int64_t n_pos, int B, int C, int HW, int32_t *symbols, int32_t *indexes,
                               float *yhat_buf, void *stream);
/* Decoder side of the same step: yhat_buf[b, pos] = symbols + mean. */
int basic_gauss_dequantize(basic_coder *c, const int32_t *symbols, const float *params, const int32_t *positions,
                           int64_t n_pos, int B, int C, int HW, float *yhat_buf, void *stream);

/* ---- context model (cbench/nn/layers/masked_conv.py:231-305, pgm_coder.py:1177-1239,1606-1638) -------- */
/* Weights in the reference's state_dict layout, host or device fp32:
 *   ctx_w [2C, C, k, k], ctx_b [2C];  m1_w [10C/3, 4C], m2_w [8C/3, 10C/3], m3_w [2C, 8C/3] (+ biases).
 * m1_w == NULL selects the merger-less variant (params = ctx + prior, pgm_coder.py:1634-1635).
 * ctx_w == NULL && m1_w == NULL: only ctx_b is used (mean-scale hyperprior, map "none"). */
int basic_ctx_create(int C, int G, int kernel_size, int device, basic_ctx **out);
void basic_ctx_destroy(basic_ctx *m);
int basic_ctx_set_weights(basic_ctx *m, const float *ctx_w, const float *ctx_b, const float *m1_w, const float *m1_b,
                          const float *m2_w, const float *m2_b, const float *m3_w, const float *m3_b);
/* The coder's INTERNAL context model (no topo_group_context_model: context_prediction + param_merger over 2G channel
 * groups, pgm_coder.py:1177-1239, :1606-1638).  The G "prior" groups carry id -1, so they see exactly each other: the
 * merger splits into an unmasked prior branch (prior -> p1 -> p2) and a context branch whose layers also read it.  The
 * caller passes the reference's matrices cut accordingly (half = bottleneck / 2; rows / columns of param_merger.{0,2,4}):
 *   m1_w = pm0[:half, :]  [half, 4C]     p1_w = pm0[half:, 2C:]   [half, 2C]
 *   m2_w = pm2[:half, :]  [half, 2half]  p2_w = pm2[half:, half:] [half, half]
 *   m3_w = pm4[:2C, :]    [2C, 2half]    (biases cut the same way).  Runs on the exact FP32 kernels. */
int basic_ctx_set_weights_internal(basic_ctx *m, const float *ctx_w, const float *ctx_b, const float *m1_w,
                                   const float *m1_b, const float *m2_w, const float *m2_b, const float *m3_w,
                                   const float *m3_b, const float *p1_w, const float *p1_b, const float *p2_w,
                                   const float *p2_b, int half);
/* Group map for the next calls: tg int32 [G, H, W] (host or device), already tiled to H x W
 * (pgm_coder.py:1299-1414).  Builds the per-stage cell and position lists on the device. */
int basic_ctx_set_map(basic_ctx *m, const int32_t *tg, int H, int W);
int basic_ctx_num_stages(basic_ctx *m);
/* Arithmetic of the context model.  BASIC_CTX_FP32: exact FP32 FMA kernel (the parity mode, default).
 * BASIC_CTX_TF32X3: tcgen05 tensor cores, every product split hi*hi + hi*lo + lo*hi (error-compensated TF32,
 * ~1e-6 relative; means / scales stay inside north_star's 1e-5); `nacc` = k-blocks (32 k each) accumulated in
 * tensor memory before the partial sum is drained into an FP32 register accumulator (shorter = more accurate).  Encoder and decoder must use the same setting (the parameters must be bit-identical on both
 * sides); layers the tensor path cannot take (tiny stages, channel groups not a multiple of 4) use FP32. */
#define BASIC_CTX_FP32 0
#define BASIC_CTX_TF32X3 1
#define BASIC_CTX_FP16X3 2
int basic_ctx_set_precision(basic_ctx *m, int precision, int nacc);
/* positions of stage g (device pointer, int32 offsets into [C,H,W]) and their count */
int basic_ctx_stage_positions(basic_ctx *m, int g, const int32_t **positions_dev, int64_t *n_pos);
/* One autoregressive step: distribution parameters of every cell of stage g, written into `params`
 * ([B, 2C, H, W] fp32, device; other cells untouched).  `buf` is the [B, C, H, W] y_hat buffer holding the
 * already coded stages, `prior` the [B, 2C, H, W] hyper-decoder output. */
int basic_ctx_stage_params(basic_ctx *m, int g, const float *buf, const float *prior, int B, float *params,
                           void *stream);

/* ---- the whole y-node path: TopoGroupPGMPriorCoder._encode_with_pgm / _pgm_generate -------------------- */
/* y, prior: fp32 [B,C,H,W] / [B,2C,H,W] host or device.  model may be NULL (params = prior).  The map must
 * have been set with basic_ctx_set_map (model != NULL) or is "none" (model == NULL).  yhat_out (optional,
 * host or device) receives the reconstruction the decoder will produce. */
int64_t basic_ypath_encode_bound(basic_coder *c, int B, int C, int H, int W, int lanes);
int basic_ypath_encode(basic_coder *c, basic_ctx *model, const float *y, const float *prior, int B, int C, int H, int W,
                       int lanes, uint8_t *out, int64_t out_cap, int64_t *out_len, float *yhat_out, void *stream);
int basic_ypath_decode(basic_coder *c, basic_ctx *model, const uint8_t *encoded, int64_t len, const float *prior, int B,
                       int C, int H, int W, int lanes, float *yhat_out, void *stream);

/* Phase timing for bench.py: with profiling enabled the y-path calls bracket their phases with CUDA events on the
 * launching stream; basic_profile_read returns (and clears) the accumulated milliseconds and span counts per
 * phase: [0] context model, [1] quantise / dequantise (+ layout copies), [2] multi-lane encode, [3] multi-lane
 * decode; arrays of 8. */
int basic_profile_enable(int on);
int basic_profile_read(double *ms, int64_t *spans);

/* Debug aid (tools/mma_bench.py): cycles[0] = issue, cycles[1] = issue + completion of `iters` back-to-back
 * tcgen05.mma (M = 128, N = n_cols; mode 0 = tf32 K 8, 1 = f16 K 16; ts = A operand in tensor memory) on one SM. */
int basic_debug_mma_bench(int mode, int ts, int n_cols, int iters, int same_acc, long long *cycles);

/* Counters for bench.py ("gpu_launches"): kernels launched by this library since the last reset. */
int64_t basic_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* BASIC_B200_H */

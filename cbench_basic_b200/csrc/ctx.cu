// Grouped / checkerboard autoregressive context model, exact-FP32 path.
//
// Reference semantics (cbench/nn/layers/masked_conv.py:102-228,287-305):
//   ctx    = conv5x5(y_hat)         tap (c, dh, dw) visible to out-group go iff tg[gc(c), h+dh, w+dw] <  tg[go, h, w]
//   m1     = conv1x1([ctx, prior])  ctx in-group j visible iff tg[j, h, w] <= tg[go, h, w]; prior always
//   m2     = conv1x1(lrelu(m1)), params = conv1x1(lrelu(m2)) with the same "<=" rule.
// The reference recomputes every position for every group; here each (channel-group, position) CELL is
// computed exactly once, at the stage its own group id says, and the intermediate activations stay in HBM
// for the later stages that are allowed to see them.  One launch per layer per stage: a tiled FP32 GEMM whose
// A operand is gathered on the fly (im2col rows of the stage's cells, masked taps skipped per tile).
// Deterministic (fixed K order, no atomics): the decoder recomputes bit-identical parameters.
#include <algorithm>

#include <cstdlib>

#include <cstdio>

#include "ctx.cuh"

namespace basic {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr float kSlope = 0.01f;  // nn.LeakyReLU default negative_slope

__global__ void __launch_bounds__(NT)
k_layer(LayerArgs a)
{
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Ws[BK][BN];
    __shared__ int s_off[BM];        // b * HW_total offset helper: b
    __shared__ int s_hw[BM];
    __shared__ uint32_t s_mask[BM];  // dense: group bits; conv: tap mask of the current input group
    __shared__ uint32_t s_or;

    const int tid = threadIdx.x;
    const int rows = a.B * a.ncells;
    const int row0 = blockIdx.x * BM;
    if (row0 >= rows) return;
    const int n0 = blockIdx.y * BN;  // relative to n_begin
    if (n0 >= a.n_count) return;

    for (int r = tid; r < BM; r += NT) {
        const int row = row0 + r;
        if (row < rows) {
            const int b = row / a.ncells, cell = a.cell_base + (row - b * a.ncells);
            s_off[r] = b;
            s_hw[r] = a.cell_hw[cell];
        } else {
            s_off[r] = -1;
            s_hw[r] = 0;
        }
    }
    const int tx = tid & 15, ty = tid >> 4;  // thread tile: rows ty*8..+7, cols tx*4..+3
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lr = tid & (BM - 1), lk0 = tid >> 7;  // A loader: row lr, k = lk0 + 2 * i  (8 loads)
    const int wn = tid & (BN - 1), wk0 = tid >> 6;  // W loader: col wn, k = wk0 + 4 * i  (4 loads)

    auto mma_tile = [&]() {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 w4 = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
    };

    if (a.is_conv) {
        const int kh = a.ksize, pad = kh / 2, cpg = a.Cin / a.G;
        const long long chw = (long long)a.Cin * a.HW;
        for (int g = 0; g < a.G; ++g) {
            __syncthreads();
            if (tid == 0) s_or = 0;
            __syncthreads();
            for (int r = tid; r < BM; r += NT) {
                const int row = row0 + r;
                uint32_t mk = 0;
                if (row < rows) {
                    const int b = row / a.ncells, cell = a.cell_base + (row - b * a.ncells);
                    mk = a.cell_tap[(size_t)cell * a.G + g];
                }
                s_mask[r] = mk;
                if (mk) atomicOr(&s_or, mk);
            }
            __syncthreads();
            const uint32_t tile_or = s_or;
            for (int tap = 0; tap < kh * kh; ++tap) {
                if (!((tile_or >> tap) & 1u)) continue;  // nobody in this tile sees the tap: skip its K block
                const int dh = tap / kh - pad, dw = tap % kh - pad;
                const int shift = dh * a.W_img + dw;
                for (int c0 = g * cpg; c0 < (g + 1) * cpg; c0 += BK) {
                    __syncthreads();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int kk = lk0 + 2 * i, c = c0 + kk;
                        float v = 0.f;
                        const int b = s_off[lr];
                        if (b >= 0 && c < (g + 1) * cpg && ((s_mask[lr] >> tap) & 1u))
                            v = a.src0.ptr[(long long)b * chw + (long long)c * a.HW + s_hw[lr] + shift];
                        As[kk][lr] = v;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kk = wk0 + 4 * i, c = c0 + kk, n = n0 + wn;
                        float v = 0.f;
                        if (c < (g + 1) * cpg && n < a.n_count)
                            v = a.wt[((size_t)tap * a.Cin + c) * a.Ntot + a.n_begin + n];
                        Ws[kk][wn] = v;
                    }
                    __syncthreads();
                    mma_tile();
                }
            }
        }
    } else {
        __syncthreads();
        if (tid == 0) s_or = 0;
        __syncthreads();
        for (int r = tid; r < BM; r += NT) {
            const int row = row0 + r;
            uint32_t mk = 0;
            if (row < rows) {
                const int b = row / a.ncells, cell = a.cell_base + (row - b * a.ncells);
                mk = a.cell_grp[cell];
            }
            s_mask[r] = mk;
            if (mk) atomicOr(&s_or, mk);
        }
        __syncthreads();
        const uint32_t tile_or = s_or;
        int kbase = 0;
        for (int si = 0; si < 2; ++si) {
            const Source s = si == 0 ? a.src0 : a.src1;
            if (!s.ptr || s.channels == 0) continue;
            const int ngroups = s.groups > 0 ? s.groups : 1;
            const int cpg = s.channels / ngroups;
            const long long chw = (long long)s.channels * a.HW;
            for (int g = 0; g < ngroups; ++g) {
                if (s.groups > 0 && !((tile_or >> g) & 1u)) continue;
                for (int c0 = g * cpg; c0 < (g + 1) * cpg; c0 += BK) {
                    __syncthreads();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int kk = lk0 + 2 * i, c = c0 + kk;
                        float v = 0.f;
                        const int b = s_off[lr];
                        if (b >= 0 && c < (g + 1) * cpg && (s.groups == 0 || ((s_mask[lr] >> g) & 1u)))
                            v = s.ptr[(long long)b * chw + (long long)c * a.HW + s_hw[lr]];
                        As[kk][lr] = v;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kk = wk0 + 4 * i, c = c0 + kk, n = n0 + wn;
                        float v = 0.f;
                        if (c < (g + 1) * cpg && n < a.n_count) v = a.wt[((size_t)kbase + c) * a.Ntot + a.n_begin + n];
                        Ws[kk][wn] = v;
                    }
                    __syncthreads();
                    mma_tile();
                }
            }
            kbase += s.channels;
        }
    }

    // epilogue: rows ty*8.., cols tx*4..
    const long long ohw = (long long)a.Ntot * a.HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= a.n_count) continue;
        const int ch = a.n_begin + n;
        const float bias = a.bias ? a.bias[ch] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = ty * 8 + i;
            const int b = s_off[r];
            if (b < 0) continue;
            const long long o = (long long)b * ohw + (long long)ch * a.HW + s_hw[r];
            float v = acc[i][j] + bias;
            if (a.add) v += a.add[o];
            if (a.lrelu) v = v > 0.f ? v : v * kSlope;
            a.out[o] = v;
        }
    }
}

// params = prior + bias (no context model weights at all, or a stage that sees nothing and has no merger)
__global__ void __launch_bounds__(256)
k_bias_prior(const float *__restrict__ prior, const float *__restrict__ bias, long long total, int HW, int C2,
             float *__restrict__ params)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)((e / HW) % C2);
        params[e] = prior[e] + (bias ? bias[ch] : 0.f);
    }
}

// Few-row stages (scanline-like maps: a stage is one position per image): the tiled kernel above would stream whole weight
// matrices through 128-row tiles that hold one row.  Here a CTA owns 32 output channels (lane = channel: the K-major
// weights are read in full 128-byte lines), gathers the <= kRowsMax masked input rows into shared memory once, and its 8
// warps split K (k = warp, warp + 8, ..., eight loads in flight each).  Deterministic: fixed k order per warp, the eight
// partial sums are added in warp order.
constexpr int kRowsMax = 4, kGemvWarps = 8, kGemvInFlight = 8;

__device__ __forceinline__ void layer_rows_cta(const LayerArgs &a, const int cta, float *sm, int *s_b, int *s_hw, int *s_cell,
                                               uint32_t *s_tapor_p)
{
    uint32_t &s_tapor = *s_tapor_p;
    const int rows = a.B * a.ncells;          // <= kRowsMax
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = cta * 32 + lane;            // relative to n_begin
    // total K in weight order: conv = taps x Cin, dense = src0 channels then src1 channels
    const int K0 = a.is_conv ? a.ksize * a.ksize * a.Cin : a.src0.channels;
    const int K = K0 + (a.is_conv || !a.src1.ptr ? 0 : a.src1.channels);
    float *A = sm;                            // [rows][K]
    float *part = sm + (size_t)rows * K;      // [kGemvWarps][kRowsMax][32]
    if (tid < kRowsMax) {
        const int row = tid;
        if (row < rows) {
            const int b = row / a.ncells, cell = a.cell_base + (row - b * a.ncells);
            s_b[row] = b; s_cell[row] = cell; s_hw[row] = a.cell_hw[cell];
        }
    }
    if (tid == 0) s_tapor = 0;
    __syncthreads();
    // ---- gather (masked elements = 0); loops are arranged so that nothing divides per element
    if (a.is_conv) {
        const int k2 = a.ksize * a.ksize, pad = a.ksize / 2, cpg = a.Cin / a.G;
        const long long chw = (long long)a.Cin * a.HW;
        for (int r = 0; r < rows; ++r) {
            const float *src = a.src0.ptr + (long long)s_b[r] * chw + s_hw[r];
            for (int tap = 0; tap < k2; ++tap) {
                const int shift = (tap / a.ksize - pad) * a.W_img + (tap % a.ksize - pad);
                uint32_t any = 0;
                for (int g = 0; g < a.G; ++g) {
                    const bool vis = (a.cell_tap[(size_t)s_cell[r] * a.G + g] >> tap) & 1u;
                    any |= vis;
                    float *dst = A + (size_t)r * K + (size_t)tap * a.Cin + g * cpg;
                    // (__ldcg: in the persistent stage kernel other CTAs wrote these during the same launch -- never through L1)
                    for (int c = tid; c < cpg; c += blockDim.x) dst[c] = vis ? __ldcg(src + (long long)(g * cpg + c) * a.HW + shift) : 0.f;
                }
                if (any && tid == 0) atomicOr(&s_tapor, 1u << tap);
            }
        }
    } else {
        for (int r = 0; r < rows; ++r) {
            const uint32_t grp = a.cell_grp[s_cell[r]];
            int kbase = 0;
            for (int si = 0; si < 2; ++si) {
                const Source sc = si == 0 ? a.src0 : a.src1;
                if (!sc.ptr || sc.channels == 0) continue;
                const int ng = sc.groups > 0 ? sc.groups : 1, cpg = sc.channels / ng;
                const float *src = sc.ptr + (long long)s_b[r] * sc.channels * a.HW + s_hw[r];
                for (int g = 0; g < ng; ++g) {
                    const bool vis = sc.groups == 0 || ((grp >> g) & 1u);
                    float *dst = A + (size_t)r * K + kbase + g * cpg;
                    for (int c = tid; c < cpg; c += blockDim.x) dst[c] = vis ? __ldcg(src + (long long)(g * cpg + c) * a.HW) : 0.f;
                }
                kbase += sc.channels;
            }
        }
    }
    __syncthreads();
    // ---- this warp's share of K, segment by segment (conv: one segment per tap, dead taps skipped; dense: one segment)
    float acc[kRowsMax];
#pragma unroll
    for (int r = 0; r < kRowsMax; ++r) acc[r] = 0.f;
    const bool n_ok = n < a.n_count;
    const float *wcol = a.wt + a.n_begin + (n_ok ? n : 0);
    const uint32_t tapor = a.is_conv ? s_tapor : 1u;
    const int nseg = a.is_conv ? a.ksize * a.ksize : 1, seglen = a.is_conv ? a.Cin : K;
    for (int seg = 0; seg < nseg; ++seg) {
        if (!((tapor >> seg) & 1u)) continue;
        const int kb = seg * seglen;
        // kGemvInFlight weight loads (128-byte lines from L2) in flight per warp: with eight the 12 - 20 CTAs of a layer drew
        // ~25 GB/s each; the order of the additions (ascending k per warp) does not depend on the batch size
        for (int c0 = warp; c0 < seglen; c0 += kGemvWarps * kGemvInFlight) {
            float wv[kGemvInFlight];
#pragma unroll
            for (int u = 0; u < kGemvInFlight; ++u) {
                const int c = c0 + u * kGemvWarps;
                wv[u] = c < seglen ? __ldg(wcol + (size_t)(kb + c) * a.Ntot) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kGemvInFlight; ++u) {
                const int c = c0 + u * kGemvWarps;
                if (c < seglen) {
#pragma unroll
                    for (int r = 0; r < kRowsMax; ++r)
                        if (r < rows) acc[r] = fmaf(A[(size_t)r * K + kb + c], wv[u], acc[r]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsMax; ++r) part[(warp * kRowsMax + r) * 32 + lane] = acc[r];
    __syncthreads();
    // ---- reduce in warp order + epilogue: thread = (row, channel)
    for (int o = tid; o < rows * 32; o += blockDim.x) {   // (rows * 32 <= 128: a whole warp per row, no divergence inside it)
        const int r = o >> 5, l = o & 31, nn = cta * 32 + l;
        const bool ok = nn < a.n_count;
        float v = 0.f;
        for (int w8 = 0; w8 < kGemvWarps; ++w8) v += part[(w8 * kRowsMax + r) * 32 + l];
        const int ch = a.n_begin + (ok ? nn : 0);
        const long long oo = ((long long)s_b[r] * a.Ntot + ch) * a.HW + s_hw[r];
        v += a.bias ? a.bias[ch] : 0.f;
        if (a.add && ok) v += a.add[oo];
        if (a.lrelu) v = v > 0.f ? v : v * kSlope;
        if (ok) a.out[oo] = v;
    }
}

__global__ void __launch_bounds__(kGemvWarps * 32)
k_layer_rows(LayerArgs a)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_b[kRowsMax], s_hw[kRowsMax], s_cell[kRowsMax];
    __shared__ uint32_t s_tapor;
    layer_rows_cta(a, blockIdx.x, sm, s_b, s_hw, s_cell, &s_tapor);
}

// out[n] = b[n] + sum_k w[n][k] * v[k], k < K0 of a row of K floats (state_dict layout), FP32 in ascending k
__global__ void k_fold_bias(const float *__restrict__ w, const float *__restrict__ b, const float *__restrict__ v, int N, int K, int K0,
                            float *__restrict__ out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = b[n];
    for (int k = 0; k < K0; ++k) acc = fmaf(w[(size_t)n * K + k], v[k], acc);
    out[n] = acc;
}

// [N][Cin][k2] (state_dict) -> [N][k2][Cin]: a tap's input channels are contiguous (dead taps are skipped as segments)
__global__ void k_conv_tap_major(const float *__restrict__ src, float *__restrict__ dst, int N, int Cin, int k2)
{
    const long long total = (long long)N * Cin * k2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e / ((long long)Cin * k2));
        const int rem = (int)(e - (long long)n * Cin * k2);
        const int c = rem / k2, tap = rem - c * k2;
        dst[((size_t)n * k2 + tap) * Cin + c] = src[e];
    }
}

// weight re-layout: src [N][K] (state_dict, conv flattened as c*k2+tap) -> dst K-major [K'][N]
__global__ void k_transpose_w(const float *__restrict__ src, float *__restrict__ dst, int N, int Cin, int k2)
{
    const long long total = (long long)N * Cin * k2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e / ((long long)Cin * k2));
        const int rem = (int)(e - (long long)n * Cin * k2);
        const int c = rem / k2, tap = rem - c * k2;
        dst[((size_t)tap * Cin + c) * N + n] = src[e];
    }
}

int upload(DevBuf &dst, const float *src, size_t count, cudaStream_t s)
{
    BASIC_TRY(dst.reserve(count * sizeof(float)));
    BASIC_CUDA(cudaMemcpyAsync(dst.p, src, count * sizeof(float), cudaMemcpyDefault, s));
    return BASIC_OK;
}

int upload_transposed(DevBuf &dst, const float *src, int N, int Cin, int k2, cudaStream_t s)
{
    const size_t count = (size_t)N * Cin * k2;
    DevBuf tmp;
    BASIC_TRY(tmp.reserve(count * sizeof(float)));
    BASIC_CUDA(cudaMemcpyAsync(tmp.p, src, count * sizeof(float), cudaMemcpyDefault, s));
    BASIC_TRY(dst.reserve(count * sizeof(float)));
    k_transpose_w<<<256, 256, 0, s>>>(tmp.as<float>(), dst.as<float>(), N, Cin, k2);
    BASIC_LAUNCHED();
    BASIC_CUDA(cudaStreamSynchronize(s));
    tmp.release();
    return BASIC_OK;
}

// stages `src` (host or device, state_dict layout) on the device and packs it for the tensor-core path
int pack_from(PackedW &dst, PackedW &dst16, const float *src, int N, int G, int is_conv, int Cin, int k2, int c_src0, int c_src1,
              cudaStream_t s)
{
    const size_t count = is_conv ? (size_t)N * Cin * k2 : (size_t)N * (c_src0 + c_src1);
    DevBuf tmp;
    BASIC_TRY(tmp.reserve(count * sizeof(float)));
    BASIC_CUDA(cudaMemcpyAsync(tmp.p, src, count * sizeof(float), cudaMemcpyDefault, s));
    int rc = pack_weights_tc(dst, tmp.as<float>(), N, G, is_conv, Cin, k2, c_src0, c_src1, 0, s);
    if (rc == BASIC_OK) rc = pack_weights_tc(dst16, tmp.as<float>(), N, G, is_conv, Cin, k2, c_src0, c_src1, 1, s);
    BASIC_CUDA(cudaStreamSynchronize(s));
    tmp.release();
    return rc;
}

}  // namespace

int ctx_set_weights(CtxModel &m, const float *ctx_w, const float *ctx_b, const float *m1_w, const float *m1_b,
                    const float *m2_w, const float *m2_b, const float *m3_w, const float *m3_b)
{
    cudaStream_t s = 0;
    const int C = m.C;
    m.kb_count.clear();
    m.c_ctx = 2 * C;
    m.c_m1 = m.c_ctx * 5 / 3;
    m.c_m2 = m.c_ctx * 4 / 3;
    m.has_conv = ctx_w != nullptr;
    m.has_merger = m1_w != nullptr;
    m.internal = false;
    if (m.has_merger && !(m1_b && m2_w && m2_b && m3_w && m3_b)) return value_error("merger weights incomplete");
    if (m.has_merger && !m.has_conv) return value_error("param merger needs the context convolution weights");
    if (m.has_merger && (m.c_m1 % m.G || m.c_m2 % m.G)) return value_error("2C*5/3 and 2C*4/3 must be divisible by channel_groups");
    if (C % m.G) return value_error("in_channels must be divisible by channel_groups");
    if (m.has_conv) {
        BASIC_TRY(upload_transposed(m.w_ctx, ctx_w, m.c_ctx, C, m.k * m.k, s));
        BASIC_TRY(pack_from(m.p_ctx, m.q_ctx, ctx_w, m.c_ctx, m.G, 1, C, m.k * m.k, 0, 0, s));
    }
    if (ctx_b) BASIC_TRY(upload(m.b_ctx, ctx_b, m.c_ctx, s)); else m.b_ctx.release();
    if (m.has_merger) {
        BASIC_TRY(upload_transposed(m.w_m1, m1_w, m.c_m1, 2 * m.c_ctx, 1, s));
        BASIC_TRY(pack_from(m.p_m1, m.q_m1, m1_w, m.c_m1, m.G, 0, 0, 1, m.c_ctx, m.c_ctx, s));
        BASIC_TRY(pack_from(m.p_m2, m.q_m2, m2_w, m.c_m2, m.G, 0, 0, 1, m.c_m1, 0, s));
        BASIC_TRY(pack_from(m.p_m3, m.q_m3, m3_w, m.c_ctx, m.G, 0, 0, 1, m.c_m2, 0, s));
        BASIC_TRY(upload(m.b_m1, m1_b, m.c_m1, s));
        BASIC_TRY(upload_transposed(m.w_m2, m2_w, m.c_m2, m.c_m1, 1, s));
        BASIC_TRY(upload(m.b_m2, m2_b, m.c_m2, s));
        BASIC_TRY(upload_transposed(m.w_m3, m3_w, m.c_ctx, m.c_m2, 1, s));
        BASIC_TRY(upload(m.b_m3, m3_b, m.c_ctx, s));
    }
    m.ws_ctx.release();
    m.ws_ctxc_key = 0xffffffffu;
    if (m.has_conv && m.has_merger && m.G == 1) {   // N-major copies for the persistent stage kernel (k_scan_stages)
        const size_t n_ctx = (size_t)m.c_ctx * C * m.k * m.k;
        DevBuf tmp;
        BASIC_TRY(tmp.reserve(n_ctx * sizeof(float)));
        BASIC_CUDA(cudaMemcpyAsync(tmp.p, ctx_w, n_ctx * sizeof(float), cudaMemcpyDefault, s));
        BASIC_TRY(m.ws_ctx.reserve(n_ctx * sizeof(float)));
        k_conv_tap_major<<<256, 256, 0, s>>>(tmp.as<float>(), m.ws_ctx.as<float>(), m.c_ctx, C, m.k * m.k);
        BASIC_LAUNCHED();
        BASIC_CUDA(cudaStreamSynchronize(s));
        tmp.release();
        BASIC_TRY(upload(m.ws_m1, m1_w, (size_t)m.c_m1 * 2 * m.c_ctx, s));
        // cells that see no neighbour (stage 0) have ctx == the convolution's bias: that half of the first merger layer is a
        // constant, folded into its bias (tensor path: the stage skips the convolution launch and half of the layer's K)
        BASIC_TRY(m.b_m1_fold.reserve((size_t)m.c_m1 * sizeof(float)));
        if (ctx_b) {
            k_fold_bias<<<(m.c_m1 + 127) / 128, 128, 0, s>>>(m.ws_m1.as<float>(), m.b_m1.as<float>(), m.b_ctx.as<float>(), m.c_m1,
                                                             2 * m.c_ctx, m.c_ctx, m.b_m1_fold.as<float>());
            BASIC_LAUNCHED();
        } else {
            BASIC_CUDA(cudaMemcpyAsync(m.b_m1_fold.p, m.b_m1.p, (size_t)m.c_m1 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        BASIC_TRY(upload(m.ws_m2, m2_w, (size_t)m.c_m2 * m.c_m1, s));
        BASIC_TRY(upload(m.ws_m3, m3_w, (size_t)m.c_ctx * m.c_m2, s));
    }
    BASIC_CUDA(cudaStreamSynchronize(s));
    return BASIC_OK;
}

// The coder's internal context model (pgm_coder.py:1177-1239, _merge_prior_params :1606-1638): masked convolutions over 2G
// channel groups -- G context groups carrying the map's ids and G prior groups with id -1 -- with the <= rule in every
// layer, all 2G out-groups computed, the first G kept at the end.  With id -1 a prior out-group sees exactly the prior
// in-groups, everywhere: the prior branch is a plain unmasked chain prior -> p1 -> p2 and the context branch is the usual
// three layers whose second source is (prior | p1 | p2).  The caller passes the matrices already cut that way:
//   m1_w [half, 4C] = rows of the context out-groups of layer 0 (columns: ctx 2C | prior 2C),   p1_w [half, 2C]
//   m2_w [half, 2*half] (columns: context branch | prior branch),                                p2_w [half, half]
//   m3_w [2C, 2*half]                                   with half = bottleneck / 2 (2C, or 4C with the expanded bottleneck)
int ctx_set_weights_internal(CtxModel &m, const float *ctx_w, const float *ctx_b, const float *m1_w, const float *m1_b,
                             const float *m2_w, const float *m2_b, const float *m3_w, const float *m3_b, const float *p1_w,
                             const float *p1_b, const float *p2_w, const float *p2_b, int half)
{
    cudaStream_t s = 0;
    const int C = m.C;
    if (!(ctx_w && ctx_b && m1_w && m1_b && m2_w && m2_b && m3_w && m3_b && p1_w && p1_b && p2_w && p2_b))
        return value_error("internal merger weights incomplete");
    if (half < 1 || half % m.G || C % m.G) return value_error("channel counts must be divisible by channel_groups");
    m.kb_count.clear();
    m.c_ctx = 2 * C;
    m.c_m1 = m.c_m2 = m.c_p = half;
    m.has_conv = m.has_merger = m.internal = true;
    m.act_B = 0;
    BASIC_TRY(upload_transposed(m.w_ctx, ctx_w, m.c_ctx, C, m.k * m.k, s));
    BASIC_TRY(upload(m.b_ctx, ctx_b, m.c_ctx, s));
    BASIC_TRY(upload_transposed(m.w_m1, m1_w, half, 2 * m.c_ctx, 1, s));
    BASIC_TRY(upload(m.b_m1, m1_b, half, s));
    BASIC_TRY(upload_transposed(m.w_m2, m2_w, half, 2 * half, 1, s));
    BASIC_TRY(upload(m.b_m2, m2_b, half, s));
    BASIC_TRY(upload_transposed(m.w_m3, m3_w, m.c_ctx, 2 * half, 1, s));
    BASIC_TRY(upload(m.b_m3, m3_b, m.c_ctx, s));
    BASIC_TRY(upload_transposed(m.w_p1, p1_w, half, m.c_ctx, 1, s));
    BASIC_TRY(upload(m.b_p1, p1_b, half, s));
    BASIC_TRY(upload_transposed(m.w_p2, p2_w, half, half, 1, s));
    BASIC_TRY(upload(m.b_p2, p2_b, half, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    return BASIC_OK;
}

// Builds the per-stage cell lists, visibility masks and coded-position lists on the host (the map is tiny:
// G*H*W integers) and uploads them.
int ctx_set_map(CtxModel &m, const int32_t *tg_any, int H, int W)
{
    const int G = m.G, HW = H * W, C = m.C, cpg = C / G, k = m.k, pad = k / 2;
    std::vector<int32_t> tg((size_t)G * HW);
    BASIC_CUDA(cudaMemcpy(tg.data(), tg_any, tg.size() * sizeof(int32_t), cudaMemcpyDefault));
    if (m.H == H && m.W == W && m.h_tg == tg && !m.stages.empty()) return BASIC_OK;  // same map as last time
    if (G > 32) return value_error("channel_groups > 32 not supported");
    int S = 0;
    for (int32_t v : tg) {
        if (v < 0) return value_error("negative topo group id");
        S = std::max(S, v + 1);
    }
    m.H = H; m.W = W; m.S = S; m.h_tg = tg;
    m.kb_count.clear();  // the k-block lists of the tensor path depend on the map
    m.stages.assign(S, CtxModel::Stage());
    for (auto &st : m.stages) {
        st.og_tap_or.assign((size_t)G * G, 0u);
        st.og_grp_or.assign((size_t)G, 0u);
    }
    // bucket cells by stage, ordered by (out-group, hw)
    std::vector<std::vector<int>> per_stage_count(S, std::vector<int>(G, 0));
    for (int g = 0; g < G; ++g)
        for (int p = 0; p < HW; ++p) per_stage_count[tg[(size_t)g * HW + p]][g]++;
    size_t cells_total = 0, pos_total = 0;
    for (int s = 0; s < S; ++s) {
        auto &st = m.stages[s];
        st.cell_off.assign(G + 1, 0);
        for (int g = 0; g < G; ++g) st.cell_off[g + 1] = st.cell_off[g] + per_stage_count[s][g];
        st.cells_at = cells_total;
        st.pos_at = pos_total;
        st.n_pos = (int64_t)st.cell_off[G] * cpg;
        cells_total += (size_t)st.cell_off[G];
        pos_total += (size_t)st.n_pos;
    }
    std::vector<int32_t> cell_hw(cells_total), positions(pos_total);
    std::vector<uint32_t> cell_tap(cells_total * G), cell_grp(cells_total);
    std::vector<int> fill(S * G, 0);
    for (int g = 0; g < G; ++g)
        for (int p = 0; p < HW; ++p) {
            const int s = tg[(size_t)g * HW + p];
            auto &st = m.stages[s];
            const size_t cell = st.cells_at + st.cell_off[g] + fill[s * G + g]++;
            cell_hw[cell] = p;
            const int h = p / W, w = p % W;
            uint32_t grp = 0;
            for (int j = 0; j < G; ++j) {
                if (tg[(size_t)j * HW + p] <= s) grp |= 1u << j;
                uint32_t taps = 0;
                for (int t = 0; t < k * k; ++t) {
                    const int hh = h + t / k - pad, ww = w + t % k - pad;
                    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                    if (tg[(size_t)j * HW + hh * W + ww] < s) taps |= 1u << t;
                }
                cell_tap[cell * G + j] = taps;
                st.tap_or |= taps;
                st.og_tap_or[(size_t)g * G + j] |= taps;
            }
            cell_grp[cell] = grp;
            st.og_grp_or[g] |= grp;
        }
    // coded positions: stage-major, then (c, hw) row-major inside an image == boolean-mask order
    for (int s = 0; s < S; ++s) {
        auto &st = m.stages[s];
        size_t at = st.pos_at;
        for (int g = 0; g < G; ++g)
            for (int c = g * cpg; c < (g + 1) * cpg; ++c)
                for (int i = st.cell_off[g]; i < st.cell_off[g + 1]; ++i)
                    positions[at++] = c * HW + cell_hw[st.cells_at + i];
    }
    {   // per stage: first cell and cell count (the persistent stage kernel of many-stage maps, one channel group)
        std::vector<int2> sc((size_t)S);
        m.max_stage_cells = 0;
        for (int s2 = 0; s2 < S; ++s2) {
            sc[(size_t)s2] = make_int2((int)m.stages[(size_t)s2].cells_at, m.stages[(size_t)s2].cell_off[G]);
            m.max_stage_cells = std::max(m.max_stage_cells, m.stages[(size_t)s2].cell_off[G]);
        }
        BASIC_TRY(m.d_stage_cells.reserve(sc.size() * sizeof(int2) + 16));
        BASIC_CUDA(cudaMemcpy(m.d_stage_cells.p, sc.data(), sc.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    BASIC_TRY(m.d_cell_hw.reserve(cell_hw.size() * 4 + 16));
    BASIC_TRY(m.d_cell_tap.reserve(cell_tap.size() * 4 + 16));
    BASIC_TRY(m.d_cell_grp.reserve(cell_grp.size() * 4 + 16));
    BASIC_TRY(m.d_positions.reserve(positions.size() * 4 + 16));
    BASIC_CUDA(cudaMemcpy(m.d_cell_hw.p, cell_hw.data(), cell_hw.size() * 4, cudaMemcpyHostToDevice));
    BASIC_CUDA(cudaMemcpy(m.d_cell_tap.p, cell_tap.data(), cell_tap.size() * 4, cudaMemcpyHostToDevice));
    BASIC_CUDA(cudaMemcpy(m.d_cell_grp.p, cell_grp.data(), cell_grp.size() * 4, cudaMemcpyHostToDevice));
    BASIC_CUDA(cudaMemcpy(m.d_positions.p, positions.data(), positions.size() * 4, cudaMemcpyHostToDevice));
    // slot order of the tensor path's activation layout: stage-major by the coding group of channel group 0, raster inside
    std::vector<int32_t> perm((size_t)HW), iperm((size_t)HW);
    {
        static const bool identity = getenv("BASIC_TC_NOPERM") && atoi(getenv("BASIC_TC_NOPERM"));  // A/B aid: raster slots
        int32_t slot = 0;
        for (int s = 0; s < S; ++s)
            for (int p = 0; p < HW; ++p)
                if (tg[p] == s) { perm[p] = slot; iperm[slot] = p; ++slot; }
        if (identity) for (int p = 0; p < HW; ++p) perm[p] = iperm[p] = p;
    }
    BASIC_TRY(m.d_perm.reserve((size_t)HW * 4 + 16));
    BASIC_TRY(m.d_iperm.reserve((size_t)HW * 4 + 16));
    BASIC_CUDA(cudaMemcpy(m.d_perm.p, perm.data(), (size_t)HW * 4, cudaMemcpyHostToDevice));
    BASIC_CUDA(cudaMemcpy(m.d_iperm.p, iperm.data(), (size_t)HW * 4, cudaMemcpyHostToDevice));
    return BASIC_OK;
}

static int launch_layer(CtxModel &m, LayerArgs a, const PackedW &pw, const PackedW &pw16, int og, const CtxModel::Stage &st, bool tc,
                        cudaStream_t stream)
{
    for (int j = 0; j < 8; ++j) a.vis_or[j] = 0;
    if (m.G <= 8) {
        if (a.is_conv) for (int j = 0; j < m.G; ++j) a.vis_or[j] = st.og_tap_or[(size_t)og * m.G + j];
        else a.vis_or[0] = st.og_grp_or[og];
    }
    const int rows = a.B * a.ncells;
    if (rows == 0 || a.n_count == 0) return BASIC_OK;
    const bool f16 = tc && m.run_precision == BASIC_CTX_FP16X3;
    const PackedW &pk = f16 ? pw16 : pw;
    a.wpack = pk.buf.as<unsigned char>();
    a.kb_total = pk.kb_total;
    a.kb_src0 = pk.kb_src0;
    a.ntile_base = og * pk.ntiles_per_group;
    a.nacc = m.nacc;
    a.mode = f16 ? 1 : 0;
    a.out_scale = f16 ? 1.f / (16.f * pk.scale) : 1.f;  // powers of two: exact
    a.range_flag = f16 ? m.range_flag.as<int>() : nullptr;
    if (tc) return launch_layer_tc(m, a, stream);
    if (rows <= kRowsMax) {
        const int K = a.is_conv ? a.ksize * a.ksize * a.Cin : a.src0.channels + (a.src1.ptr ? a.src1.channels : 0);
        const size_t smem = ((size_t)rows * K + (size_t)kGemvWarps * kRowsMax * 32) * sizeof(float);
        if (smem <= 200 * 1024) {
            static PerDeviceOnce attr_once;
            if (attr_once.first()) {
                BASIC_CUDA(cudaFuncSetAttribute(k_layer_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            }
            k_layer_rows<<<(a.n_count + 31) / 32, kGemvWarps * 32, smem, stream>>>(a);
            BASIC_LAUNCHED();
            return BASIC_OK;
        }
    }
    dim3 grid((rows + BM - 1) / BM, (a.n_count + BN - 1) / BN);
    k_layer<<<grid, NT, 0, stream>>>(a);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

bool ctx_uses_tc(const CtxModel &m, int B) { return tc_model_eligible(m, B); }
int ctx_precision(const CtxModel &m) { return m.precision; }
int ctx_run_precision(const CtxModel &m) { return m.run_precision; }
const int32_t *ctx_perm(const CtxModel &m) { return m.d_perm.as<int32_t>(); }
// blocked channels-last copy in the operand format of the mode the next stage calls run in (floats, or split16)
int ctx_to_cl(CtxModel &m, const float *src, float *dst, int B, int channels, cudaStream_t s)
{
    const bool f16 = m.run_precision == BASIC_CTX_FP16X3;
    if (f16) BASIC_TRY(m.range_flag.reserve(16));
    return launch_nchw_to_cl(src, dst, B, channels, m.H * m.W, m.d_iperm.as<int32_t>(), s, f16 ? 1 : 0,
                             f16 ? m.range_flag.as<int>() : nullptr);
}
void ctx_set_run_precision(CtxModel &m, int p)
{
    if (p == BASIC_CTX_FP16X3 && !tc_fp16_ok(m)) p = BASIC_CTX_TF32X3;  // (channel counts not in 8-channel chunk pairs)
    if (m.run_precision != p) m.act_B = 0;  // (FP32 and the tensor modes keep their activations in different layouts)
    m.run_precision = p;
}
// FP16X3 range flag: cleared before a pass, read (with a stream sync) after it
int ctx_range_flag_clear(CtxModel &m, cudaStream_t s)
{
    BASIC_TRY(m.range_flag.reserve(16));
    BASIC_CUDA(cudaMemsetAsync(m.range_flag.p, 0, 16, s));
    return BASIC_OK;
}
int ctx_range_flag_read(CtxModel &m, cudaStream_t s, int *flag)
{
    *flag = 0;
    if (!m.range_flag.p) return BASIC_OK;
    BASIC_CUDA(cudaMemcpyAsync(flag, m.range_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    return BASIC_OK;
}
// Queues the copy of the flag into pinned host memory; the value is valid after the caller's next synchronisation of `s`.
int ctx_range_flag_copy(CtxModel &m, cudaStream_t s, int *pinned_flag)
{
    *pinned_flag = 0;
    if (!m.range_flag.p) return BASIC_OK;
    BASIC_CUDA(cudaMemcpyAsync(pinned_flag, m.range_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    return BASIC_OK;
}
size_t ctx_cl_elems(int B, int channels, int HW) { return cl_elems(B, channels, HW); }

// One autoregressive step (see include/basic_b200.h basic_ctx_stage_params).  buf / prior are NCHW; the tensor path
// reads channels-last copies: the caller's (buf_cl / prior_cl, kept up to date by the y-path driver) or, when they
// are NULL (the public stage API), copies made here.
int ctx_stage_params(CtxModel &m, int g, const float *buf, const float *prior, int B, float *params, cudaStream_t stream,
                     const float *buf_cl, const float *prior_cl, bool params_cl)
{
    if (g < 0 || g >= m.S) return value_error("stage out of range");
    const int HW = m.H * m.W, G = m.G;
    const auto &st = m.stages[g];
    // no context weights, or a merger-less model whose single stage sees no neighbour (map "none": conv = bias):
    // params = prior + bias, elementwise over the whole tensor, only needed once (stage 0)
    if (!m.has_conv || (!m.has_merger && m.S == 1 && st.tap_or == 0)) {
        if (g == 0) {
            const long long total = (long long)B * m.c_ctx * HW;
            k_bias_prior<<<m.sm_count * 8, 256, 0, stream>>>(prior, m.b_ctx.p ? m.b_ctx.as<float>() : nullptr, total, HW,
                                                            m.c_ctx, params);
            BASIC_LAUNCHED();
        }
        return BASIC_OK;
    }
    const bool tc = tc_model_eligible(m, B);
    if (tc && !m.range_flag.p) {
        BASIC_TRY(m.range_flag.reserve(16));
        BASIC_CUDA(cudaMemsetAsync(m.range_flag.p, 0, 16, stream));
    }
    if (m.act_B < B || !(tc ? m.cl_ctx.p : m.a_ctx.p)) {
        DevBuf &x0 = tc ? m.cl_ctx : m.a_ctx, &x1 = tc ? m.cl_m1 : m.a_m1, &x2 = tc ? m.cl_m2 : m.a_m2;
        BASIC_TRY(x0.reserve(cl_elems(B, m.c_ctx, HW) * sizeof(float)));  // (the padded size also covers NCHW)
        if (m.has_merger) {
            BASIC_TRY(x1.reserve(cl_elems(B, m.c_m1, HW) * sizeof(float)));
            BASIC_TRY(x2.reserve(cl_elems(B, m.c_m2, HW) * sizeof(float)));
        }
        if (m.internal) {
            BASIC_TRY(m.a_p1.reserve(cl_elems(B, m.c_p, HW) * sizeof(float)));
            BASIC_TRY(m.a_p2.reserve(cl_elems(B, m.c_p, HW) * sizeof(float)));
        }
        m.act_B = B;
    }
    const int f16 = tc && m.run_precision == BASIC_CTX_FP16X3;
    if (tc && !buf_cl) {
        BASIC_TRY(m.cl_buf.reserve(cl_elems(B, m.C, HW) * sizeof(float)));
        BASIC_TRY(launch_nchw_to_cl(buf, m.cl_buf.as<float>(), B, m.C, HW, m.d_iperm.as<int32_t>(), stream, f16, m.range_flag.as<int>()));
        buf_cl = m.cl_buf.as<float>();
    }
    if (tc && !prior_cl && (m.has_merger || true)) {
        BASIC_TRY(m.cl_prior.reserve(cl_elems(B, m.c_ctx, HW) * sizeof(float)));
        BASIC_TRY(launch_nchw_to_cl(prior, m.cl_prior.as<float>(), B, m.c_ctx, HW, m.d_iperm.as<int32_t>(), stream, f16, m.range_flag.as<int>()));
        prior_cl = m.cl_prior.as<float>();
    }
    // tensor path: the parameters leave the last layer in blocked channels-last floats (full-line stores); the y-path
    // driver reads them like that (params_cl), the public stage API gets an NCHW copy
    float *params_out = params;
    if (tc && !params_cl) {
        const size_t need = cl_elems(B, m.c_ctx, HW) * sizeof(float);
        if (m.cl_params.cap < need) {
            BASIC_TRY(m.cl_params.reserve(need));
            BASIC_CUDA(cudaMemsetAsync(m.cl_params.p, 0, need, stream));
        }
        params_out = m.cl_params.as<float>();
    }
    float *act0 = tc ? m.cl_ctx.as<float>() : m.a_ctx.as<float>();
    float *act1 = tc ? m.cl_m1.as<float>() : m.a_m1.as<float>();
    float *act2 = tc ? m.cl_m2.as<float>() : m.a_m2.as<float>();
    const int cl = tc ? 1 : 0;
    auto base_args = [&](int og, int ncells) {
        LayerArgs a = {};
        a.cell_hw = m.d_cell_hw.as<int32_t>();
        a.perm = m.d_perm.as<int32_t>();
        a.cell_tap = m.d_cell_tap.as<uint32_t>();
        a.cell_grp = m.d_cell_grp.as<uint32_t>();
        a.ncells = ncells;
        a.cell_base = (int)(st.cells_at + st.cell_off[og]);
        a.B = B; a.HW = HW; a.W_img = m.W; a.H_img = m.H; a.G = G;
        return a;
    };
    static const bool fold_off = [] { const char *e = getenv("BASIC_CTX_FOLD"); return e && e[0] == '0'; }();  // A/B switch
    const bool fold0 = tc && G == 1 && m.has_merger && st.tap_or == 0 && m.b_m1_fold.p && !fold_off;
    for (int og = 0; og < G && !fold0; ++og) {
        const int ncells = st.cell_off[og + 1] - st.cell_off[og];
        if (ncells == 0) continue;
        LayerArgs a = base_args(og, ncells);
        a.is_conv = 1; a.ksize = m.k; a.Cin = m.C;
        a.src0 = Source{tc ? buf_cl : buf, m.C, G, cl};
        a.src1 = Source{nullptr, 0, 0, 0};
        a.wt = m.w_ctx.as<float>(); a.bias = m.b_ctx.as<float>();
        a.Ntot = m.c_ctx; a.n_begin = og * (m.c_ctx / G); a.n_count = m.c_ctx / G;
        if (m.has_merger) {
            a.out = act0; a.out_cl = cl; a.add = nullptr;
        } else {  // params = ctx + prior: written straight to the NCHW parameter tensor
            a.out = params_out; a.out_cl = cl; a.out_f32 = 1; a.add = prior;
        }
        a.lrelu = 0;
        a.tap_or = st.tap_or;
        a.list_key = (g * G + og) * 4;
        BASIC_TRY(launch_layer(m, a, m.p_ctx, m.q_ctx, og, st, tc, stream));
    }
    auto finish = [&]() -> int {  // public stage API on the tensor path: NCHW copy of the parameters
        if (tc && !params_cl) return launch_cl_to_nchw(m.cl_params.as<float>(), params, B, m.c_ctx, HW, m.d_iperm.as<int32_t>(), stream);
        return BASIC_OK;
    };
    if (!m.has_merger) return finish();
    // internal merger: the prior branch at the positions of this stage's cells (unmasked 1x1 layers over the prior; a
    // position shared by several channel groups of the stage is simply computed once per group)
    if (m.internal) {
        for (int pl = 0; pl < 2; ++pl)
            for (int og = 0; og < G; ++og) {
                const int ncells = st.cell_off[og + 1] - st.cell_off[og];
                if (ncells == 0) continue;
                LayerArgs a = base_args(og, ncells);
                a.is_conv = 0;
                a.src0 = pl == 0 ? Source{prior, m.c_ctx, 0, 0} : Source{m.a_p1.as<float>(), m.c_p, 0, 0};
                a.src1 = Source{nullptr, 0, 0, 0};
                a.wt = pl == 0 ? m.w_p1.as<float>() : m.w_p2.as<float>();
                a.bias = pl == 0 ? m.b_p1.as<float>() : m.b_p2.as<float>();
                a.Ntot = m.c_p; a.n_begin = 0; a.n_count = m.c_p;
                a.out = pl == 0 ? m.a_p1.as<float>() : m.a_p2.as<float>();
                a.out_cl = 0; a.lrelu = 1;
                a.list_key = 0;
                BASIC_TRY(launch_layer(m, a, m.p_m1, m.q_m1, og, st, false, stream));
            }
    }
    // the three 1x1 layers; within a stage, layer L+1 of a cell may read layer L of ANOTHER channel group of the
    // same stage at the same position, hence one launch wave per layer
    for (int layer = 1; layer <= 3; ++layer) {
        for (int og = 0; og < G; ++og) {
            const int ncells = st.cell_off[og + 1] - st.cell_off[og];
            if (ncells == 0) continue;
            LayerArgs a = base_args(og, ncells);
            a.is_conv = 0;
            if (layer == 1) {
                a.src0 = Source{act0, m.c_ctx, G, cl};
                a.src1 = Source{tc ? prior_cl : prior, m.c_ctx, 0, cl};
                a.wt = m.w_m1.as<float>(); a.bias = fold0 ? m.b_m1_fold.as<float>() : m.b_m1.as<float>();
                a.fold_src0 = fold0 ? 1 : 0;
                a.Ntot = m.c_m1; a.out = act1; a.out_cl = cl; a.lrelu = 1;
            } else if (layer == 2) {
                a.src0 = Source{act1, m.c_m1, G, cl};
                if (m.internal) a.src1 = Source{m.a_p1.as<float>(), m.c_p, 0, 0};
                a.wt = m.w_m2.as<float>(); a.bias = m.b_m2.as<float>();
                a.Ntot = m.c_m2; a.out = act2; a.out_cl = cl; a.lrelu = 1;
            } else {
                a.src0 = Source{act2, m.c_m2, G, cl};
                if (m.internal) a.src1 = Source{m.a_p2.as<float>(), m.c_p, 0, 0};
                a.wt = m.w_m3.as<float>(); a.bias = m.b_m3.as<float>();
                a.Ntot = m.c_ctx; a.out = params_out; a.out_cl = cl; a.out_f32 = 1; a.lrelu = 0;
            }
            a.n_begin = og * (a.Ntot / G);
            a.n_count = a.Ntot / G;
            a.list_key = (g * G + og) * 4 + layer;
            BASIC_TRY(launch_layer(m, a, layer == 1 ? m.p_m1 : layer == 2 ? m.p_m2 : m.p_m3,
                                   layer == 1 ? m.q_m1 : layer == 2 ? m.q_m2 : m.q_m3, og, st, tc, stream));
        }
    }
    return finish();
}

CtxModel *ctx_new(int C, int G, int k, int device, int sm_count)
{
    CtxModel *m = new CtxModel();
    m->C = C; m->G = G; m->k = k; m->device = device; m->sm_count = sm_count;
    m->c_ctx = 2 * C;
    return m;
}

void ctx_delete(CtxModel *m)
{
    if (!m) return;
    DevBuf *bufs[] = {&m->w_ctx, &m->b_ctx, &m->w_m1, &m->b_m1, &m->w_m2, &m->b_m2, &m->w_m3, &m->b_m3, &m->d_cell_hw,
                      &m->d_cell_tap, &m->d_cell_grp, &m->d_positions, &m->a_ctx, &m->a_m1, &m->a_m2,
                      &m->p_ctx.buf, &m->p_m1.buf, &m->p_m2.buf, &m->p_m3.buf, &m->cl_ctx, &m->cl_m1, &m->cl_m2, &m->cl_buf,
                      &m->cl_prior, &m->q_ctx.buf, &m->q_m1.buf, &m->q_m2.buf, &m->q_m3.buf, &m->range_flag, &m->kb_pool, &m->d_perm, &m->d_iperm, &m->cl_params,
                      &m->w_p1, &m->b_p1, &m->w_p2, &m->b_p2, &m->a_p1, &m->a_p2, &m->d_stage_cells, &m->scan_barrier, &m->scan_ws, &m->scan_cs, &m->ws_ctxc, &m->ws_ctx, &m->ws_m1, &m->ws_m2, &m->ws_m3, &m->b_m1_fold};
    for (DevBuf *b : bufs) b->release();
    delete m;
}

int ctx_num_stages(const CtxModel &m) { return m.S; }

int ctx_set_precision(CtxModel &m, int precision, int nacc)
{
    if (precision != BASIC_CTX_FP32 && precision != BASIC_CTX_TF32X3 && precision != BASIC_CTX_FP16X3)
        return value_error("unknown context-model precision");
    if (nacc < 1 || nacc > 64) return value_error("segment length must be 1..64 k-blocks");
    if (m.precision != precision) m.act_B = 0;  // the two paths keep their activations in different layouts
    // 3xFP16 needs channel counts in 8-channel chunk pairs; otherwise the model runs (and says so: ctx_precision) in 3xTF32
    if (precision == BASIC_CTX_FP16X3 && m.c_ctx && !tc_fp16_ok(m)) precision = BASIC_CTX_TF32X3;
    m.precision = precision;
    m.run_precision = precision;
    m.nacc = nacc;
    return BASIC_OK;
}

int ctx_stage_positions(const CtxModel &m, int g, const int32_t **positions_dev, int64_t *n_pos)
{
    if (g < 0 || g >= m.S) return value_error("stage out of range");
    *positions_dev = m.d_positions.as<int32_t>() + m.stages[g].pos_at;
    *n_pos = m.stages[g].n_pos;
    return BASIC_OK;
}

int ctx_dims(const CtxModel &m, int *C, int *G, int *H, int *W)
{
    *C = m.C; *G = m.G; *H = m.H; *W = m.W;
    return BASIC_OK;
}

}  // namespace basic

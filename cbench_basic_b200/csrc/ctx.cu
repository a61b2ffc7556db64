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
#include "rans_lanes.cuh"

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

// What the last layer's epilogue does with its (mean, scale) pairs when the stage kernel also quantises (k_scan_stages):
// the arithmetic of gauss.cu's k_quantize_index / k_dequantize, element for element.
struct RowsQuant {
    const float *y;            // encoder: the latents [B, C, HW]; NULL = decoder (indexes only)
    float *buf;                // encoder: y_hat written back for the later stages
    int32_t *sym, *idx;        // this stage's slice of the stream: element (b, c, cell i) at b * (C * ncells) + c * ncells + i
    const float *scale_table;
    int n_scales, C;
};

__device__ inline int scale_index_dev(float sigma, const float *__restrict__ tab, int n)
{
    if (!(fabsf(sigma) <= 3.402823466e38f)) return 0;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] < sigma) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == n) return n - 1;
    const float d0 = fabsf(__fsub_rn(sigma, tab[lo - 1])), d1 = fabsf(__fsub_rn(sigma, tab[lo]));
    return d0 <= d1 ? lo - 1 : lo;
}

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

// ---- many-stage maps (scanline: one position per stage, 1536 stages for a Kodak-shape image; the serial JointAR coder):
// one PERSISTENT kernel walks the stages.  Per stage the four layers of the context model are matrix-vector products over
// a handful of rows (rows = batch x cells of the stage), so what a stage costs is latency, not arithmetic:
//  * every weight the kernel needs stays resident in shared memory for its whole life -- CTA c of the grid owns the
//    output-channel pairs c, c + grid, c + 2 grid, ... of every layer (N-major rows, the convolution only with the taps some
//    stage of the map can see): 113 KB at C = 192 on 148 CTAs;
//  * what CTAs exchange -- the layer outputs of the current stage and y_hat -- travels as {value, tag} words (8 bytes, written
//    and read as one access, so a matching tag IS the value's arrival: the "LL" protocol of collective libraries).  Layer
//    outputs carry a tag that counts (stage, layer) steps, y_hat the id of the coding call.  A consumer polls the words it
//    gathers until their tags match: no grid barrier, no fence, and a CTA never waits for more than the data it reads.  (A
//    grid barrier per layer cost 2.4 k cycles of the 5.4 k a layer took.)  Buffers are reused every stage; a producer cannot
//    overwrite a word a consumer still needs because its own next input depends on that consumer's output (a CTA that owns
//    no channel of a layer does not gather for it);
//  * vectors are contiguous (per-stage [row][N] outputs, position-major y_hat and prior): a gather is a few 128-bit loads
//    per thread, where the NCHW gather asked L2 for one sector per float from 148 CTAs at once.
// The multiply itself: a warp per channel pair, lanes striding over K, four partial sums per output, butterfly.  The last
// layer's pair is (mean, scale) of one latent channel, so the quantiser runs in its epilogue.  The encoder knows y: its
// whole pass is ONE launch; the decoder launches the kernel once per stage (the coder sits between two stages) and the
// launch first turns the previous stage's symbols into y_hat.  Deterministic (fixed summation order); one channel group (G = 1).
constexpr int kScanWarps = 16, kScanRows = 4, kScanMaxRows = 32;   // (launched for <= kScanRows rows per stage; more rows go to k_scan_blocks:
                                                                   // here every CTA would gather every row)

// The decoder's single launch: the multi-lane coder's chunk warps live inside the stage kernel.  Chunk k belongs to warp
// 7 - k / grid of CTA k % grid; it keeps its 32 lane states and its word position in registers from stage to stage, takes
// the scale indexes (and means) of its share of the stage's slice as tagged words from the CTAs that computed them, decodes,
// and publishes y_hat = symbol + mean as tagged words -- which is what the next stage's convolution waits for.
constexpr int kScanDecSlots = 4;   // warps 7 .. 4 of a CTA
struct ScanDecode {
    const unsigned char *blob;   // coder tables (rans tables blob in global memory, read through the read-only path)
    size_t blob_bytes, meta_bytes, cdf16_bytes;
    int T, precision, bypass;
    const unsigned char *seg;    // the segment (device)
    long long seg_cap;
    int seg_slices, n_chunks;
    const int32_t *chunk_syms;   // per slice
    uint4 *idx_t;                // [slice element] {scale index, tag, mean, tag} of the current stage
    int *status;
};

struct ScanArgs {
    const float *w[4];           // N-major: conv [2C][k2][C], dense [N][K]
    const float *bias[4];
    int N[4], K[4];              // K[0] = ntaps * C: the convolution's K is compact (only the taps of `taps`, in that order)
    int pairs[4];                // channel pairs a CTA owns per layer = ceil(N / 2 / gridDim.x) <= kScanWarps
    int ntaps;
    unsigned char taps[25];
    int shift[25];               // offset of tap taps[t] relative to the centre: dy * W + dx
    int C, ksize, HW, W_img, B;
    const int2 *stage_cells;     // per stage: first cell, cells
    const int32_t *cell_hw;
    const uint32_t *cell_tap, *cell_grp;
    float *buf;                  // y_hat [B, C, HW]
    uint2 *yhat_pm;              // ... and its position-major copy [B, HW, C] of {value, call tag}: a tap is C contiguous words
    const float *prior_pm;       // prior, position-major [B, HW, 2C]
    uint2 *vec[4];               // outputs of the four layers for the rows of the CURRENT stage, [row][N] of {value, step tag}
                                 // (one channel group: a layer only reads its own cell's previous layer)
    float *params;               // [B, 2C, HW]
    int g0, g1;                  // stages [g0, g1)
    RowsQuant qz;                // sym / idx point at the slice of stage g0
    const int32_t *dq_sym;       // decoder: the symbols of stage g0 - 1, dequantised into buf before anything else; else NULL
    uint32_t step0;              // tag of (stage g, layer L) = step0 + 4 (g - g0) + L + 1: monotonic over launches
    uint32_t call_tag;           // tag of every y_hat word of this coding call
    long long *timing;           // SCAN_TIMING builds
    ScanDecode dec;              // n_chunks > 0: the decoder's single launch
    int dctas;                   // k_scan_stages<DEC>: the last dctas CTAs of the grid only decode (tables in their shared memory)
    int CB;                      // k_scan_blocks: channel blocks per row block (grid = row blocks x CB)
    const float *wc;             // k_scan_blocks: the convolution's weights with only the visible taps, N-major [2C][ntaps * C]
};

// One 16-byte load of two {value, tag} words (L2, never L1).  A gather issues a batch of these and only then looks at the tags,
// re-reading the words that have not arrived yet: the loads of a batch overlap instead of costing an L2 round trip each.
__device__ __forceinline__ uint4 ll_ld(const uint2 *p)
{
    uint4 q;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    return q;
}

__device__ __forceinline__ float2 ll_wait(uint4 q, const uint2 *p, uint32_t tag)
{
    int spins = 0;
    while (q.y != tag || q.w != tag) {
        q = ll_ld(p);
        if (++spins > (1 << 24)) asm volatile("trap;");   // a producer died: fail the launch instead of hanging the GPU
    }
    return make_float2(__uint_as_float(q.x), __uint_as_float(q.z));
}

constexpr int kScanBatch = 5;   // 16-byte loads in flight per thread (12 taps x 96 channel pairs / 256 threads = 4.5)

// Two neighbouring output channels (weight rows w, w + K in shared memory) against R input vectors A[r][K], one of kScanSplit
// interleaved parts of K: lanes stride over the part in 128-bit steps, four partial sums per output (x, y, z, w components),
// then a butterfly; lane 0 leaves the 2 R sums in part[r][0 / 1].  The order of the additions depends on nothing but K.
constexpr int kScanSplit = 4;

template <int R>
__device__ __forceinline__ void scan_pair_part(const float *__restrict__ wrow, const float *__restrict__ A, int K, int q, int lane, float *part)
{
    const int K4 = K >> 2;
    const float4 *w0 = reinterpret_cast<const float4 *>(wrow), *w1 = w0 + K4;
    const float4 *A4 = reinterpret_cast<const float4 *>(A);
    float4 a0[R], a1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = q * 32 + lane; i < K4; i += 32 * kScanSplit) {
        const float4 x0 = w0[i], x1 = w1[i];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 av = A4[r * K4 + i];
            a0[r].x = fmaf(av.x, x0.x, a0[r].x); a0[r].y = fmaf(av.y, x0.y, a0[r].y);
            a0[r].z = fmaf(av.z, x0.z, a0[r].z); a0[r].w = fmaf(av.w, x0.w, a0[r].w);
            a1[r].x = fmaf(av.x, x1.x, a1[r].x); a1[r].y = fmaf(av.y, x1.y, a1[r].y);
            a1[r].z = fmaf(av.z, x1.z, a1[r].z); a1[r].w = fmaf(av.w, x1.w, a1[r].w);
        }
    }
    float v0[R], v1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        v0[r] = (a0[r].x + a0[r].y) + (a0[r].z + a0[r].w);
        v1[r] = (a1[r].x + a1[r].y) + (a1[r].z + a1[r].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v0[r] += __shfl_xor_sync(0xffffffffu, v0[r], o);
            v1[r] += __shfl_xor_sync(0xffffffffu, v1[r], o);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) { part[2 * r] = v0[r]; part[2 * r + 1] = v1[r]; }
    }
}

// One chunk's share [dbase, dbase + m) of the current stage's slice, decoded by its warp: the coding steps of k_bls_decode
// (rans_lanes.cu: local symbol j -> lane (j % 128) / 4, step (j / 128) * 4 + j % 4; renormalising lanes take consecutive words in
// lane order; bypass_precision 4 escapes), operands arriving as tagged words, stream words out of a 128-word window in
// shared memory (beyond it: global), tables through the read-only path.
template <bool SM>
__device__ __forceinline__ void scan_decode_share(const ScanArgs &S, const Tab<SM> &tb, const uint32_t *__restrict__ units, const uint32_t *win,
                                               uint32_t wbase, uint32_t wend, uint32_t &x, uint32_t &wp, int &st, long long dbase, int m,
                                               int lane, uint32_t tag, int cells, const int *s_hw)
{
    const unsigned lt_mask = (1u << lane) - 1;
    const int prec = S.dec.precision;
    const uint32_t pmask = (1u << prec) - 1;
    const uint16_t *win16 = reinterpret_cast<const uint16_t *>(win);
    const uint16_t *words16 = reinterpret_cast<const uint16_t *>(units);
    auto word_at = [&](uint32_t at) -> uint32_t {
        const uint32_t off = at - wbase;
        return off < 128u ? win16[off] : __ldg(words16 + at);
    };
    auto refill = [&](bool need) {
        const unsigned nm = __ballot_sync(0xffffffffu, need);
        if (need) {
            const uint32_t at = wp + __popc(nm & lt_mask);
            uint32_t word = 0;
            if (at < wend) word = word_at(at); else st |= 4;
            x = (x << 16) | word;
        }
        wp += __popc(nm);
    };
    const int C = S.C, per_b = C * cells;
    const int nblocks = (m + 127) >> 7;
    for (int blk = 0; blk < nblocks; ++blk) {
        const int j0 = blk * 128 + lane * 4;
        // operands of the block's four symbols: {scale index, tag, mean, tag}, all loads first
        uint4 op[4];
        uint4 mt[4];
#ifdef SCAN_TIMING
        const long long tw0 = clock64();
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (j0 + q < m) op[q] = ll_ld(reinterpret_cast<const uint2 *>(S.dec.idx_t + dbase + j0 + q));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = 0;
            if (j0 + q < m) {
                int spins = 0;
                while (op[q].y != tag || op[q].w != tag) {
                    op[q] = ll_ld(reinterpret_cast<const uint2 *>(S.dec.idx_t + dbase + j0 + q));
                    if (++spins > (1 << 24)) asm volatile("trap;");
                }
                c = (int)op[q].x;
                if ((uint32_t)c >= (uint32_t)S.dec.T) { st |= 1; c = 0; }
            }
            mt[q] = tb.meta_at(c);  // cdf_base | lut_base | cdf_size, lut_shift | offset
        }
#ifdef SCAN_TIMING
        __syncwarp();
        if (dbase == 0 && lane == 0) S.timing[9] += clock64() - tw0;   // chunk 0: waiting for its operands
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool active = j0 + q < m;
            const typename Tab<SM>::addr_t cd = tb.cdf_at(mt[q].x);
            const int nsyms = (int)(mt[q].z & 0xffffu) - 1, maxv = nsyms - 1;
            const uint32_t cum = x & pmask;
            // bucket LUT -> four CDF entries at once (one round trip; the 16-bit CDF stores 2^16 as 0: guarded by nsyms) -> select;
            // further probes only in tails of width-1 symbols
            int s = (int)Tab<SM>::template ld16<0>(tb.lut_at(mt[q].y + (cum >> ((mt[q].z >> 16) & 0xffu))));
            const typename Tab<SM>::addr_t e = cd + 2 * s;
            const uint32_t c0 = Tab<SM>::template ld16<0>(e), c1 = Tab<SM>::template ld16<2>(e), c2 = Tab<SM>::template ld16<4>(e), c3 = Tab<SM>::template ld16<6>(e);
            const bool a1 = s + 1 < nsyms && c1 <= cum;
            const bool a2 = a1 && s + 2 < nsyms && c2 <= cum;
            const bool a3 = a2 && s + 3 < nsyms && c3 <= cum;
            uint32_t start = a2 ? c2 : a1 ? c1 : c0, next = a2 ? c3 : a1 ? c2 : c1;
            s += (int)a1 + (int)a2;
            if (a3) {
                ++s;
                while (s + 1 < nsyms && Tab<SM>::template ld16<2>(cd + 2 * s) <= cum) ++s;
                start = Tab<SM>::template ld16<0>(cd + 2 * s);
                next = Tab<SM>::template ld16<2>(cd + 2 * s);
            }
            const uint32_t freq = (uint16_t)(next - start);
            if (active) x = freq * (x >> prec) + cum - start;
            refill(active && x < kRansL);
            int32_t value = s;
            const bool esc = active && S.dec.bypass && s == maxv;
            // bypass_precision 4: the first unit starts with the digit count nb (<= 8 for a 32-bit payload, one count token),
            // followed by the digits, least significant first, four tokens per unit
            if (__any_sync(0xffffffffu, esc)) {
                bool in = esc, first = true;
                uint32_t nb = 0, raw = 0, jj = 0;
                while (__any_sync(0xffffffffu, in)) {
                    const bool was = in;
                    if (in) {
                        uint32_t cnt, used, bits = x;
                        if (first) {
                            nb = x & 15u;
                            if (nb > 8) { st |= 4; nb = 0; }  // no encoder writes this
                            cnt = min(3u, nb);
                            used = cnt + 1;
                            bits = x >> 4;
                            first = false;
                        } else {
                            cnt = min(4u, nb - jj);
                            used = cnt;
                        }
                        raw |= (bits & ((1u << (4 * cnt)) - 1)) << (4 * jj);
                        jj += cnt;
                        x >>= 4 * used;
                        in = jj < nb;
                    }
                    refill(was && x < kRansL);
                }
                if (esc) {
                    const int32_t v2 = (int32_t)(raw >> 1);
                    value = (raw & 1) ? -v2 - 1 : v2 + maxv;
                }
            }
            if (active) {
                const int32_t sym = value + (int32_t)mt[q].w;
                // pgm_coder.py:973-975 (sym + mean), then _data_postprocess x * 1 + 0 (turns -0.0 into +0.0)
                const float v = __fadd_rn(__fadd_rn((float)sym, __uint_as_float(op[q].z)), 0.0f);
                const unsigned e = (unsigned)dbase + (unsigned)(j0 + q);   // (a slice holds < 2^31 elements: 32-bit divisions)
                const unsigned b = e / (unsigned)per_b, r2 = e - b * (unsigned)per_b;
                const unsigned c = cells == 1 ? r2 : r2 / (unsigned)cells, i = cells == 1 ? 0u : r2 - c * (unsigned)cells;
                const int hw = s_hw[b * cells + i];
                S.yhat_pm[((long long)b * S.HW + hw) * C + c] = make_uint2(__float_as_uint(v), S.call_tag);
                S.buf[((long long)b * C + c) * S.HW + hw] = v;
            }
        }
    }
}

template <bool DEC>   // DEC: the decoder's single launch (chunk warps inside the kernel)
__global__ void __launch_bounds__(kScanWarps * 32)
k_scan_stages(const __grid_constant__ ScanArgs S)
{
    extern __shared__ __align__(16) float smem[];   // resident weights of the four layers | A [kScanRows][Kmax]
    __shared__ int sr_b[2][kScanMaxRows], sr_hw[2][kScanMaxRows], sr_i[2][kScanMaxRows];   // the rows of the current / next stage
    __shared__ uint32_t sr_tap[2][kScanMaxRows], sr_grp[2][kScanMaxRows];
    __shared__ float s_part[kScanWarps * kScanSplit][kScanRows][2];   // partial sums of (owned pair, K part)
    __shared__ uint32_t s_win[kScanDecSlots][64];   // decoder warps: the next 128 stream words of their chunk
    __shared__ float s_tab[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, nctas = gridDim.x - (DEC ? S.dctas : 0);   // nctas: the CTAs that own weights
    const int k2 = S.ksize * S.ksize, C = S.C;
    if (DEC && cta >= nctas) {
        // ---- a decoder CTA: the coder tables in ITS shared memory (the 113 KB of weights leave no room for them beside a weight
        // CTA's, and table lookups out of L2 made a coding step 2.5 k cycles), one chunk per warp, nothing else to do
        unsigned char *sm8 = reinterpret_cast<unsigned char *>(smem);
        const Tab<true> tb = stage_tables<true>(S.dec.blob, S.dec.blob_bytes, S.dec.meta_bytes, S.dec.cdf16_bytes, sm8);
        uint32_t *win = reinterpret_cast<uint32_t *>(sm8 + ((S.dec.blob_bytes + 15) & ~(size_t)15)) + warp * 64;
        const int dk = (cta - nctas) + warp * S.dctas;
        const bool on = dk < S.dec.n_chunks;
        uint32_t dx = 0, dwp = 0, dwend = 0, dwbase = 0;
        int dst = 0;
        const uint32_t *d_units = nullptr;
        if (on) {
            const uint32_t *end_word = reinterpret_cast<const uint32_t *>(S.dec.seg) + 2 + S.dec.seg_slices;
            const uint32_t *states = end_word + S.dec.n_chunks;
            const long long words_at = kSegHdr + 4ll * S.dec.seg_slices + 4ll * S.dec.n_chunks + 128ll * S.dec.n_chunks;
            d_units = reinterpret_cast<const uint32_t *>(S.dec.seg + words_at);
            dwend = end_word[dk];
            dwp = dk ? end_word[dk - 1] : 0;
            if (dwend < dwp || words_at + 2ll * dwend > S.dec.seg_cap) { dst |= 4; dwend = dwp = 0; }  // corrupt directory
            dx = states[(size_t)dk * 32 + lane];
        }
        uint32_t step = S.step0;
        for (int g = S.g0; g < S.g1; ++g) {
            const int2 sc = S.stage_cells[g];
            for (int row = tid; row < S.B * sc.y; row += blockDim.x) sr_hw[0][row] = S.cell_hw[sc.x + (row % sc.y)];
            __syncthreads();
            step += 4;
            if (on) {
                const long long n = (long long)S.B * C * sc.y, cs = S.dec.chunk_syms[g], dbase = (long long)dk * cs, rem = n - dbase;
                const int m = (int)(rem <= 0 ? 0 : rem < cs ? rem : cs);
                if (m > 0) {
                    dwbase = dwp & ~1u;
                    const uint32_t u0 = dwbase >> 1, u_lim = (dwend + 1) >> 1;
                    win[lane] = u0 + lane < u_lim ? __ldg(d_units + u0 + lane) : 0u;
                    win[lane + 32] = u0 + 32 + lane < u_lim ? __ldg(d_units + u0 + 32 + lane) : 0u;
                    __syncwarp();
#ifdef SCAN_TIMING
                    const long long td0 = clock64();
#endif
                    scan_decode_share<true>(S, tb, d_units, win, dwbase, dwend, dx, dwp, dst, dbase, m, lane, step, sc.y, sr_hw[0]);
#ifdef SCAN_TIMING
                    if (dk == 0 && lane == 0) S.timing[8] += clock64() - td0;
#endif
                }
            }
            __syncthreads();
        }
        if (on) {
            if (dwp != dwend && lane == 0) dst |= 4;
            if (dst) atomicOr(S.dec.status, dst);
        }
        return;
    }
    int woff[5];
    woff[0] = 0;
#pragma unroll
    for (int L = 0; L < 4; ++L) woff[L + 1] = woff[L] + S.pairs[L] * 2 * S.K[L];
    float *A = smem + woff[4];
    for (int i = tid; i < S.qz.n_scales && i < 256; i += blockDim.x) s_tab[i] = S.qz.scale_table[i];
    // ---- this CTA's weight rows
#pragma unroll 1
    for (int L = 0; L < 4; ++L) {
        const int K4 = S.K[L] >> 2, C4 = C >> 2;
        for (int j = 0; j < S.pairs[L]; ++j) {
            const int n0 = 2 * (j * nctas + cta);
            if (n0 >= S.N[L]) continue;
            float4 *dst = reinterpret_cast<float4 *>(smem + woff[L] + j * 2 * S.K[L]);
            if (L == 0) {
                const float4 *src = reinterpret_cast<const float4 *>(S.w[0] + (size_t)n0 * k2 * C);
                for (int i = tid; i < 2 * K4; i += blockDim.x) {
                    const int o = i >= K4, q = i - o * K4, slot = q / C4, c4 = q - slot * C4;
                    dst[i] = __ldg(src + (size_t)(o * k2 + S.taps[slot]) * C4 + c4);
                }
            } else {
                const float4 *src = reinterpret_cast<const float4 *>(S.w[L] + (size_t)n0 * S.K[L]);
                for (int i = tid; i < 2 * K4; i += blockDim.x) dst[i] = __ldg(src + i);
            }
        }
    }
    RowsQuant qz = S.qz;
#ifdef SCAN_TIMING
    long long tk[5] = {0, 0, 0, 0, 0}, t_a = clock64(), t_b;   // (unused), gather, multiply, chunk sync, stage sync
#define SCAN_T(i) do { t_b = clock64(); tk[i] += t_b - t_a; t_a = t_b; } while (0)
#else
#define SCAN_T(i) do { } while (0)
#endif
    if (S.dq_sym) {   // decoder: y_hat of the previous stage (every CTA writes the same words and reads its own back)
        const int2 pc = S.stage_cells[S.g0 - 1];
        const int per_b = C * pc.y;
        for (int e = tid; e < S.B * per_b; e += blockDim.x) {
            const int row = e / C, c = e - row * C, b = row / pc.y, i = row - b * pc.y;
            const int hw = S.cell_hw[pc.x + i];
            const float mean = __uint_as_float(__ldcg(&S.vec[3][(size_t)row * 2 * C + 2 * c].x));   // (the previous launch left its parameters there)
            // pgm_coder.py:973-975 (sym + mean), then _data_postprocess x * 1 + 0 (turns -0.0 into +0.0)
            const float v = __fadd_rn(__fadd_rn((float)S.dq_sym[(size_t)b * per_b + (size_t)c * pc.y + i], mean), 0.0f);
            S.buf[((long long)b * C + c) * S.HW + hw] = v;
            S.yhat_pm[((long long)b * S.HW + hw) * C + c] = make_uint2(__float_as_uint(v), S.call_tag);
        }
    }
    __syncthreads();
    SCAN_T(0);
    // ---- decoder warps (single-launch decoding)
    const int dslot = kScanWarps - 1 - warp;
    const int dk = dslot * nctas + cta;                       // this warp's chunk
    const bool dec_warp = DEC && S.dctas == 0 && dslot < kScanDecSlots && dk < S.dec.n_chunks;   // (no decoder CTAs: chunk warps beside the weights)
    uint32_t dx = 0, dwp = 0, dwend = 0, dwbase = 0;
    int dst = 0;
    const uint32_t *d_units = nullptr;
    Tab<false> dtb;
    if (DEC && dec_warp) {
        const uint32_t *end_word = reinterpret_cast<const uint32_t *>(S.dec.seg) + 2 + S.dec.seg_slices;
        const uint32_t *states = end_word + S.dec.n_chunks;
        const long long words_at = kSegHdr + 4ll * S.dec.seg_slices + 4ll * S.dec.n_chunks + 128ll * S.dec.n_chunks;
        d_units = reinterpret_cast<const uint32_t *>(S.dec.seg + words_at);
        dwend = end_word[dk];
        dwp = dk ? end_word[dk - 1] : 0;
        if (dwend < dwp || words_at + 2ll * dwend > S.dec.seg_cap) { dst |= 4; dwend = dwp = 0; }  // corrupt directory
        dx = states[(size_t)dk * 32 + lane];
        dtb.init(S.dec.blob, nullptr, S.dec.meta_bytes, S.dec.cdf16_bytes);
    }
    uint32_t step = S.step0;
    // the rows of a stage: image, cell, position, visibility -- loaded one stage ahead
    auto load_rows = [&](int g, int slot) {
        const int2 sc = S.stage_cells[g];
        for (int row = tid; row < S.B * sc.y; row += blockDim.x) {
            const int b = row / sc.y, i = row - b * sc.y, cell = sc.x + i;
            sr_b[slot][row] = b; sr_i[slot][row] = i; sr_hw[slot][row] = S.cell_hw[cell]; sr_tap[slot][row] = S.cell_tap[cell];
            sr_grp[slot][row] = S.cell_grp[cell];
        }
    };
    load_rows(S.g0, S.g0 & 1);
    __syncthreads();
    for (int g = S.g0; g < S.g1; ++g) {
        const int2 sc = S.stage_cells[g];
        const int rows_total = S.B * sc.y;
        const long long slice = (long long)S.B * qz.C * sc.y;
        const int *s_b = sr_b[g & 1], *s_hw = sr_hw[g & 1], *s_i = sr_i[g & 1];
        const uint32_t *s_tap = sr_tap[g & 1], *s_grp = sr_grp[g & 1];
        if (g + 1 < S.g1) load_rows(g + 1, (g + 1) & 1);   // (its last readers passed the barrier that ended stage g - 1)
        if (DEC && dec_warp) {   // the next 128 words of the chunk: in shared memory long before the stage's indexes arrive
            dwbase = dwp & ~1u;
            const uint32_t u0 = dwbase >> 1, u_lim = (dwend + 1) >> 1;
            s_win[dslot][lane] = u0 + lane < u_lim ? __ldg(d_units + u0 + lane) : 0u;
            s_win[dslot][lane + 32] = u0 + 32 + lane < u_lim ? __ldg(d_units + u0 + 32 + lane) : 0u;
            __syncwarp();
        }
#pragma unroll 1
        for (int L = 0; L < 4; ++L) {
            ++step;
            const int K = S.K[L], N = S.N[L];
            if (2 * cta >= N) continue;   // this CTA owns no channel of the layer: it neither gathers nor waits (CTA-uniform)
            const int owned = min(S.pairs[L], (N / 2 - cta + nctas - 1) / nctas);   // pairs cta, cta + nctas, ... < N / 2
            const int n0 = 2 * (warp * nctas + cta);
            const bool mine = warp < owned;
            float bias0 = 0.f, bias1 = 0.f, yv = 0.f;
            if (mine) { bias0 = S.bias[L][n0]; bias1 = S.bias[L][n0 + 1]; }
            if (L == 3 && mine && qz.y && lane < rows_total && rows_total <= kScanRows)   // (off the chain: it waits in DRAM while the layer runs)
                yv = qz.y[((long long)s_b[lane] * qz.C + (n0 >> 1)) * S.HW + s_hw[lane]];
#pragma unroll 1
            for (int r0 = 0; r0 < rows_total; r0 += kScanRows) {
                const int rows = min(kScanRows, rows_total - r0);
                // ---- gather the input vectors: contiguous 128-bit loads out of L2, polled until the tags match.  The rows of the
                // chunk form ONE index space (row-major pairs of floats), so that a thread's loads of all rows are in flight together
                {
                    float2 *A2 = reinterpret_cast<float2 *>(A);
                    const int C2 = C >> 1, K2 = K >> 1, total = rows * K2;
                    if (L == 0) {
                        for (int base = tid; base < total; base += blockDim.x * kScanBatch) {
                            uint4 q[kScanBatch];
                            const uint2 *ptr[kScanBatch];
#pragma unroll
                            for (int u = 0; u < kScanBatch; ++u) {
                                const int o = base + u * blockDim.x;
                                ptr[u] = nullptr;
                                if (o < total) {
                                    const int r = o / K2, i = o - r * K2, t = i / C2, c2 = i - t * C2;
                                    if ((s_tap[r0 + r] >> S.taps[t]) & 1u) {
                                        ptr[u] = S.yhat_pm + ((long long)s_b[r0 + r] * S.HW + s_hw[r0 + r] + S.shift[t]) * C + 2 * c2;
                                        q[u] = ll_ld(ptr[u]);
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < kScanBatch; ++u) {
                                const int o = base + u * blockDim.x;
                                if (o < total) A2[o] = ptr[u] ? ll_wait(q[u], ptr[u], S.call_tag) : make_float2(0.f, 0.f);
                            }
                        }
                    } else {
                        const int c0n2 = S.N[L - 1] >> 1, p2 = K2 - c0n2;   // (p2 > 0 only at L == 1: the prior, always visible)
                        for (int base = tid; base < total; base += blockDim.x * 2) {
                            uint4 q[2];
                            float2 pv[2];
                            const uint2 *ptr[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int o = base + u * blockDim.x;
                                ptr[u] = nullptr;
                                pv[u] = make_float2(0.f, 0.f);
                                if (o < total) {
                                    const int r = o / K2, i = o - r * K2;
                                    if (i < c0n2) {
                                        if (s_grp[r0 + r] & 1u) {
                                            ptr[u] = S.vec[L - 1] + (size_t)(r0 + r) * S.N[L - 1] + 2 * i;
                                            q[u] = ll_ld(ptr[u]);
                                        }
                                    } else {
                                        pv[u] = __ldg(reinterpret_cast<const float2 *>(S.prior_pm + ((long long)s_b[r0 + r] * S.HW + s_hw[r0 + r]) * (2 * p2)) + (i - c0n2));
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int o = base + u * blockDim.x;
                                if (o < total) A2[o] = ptr[u] ? ll_wait(q[u], ptr[u], step - 1) : pv[u];
                            }
                        }
                    }
                }
                __syncthreads();
                SCAN_T(1);
                // ---- (owned pair, K part) items over the warps
                for (int item = warp; item < owned * kScanSplit; item += kScanWarps) {
                    const int j = item / kScanSplit, q = item - j * kScanSplit;
                    const float *wrow = smem + woff[L] + j * 2 * K;
                    float *part = &s_part[item][0][0];
                    switch (rows) {
                    case 1: scan_pair_part<1>(wrow, A, K, q, lane, part); break;
                    case 2: scan_pair_part<2>(wrow, A, K, q, lane, part); break;
                    case 3: scan_pair_part<3>(wrow, A, K, q, lane, part); break;
                    default: scan_pair_part<4>(wrow, A, K, q, lane, part); break;
                    }
                }
                SCAN_T(2);
                __syncthreads();
                SCAN_T(3);
                // ---- warp j, lane r finishes row r of pair j
                if (mine && lane < rows) {
                    const int row = r0 + lane;
                    const float(*pp)[kScanRows][2] = &s_part[warp * kScanSplit];
                    float m0 = (pp[0][lane][0] + pp[1][lane][0]) + (pp[2][lane][0] + pp[3][lane][0]);
                    float m1 = (pp[0][lane][1] + pp[1][lane][1]) + (pp[2][lane][1] + pp[3][lane][1]);
                    m0 += bias0;
                    m1 += bias1;
                    if (L == 1 || L == 2) {
                        m0 = m0 > 0.f ? m0 : m0 * kSlope;
                        m1 = m1 > 0.f ? m1 : m1 * kSlope;
                    }
                    *reinterpret_cast<uint4 *>(S.vec[L] + (size_t)row * N + n0) = make_uint4(__float_as_uint(m0), step, __float_as_uint(m1), step);
                    if (L == 3) {   // (mean, scale) of latent channel n0 / 2: scale index, symbol, y_hat
                        const int c = n0 >> 1, b = s_b[row];
                        const long long e = (long long)b * qz.C * sc.y + (long long)c * sc.y + s_i[row];
                        const long long yo = ((long long)b * qz.C + c) * S.HW + s_hw[row];
                        if (qz.y) {
                            if (rows_total > kScanRows) yv = qz.y[yo];
                            const float sq = rintf(__fsub_rn(yv, m0));  // torch.round: half to even
                            const float yh = __fadd_rn(sq, m0);
                            S.yhat_pm[((long long)b * S.HW + s_hw[row]) * qz.C + c] = make_uint2(__float_as_uint(yh), S.call_tag);   // (first: the next stage waits for it)
                            qz.sym[e] = (int32_t)sq;
                            qz.buf[yo] = yh;
                        }
                        const int si = scale_index_dev(m1, s_tab, qz.n_scales);
                        if (DEC) S.dec.idx_t[e] = make_uint4((uint32_t)si, step, __float_as_uint(m0), step);   // (first: a chunk warp waits for it)
                        qz.idx[e] = si;
                        const long long oo = ((long long)b * N + n0) * S.HW + s_hw[row];
                        S.params[oo] = m0;
                        S.params[oo + S.HW] = m1;
                    }
                }
                SCAN_T(4);
            }
        }
        if (DEC && dec_warp) {
            const long long n = slice, cs = S.dec.chunk_syms[g], dbase = (long long)dk * cs, rem = n - dbase;
            const int m = (int)(rem <= 0 ? 0 : rem < cs ? rem : cs);
#ifdef SCAN_TIMING
            const long long td0 = clock64();
#endif
            if (m > 0)
                scan_decode_share<false>(S, dtb, d_units, s_win[dslot], dwbase, dwend, dx, dwp, dst, dbase, m, lane, step, sc.y, s_hw);
#ifdef SCAN_TIMING
            if (dk == 0 && lane == 0) S.timing[8] += clock64() - td0;   // chunk 0: operand wait + decode of its share, all stages
#endif
        }
        __syncthreads();   // the row records of stage g + 1 are complete, those of stage g free
        SCAN_T(0);
        qz.sym += slice;
        qz.idx += slice;
    }
    if (DEC && dec_warp) {
        if (dwp != dwend && lane == 0) dst |= 4;   // the chunk's words must be used up exactly
        if (dst) atomicOr(S.dec.status, dst);
    }
#ifdef SCAN_TIMING
    if (tid == 0 && cta == 0)
        for (int i = 0; i < 5; ++i) S.timing[i] = tk[i];
#endif
}

// ---- the same walk for stages of 5 .. 128 rows (a batch of scanline images: configs[3]'s scanline level codes 64 crops at
// once): a 2-D decomposition.  The rows of a stage are cut into blocks of kBlkRows; CTA (rb, cb) of a row-blocks x CB grid
// computes, for the rows of block rb, the channel pairs cb, cb + CB, ... of every layer.  A CTA owns 1 / CB of the weights -- too
// much to keep resident -- so each warp streams the two weight rows of a pair from L2 once per stage against all rows of the
// block held in shared memory (k_scan_stages with R rows would gather every row in every CTA: 322 MB of L2 traffic per stage at
// 64 rows; here: weights x row blocks + vectors x CB = ~100 MB).  Exchange, tags, quantiser and the in-kernel chunk decoder are
// those of k_scan_stages; vectors only travel inside a row block.
constexpr int kBlkRows = 8, kBlkMaxRows = 128, kBlkInFlight = 4, kBlkWarps = 16;   // (16 rows per block: 9.3 -> 13.9 ms on configs[3]'s scanline level)

template <int R>
__device__ __forceinline__ void blk_pair(const float *__restrict__ w0g, const float *__restrict__ w1g, const float *__restrict__ A, int K,
                                         int lane, float &m0, float &m1)
{
    const int K4 = K >> 2;
    const float4 *w0 = reinterpret_cast<const float4 *>(w0g), *w1 = reinterpret_cast<const float4 *>(w1g);
    const float4 *A4 = reinterpret_cast<const float4 *>(A);
    float a0[R], a1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a0[r] = a1[r] = 0.f;
    for (int i0 = lane; i0 < K4; i0 += 32 * kBlkInFlight) {
        float4 x0[kBlkInFlight], x1[kBlkInFlight];
#pragma unroll
        for (int u = 0; u < kBlkInFlight; ++u) {
            const int i = i0 + 32 * u;
            const bool ok = i < K4;
            x0[u] = ok ? __ldg(w0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            x1[u] = ok ? __ldg(w1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kBlkInFlight; ++u) {
            const int i = i0 + 32 * u;
            if (i < K4) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 av = A4[r * K4 + i];
                    a0[r] = fmaf(av.x, x0[u].x, a0[r]); a0[r] = fmaf(av.y, x0[u].y, a0[r]);
                    a0[r] = fmaf(av.z, x0[u].z, a0[r]); a0[r] = fmaf(av.w, x0[u].w, a0[r]);
                    a1[r] = fmaf(av.x, x1[u].x, a1[r]); a1[r] = fmaf(av.y, x1[u].y, a1[r]);
                    a1[r] = fmaf(av.z, x1[u].z, a1[r]); a1[r] = fmaf(av.w, x1[u].w, a1[r]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a0[r] += __shfl_xor_sync(0xffffffffu, a0[r], o);
            a1[r] += __shfl_xor_sync(0xffffffffu, a1[r], o);
        }
    }
    m0 = a0[0];
    m1 = a1[0];
#pragma unroll
    for (int r = 1; r < R; ++r) if (lane == r) { m0 = a0[r]; m1 = a1[r]; }
}

template <bool DEC>
__global__ void __launch_bounds__(kBlkWarps * 32)
k_scan_blocks(const __grid_constant__ ScanArgs S)
{
    extern __shared__ __align__(16) float A[];   // [kBlkRows][Kmax]
    __shared__ int sr_b[2][kBlkMaxRows], sr_hw[2][kBlkMaxRows], sr_i[2][kBlkMaxRows];   // the rows of the current / next stage (all of them: the chunk decoder needs any)
    __shared__ uint32_t sr_tap[2][kBlkMaxRows], sr_grp[2][kBlkMaxRows];
    __shared__ uint32_t s_win[kScanDecSlots][64];
    __shared__ float s_tab[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, nctas = gridDim.x, CB = S.CB, rb = cta / CB, cb = cta - rb * CB;
    const int C = S.C;
    for (int i = tid; i < S.qz.n_scales && i < 256; i += blockDim.x) s_tab[i] = S.qz.scale_table[i];
    RowsQuant qz = S.qz;
    // ---- decoder warps (single-launch decoding), as in k_scan_stages
    const int dslot = kBlkWarps - 1 - warp;
    const int dk = dslot * nctas + cta;
    const bool dec_warp = DEC && dslot < kScanDecSlots && dk < S.dec.n_chunks;
    uint32_t dx = 0, dwp = 0, dwend = 0, dwbase = 0;
    int dst = 0;
    const uint32_t *d_units = nullptr;
    Tab<false> dtb;
    if (DEC && dec_warp) {
        const uint32_t *end_word = reinterpret_cast<const uint32_t *>(S.dec.seg) + 2 + S.dec.seg_slices;
        const uint32_t *states = end_word + S.dec.n_chunks;
        const long long words_at = kSegHdr + 4ll * S.dec.seg_slices + 4ll * S.dec.n_chunks + 128ll * S.dec.n_chunks;
        d_units = reinterpret_cast<const uint32_t *>(S.dec.seg + words_at);
        dwend = end_word[dk];
        dwp = dk ? end_word[dk - 1] : 0;
        if (dwend < dwp || words_at + 2ll * dwend > S.dec.seg_cap) { dst |= 4; dwend = dwp = 0; }
        dx = states[(size_t)dk * 32 + lane];
        dtb.init(S.dec.blob, nullptr, S.dec.meta_bytes, S.dec.cdf16_bytes);
    }
    if (S.dq_sym) {   // decoder launched stage by stage: y_hat of the previous stage first (every CTA writes the same words)
        const int2 pc = S.stage_cells[S.g0 - 1];
        const int per_b = C * pc.y;
        for (int e = tid; e < S.B * per_b; e += blockDim.x) {
            const int row = e / C, c = e - row * C, b = row / pc.y, i = row - b * pc.y;
            const int hw = S.cell_hw[pc.x + i];
            const float mean = __uint_as_float(__ldcg(&S.vec[3][(size_t)row * 2 * C + 2 * c].x));
            const float v = __fadd_rn(__fadd_rn((float)S.dq_sym[(size_t)b * per_b + (size_t)c * pc.y + i], mean), 0.0f);
            S.buf[((long long)b * C + c) * S.HW + hw] = v;
            S.yhat_pm[((long long)b * S.HW + hw) * C + c] = make_uint2(__float_as_uint(v), S.call_tag);
        }
        __syncthreads();
    }
    uint32_t step = S.step0;
    auto load_rows = [&](int g, int slot) {
        const int2 sc = S.stage_cells[g];
        for (int row = tid; row < S.B * sc.y; row += blockDim.x) {
            const int b = row / sc.y, i = row - b * sc.y, cell = sc.x + i;
            sr_b[slot][row] = b; sr_i[slot][row] = i; sr_hw[slot][row] = S.cell_hw[cell]; sr_tap[slot][row] = S.cell_tap[cell];
            sr_grp[slot][row] = S.cell_grp[cell];
        }
    };
    load_rows(S.g0, S.g0 & 1);
    __syncthreads();
    for (int g = S.g0; g < S.g1; ++g) {
        const int2 sc = S.stage_cells[g];
        const int rows_total = S.B * sc.y;
        const long long slice = (long long)S.B * qz.C * sc.y;
        const int *s_b = sr_b[g & 1], *s_hw = sr_hw[g & 1], *s_i = sr_i[g & 1];
        const uint32_t *s_tap = sr_tap[g & 1], *s_grp = sr_grp[g & 1];
        if (g + 1 < S.g1) load_rows(g + 1, (g + 1) & 1);
        if (DEC && dec_warp) {
            dwbase = dwp & ~1u;
            const uint32_t u0 = dwbase >> 1, u_lim = (dwend + 1) >> 1;
            s_win[dslot][lane] = u0 + lane < u_lim ? __ldg(d_units + u0 + lane) : 0u;
            s_win[dslot][lane + 32] = u0 + 32 + lane < u_lim ? __ldg(d_units + u0 + 32 + lane) : 0u;
            __syncwarp();
        }
        const int r0 = rb * kBlkRows, rows = min(kBlkRows, rows_total - r0);   // this CTA's rows of the stage (<= 0: none)
#pragma unroll 1
        for (int L = 0; L < 4; ++L) {
            ++step;
            const int K = S.K[L], N = S.N[L], npairs = N >> 1;
            if (rows <= 0 || cb >= npairs) continue;   // nothing of this layer is computed here: neither gather nor wait (CTA-uniform)
            // ---- gather the rows of the block (as k_scan_stages: one index space, loads of a batch in flight together)
            {
                float2 *A2 = reinterpret_cast<float2 *>(A);
                const int C2 = C >> 1, K2 = K >> 1, total = rows * K2;
                if (L == 0) {
                    for (int base = tid; base < total; base += blockDim.x * kScanBatch) {
                        uint4 q[kScanBatch];
                        const uint2 *ptr[kScanBatch];
#pragma unroll
                        for (int u = 0; u < kScanBatch; ++u) {
                            const int o = base + u * blockDim.x;
                            ptr[u] = nullptr;
                            if (o < total) {
                                const int r = o / K2, i = o - r * K2, t = i / C2, c2 = i - t * C2;
                                if ((s_tap[r0 + r] >> S.taps[t]) & 1u) {
                                    ptr[u] = S.yhat_pm + ((long long)s_b[r0 + r] * S.HW + s_hw[r0 + r] + S.shift[t]) * C + 2 * c2;
                                    q[u] = ll_ld(ptr[u]);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kScanBatch; ++u) {
                            const int o = base + u * blockDim.x;
                            if (o < total) A2[o] = ptr[u] ? ll_wait(q[u], ptr[u], S.call_tag) : make_float2(0.f, 0.f);
                        }
                    }
                } else {
                    const int c0n2 = S.N[L - 1] >> 1, p2 = K2 - c0n2;
                    for (int base = tid; base < total; base += blockDim.x * 4) {
                        uint4 q[4];
                        float2 pv[4];
                        const uint2 *ptr[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int o = base + u * blockDim.x;
                            ptr[u] = nullptr;
                            pv[u] = make_float2(0.f, 0.f);
                            if (o < total) {
                                const int r = o / K2, i = o - r * K2;
                                if (i < c0n2) {
                                    if (s_grp[r0 + r] & 1u) {
                                        ptr[u] = S.vec[L - 1] + (size_t)(r0 + r) * S.N[L - 1] + 2 * i;
                                        q[u] = ll_ld(ptr[u]);
                                    }
                                } else {
                                    pv[u] = __ldg(reinterpret_cast<const float2 *>(S.prior_pm + ((long long)s_b[r0 + r] * S.HW + s_hw[r0 + r]) * (2 * p2)) + (i - c0n2));
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int o = base + u * blockDim.x;
                            if (o < total) A2[o] = ptr[u] ? ll_wait(q[u], ptr[u], step - 1) : pv[u];
                        }
                    }
                }
            }
            __syncthreads();
            // ---- a warp per owned channel pair, all rows of the block at once (rows past `rows` hold stale values: computed, never used)
            const float *wl = L == 0 ? S.wc : S.w[L];
            for (int p = cb + warp * CB; p < npairs; p += kBlkWarps * CB) {
                const int n0 = 2 * p;
                float m0, m1;
                if (rows > 4) blk_pair<8>(wl + (size_t)n0 * K, wl + (size_t)(n0 + 1) * K, A, K, lane, m0, m1);
                else blk_pair<4>(wl + (size_t)n0 * K, wl + (size_t)(n0 + 1) * K, A, K, lane, m0, m1);
                if (lane < rows) {   // lane r finishes row r
                    const int row = r0 + lane;
                    m0 += S.bias[L][n0];
                    m1 += S.bias[L][n0 + 1];
                    if (L == 1 || L == 2) {
                        m0 = m0 > 0.f ? m0 : m0 * kSlope;
                        m1 = m1 > 0.f ? m1 : m1 * kSlope;
                    }
                    *reinterpret_cast<uint4 *>(S.vec[L] + (size_t)row * N + n0) = make_uint4(__float_as_uint(m0), step, __float_as_uint(m1), step);
                    if (L == 3) {   // (mean, scale) of latent channel p: scale index, symbol, y_hat
                        const int c = p, b = s_b[row];
                        const long long e = (long long)b * qz.C * sc.y + (long long)c * sc.y + s_i[row];
                        const long long yo = ((long long)b * qz.C + c) * S.HW + s_hw[row];
                        if (qz.y) {
                            const float sq = rintf(__fsub_rn(qz.y[yo], m0));  // torch.round: half to even
                            const float yh = __fadd_rn(sq, m0);
                            S.yhat_pm[((long long)b * S.HW + s_hw[row]) * qz.C + c] = make_uint2(__float_as_uint(yh), S.call_tag);
                            qz.sym[e] = (int32_t)sq;
                            qz.buf[yo] = yh;
                        }
                        const int si = scale_index_dev(m1, s_tab, qz.n_scales);
                        if (DEC) S.dec.idx_t[e] = make_uint4((uint32_t)si, step, __float_as_uint(m0), step);
                        qz.idx[e] = si;
                        const long long oo = ((long long)b * N + n0) * S.HW + s_hw[row];
                        S.params[oo] = m0;
                        S.params[oo + S.HW] = m1;
                    }
                }
            }
            __syncthreads();   // A is rewritten by the next layer's gather
        }
        if (DEC && dec_warp) {
            const long long n = slice, cs = S.dec.chunk_syms[g], dbase = (long long)dk * cs, rem = n - dbase;
            const int m = (int)(rem <= 0 ? 0 : rem < cs ? rem : cs);
            if (m > 0)
                scan_decode_share<false>(S, dtb, d_units, s_win[dslot], dwbase, dwend, dx, dwp, dst, dbase, m, lane, step, sc.y, s_hw);
        }
        __syncthreads();   // the row records of stage g + 1 are complete, those of stage g free
        qz.sym += slice;
        qz.idx += slice;
    }
    if (DEC && dec_warp) {
        if (dwp != dwend && lane == 0) dst |= 4;
        if (dst) atomicOr(S.dec.status, dst);
    }
}

// the visible taps of the N-major convolution weights, packed: dst[n][t * C + c] = src[n][taps[t] * C + c]
__global__ void k_compact_taps(const float *__restrict__ src, float *__restrict__ dst, int N, int C, int k2, int ntaps, uint32_t tap_union)
{
    __shared__ int taps[32];
    if (threadIdx.x == 0) {
        int n = 0;
        for (int t = 0; t < k2; ++t) if ((tap_union >> t) & 1u) taps[n++] = t;
    }
    __syncthreads();
    const long long total = (long long)N * ntaps * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int t = (int)(r % ntaps), n = (int)(r / ntaps);
        dst[i] = src[((size_t)n * k2 + taps[t]) * C + c];
    }
}

// [B][channels][HW] -> [B][HW][channels] (the stage kernel's position-major copy of the prior)
__global__ void __launch_bounds__(256)
k_to_position_major(const float *__restrict__ src, float *__restrict__ dst, int channels, int HW)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8)
        tile[j][tx] = (c0 + j < channels && p0 + tx < HW) ? src[((size_t)b * channels + c0 + j) * HW + p0 + tx] : 0.f;
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (p0 + j < HW && c0 + tx < channels) dst[((size_t)b * HW + p0 + j) * channels + c0 + tx] = tile[tx][j];
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

// ---- the persistent stage kernel (k_scan_stages): when it applies and how it is launched
static int scan_ctas(const CtxModel &m)
{
    static const int forced = [] { const char *e = getenv("BASIC_SCAN_CTAS"); return e ? atoi(e) : 0; }();  // sweeps
    return forced > 0 ? std::min(forced, m.sm_count) : m.sm_count;   // one CTA per SM: all of them must be resident (grid barrier)
}

static size_t scan_smem(const CtxModel &m, int ntaps)
{
    const int nctas = scan_ctas(m);
    const int K[4] = {ntaps * m.C, 2 * m.c_ctx, m.c_m1, m.c_m2}, N[4] = {m.c_ctx, m.c_m1, m.c_m2, m.c_ctx};
    size_t fl = 0;
    int kmax = 0;
    for (int L = 0; L < 4; ++L) {
        fl += (size_t)((N[L] / 2 + nctas - 1) / nctas) * 2 * K[L];
        kmax = std::max(kmax, K[L]);
    }
    return (fl + (size_t)kScanRows * kmax) * sizeof(float);
}

static int scan_ntaps(const CtxModel &m)   // taps some stage of the current map can see
{
    uint32_t u = 0;
    for (const auto &st : m.stages) u |= st.tap_or;
    int n = 0;
    for (; u; u &= u - 1) ++n;
    return std::max(n, 1);
}

static size_t blk_smem(const CtxModel &m, int ntaps)
{
    const int kmax = std::max(std::max(ntaps * m.C, 2 * m.c_ctx), std::max(m.c_m1, m.c_m2));
    return (size_t)kBlkRows * kmax * sizeof(float);
}

bool ctx_scan_supported(const CtxModel &m, int B)
{
    static const bool off = [] { const char *e = getenv("BASIC_SCAN_KERNEL"); return e && e[0] == '0'; }();  // A/B switch
    if (off || !m.has_conv || !m.has_merger || m.internal || m.G != 1 || m.S < 8 || !m.d_stage_cells.p || !m.ws_ctx.p) return false;
    // (up to kScanMaxRows rows per stage the tensor path would run 128-row tiles that are mostly empty, four launches per stage;
    // measured at 64 rows per stage: 258 us here against 160 us there)
    if ((long long)B * m.max_stage_cells > kBlkMaxRows || m.k > 5) return false;
    if (m.C % 4 || m.c_m1 % 4 || m.c_m2 % 4 || (m.c_m1 | m.c_m2 | m.c_ctx) & 1) return false;   // 128-bit weight loads, channel pairs
    if ((long long)B * m.max_stage_cells > kScanRows)   // row blocks (k_scan_blocks): only the rows of a block live in shared memory
        return blk_smem(m, scan_ntaps(m)) <= 216 * 1024;
    const int nctas = scan_ctas(m);
    if ((std::max(m.c_m1, std::max(m.c_m2, m.c_ctx)) / 2 + nctas - 1) / nctas > kScanWarps) return false;   // a warp per owned pair
    return scan_smem(m, m.k * m.k) <= 216 * 1024;   // (+ 8 KB of static shared memory: 227 KB per CTA)
}

// Stages [g0, g1): parameters into `params` (NCHW), scale indexes (and, with y, symbols + the y_hat write-back into buf) into
// the stream slices starting at idx / sym.  One launch.  dq_sym (decoder): symbols of stage g0 - 1, dequantised into buf first.
// grid of the stage kernels for `rows` rows per stage: k_scan_stages (one CTA per SM) up to kScanRows rows, else k_scan_blocks
// (row blocks x channel blocks)
static void scan_grid(const CtxModel &m, int rows, int *nctas, int *CB)
{
    if (rows <= kScanRows) { *nctas = scan_ctas(m); *CB = 0; return; }
    const int RB = (rows + kBlkRows - 1) / kBlkRows;
    *CB = std::max(1, std::min(scan_ctas(m) / RB, std::max(m.c_m1, std::max(m.c_m2, m.c_ctx)) / 2));
    *nctas = RB * *CB;
}

bool ctx_scan_decode_supported(const CtxModel &m, int B, int n_chunks, int bypass_precision, int freq_precision)
{
    static const bool off = [] { const char *e = getenv("BASIC_SCAN_DECODE"); return e && e[0] == '0'; }();  // A/B switch
    int nctas, CB;
    scan_grid(m, B * m.max_stage_cells, &nctas, &CB);
    return !off && n_chunks > 0 && n_chunks <= kScanDecSlots * nctas && bypass_precision == 4 && freq_precision <= 16;
}

int ctx_scan_run(CtxModel &m, int g0, int g1, float *buf, const float *prior, int B, float *params, const float *y, int32_t *sym,
                 int32_t *idx, const float *d_scale_table, int n_scales, cudaStream_t stream, const int32_t *dq_sym,
                 const ScanDecodeHost *dec)
{
    if (g0 < 0 || g1 > m.S || g0 >= g1) return value_error("stage range out of bounds");
    if (dq_sym && g0 == 0) return value_error("no stage precedes stage 0");
    const int HW = m.H * m.W;
    // workspace: tagged layer outputs [kScanMaxRows][N] x 4 | tagged position-major y_hat [B][HW][C] | position-major prior
    const size_t vec_w = (size_t)kBlkMaxRows * (2 * m.c_ctx + m.c_m1 + m.c_m2), yh_w = (size_t)B * HW * m.C;
    const size_t idx_w = (size_t)kBlkMaxRows * m.C;   // {scale index, tag, mean, tag} per element of a stage's slice
    const size_t ws_bytes = (vec_w + yh_w) * sizeof(uint2) + idx_w * sizeof(uint4) + (size_t)B * HW * 2 * m.C * sizeof(float) + 64;
    int nctas, CB;
    scan_grid(m, B * m.max_stage_cells, &nctas, &CB);
    if (nctas > m.sm_count) return value_error("stage kernel: more rows per stage than the grid can hold");
    const uint32_t steps = 4u * (uint32_t)(g1 - g0);
    if (ws_bytes > m.scan_ws.cap || m.scan_nctas != nctas || m.scan_step > 0x7fff0000u - steps || m.scan_call > 0x7fff0000u) {
        // (re)allocated, or the tags are about to wrap: every tag back to "never written"
        if (g0 != 0) return value_error("stage kernel: workspace changed in the middle of a coding call");
        BASIC_TRY(m.scan_ws.reserve(ws_bytes));
        BASIC_CUDA(cudaMemsetAsync(m.scan_ws.p, 0, m.scan_ws.cap, stream));
        m.scan_step = 0;
        m.scan_call = 0;
        m.scan_nctas = nctas;
    }
    uint2 *vec = m.scan_ws.as<uint2>(), *yhat_pm = vec + vec_w;
    uint4 *idx_t = reinterpret_cast<uint4 *>(yhat_pm + yh_w);
    float *prior_pm = reinterpret_cast<float *>(idx_t + idx_w);
    if (g0 == 0) {   // a new coding call (both the encoder's one launch and the decoder's first start here)
        ++m.scan_call;
        const dim3 grid((HW + 31) / 32, (2 * m.C + 31) / 32, B);
        k_to_position_major<<<grid, 256, 0, stream>>>(prior, prior_pm, 2 * m.C, HW);
        BASIC_LAUNCHED();
    }
    ScanArgs S = {};
    // the taps ANY stage of the map can see, whatever [g0, g1) is: the compact K order -- and with it the order of the
    // additions -- must be the same in the encoder's one launch and the decoder's per-stage launches (masked taps add +0)
    uint32_t tap_union = 0;
    for (const auto &st : m.stages) tap_union |= st.tap_or;
    for (int t = 0; t < m.k * m.k; ++t)
        if ((tap_union >> t) & 1u) {
            S.shift[S.ntaps] = (t / m.k - m.k / 2) * m.W + (t % m.k - m.k / 2);
            S.taps[S.ntaps++] = (unsigned char)t;
        }
    S.w[0] = m.ws_ctx.as<float>(); S.w[1] = m.ws_m1.as<float>(); S.w[2] = m.ws_m2.as<float>(); S.w[3] = m.ws_m3.as<float>();
    S.bias[0] = m.b_ctx.as<float>(); S.bias[1] = m.b_m1.as<float>(); S.bias[2] = m.b_m2.as<float>(); S.bias[3] = m.b_m3.as<float>();
    S.N[0] = m.c_ctx; S.N[1] = m.c_m1; S.N[2] = m.c_m2; S.N[3] = m.c_ctx;
    S.K[0] = S.ntaps * m.C; S.K[1] = 2 * m.c_ctx; S.K[2] = m.c_m1; S.K[3] = m.c_m2;
    // single-launch decoding on k_scan_stages: a few CTAs of the grid hold the coder tables instead of weights and do nothing
    // but decode (one chunk per warp), when the chunks fit them and the weights still fit the remaining CTAs
    constexpr int kDecCtas = 4;
    int wctas = nctas;
    size_t dec_smem = 0;
    if (dec && CB == 0 && dec->n_chunks <= kDecCtas * kScanWarps && nctas > 4 * kDecCtas) {
        static const bool off = [] { const char *e = getenv("BASIC_SCAN_DEC_CTAS"); return e && e[0] == '0'; }();  // A/B switch
        const size_t need = ((dec->tables->blob_bytes + 15) & ~(size_t)15) + (size_t)kScanWarps * 256 + 16;
        int worst = 0;
        for (int L = 0; L < 4; ++L) worst = std::max(worst, (S.N[L] / 2 + (nctas - kDecCtas) - 1) / (nctas - kDecCtas));
        if (!off && need <= 216 * 1024 && worst <= kScanWarps) {
            wctas = nctas - kDecCtas;
            dec_smem = need;
        }
    }
    for (int L = 0; L < 4; ++L) S.pairs[L] = (S.N[L] / 2 + wctas - 1) / wctas;
    S.C = m.C; S.ksize = m.k; S.HW = HW; S.W_img = m.W; S.B = B;
    S.stage_cells = m.d_stage_cells.as<int2>();
    S.cell_hw = m.d_cell_hw.as<int32_t>();
    S.cell_tap = m.d_cell_tap.as<uint32_t>();
    S.cell_grp = m.d_cell_grp.as<uint32_t>();
    S.buf = buf; S.yhat_pm = yhat_pm; S.prior_pm = prior_pm;
    S.vec[0] = vec;
    S.vec[1] = S.vec[0] + (size_t)kBlkMaxRows * m.c_ctx;
    S.vec[2] = S.vec[1] + (size_t)kBlkMaxRows * m.c_m1;
    S.vec[3] = S.vec[2] + (size_t)kBlkMaxRows * m.c_m2;
    S.params = params;
    S.g0 = g0; S.g1 = g1;
    S.qz = RowsQuant{y, buf, sym, idx, d_scale_table, n_scales, m.C};
    S.dq_sym = dq_sym;
    S.step0 = m.scan_step;
    S.call_tag = m.scan_call;
    if (dec) {   // single-launch decoding: the coder's chunk warps run inside the kernel
        if (g0 != 0 || g1 != m.S || y || dq_sym) return value_error("in-kernel decoding covers the whole map in one launch");
        BASIC_TRY(m.scan_cs.reserve((size_t)m.S * sizeof(int32_t)));
        BASIC_CUDA(cudaMemcpyAsync(m.scan_cs.p, dec->chunk_syms, (size_t)m.S * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        const RansTables &tb = *dec->tables;
        S.dec.blob = tb.blob.as<unsigned char>();
        S.dec.blob_bytes = tb.blob_bytes;
        S.dec.meta_bytes = tb.meta_bytes;
        S.dctas = nctas - wctas;
        S.dec.cdf16_bytes = tb.cdf16_bytes;
        S.dec.T = tb.T;
        S.dec.precision = tb.precision;
        S.dec.bypass = dec->bypass;
        S.dec.seg = dec->seg;
        S.dec.seg_cap = dec->seg_cap;
        S.dec.seg_slices = m.S;
        S.dec.n_chunks = dec->n_chunks;
        S.dec.chunk_syms = m.scan_cs.as<int32_t>();
        S.dec.idx_t = idx_t;
        S.dec.status = dec->status;
    }
#ifdef SCAN_TIMING
    if (!m.scan_barrier.p) {
        BASIC_TRY(m.scan_barrier.reserve(256));
        BASIC_CUDA(cudaMemset(m.scan_barrier.p, 0, 256));
    }
    S.timing = m.scan_barrier.as<long long>();
#endif
    // The stage kernels need their whole grid resident (CTAs wait for each other's words).  Two of them from two streams could
    // each get a part of the SMs and wait forever, so launches on one device are chained through an event: the next one starts
    // when the previous one has finished, whatever streams they are on.
    static cudaEvent_t scan_done[64] = {};
    int cur_dev = 0;
    BASIC_CUDA(cudaGetDevice(&cur_dev));
    cudaEvent_t &ev_done = scan_done[cur_dev & 63];
    if (!ev_done) BASIC_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    else BASIC_CUDA(cudaStreamWaitEvent(stream, ev_done, 0));
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        BASIC_CUDA(cudaFuncSetAttribute(k_scan_stages<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        BASIC_CUDA(cudaFuncSetAttribute(k_scan_stages<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        BASIC_CUDA(cudaFuncSetAttribute(k_scan_blocks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        BASIC_CUDA(cudaFuncSetAttribute(k_scan_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    }
    if (CB > 0) {   // row blocks: the convolution's visible taps packed once per (weights, map)
        if (m.ws_ctxc_key != tap_union || !m.ws_ctxc.p) {
            BASIC_TRY(m.ws_ctxc.reserve((size_t)m.c_ctx * S.ntaps * m.C * sizeof(float) + 16));
            k_compact_taps<<<256, 256, 0, stream>>>(m.ws_ctx.as<float>(), m.ws_ctxc.as<float>(), m.c_ctx, m.C, m.k * m.k, S.ntaps, tap_union);
            BASIC_LAUNCHED();
            m.ws_ctxc_key = tap_union;
        }
        S.CB = CB;
        S.wc = m.ws_ctxc.as<float>();
        if (dec) k_scan_blocks<true><<<nctas, kBlkWarps * 32, blk_smem(m, S.ntaps), stream>>>(S);
        else k_scan_blocks<false><<<nctas, kBlkWarps * 32, blk_smem(m, S.ntaps), stream>>>(S);
    } else if (dec) {
        // (resident weights laid out for wctas owners)
        size_t fl = 0;
        int kmax = 0;
        for (int L = 0; L < 4; ++L) { fl += (size_t)S.pairs[L] * 2 * S.K[L]; kmax = std::max(kmax, S.K[L]); }
        const size_t w_smem = (fl + (size_t)kScanRows * kmax) * sizeof(float);
        if (std::max(w_smem, dec_smem) > 216 * 1024) return value_error("stage kernel: shared memory");
        k_scan_stages<true><<<nctas, kScanWarps * 32, std::max(w_smem, dec_smem), stream>>>(S);
    }
    else k_scan_stages<false><<<nctas, kScanWarps * 32, scan_smem(m, S.ntaps), stream>>>(S);
    BASIC_LAUNCHED();
    BASIC_CUDA(cudaEventRecord(ev_done, stream));
    m.scan_step += steps;
#ifdef SCAN_TIMING
    if (g1 - g0 > 4) {
        long long tk[10];
        BASIC_CUDA(cudaStreamSynchronize(stream));
        BASIC_CUDA(cudaMemcpy(tk, m.scan_barrier.as<long long>(), sizeof(tk), cudaMemcpyDeviceToHost));
        BASIC_CUDA(cudaMemset(m.scan_barrier.p, 0, 256));
        const long long ns = g1 - g0;
        FILE *tf = fopen("gpurun_out/scan_timing.txt", "a");
        if (!tf) tf = stderr;
        fprintf(tf, "scan cta 0: %lld stages, rows %d; cycles per stage: stage sync %lld  gather + wait %lld  multiply %lld  sync %lld  epilogue %lld\n",
                ns, B * m.max_stage_cells, tk[0] / ns, tk[1] / ns, tk[2] / ns, tk[3] / ns, tk[4] / ns);
        fprintf(tf, "   chunk 0's warp: operand wait + decode %lld cycles per stage, of which waiting for operands %lld\n", tk[8] / ns, tk[9] / ns);
        if (tf != stderr) fclose(tf);
    }
#endif
    return BASIC_OK;
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

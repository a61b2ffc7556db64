// Gaussian-conditional quantisation + scale index (reference: pgm_coder.py:802-821 _select_best_indexes,
// torch_ans.py:105-159 _data_preprocess "uniform" quantiser, pgm_coder.py:927-941 / :965-978 the per-group
// gather / scatter).  Elementwise: a thread takes four consecutive elements of the stream -- position list, symbols and
// indexes move with 128-bit accesses, one division per quad -- and gathers their parameters / latents.
#include "common.cuh"

namespace basic {

namespace {

// Where a scale sits in a log-spaced table, as a guess: index ~ (log2(sigma) - lg_min) * inv_step.  The search below starts there
// and walks to the exact answer, so a table that is not log-spaced only costs steps (inv_step = 0: a linear scan).
struct TableGuess {
    float lg_min, inv_step;
};

__device__ inline TableGuess table_guess(const float *__restrict__ tab, int n)
{
    TableGuess g = {0.f, 0.f};
    if (n >= 2 && tab[0] > 0.f && tab[n - 1] > tab[0]) {
        g.lg_min = __log2f(tab[0]);
        g.inv_step = (float)(n - 1) / (__log2f(tab[n - 1]) - g.lg_min);
    }
    return g;
}

// argmin_t |sigma - table[t]| in float32, first minimum (torch.argmin on CPU returns the first).
__device__ inline int scale_index(float sigma, const float *__restrict__ tab, int n, TableGuess tg)
{
    if (!(fabsf(sigma) <= 3.402823466e38f)) return 0;  // NaN / inf: every distance is NaN / inf -> index 0
    // lo = first t with tab[t] >= sigma (what a binary search returns): guessed, then corrected -- zero or one step each way
    // on the reference's table, six dependent probes saved
    int lo = 0;
    if (sigma > tab[0]) {
        lo = (int)((__log2f(sigma) - tg.lg_min) * tg.inv_step) + 1;
        lo = lo < 1 ? 1 : lo > n ? n : lo;
        while (lo > 0 && tab[lo - 1] >= sigma) --lo;
        while (lo < n && tab[lo] < sigma) ++lo;
    }
    if (lo == 0) return 0;
    if (lo == n) return n - 1;
    const float d0 = fabsf(__fsub_rn(sigma, tab[lo - 1])), d1 = fabsf(__fsub_rn(sigma, tab[lo]));
    return d0 <= d1 ? lo - 1 : lo;
}

// One coded element: where its parameters and its latent live.  Streams (symbols / indexes, and the position list) are
// contiguous: a thread takes FOUR consecutive elements and moves them with 128-bit accesses when they belong to one image.
struct ElemAddr {
    long long yo;     // offset of the latent in y / y_hat
    float mean, sigma;
};

// p / HW for 0 <= p < 2^31 with m = ceil(2^32 / HW): the high word of p * m overshoots by at most one
__device__ inline int div_hw(int p, int HW, unsigned m)
{
    unsigned q = __umulhi((unsigned)p, m);
    if (q * (unsigned)HW > (unsigned)p) --q;
    return (int)q;
}

__device__ inline ElemAddr elem_addr(const float *__restrict__ params, long long b, int p, int C, int HW, long long chw, int params_cl,
                                     const int32_t *__restrict__ perm, bool want_sigma, unsigned hw_magic)
{
    ElemAddr r;
    const int c = HW == 1 ? p : div_hw(p, HW, hw_magic);
    r.yo = b * chw + p;
    if (params_cl) {  // blocked channels-last parameters of the tensor path (ctx.cuh): (mean, scale) are neighbours
        const int hw = perm[p - c * HW];  // slot of the position
        const float2 ms = *reinterpret_cast<const float2 *>(
            params + ((((long long)b * ((HW + 31) >> 5) + (hw >> 5)) * (C >> 1) + (c >> 1)) * 32 + (hw & 31)) * 4 + (c & 1) * 2);
        r.mean = ms.x;
        r.sigma = ms.y;
    } else {
        const long long po = b * 2 * chw + p + (long long)c * HW;  // channel 2c (mean); scale is HW further
        r.mean = params[po];
        r.sigma = want_sigma ? params[po + HW] : 0.f;
    }
    return r;
}

__global__ void __launch_bounds__(256)
k_quantize_index(const float *__restrict__ y, const float *__restrict__ params, const int32_t *__restrict__ positions,
                 long long n_pos, int B, int C, int HW, const float *__restrict__ scale_table, int n_scales,
                 int32_t *__restrict__ symbols, int32_t *__restrict__ indexes, float *__restrict__ yhat, int params_cl,
                 const int32_t *__restrict__ perm)
{
    __shared__ float tab[256];
    for (int i = threadIdx.x; i < n_scales; i += blockDim.x) tab[i] = scale_table[i];
    __syncthreads();
    const TableGuess tg = table_guess(tab, n_scales);
    const unsigned hw_magic = HW > 1 ? (unsigned)((0x100000000ull + (unsigned)HW - 1) / (unsigned)HW) : 0u;
    const long long total = (long long)B * n_pos;
    const long long chw = (long long)C * HW;
    // quads never straddle two images when n_pos is a multiple of 4; the streams must be 16-byte aligned
    const bool vec = (n_pos & 3) == 0 && positions &&
                     ((reinterpret_cast<uintptr_t>(positions) | reinterpret_cast<uintptr_t>(indexes) | reinterpret_cast<uintptr_t>(symbols)) & 15) == 0;
    const long long quads = (total + 3) >> 2;
    const bool small = total < 0x7fffffffll;
    for (long long qd = blockIdx.x * (long long)blockDim.x + threadIdx.x; qd < quads; qd += (long long)gridDim.x * blockDim.x) {
        const long long e0 = qd << 2;
        const long long b = small ? (long long)((unsigned)e0 / (unsigned)n_pos) : e0 / n_pos;  // a 64-bit division is ~100 instructions
        const long long k0 = e0 - b * n_pos;
        int p[4];
        int32_t ix[4], sy[4];
        if (vec) {
            const int4 pp = __ldg(reinterpret_cast<const int4 *>(positions + k0));
            p[0] = pp.x; p[1] = pp.y; p[2] = pp.z; p[3] = pp.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const ElemAddr a = elem_addr(params, b, p[j], C, HW, chw, params_cl, perm, true, hw_magic);
                ix[j] = scale_index(a.sigma, tab, n_scales, tg);
                if (y) {
                    const float s = rintf(__fsub_rn(y[a.yo], a.mean));  // torch.round: half to even
                    sy[j] = (int32_t)s;
                    if (yhat) yhat[a.yo] = __fadd_rn(s, a.mean);
                }
            }
            *reinterpret_cast<int4 *>(indexes + e0) = make_int4(ix[0], ix[1], ix[2], ix[3]);
            if (y) *reinterpret_cast<int4 *>(symbols + e0) = make_int4(sy[0], sy[1], sy[2], sy[3]);
        } else {
            for (int j = 0; j < 4 && e0 + j < total; ++j) {
                const long long e = e0 + j, bb = e / n_pos, k = e - bb * n_pos;
                const int pj = positions ? positions[k] : (int)k;
                const ElemAddr a = elem_addr(params, bb, pj, C, HW, chw, params_cl, perm, true, hw_magic);
                indexes[e] = scale_index(a.sigma, tab, n_scales, tg);
                if (y) {
                    const float s = rintf(__fsub_rn(y[a.yo], a.mean));
                    symbols[e] = (int32_t)s;
                    if (yhat) yhat[a.yo] = __fadd_rn(s, a.mean);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_dequantize(const int32_t *__restrict__ symbols, const float *__restrict__ params, const int32_t *__restrict__ positions,
             long long n_pos, int B, int C, int HW, float *__restrict__ yhat, int params_cl, const int32_t *__restrict__ perm)
{
    const long long total = (long long)B * n_pos;
    const long long chw = (long long)C * HW;
    const unsigned hw_magic = HW > 1 ? (unsigned)((0x100000000ull + (unsigned)HW - 1) / (unsigned)HW) : 0u;
    const bool vec = (n_pos & 3) == 0 && positions &&
                     ((reinterpret_cast<uintptr_t>(positions) | reinterpret_cast<uintptr_t>(symbols)) & 15) == 0;
    const long long quads = (total + 3) >> 2;
    for (long long qd = blockIdx.x * (long long)blockDim.x + threadIdx.x; qd < quads; qd += (long long)gridDim.x * blockDim.x) {
        const long long e0 = qd << 2;
        if (vec) {
            const long long b = total < 0x7fffffffll ? (long long)((unsigned)e0 / (unsigned)n_pos) : e0 / n_pos, k0 = e0 - b * n_pos;
            const int4 pp = __ldg(reinterpret_cast<const int4 *>(positions + k0));
            const int4 ss = *reinterpret_cast<const int4 *>(symbols + e0);
            const int p[4] = {pp.x, pp.y, pp.z, pp.w}, sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const ElemAddr a = elem_addr(params, b, p[j], C, HW, chw, params_cl, perm, false, hw_magic);
                // pgm_coder.py:973-975 (sym + mean), then _data_postprocess x * 1 + 0 (turns -0.0 into +0.0)
                yhat[a.yo] = __fadd_rn(__fadd_rn((float)sv[j], a.mean), 0.0f);
            }
        } else {
            for (int j = 0; j < 4 && e0 + j < total; ++j) {
                const long long e = e0 + j, bb = e / n_pos, k = e - bb * n_pos;
                const int pj = positions ? positions[k] : (int)k;
                const ElemAddr a = elem_addr(params, bb, pj, C, HW, chw, params_cl, perm, false, hw_magic);
                yhat[a.yo] = __fadd_rn(__fadd_rn((float)symbols[e], a.mean), 0.0f);
            }
        }
    }
}

}  // namespace

int launch_quantize_index(const float *y, const float *params, const int32_t *positions, int64_t n_pos, int B, int C, int HW,
                          const float *d_scale_table, int n_scales, int32_t *symbols, int32_t *indexes, float *yhat,
                          int sm_count, cudaStream_t stream, int params_cl, const int32_t *perm)
{
    const long long total = (long long)B * n_pos;
    if (total == 0) return BASIC_OK;
    long long blocks = ((total + 3) / 4 + 255) / 256;   // four elements per thread
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    k_quantize_index<<<(int)blocks, 256, 0, stream>>>(y, params, positions, n_pos, B, C, HW, d_scale_table, n_scales, symbols,
                                                      indexes, yhat, params_cl, perm);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int launch_dequantize(const int32_t *symbols, const float *params, const int32_t *positions, int64_t n_pos, int B, int C,
                      int HW, float *yhat, int sm_count, cudaStream_t stream, int params_cl, const int32_t *perm)
{
    const long long total = (long long)B * n_pos;
    if (total == 0) return BASIC_OK;
    long long blocks = ((total + 3) / 4 + 255) / 256;   // four elements per thread
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    k_dequantize<<<(int)blocks, 256, 0, stream>>>(symbols, params, positions, n_pos, B, C, HW, yhat, params_cl, perm);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

// Gaussian-conditional quantisation + scale index (reference: pgm_coder.py:802-821 _select_best_indexes,
// torch_ans.py:105-159 _data_preprocess "uniform" quantiser, pgm_coder.py:927-941 / :965-978 the per-group
// gather / scatter).  Elementwise, HBM-bound: one thread per coded element, coalesced along the position list.
#include "common.cuh"

namespace basic {

namespace {

// argmin_t |sigma - table[t]| in float32, first minimum (torch.argmin on CPU returns the first).
__device__ inline int scale_index(float sigma, const float *__restrict__ tab, int n)
{
    if (!(fabsf(sigma) <= 3.402823466e38f)) return 0;  // NaN / inf: every distance is NaN / inf -> index 0
    int lo = 0, hi = n;                                // first t with tab[t] >= sigma
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] < sigma) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == n) return n - 1;
    const float d0 = fabsf(__fsub_rn(sigma, tab[lo - 1])), d1 = fabsf(__fsub_rn(sigma, tab[lo]));
    return d0 <= d1 ? lo - 1 : lo;
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
    const long long total = (long long)B * n_pos;
    const long long chw = (long long)C * HW;
    const bool small = total < 0x7fffffffll && (long long)B * 2 * chw < 0x7fffffffll;   // 32-bit divisions (a 64-bit one is ~100 instructions)
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long b, k;
        if (small) {
            const unsigned bb = (unsigned)e / (unsigned)n_pos;
            b = bb;
            k = (unsigned)e - bb * (unsigned)n_pos;
        } else {
            b = e / n_pos;
            k = e - b * n_pos;
        }
        const int p = positions ? positions[k] : (int)k;
        const int c = p / HW;
        float mean, sigma;
        if (params_cl) {  // blocked channels-last parameters of the tensor path (ctx.cuh): (mean, scale) are neighbours
            const int hw = perm[p - c * HW];  // slot of the position
            const float2 ms = *reinterpret_cast<const float2 *>(
                params + ((((long long)b * ((HW + 31) >> 5) + (hw >> 5)) * (C >> 1) + (c >> 1)) * 32 + (hw & 31)) * 4 + (c & 1) * 2);
            mean = ms.x;
            sigma = ms.y;
        } else {
            const long long po = b * 2 * chw + p + (long long)c * HW;  // channel 2c (mean); scale is HW further
            mean = params[po];
            sigma = params[po + HW];
        }
        indexes[e] = scale_index(sigma, tab, n_scales);
        if (y) {
            const float s = rintf(__fsub_rn(y[b * chw + p], mean));  // torch.round: half to even
            symbols[e] = (int32_t)s;
            if (yhat) yhat[b * chw + p] = __fadd_rn(s, mean);
        }
    }
}

__global__ void __launch_bounds__(256)
k_dequantize(const int32_t *__restrict__ symbols, const float *__restrict__ params, const int32_t *__restrict__ positions,
             long long n_pos, int B, int C, int HW, float *__restrict__ yhat, int params_cl, const int32_t *__restrict__ perm)
{
    const long long total = (long long)B * n_pos;
    const long long chw = (long long)C * HW;
    const bool small = total < 0x7fffffffll;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long b, k;
        if (small) {
            const unsigned bb = (unsigned)e / (unsigned)n_pos;
            b = bb;
            k = (unsigned)e - bb * (unsigned)n_pos;
        } else {
            b = e / n_pos;
            k = e - b * n_pos;
        }
        const int p = positions ? positions[k] : (int)k;
        const int c = p / HW;
        const int hw = params_cl ? perm[p - c * HW] : 0;  // slot of the position
        const float mean = params_cl ? params[((((long long)b * ((HW + 31) >> 5) + (hw >> 5)) * (C >> 1) + (c >> 1)) * 32 + (hw & 31)) * 4 + (c & 1) * 2]
                                     : params[b * 2 * chw + p + (long long)c * HW];
        // pgm_coder.py:973-975 (sym + mean), then _data_postprocess x * 1 + 0 (turns -0.0 into +0.0)
        yhat[b * chw + p] = __fadd_rn(__fadd_rn((float)symbols[e], mean), 0.0f);
    }
}

}  // namespace

int launch_quantize_index(const float *y, const float *params, const int32_t *positions, int64_t n_pos, int B, int C, int HW,
                          const float *d_scale_table, int n_scales, int32_t *symbols, int32_t *indexes, float *yhat,
                          int sm_count, cudaStream_t stream, int params_cl, const int32_t *perm)
{
    const long long total = (long long)B * n_pos;
    if (total == 0) return BASIC_OK;
    long long blocks = (total + 255) / 256;
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
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    k_dequantize<<<(int)blocks, 256, 0, stream>>>(symbols, params, positions, n_pos, B, C, HW, yhat, params_cl, perm);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

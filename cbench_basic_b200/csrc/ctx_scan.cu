// Many-stage maps (scanline, zigzag, the serial JointAR coder): the context model walked stage by stage INSIDE one persistent
// kernel -- exact FP32, one channel group.  k_scan_stages (<= 4 rows per stage: every weight resident in shared memory),
// k_scan_blocks (5 .. 128 rows: row blocks x channel blocks, weights streamed once per stage), the in-kernel chunk decoder of the
// decoder's single launch, and the host side that picks and launches them (ctx_scan_supported / ctx_scan_run).
// Reference semantics: cbench/nn/layers/masked_conv.py:102-228, :287-305 evaluated once per coding group by
// pgm_coder.py:912-981; DESIGN.md section 5.3 has the measurements behind every design step.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "ctx.cuh"
#include "rans_lanes.cuh"

namespace basic {

namespace {

constexpr float kSlope = 0.01f;  // nn.LeakyReLU default negative_slope

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

}  // namespace

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

}  // namespace basic

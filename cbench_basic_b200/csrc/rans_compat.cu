// lanes = 1 compatibility coder: reproduces the reference rANS64 bitstream byte for byte
// (cbench/csrc/ans/rans64.cpp:203-361 encode, :389-598 decode; ryg rans64.h:59-142).
// The state recurrence is strictly serial, so one thread walks the stream while the rest of the CTA does
// the parallel part: symbol -> (start, freq, exact reciprocal) lookup and operand staging through shared
// memory.  This mode exists for bitstream parity; the throughput path is rans_lanes.cu.
#include "common.cuh"

namespace basic {

namespace {

constexpr int kTile = 2048;     // decode staging tile (2 x 8 KB of shared memory)
constexpr int kEncTile = 1024;  // encode staging tile (32 KB of EncSym)
constexpr int kThreads = 256;
constexpr unsigned long long kL64 = 1ull << 31;

struct EncSym {            // per-symbol operands prepared in parallel
    unsigned long long rcp;  // Alverson reciprocal (rans64.h:167-245)
    uint32_t bias, cmpl;     // x_new = x + bias + q * cmpl
    uint32_t freq;           // for x_max
    uint32_t shift_esc;      // bits 0..7 rcp_shift, bit 8 escape flag
    uint32_t raw;            // escape payload
    uint32_t pad;
};

__device__ inline void enc_sym_init(EncSym &s, uint32_t start, uint32_t freq, uint32_t prec)
{
    s.freq = freq;
    s.cmpl = (1u << prec) - freq;
    if (freq < 2) {
        s.rcp = ~0ull;
        s.shift_esc = 0;
        s.bias = start + (1u << prec) - 1;
    } else {
        uint32_t shift = 0;
        while (freq > (1u << shift)) shift++;
        unsigned long long x0 = freq - 1, x1 = 1ull << (shift + 31);
        unsigned long long t1 = x1 / freq;
        x0 += (x1 % freq) << 32;
        unsigned long long t0 = x0 / freq;
        s.rcp = t0 + (t1 << 32);
        s.shift_esc = shift - 1;
        s.bias = start;
    }
}

// Tables of the lanes=1 kernels: staged into dynamic shared memory when they fit (the serial thread's chain is
// record -> LUT -> CDF -> CDF, four dependent loads per symbol: out of L2 that was most of the decoder's time).
__device__ inline TableView stage_compat_tables(const void *blob, size_t blob_bytes, size_t meta_bytes, size_t cdf16_bytes,
                                                int to_smem, unsigned char *dyn)
{
    if (!to_smem) return make_view(blob, meta_bytes, cdf16_bytes);
    const uint4 *src = reinterpret_cast<const uint4 *>(blob);
    uint4 *dst = reinterpret_cast<uint4 *>(dyn);
    for (size_t i = threadIdx.x; i < blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return make_view(dyn, meta_bytes, cdf16_bytes);
}

// One CTA per stream (blockIdx.x = stream b of a batch of equally long streams: the z node codes one stream per image).
// Words are written back-to-front into out_words[b * cap_words .. (b + 1) * cap_words); first_word[b] receives the index
// (inside that region) of the first word, status: bit0 = index out of range, bit1 = symbol out of range with bypass off,
// bit2 = capacity exceeded.
__global__ void __launch_bounds__(kThreads)
k_rans64_encode(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, long long n,
                const void *__restrict__ blob, size_t blob_bytes, size_t meta_bytes, size_t cdf16_bytes, int tables_in_smem, int T,
                int precision, int bypass, int bypass_precision, uint32_t *__restrict__ out_words, long long cap_words,
                long long *first_word, int *status)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    symbols += (long long)blockIdx.x * n;
    indexes += (long long)blockIdx.x * n;
    out_words += (long long)blockIdx.x * cap_words;
    first_word += blockIdx.x;
    __shared__ EncSym sm[kEncTile];
    const TableView tv = stage_compat_tables(blob, blob_bytes, meta_bytes, cdf16_bytes, tables_in_smem, dyn_smem);
    const int tid = threadIdx.x;
    unsigned long long x = kL64;
    long long p = cap_words;
    int st = 0;
    const uint32_t bp = (uint32_t)bypass_precision, maxb = (1u << bp) - 1;
    const unsigned long long xmax_bits = ((kL64 >> 16) << 32) * (unsigned long long)(1u << (16 - bp));  // rans64.cpp:37-38
    for (long long hi = n; hi > 0; hi -= kEncTile) {
        const long long lo = hi - kEncTile > 0 ? hi - kEncTile : 0;
        const int cnt = (int)(hi - lo);
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) {
            int32_t c = indexes[lo + k];
            if ((uint32_t)c >= (uint32_t)T) { st |= 1; c = 0; }
            const TableMeta m = tv.meta[c];
            const int32_t maxv = (int32_t)m.cdf_size - 2;
            int32_t v = symbols[lo + k] - m.offset;
            uint32_t raw = 0, esc = 0;
            if (bypass) {
                if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = maxv; }
                else if (v >= maxv) { raw = (uint32_t)(2 * (v - maxv)); v = maxv; }
                esc = v == maxv;
            } else if (v < 0 || v > maxv) { st |= 2; v = 0; }
            const uint32_t start = tv.cdf[m.cdf_base + v];
            const uint32_t freq = (uint16_t)(tv.cdf[m.cdf_base + v + 1] - start);
            EncSym s;
            enc_sym_init(s, start, freq, (uint32_t)precision);
            s.shift_esc |= esc << 8;
            s.raw = raw;
            s.pad = 0;
            sm[k] = s;
        }
        __syncthreads();
        if (tid == 0) {
            EncSym nxt = sm[cnt - 1];  // operands of the next symbol are fetched while the current one is coded
            for (int k = cnt - 1; k >= 0; --k) {
                const EncSym s = nxt;
                if (k > 0) nxt = sm[k - 1];
                if (s.shift_esc & 0x100) {  // escape tokens, pushed last to first (rans64.cpp:293-335)
                    const uint32_t raw = s.raw;
                    int nd = 0;
                    while (nd * (int)bp < 32 && (raw >> (nd * bp)) != 0) ++nd;
                    const int ncnt = nd / (int)maxb + 1, ntok = ncnt + nd;
                    for (int u = ntok - 1; u >= 0; --u) {
                        uint32_t tok;
                        if (u >= ncnt) tok = (raw >> ((u - ncnt) * bp)) & maxb;
                        else tok = u < ncnt - 1 ? maxb : (uint32_t)(nd - (ncnt - 1) * (int)maxb);
                        if (x >= xmax_bits) {
                            if (p > 0) out_words[--p] = (uint32_t)x; else st |= 4;
                            x >>= 32;
                        }
                        x = (x << bp) | tok;
                    }
                }
                const unsigned long long x_max = (unsigned long long)s.freq << (63 - precision);  // ((L >> prec) << 32) * freq
                if (x >= x_max) {
                    if (p > 0) out_words[--p] = (uint32_t)x; else st |= 4;
                    x >>= 32;
                }
                const unsigned long long q = __umul64hi(x, s.rcp) >> (s.shift_esc & 0xff);
                x = x + s.bias + q * s.cmpl;
            }
        }
    }
    if (st) atomicOr(status, st);
    if (tid == 0) {
        if (p >= 2) {
            p -= 2;
            out_words[p] = (uint32_t)x;
            out_words[p + 1] = (uint32_t)(x >> 32);
        } else atomicOr(status, 4);
        *first_word = p;
    }
}

struct DecState {
    unsigned long long x;
    long long pos;  // next word to read
};

// rans64.cpp:49-65 with the next word of the stream held in a register (`nw` = words[pos], fetched when the previous one
// was consumed): the serial thread never waits for a global load inside a renormalisation.
__device__ inline void pull_word(unsigned long long &x, const uint32_t *__restrict__ w, long long &pos, long long nwords,
                                 uint32_t &nw, int &st)
{
    if (pos >= nwords) st |= 4;
    x = (x << 32) | nw;
    ++pos;
    nw = pos < nwords ? w[pos] : 0u;
}

__device__ inline uint32_t getbits64(unsigned long long &x, const uint32_t *__restrict__ w, long long &pos,
                                     long long nwords, uint32_t nb, uint32_t &nw, int &st)
{
    const uint32_t v = (uint32_t)(x & ((1u << nb) - 1));
    x >>= nb;
    if (x < kL64) pull_word(x, w, pos, nwords, nw, st);
    return v;
}

// One CTA per stream: thread 0 walks the stream, everybody stages indexes in / symbols out through shared memory.
// Batch of equally long symbol runs (blockIdx.x = b): stream b is words[stream_off[b] .. + stream_nwords[b]) (both arrays
// NULL for a single stream of `nwords` words), its state state[b], its operands indexes / out + b * n.
// The serial chain per symbol is  LUT -> four CDF entries at once -> 64-bit multiply-add -> (word from a register);  the
// table record of the NEXT symbol is fetched while the current one is decoded (common.cuh Tab<SM>, ld.shared when the
// image fits).
template <bool SM>
__global__ void __launch_bounds__(kThreads)
k_rans64_decode(const uint32_t *__restrict__ words, long long nwords, const long long *__restrict__ stream_off,
                const long long *__restrict__ stream_nwords, DecState *state, int init_state,
                const int32_t *__restrict__ indexes, long long n, const void *__restrict__ blob, size_t blob_bytes, size_t meta_bytes,
                size_t cdf16_bytes, int T, int precision, int bypass, int bypass_precision, int32_t *__restrict__ out, int *status)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    if (stream_off) { words += stream_off[blockIdx.x]; nwords = stream_nwords[blockIdx.x]; }
    state += blockIdx.x;
    indexes += (long long)blockIdx.x * n;
    out += (long long)blockIdx.x * n;
    __shared__ int32_t s_idx[kTile];
    __shared__ int32_t s_out[kTile];
    if (SM) {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(dyn_smem);
        for (size_t i = threadIdx.x; i < blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    Tab<SM> tb;
    tb.init(blob, dyn_smem, meta_bytes, cdf16_bytes);
    const int tid = threadIdx.x;
    unsigned long long x = 0;
    long long pos = 0;
    uint32_t nw = 0;
    int st = 0;
    if (tid == 0) {
        if (init_state) {  // set_stream: rans64.hpp:104-111, Rans64DecInit rans64.h:106-115
            if (nwords >= 2) { x = (unsigned long long)words[0] | ((unsigned long long)words[1] << 32); pos = 2; }
            else st |= 4;
        } else { x = state->x; pos = state->pos; }
        nw = pos < nwords ? words[pos] : 0u;
    }
    const uint32_t bp = (uint32_t)bypass_precision, maxb = (1u << bp) - 1;
    const uint32_t pmask = (1u << precision) - 1;
    for (long long lo = 0; lo < n; lo += kTile) {
        const int cnt = (int)(n - lo < kTile ? n - lo : kTile);
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) {
            int32_t c = indexes[lo + k];
            if ((uint32_t)c >= (uint32_t)T) { st |= 1; c = 0; }
            s_idx[k] = c;
        }
        __syncthreads();
        if (tid == 0) {
            uint4 rec = tb.meta_at(s_idx[0]);  // cdf_base | lut_base | cdf_size, lut_shift | offset
            for (int k = 0; k < cnt; ++k) {
                const uint4 m = rec;
                if (k + 1 < cnt) rec = tb.meta_at(s_idx[k + 1]);
                const typename Tab<SM>::addr_t cd = tb.cdf_at(m.x);
                const int nsyms = (int)(m.z & 0xffffu) - 1, maxv = nsyms - 1;
                const uint32_t cum = (uint32_t)x & pmask;
                int s = (int)Tab<SM>::template ld16<0>(tb.lut_at(m.y + (cum >> ((m.z >> 16) & 0xffu))));
                const typename Tab<SM>::addr_t e = cd + 2 * s;
                const uint32_t c0 = Tab<SM>::template ld16<0>(e), c1 = Tab<SM>::template ld16<2>(e);
                const uint32_t c2 = Tab<SM>::template ld16<4>(e), c3 = Tab<SM>::template ld16<6>(e);
                const bool a1 = s + 1 < nsyms && c1 <= cum;
                const bool a2 = a1 && s + 2 < nsyms && c2 <= cum;
                const bool a3 = a2 && s + 3 < nsyms && c3 <= cum;
                uint32_t start = a2 ? c2 : a1 ? c1 : c0, next = a2 ? c3 : a1 ? c2 : c1;
                s += (int)a1 + (int)a2;
                if (a3) {  // rans64.cpp:456-460 walks the table linearly; so does a bucket of width-1 symbols
                    ++s;
                    while (s + 1 < nsyms && Tab<SM>::template ld16<2>(cd + 2 * s) <= cum) ++s;
                    start = Tab<SM>::template ld16<0>(cd + 2 * s);
                    next = Tab<SM>::template ld16<2>(cd + 2 * s);
                }
                const uint32_t freq = (uint16_t)(next - start);
                x = (unsigned long long)freq * (x >> precision) + cum - start;
                if (x < kL64) pull_word(x, words, pos, nwords, nw, st);
                int32_t value = s;
                if (bypass && s == maxv) {
                    uint32_t val = getbits64(x, words, pos, nwords, bp, nw, st), nb = val;
                    while (val == maxb && nb < 64) { val = getbits64(x, words, pos, nwords, bp, nw, st); nb += val; }
                    uint32_t raw = 0;
                    for (uint32_t j = 0; j < nb; ++j) {
                        val = getbits64(x, words, pos, nwords, bp, nw, st);
                        if (j * bp < 32) raw |= val << (j * bp);
                    }
                    value = (int32_t)(raw >> 1);
                    value = (raw & 1) ? -value - 1 : value + maxv;
                }
                s_out[k] = value + (int32_t)m.w;
            }
        }
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) out[lo + k] = s_out[k];
    }
    if (st) atomicOr(status, st);
    if (tid == 0) { state->x = x; state->pos = pos; }
}

// ---- in-coder autoregressive table lookup (ans_interface.hpp:58-105, the table branch; rans64.cpp:259-263, :439-443) ----
// The table an element is coded with is  ar_table[ar_index][index][v0]( [v1] )  with  v_k = off_k[i] > 0 ? symbol[i - off_k[i]] + 1 : 0:
// it depends on up to two earlier symbols of the same array.  The encoder knows every symbol: a parallel prepass turns
// (index, neighbours) into effective table indexes.  The decoder learns the neighbours one by one: the serial thread of the
// lanes = 1 decoder does the lookup itself, symbol by symbol.
struct ArParams {
    const int32_t *table;      // [A][I][D1] or [A][I][D1][D2]
    int A, I, D1, D2;          // D2 = 0: one neighbour
    const int32_t *ar_indexes; // [n] or NULL (= 0)
    const int32_t *off0, *off1;// [n] distances back (off1 NULL with one neighbour)
};

__device__ inline int32_t ar_lookup(const ArParams &P, int32_t index, long long i, const int32_t *__restrict__ symbols, int &st)
{
    const int32_t a = P.ar_indexes ? P.ar_indexes[i] : 0;
    const int32_t o0 = P.off0[i];
    int32_t v0 = o0 > 0 ? (i - o0 >= 0 ? symbols[i - o0] + 1 : (st |= 1, 0)) : 0;
    int32_t v1 = 0;
    if (P.D2) {
        const int32_t o1 = P.off1[i];
        v1 = o1 > 0 ? (i - o1 >= 0 ? symbols[i - o1] + 1 : (st |= 1, 0)) : 0;
    }
    // the reference indexes its nested vectors unchecked (undefined behaviour out of range): here that is an error
    if ((uint32_t)a >= (uint32_t)P.A || (uint32_t)index >= (uint32_t)P.I || (uint32_t)v0 >= (uint32_t)P.D1 ||
        (P.D2 && (uint32_t)v1 >= (uint32_t)P.D2)) { st |= 1; return 0; }
    const long long at = ((long long)a * P.I + index) * P.D1 + v0;
    return P.D2 ? P.table[at * P.D2 + v1] : P.table[at];
}

__global__ void __launch_bounds__(256)
k_ar_effective_indexes(ArParams P, const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, long long n,
                       int32_t *__restrict__ out, int *status)
{
    int st = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = ar_lookup(P, indexes[i], i, symbols, st);
    if (st) atomicOr(status, st);
}

// lanes = 1 decoder with the lookup inside the chain: ONE thread (the dependency is symbol to symbol); tables from L2.
__global__ void __launch_bounds__(32)
k_rans64_decode_ar(const uint32_t *__restrict__ words, long long nwords, ArParams P, const int32_t *__restrict__ indexes, long long n,
                   const void *__restrict__ blob, size_t meta_bytes, size_t cdf16_bytes, int T, int precision, int bypass,
                   int bypass_precision, int32_t *__restrict__ out, int *status)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const TableView tv = make_view(blob, meta_bytes, cdf16_bytes);
    int st = 0;
    unsigned long long x = 0;
    long long pos = 0;
    if (nwords >= 2) { x = (unsigned long long)words[0] | ((unsigned long long)words[1] << 32); pos = 2; } else st |= 4;
    uint32_t nw = pos < nwords ? words[pos] : 0u;
    const uint32_t bp = (uint32_t)bypass_precision, maxb = (1u << bp) - 1, pmask = (1u << precision) - 1;
    for (long long i = 0; i < n; ++i) {
        int32_t c = ar_lookup(P, indexes[i], i, out, st);
        if ((uint32_t)c >= (uint32_t)T) { st |= 1; c = 0; }
        const TableMeta m = tv.meta[c];
        const int nsyms = (int)m.cdf_size - 1, maxv = nsyms - 1;
        const uint16_t *cdf = tv.cdf + m.cdf_base;
        const uint32_t cum = (uint32_t)x & pmask;
        int s2 = (int)tv.lut[m.lut_base + (cum >> m.lut_shift)];
        while (s2 + 1 < nsyms && cdf[s2 + 1] <= cum) ++s2;   // (entries 1 .. nsyms - 1 are below 2^16; the total, stored as 0, is never probed)
        const uint32_t start = cdf[s2], freq = (uint16_t)(cdf[s2 + 1] - start);
        x = (unsigned long long)freq * (x >> precision) + cum - start;
        if (x < kL64) pull_word(x, words, pos, nwords, nw, st);
        int32_t value = s2;
        if (bypass && s2 == maxv) {
            uint32_t val = getbits64(x, words, pos, nwords, bp, nw, st), nb = val;
            while (val == maxb && nb < 64) { val = getbits64(x, words, pos, nwords, bp, nw, st); nb += val; }
            uint32_t raw = 0;
            for (uint32_t j = 0; j < nb; ++j) {
                val = getbits64(x, words, pos, nwords, bp, nw, st);
                if (j * bp < 32) raw |= val << (j * bp);
            }
            value = (int32_t)(raw >> 1);
            value = (raw & 1) ? -value - 1 : value + maxv;
        }
        out[i] = value + m.offset;
    }
    if (st) atomicOr(status, st);
}

}  // namespace

int launch_ar_effective_indexes(const int32_t *d_table, int A, int I, int D1, int D2, const int32_t *d_ar_idx, const int32_t *d_off0,
                                const int32_t *d_off1, const int32_t *d_sym, const int32_t *d_idx, int64_t n, int32_t *d_out,
                                int *d_status, cudaStream_t stream)
{
    if (n <= 0) return BASIC_OK;
    const ArParams P = {d_table, A, I, D1, D2, d_ar_idx, d_off0, d_off1};
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    k_ar_effective_indexes<<<blocks, 256, 0, stream>>>(P, d_sym, d_idx, n, d_out, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int launch_rans64_decode_ar(const RansTables &tb, const uint32_t *d_words, int64_t nwords, const int32_t *d_table, int A, int I, int D1,
                            int D2, const int32_t *d_ar_idx, const int32_t *d_off0, const int32_t *d_off1, const int32_t *d_idx,
                            int64_t n, int bypass, int bypass_precision, int32_t *d_out, int *d_status, cudaStream_t stream)
{
    const ArParams P = {d_table, A, I, D1, D2, d_ar_idx, d_off0, d_off1};
    k_rans64_decode_ar<<<1, 32, 0, stream>>>(d_words, nwords, P, d_idx, n, tb.blob.p, tb.meta_bytes, tb.cdf16_bytes, tb.T, tb.precision,
                                             bypass, bypass_precision, d_out, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// the table image goes to shared memory when it fits beside the kernel's static tiles
static int compat_table_smem(const RansTables &tb, size_t static_bytes)
{
    return tb.blob_bytes + static_bytes + 2048 <= 227 * 1024 ? (int)tb.blob_bytes : 0;
}
static int compat_attrs()
{
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        BASIC_CUDA(cudaFuncSetAttribute(k_rans64_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048 - (int)(sizeof(EncSym) * kEncTile)));
        BASIC_CUDA(cudaFuncSetAttribute(k_rans64_decode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048 - (int)(2 * sizeof(int32_t) * kTile)));
    }
    return BASIC_OK;
}

// n_streams equally long runs of n symbols each, coded as independent streams by one CTA each; stream b's words end at
// d_words + (b + 1) * cap_words and start at d_first[b] inside its region.
int launch_rans64_encode(const RansTables &tb, const int32_t *d_sym, const int32_t *d_idx, int64_t n, int bypass,
                         int bypass_precision, uint32_t *d_words, int64_t cap_words, long long *d_first, int *d_status,
                         cudaStream_t stream, int n_streams)
{
    BASIC_TRY(compat_attrs());
    if (n_streams < 1) return BASIC_OK;
    const int smem = compat_table_smem(tb, sizeof(EncSym) * kEncTile);
    k_rans64_encode<<<n_streams, kThreads, smem, stream>>>(d_sym, d_idx, n, tb.blob.p, tb.blob_bytes, tb.meta_bytes, tb.cdf16_bytes,
                                                           smem > 0, tb.T, tb.precision, bypass, bypass_precision, d_words,
                                                           cap_words, d_first, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// d_off / d_nwords: per-stream word offsets and lengths of a batch (device arrays), or NULL for one stream of `nwords`.
int launch_rans64_decode(const RansTables &tb, const uint32_t *d_words, int64_t nwords, void *d_state, int init_state,
                         const int32_t *d_idx, int64_t n, int bypass, int bypass_precision, int32_t *d_out, int *d_status,
                         cudaStream_t stream, int n_streams, const long long *d_off, const long long *d_nwords)
{
    BASIC_TRY(compat_attrs());
    if (n_streams < 1) return BASIC_OK;
    const int smem = compat_table_smem(tb, 2 * sizeof(int32_t) * kTile);
    auto kern = smem > 0 ? k_rans64_decode<true> : k_rans64_decode<false>;
    kern<<<n_streams, kThreads, smem, stream>>>(d_words, nwords, d_off, d_nwords, reinterpret_cast<DecState *>(d_state), init_state, d_idx,
                                                n, tb.blob.p, tb.blob_bytes, tb.meta_bytes, tb.cdf16_bytes, tb.T, tb.precision, bypass,
                                                bypass_precision, d_out, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

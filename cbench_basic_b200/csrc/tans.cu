// tANS (FSE-derived) coder, lanes = 1: table construction and the reference bitstream on the device
// (reference cbench/csrc/ans/tans.cpp:17-368 tables, :534-690 encode, :722-822 decode;
//  csrc/FSE/bitstream.h:185-400 bit IO).  The table build and the state walk are serial per table / per
// stream, exactly as in the reference; the parallel part (operand staging, escape mapping) is done by the
// rest of the CTA.  tANS is not on the BaSIC default path (freq_precision = 16 > TANS_MAX_TABLELOG, SURVEY
// hard part 8); it is here for API completeness of the cbench.ans replacement.
#include "common.cuh"

namespace basic {

struct TansDEntry { uint32_t newState; uint16_t symbol; uint16_t nbBits; };  // tans.hpp:101-106

struct TansTables {
    int T = 0, tableLog = 11, bypass = 0, bypass_precision = 4, max_nsym = 0, role = 0;
    bool has_c = false, has_d = false;
    DevBuf nsym, offsets, ct_state, ct_nb, ct_fs, dt, dt_fast, scratch, err;
    // bypass tables are stored as table index T
};

namespace {

__device__ inline unsigned hb32(uint32_t v) { return 31 - __clz(v); }

// tans.cpp:26-100
__device__ int normalize_m2(short *norm, uint32_t tableLog, const uint32_t *count, unsigned long long total, uint32_t maxSV)
{
    uint32_t distributed = 0, toDistribute;
    const uint32_t lowThreshold = (uint32_t)(total >> tableLog);
    uint32_t lowOne = (uint32_t)((total * 3) >> (tableLog + 1));
    for (uint32_t s = 0; s <= maxSV; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = -1; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = -2;
    }
    toDistribute = (1u << tableLog) - distributed;
    if ((total / toDistribute) > lowOne) {
        lowOne = (uint32_t)((total * 3) / (toDistribute * 2));
        for (uint32_t s = 0; s <= maxSV; s++)
            if (norm[s] == -2 && count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; }
        toDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSV + 1) {
        uint32_t maxV = 0, maxC = 0;
        for (uint32_t s = 0; s <= maxSV; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (short)toDistribute;
        return 0;
    }
    const unsigned long long vStepLog = 62 - tableLog, mid = (1ull << (vStepLog - 1)) - 1;
    const unsigned long long rStep = ((((unsigned long long)1 << vStepLog) * toDistribute) + mid) / total;
    unsigned long long tmpTotal = mid;
    for (uint32_t s = 0; s <= maxSV; s++) {
        if (norm[s] == -2) {
            const unsigned long long end = tmpTotal + (count[s] * rStep);
            const uint32_t weight = (uint32_t)(end >> vStepLog) - (uint32_t)(tmpTotal >> vStepLog);
            if (weight < 1) return 1;
            norm[s] = (short)weight;
            tmpTotal = end;
        }
    }
    return 0;
}

// tans.cpp:17-23,102-148
__device__ int normalize_count(short *norm, uint32_t tableLog, const uint32_t *count, int nsym)
{
    const uint32_t rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const uint32_t maxSV = (uint32_t)nsym - 1;
    unsigned long long total = 0;
    for (int i = 0; i < nsym; ++i) total += count[i];
    if (total == 0) return 1;
    {
        const uint32_t a = hb32((uint32_t)(total - 1)) + 1, b = hb32(maxSV) + 2;
        if (tableLog < (a < b ? a : b)) return 1;
    }
    const unsigned long long scale = 62 - tableLog, step = ((unsigned long long)1 << 62) / total, vStep = 1ull << (scale - 20);
    int still = 1 << tableLog;
    uint32_t largest = 0;
    short largestP = 0;
    const uint32_t lowThreshold = (uint32_t)(total >> tableLog);
    for (uint32_t s = 0; s <= maxSV; s++) {
        if (count[s] == total) return 1;  // reference "rle special case" leaves norm undefined; refused here
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = -1; still--; }
        else {
            short proba = (short)((count[s] * step) >> scale);
            if (proba < 8) {
                const unsigned long long restToBeat = vStep * rtb[proba];
                proba += (count[s] * step) - ((unsigned long long)proba << scale) > restToBeat;
            }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba;
            still -= proba;
        }
    }
    if (-still >= (norm[largest] >> 1)) return normalize_m2(norm, tableLog, count, total, maxSV);
    norm[largest] += (short)still;
    return 0;
}

struct BuildArgs {
    const int32_t *freqs; int M; const int32_t *nsym; int T, tableLog, max_nsym, bypass, nbypass, build_c, build_d;
    uint16_t *ct_state; uint32_t *ct_nb; int32_t *ct_fs; TansDEntry *dt; int32_t *dt_fast;
    uint32_t *scratch;  // per table: count[max_nsym] | cumul[max_nsym + 2] | tableSymbol[tsz] (u32 each) | norm
    int *err;
};

// One CTA per table (table index T is the uniform bypass table).  The count normalisation (tans.cpp:102-148, a few hundred
// sequential operations) stays on one thread; the spread and the table fills are parallel:
//  * spread (tans.cpp:176-190 / :288-300): position j * step mod size for j = 0, 1, ... is a permutation (step is odd), and the
//    reference skips the positions above `high` (reserved for the low-probability symbols): the k-th symbol occurrence lands on
//    the k-th j whose position is <= high.  A block-wide prefix sum over "position j is usable" gives every j its k, a binary
//    search in the cumulative counts its symbol.
//  * C- and D-table fill: the reference walks the states u in increasing order and gives each symbol's occurrences consecutive
//    slots (stateTable[cumul[s]++], symbolNext[s]++): one thread per symbol walks the spread table and numbers its own.
constexpr int kBuildThreads = 256;

__global__ void __launch_bounds__(kBuildThreads)
k_tans_build(BuildArgs a)
{
    const int t = blockIdx.x, tid = threadIdx.x;
    const int tsz = 1 << a.tableLog;
    const int n = t < a.T ? a.nsym[t] : a.nbypass;
    if (t == a.T && !a.bypass) return;
    const size_t per = (size_t)a.max_nsym * 3 + 2 + tsz + 8;
    uint32_t *count = a.scratch + (size_t)t * per;
    uint32_t *cumul = count + a.max_nsym;
    uint32_t *tableSymbol = cumul + a.max_nsym + 2;
    short *norm = reinterpret_cast<short *>(tableSymbol + tsz);
    __shared__ int s_err;
    __shared__ uint32_t s_high, s_scan[kBuildThreads];
    for (int i = tid; i < n; i += kBuildThreads) count[i] = t < a.T ? (uint32_t)a.freqs[(size_t)t * a.M + i] : 1u;
    if (tid == 0) s_err = 0;
    __syncthreads();
    const uint32_t tableLog = (uint32_t)a.tableLog, tableSize = (uint32_t)tsz, tableMask = tableSize - 1;
    const uint32_t step = (tableSize >> 1) + (tableSize >> 3) + 3;
    if (tid == 0) {
        if (normalize_count(norm, tableLog, count, n)) s_err = 1;
        else {
            // cumulative counts of the spread symbols (norm > 0) and the low-probability symbols' places at the top
            uint32_t high = tableSize - 1, run = 0;
            for (int u = 0; u < n; ++u) {
                cumul[u] = run;                                  // first occurrence number of symbol u
                if (norm[u] == -1) tableSymbol[high--] = (uint32_t)u;
                else if (norm[u] > 0) run += (uint32_t)norm[u];
            }
            cumul[n] = run;
            s_high = high;
            if (run != high + 1) s_err = 1;                      // (the reference's "pos != 0" check)
        }
    }
    __syncthreads();
    if (s_err) { if (tid == 0) a.err[t] = 1; return; }
    const uint32_t high = s_high;
    // ---- spread: j -> position; usable positions numbered in j order
    const uint32_t per_thread = (tableSize + kBuildThreads - 1) / kBuildThreads;
    const uint32_t j0 = tid * per_thread, j1 = min(j0 + per_thread, tableSize);
    uint32_t local = 0;
    for (uint32_t j = j0; j < j1; ++j) local += ((j * step) & tableMask) <= high;
    s_scan[tid] = local;
    __syncthreads();
    if (tid == 0) { uint32_t run = 0; for (int i = 0; i < kBuildThreads; ++i) { const uint32_t v = s_scan[i]; s_scan[i] = run; run += v; } }
    __syncthreads();
    uint32_t k = s_scan[tid];
    for (uint32_t j = j0; j < j1; ++j) {
        const uint32_t pos = (j * step) & tableMask;
        if (pos > high) continue;
        int lo = 0, hi = n - 1;                                  // the symbol whose occurrences cover number k: last u with cumul[u] <= k, norm > 0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (cumul[mid] <= k) lo = mid; else hi = mid - 1;
        }
        while (norm[lo] <= 0) --lo;                              // (symbols without spread occurrences share their successor's start)
        tableSymbol[pos] = (uint32_t)lo;
        ++k;
    }
    __syncthreads();
    if (a.build_c) {  // tans.cpp:150-228
        uint16_t *stateTable = a.ct_state + (size_t)t * tsz;
        uint32_t *nb = a.ct_nb + (size_t)t * a.max_nsym;
        int32_t *fs = a.ct_fs + (size_t)t * a.max_nsym;
        // slot of a symbol's first state = states of all symbols before it (a low-probability symbol has one)
        if (tid == 0) {
            uint32_t total = 0;
            for (int s2 = 0; s2 < n; s2++) {
                count[s2] = total;                               // (count is free now: first stateTable slot of the symbol)
                if (norm[s2] == 0) { nb[s2] = 0; fs[s2] = 0; }
                else if (norm[s2] == -1 || norm[s2] == 1) { nb[s2] = (tableLog << 16) - (1u << tableLog); fs[s2] = (int32_t)total - 1; total++; }
                else {
                    const uint32_t maxBitsOut = tableLog - hb32((uint32_t)(norm[s2] - 1));
                    const uint32_t minStatePlus = (uint32_t)norm[s2] << maxBitsOut;
                    nb[s2] = (maxBitsOut << 16) - minStatePlus;
                    fs[s2] = (int32_t)total - norm[s2];
                    total += (uint32_t)norm[s2];
                }
            }
        }
        __syncthreads();
        for (int s2 = tid; s2 < n; s2 += kBuildThreads) {
            if (norm[s2] == 0) continue;
            uint32_t slot = count[s2];
            for (uint32_t u = 0; u < tableSize; u++)
                if (tableSymbol[u] == (uint32_t)s2) stateTable[slot++] = (uint16_t)(tableSize + u);
        }
    }
    if (a.build_d) {  // tans.cpp:262-318
        TansDEntry *table = a.dt + (size_t)t * tsz;
        if (tid == 0) {
            int fast = 1;
            const short largeLimit = (short)(1 << (tableLog - 1));
            for (int s2 = 0; s2 < n; s2++) if (norm[s2] >= largeLimit) fast = 0;
            a.dt_fast[t] = fast;
        }
        for (int s2 = tid; s2 < n; s2 += kBuildThreads) {
            if (norm[s2] == 0) continue;
            uint32_t next = norm[s2] == -1 ? 1u : (uint32_t)(uint16_t)norm[s2];
            for (uint32_t u = 0; u < tableSize; u++) {
                if (tableSymbol[u] != (uint32_t)s2) continue;
                const uint32_t nbBits = tableLog - hb32(next);
                table[u].symbol = (uint16_t)s2;
                table[u].nbBits = (uint16_t)(uint8_t)nbBits;
                table[u].newState = (uint32_t)((uint16_t)next << nbBits) - tableSize;
                ++next;
            }
        }
    }
}

constexpr int kTile = 1024, kThreads = 256;

struct CodeArgs {
    int T, tableLog, bypass, bypass_precision, max_nsym;
    const int32_t *nsym, *offsets;
    const uint16_t *ct_state; const uint32_t *ct_nb; const int32_t *ct_fs;
    const TansDEntry *dt; const int32_t *dt_fast;
};

// status bits: 1 index range, 4 stream, 8 FSE error (dst too small on encode / missing end mark on decode)
__global__ void __launch_bounds__(kThreads)
k_tans_encode(CodeArgs a, const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, long long n,
              uint8_t *__restrict__ out, long long cap, long long *out_len, int *status)
{
    __shared__ int32_t s_val[kTile], s_tab[kTile];
    __shared__ uint32_t s_raw[kTile];
    const int tid = threadIdx.x;
    const size_t tsz = (size_t)1 << a.tableLog;
    if (cap <= 8) { if (tid == 0) { atomicOr(status, 8); *out_len = 0; } return; }  // BIT_initCStream
    unsigned long long container = 0, state = 1ull << a.tableLog;
    int bitPos = 0, st = 0;
    long long ptr = 0;
    const long long end = cap - 8;
    const uint32_t bp = (uint32_t)a.bypass_precision, maxb = (1u << bp) - 1;
    auto flush = [&]() {
        const int nb = bitPos >> 3;
        for (int i = 0; i < nb; ++i) out[ptr + i] = (uint8_t)(container >> (8 * i));
        ptr += nb;
        if (ptr > end) ptr = end;
        bitPos &= 7;
        container = nb >= 8 ? 0 : container >> (nb * 8);
    };
    auto put = [&](const uint16_t *stt, uint32_t dnb, int32_t dfs) {  // Tans_encodeSymbol + BIT_flushBits
        const uint32_t nbOut = (uint32_t)((state + dnb) >> 16);
        container |= (state & ((1ull << nbOut) - 1)) << bitPos;
        bitPos += (int)nbOut;
        state = stt[(long long)(state >> nbOut) + dfs];
        flush();
    };
    for (long long hi = n; hi > 0; hi -= kTile) {
        const long long lo = hi - kTile > 0 ? hi - kTile : 0;
        const int cnt = (int)(hi - lo);
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) {
            int32_t c = indexes[lo + k];
            if ((uint32_t)c >= (uint32_t)a.T) { st |= 1; c = 0; }
            const int32_t maxv = a.nsym[c] - 1;
            int32_t v = symbols[lo + k] - a.offsets[c];
            uint32_t raw = 0;
            if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = maxv; }       // tans.cpp:618-625 (always, bypass or not)
            else if (v >= maxv) { raw = (uint32_t)(2 * (v - maxv)); v = maxv; }
            s_val[k] = v; s_tab[k] = c; s_raw[k] = raw;
        }
        __syncthreads();
        if (tid == 0) {
            for (int k = cnt - 1; k >= 0; --k) {
                const int c = s_tab[k], v = s_val[k];
                if (a.bypass && v == a.nsym[c] - 1) {
                    const uint32_t raw = s_raw[k];
                    int nd = 0;
                    while (nd * (int)bp < 32 && (raw >> (nd * bp)) != 0) ++nd;
                    const int ncnt = nd / (int)maxb + 1, ntok = ncnt + nd;
                    for (int u = ntok - 1; u >= 0; --u) {
                        uint32_t tok;
                        if (u >= ncnt) tok = (raw >> ((u - ncnt) * bp)) & maxb;
                        else tok = u < ncnt - 1 ? maxb : (uint32_t)(nd - (ncnt - 1) * (int)maxb);
                        put(a.ct_state + (size_t)a.T * tsz, a.ct_nb[(size_t)a.T * a.max_nsym + tok], a.ct_fs[(size_t)a.T * a.max_nsym + tok]);
                    }
                }
                put(a.ct_state + (size_t)c * tsz, a.ct_nb[(size_t)c * a.max_nsym + v], a.ct_fs[(size_t)c * a.max_nsym + v]);
            }
        }
    }
    if (st) atomicOr(status, st);
    if (tid == 0) {
        container |= (state & ((1ull << a.tableLog) - 1)) << bitPos;  // Tans_flushCState
        bitPos += a.tableLog;
        flush();
        container |= 1ull << bitPos;  // BIT_closeCStream: end mark
        bitPos += 1;
        flush();
        if (ptr >= end) *out_len = 0;
        else {
            if (bitPos > 0) out[ptr] = (uint8_t)container;
            *out_len = ptr + (bitPos > 0);
        }
    }
}

__device__ inline unsigned long long load_le64(const uint8_t *p)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const unsigned long long *base = reinterpret_cast<const unsigned long long *>(addr & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(addr & 7) * 8;
    const unsigned long long lo = base[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (base[1] << (64 - sh));
}

__global__ void __launch_bounds__(kThreads)
k_tans_decode(CodeArgs a, const uint8_t *__restrict__ src, long long len, const int32_t *__restrict__ indexes, long long n,
              int32_t *__restrict__ out, int *status)
{
    __shared__ int32_t s_idx[kTile], s_out[kTile];
    const int tid = threadIdx.x;
    const size_t tsz = (size_t)1 << a.tableLog;
    unsigned long long container = 0, state = 0;
    unsigned consumed = 0;
    long long ptr = 0;  // byte offset of the container inside src
    int st = 0;
    auto reload = [&]() {  // bitstream.h:361-389
        if (consumed > 64) return;
        if (ptr >= 8) { ptr -= consumed >> 3; consumed &= 7; container = load_le64(src + ptr); return; }
        if (ptr == 0) return;
        unsigned nb = consumed >> 3;
        if (ptr - (long long)nb < 0) nb = (unsigned)ptr;
        ptr -= nb; consumed -= nb * 8; container = load_le64(src + ptr);
    };
    auto read = [&](unsigned nb, int fast) -> unsigned long long {
        unsigned long long v;
        if (fast) v = (container << (consumed & 63)) >> ((64 - nb) & 63);
        else v = ((container << (consumed & 63)) >> 1) >> ((63 - nb) & 63);
        consumed += nb;
        return v;
    };
    auto get = [&](const TansDEntry *t, int fast) -> uint32_t {
        const TansDEntry e = t[state];
        state = e.newState + read(e.nbBits, fast);
        return e.symbol;
    };
    if (tid == 0) {
        const uint8_t last = src[len - 1];
        if (last == 0) st |= 8;
        if (len >= 8) { ptr = len - 8; container = load_le64(src + ptr); consumed = 8 - hb32(last ? last : 1); }
        else {
            ptr = 0;
            for (long long k = 0; k < len; ++k) container += (unsigned long long)src[k] << (8 * k);
            consumed = 8 - hb32(last ? last : 1) + (unsigned)(8 - len) * 8;
        }
        state = read((unsigned)a.tableLog, 0);
        reload();
    }
    const uint32_t bp = (uint32_t)a.bypass_precision, maxb = (1u << bp) - 1;
    for (long long lo = 0; lo < n; lo += kTile) {
        const int cnt = (int)(n - lo < kTile ? n - lo : kTile);
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) {
            int32_t c = indexes[lo + k];
            if ((uint32_t)c >= (uint32_t)a.T) { st |= 1; c = 0; }
            s_idx[k] = c;
        }
        __syncthreads();
        if (tid == 0 && !(st & 8)) {
            const TansDEntry *bt = a.dt + (size_t)a.T * tsz;
            const int bfast = a.bypass ? a.dt_fast[a.T] : 0;
            for (int k = 0; k < cnt; ++k) {
                const int c = s_idx[k];
                const int32_t maxv = a.nsym[c] - 1;
                reload();
                int32_t value = (int32_t)get(a.dt + (size_t)c * tsz, a.dt_fast[c]);
                if (a.bypass && value == maxv) {
                    uint32_t val = get(bt, bfast), nb = val;
                    while (val == maxb && nb < 64) { val = get(bt, bfast); nb += val; }
                    uint32_t raw = 0;
                    for (uint32_t j = 0; j < nb; ++j) { val = get(bt, bfast); if (j * bp < 32) raw |= val << (j * bp); }
                    value = (int32_t)(raw >> 1);
                    value = (raw & 1) ? -value - 1 : value + maxv;
                }
                s_out[k] = value + a.offsets[c];
            }
        }
        __syncthreads();
        for (int k = tid; k < cnt; k += kThreads) out[lo + k] = s_out[k];
    }
    if (st) atomicOr(status, st);
}

CodeArgs code_args(const TansTables &t)
{
    CodeArgs a;
    a.T = t.T; a.tableLog = t.tableLog; a.bypass = t.bypass; a.bypass_precision = t.bypass_precision; a.max_nsym = t.max_nsym;
    a.nsym = t.nsym.as<int32_t>(); a.offsets = t.offsets.as<int32_t>();
    a.ct_state = t.ct_state.as<uint16_t>(); a.ct_nb = t.ct_nb.as<uint32_t>(); a.ct_fs = t.ct_fs.as<int32_t>();
    a.dt = t.dt.as<TansDEntry>(); a.dt_fast = t.dt_fast.as<int32_t>();
    return a;
}

}  // namespace

TansTables *tans_new() { return new TansTables(); }

void tans_delete(TansTables *t)
{
    if (!t) return;
    DevBuf *bufs[] = {&t->nsym, &t->offsets, &t->ct_state, &t->ct_nb, &t->ct_fs, &t->dt, &t->dt_fast, &t->scratch, &t->err};
    for (DevBuf *b : bufs) b->release();
    delete t;
}

// role: 0 = both directions, 1 = encoder (C tables only), 2 = decoder (D tables only; checks tans.cpp:273-274)
int tans_init(TansTables &t, const int32_t *freqs, int T, int M, const int32_t *nsym, const int32_t *offsets, unsigned tableLog,
              int bypass, unsigned bypass_precision, int role, cudaStream_t s)
{
    if (T <= 0 || M <= 0) return value_error("freqs should be 2-dimensional with shape (num_symbols.size(), >num_symbols.max())");
    if (tableLog == 0) tableLog = 11;
    if (tableLog > 20) return value_error("tableLog requires too much memory : unsupported");
    const bool want_d = role != 1, want_c = role != 2;
    int maxn = 1 << bypass_precision;
    for (int i = 0; i < T; ++i) {
        if (nsym[i] < 1 || nsym[i] > M) return value_error("num_symbols out of range of freqs");
        maxn = nsym[i] > maxn ? nsym[i] : maxn;
        if (want_d && role == 2 && (unsigned)(nsym[i] - 1) > 65534u) return value_error("Unsupported max Symbol Value : too large");
    }
    if (role == 2 && tableLog > 12) return value_error("tableLog requires too much memory : unsupported");
    t.T = T; t.tableLog = (int)tableLog; t.bypass = bypass; t.bypass_precision = (int)bypass_precision; t.max_nsym = maxn; t.role = role;
    t.has_c = want_c;
    t.has_d = want_d && tableLog <= 12;
    const size_t tsz = (size_t)1 << tableLog, TT = (size_t)T + 1;
    DevBuf d_freqs;
    BASIC_TRY(d_freqs.reserve((size_t)T * M * 4));
    BASIC_CUDA(cudaMemcpyAsync(d_freqs.p, freqs, (size_t)T * M * 4, cudaMemcpyHostToDevice, s));
    BASIC_TRY(t.nsym.reserve(T * 4));
    BASIC_TRY(t.offsets.reserve(T * 4));
    BASIC_CUDA(cudaMemcpyAsync(t.nsym.p, nsym, T * 4, cudaMemcpyHostToDevice, s));
    BASIC_CUDA(cudaMemcpyAsync(t.offsets.p, offsets, T * 4, cudaMemcpyHostToDevice, s));
    if (t.has_c) {
        BASIC_TRY(t.ct_state.reserve(TT * tsz * 2));
        BASIC_TRY(t.ct_nb.reserve(TT * maxn * 4));
        BASIC_TRY(t.ct_fs.reserve(TT * maxn * 4));
    }
    if (t.has_d) BASIC_TRY(t.dt.reserve(TT * tsz * sizeof(TansDEntry)));
    BASIC_TRY(t.dt_fast.reserve(TT * 4));
    const size_t per = (size_t)maxn * 3 + 2 + tsz + 8 + (size_t)maxn;  // + norm (shorts, over-allocated)
    BASIC_TRY(t.scratch.reserve(TT * per * 4));
    BASIC_TRY(t.err.reserve(TT * 4));
    BASIC_CUDA(cudaMemsetAsync(t.err.p, 0, TT * 4, s));
    BuildArgs a;
    a.freqs = d_freqs.as<int32_t>(); a.M = M; a.nsym = t.nsym.as<int32_t>(); a.T = T; a.tableLog = (int)tableLog; a.max_nsym = maxn;
    a.bypass = bypass; a.nbypass = 1 << bypass_precision; a.build_c = t.has_c; a.build_d = t.has_d;
    a.ct_state = t.ct_state.as<uint16_t>(); a.ct_nb = t.ct_nb.as<uint32_t>(); a.ct_fs = t.ct_fs.as<int32_t>();
    a.dt = t.dt.as<TansDEntry>(); a.dt_fast = t.dt_fast.as<int32_t>(); a.scratch = t.scratch.as<uint32_t>(); a.err = t.err.as<int>();
    k_tans_build<<<(unsigned)TT, kBuildThreads, 0, s>>>(a);
    BASIC_LAUNCHED();
    std::vector<int> err(TT);
    BASIC_CUDA(cudaMemcpyAsync(err.data(), t.err.p, TT * 4, cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    d_freqs.release();
    for (size_t i = 0; i < TT; ++i) if (err[i]) return value_error("Error (generic)");
    return BASIC_OK;
}

int tans_encode(TansTables &t, const int32_t *d_sym, const int32_t *d_idx, int64_t n, uint8_t *d_out, int64_t cap, long long *d_len,
                int *d_status, cudaStream_t s)
{
    if (!t.has_c) return value_error("this tANS object holds no encoder tables");
    k_tans_encode<<<1, kThreads, 0, s>>>(code_args(t), d_sym, d_idx, n, d_out, cap, d_len, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int tans_decode(TansTables &t, const uint8_t *d_enc, int64_t len, const int32_t *d_idx, int64_t n, int32_t *d_out, int *d_status,
                cudaStream_t s)
{
    if (!t.has_d) return value_error("tableLog requires too much memory : unsupported");
    k_tans_decode<<<1, kThreads, 0, s>>>(code_args(t), d_enc, len, d_idx, n, d_out, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

/*
 * ans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's native entropy coder (cbench/csrc/ans) plus the CPU
 * specification of this repo's own multi-lane stream format.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product path
 * (cbench_basic_b200/) never does.
 *
 * Parity status: PINNED.  Every function below is checked byte-for-byte / value-for-value against the
 * unmodified reference compiled into oracle/_ref (tests/test_oracle_vs_ref.py, runs wherever
 * oracle/_ref exists) and against committed golden vectors produced by that reference
 * (tests/golden/, tests/test_oracle_golden.py).  The reference has no golden vectors of its own
 * (SURVEY.md section 4).
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_GENERIC (-1)        /* FSE "Error (generic)"                     */
#define ORC_ERR_DST_TOO_SMALL (-2)  /* FSE "Destination buffer is too small"     */
#define ORC_ERR_SRC_SIZE (-3)       /* FSE "Src size incorrect"                  */
#define ORC_ERR_TABLELOG_LARGE (-4) /* FSE "tableLog requires too much memory"   */
#define ORC_ERR_MAXSYM_LARGE (-5)   /* FSE "Unsupported max Symbol Value"        */
#define ORC_ERR_RANGE (-6)          /* symbol outside the table with bypass off  */
#define ORC_ERR_CAPACITY (-7)       /* caller's output buffer too small          */

/* ------------------------------------------------------------------------------------------------
 * 1. Quantised CDF construction
 * ---------------------------------------------------------------------------------------------- */

/* cbench/csrc/ans/rans64.cpp:69-126 pmf_to_quantized_cdf.  cdf has n+1 entries. */
int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, int32_t *cdf)
{
    cdf[0] = 0;
    for (int i = 0; i < n; ++i)
        cdf[i + 1] = (int32_t)roundf(pmf[i] * (float)(1 << precision)); /* std::round(float) -> int32 */
    /* std::accumulate(..., 0) sums in int, assigned to uint32 */
    int32_t acc = 0;
    for (int i = 0; i <= n; ++i) acc += cdf[i];
    const uint32_t total = (uint32_t)acc;
    if (total == 0) return ORC_ERR_GENERIC; /* reference divides by zero */
    for (int i = 0; i <= n; ++i) {
        /* (static_cast<uint64_t>(1 << precision) * p) / total, p int32 promoted to uint64 */
        uint64_t num = (uint64_t)(1 << precision) * (uint64_t)(int64_t)cdf[i];
        cdf[i] = (int32_t)(num / total);
    }
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1]; /* std::partial_sum */
    cdf[n] = 1 << precision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            uint32_t best_freq = ~0u;
            int best_steal = -1;
            for (int j = 0; j < n; ++j) {
                uint32_t freq = (uint32_t)(cdf[j + 1] - cdf[j]);
                if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
            }
            if (best_steal < 0) return ORC_ERR_GENERIC; /* reference: assert (compiled out) */
            if (best_steal < i) { for (int j = best_steal + 1; j <= i; ++j) cdf[j]--; }
            else                { for (int j = i + 1; j <= best_steal; ++j) cdf[j]++; }
        }
    }
    return ORC_OK;
}

/* cbench/csrc/ans/rans64.cpp:128-159 Rans64Base::init_params.
 * freqs [T, M] int32 row-major, nsym [T].  cdfs_out [T, stride] zero padded (stride >= max nsym + 2),
 * sizes_out[t] = nsym[t] + 2.  Float arithmetic is float32 with the reference's evaluation order. */
int orc_rans64_init_params(const int32_t *freqs, int T, int M, const int32_t *nsym, int precision,
                           int32_t *cdfs_out, int stride, int32_t *sizes_out)
{
    const float tail_mass = 1.f;
    for (int t = 0; t < T; ++t) {
        const int n = nsym[t];
        if (n < 0 || n > M || n + 2 > stride) return ORC_ERR_GENERIC;
        float *pmf = (float *)malloc(sizeof(float) * (size_t)(n + 1));
        float s = 0.0f;                                   /* std::accumulate(int..., 0.0f) */
        for (int i = 0; i < n; ++i) s = s + (float)freqs[(size_t)t * M + i];
        const float freq_total = s + tail_mass;
        pmf[n] = tail_mass / freq_total;
        for (int i = 0; i < n; ++i) pmf[i] = (float)freqs[(size_t)t * M + i] / freq_total;
        int32_t *row = cdfs_out + (size_t)t * stride;
        memset(row, 0, sizeof(int32_t) * (size_t)stride);
        int rc = orc_pmf_to_quantized_cdf(pmf, n + 1, precision, row);
        free(pmf);
        if (rc) return rc;
        sizes_out[t] = n + 2;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * 2. rANS64, lanes = 1 (the reference bitstream)
 * ---------------------------------------------------------------------------------------------- */
#define R64_L (1ull << 31) /* rans64.h:59 */

typedef struct {
    int T, stride, precision, bypass, bypass_precision;
    const int32_t *cdfs, *sizes, *offsets;
} orc_rans64_tables;

/* number of base-2^bp digits of raw (rans64.cpp:298-301) */
static int n_digits(uint32_t raw, int bp)
{
    int n = 0;
    while ((n * bp) < 32 && (raw >> (n * bp)) != 0) ++n;
    return n;
}

/* Escape token list in DECODER order (rans64.cpp:303-321): count in unary base (2^bp - 1), then the
 * digits LSB first.  Returns number of tokens. */
static int escape_tokens(uint32_t raw, int bp, uint32_t *tok)
{
    const uint32_t maxv = (1u << bp) - 1;
    int n = n_digits(raw, bp), k = 0;
    int32_t val = n;
    while (val >= (int32_t)maxv) { tok[k++] = maxv; val -= (int32_t)maxv; }
    tok[k++] = (uint32_t)val;
    for (int j = 0; j < n; ++j) tok[k++] = (raw >> (j * bp)) & maxv;
    return k;
}

/* rans64.cpp:203-361 Rans64Encoder::encode_with_indexes (cache = false).  Words are produced
 * back-to-front into a scratch of `cap_words` uint32; on return *out points INTO scratch at the first
 * word and *out_words is the length.  rans64.h:77-103 for Put/Flush, rans64.cpp:29-47 for PutBits. */
int orc_rans64_encode(const orc_rans64_tables *tb, const int32_t *symbols, const int32_t *indexes,
                      int64_t n, uint32_t *scratch, int64_t cap_words, int64_t *first_word)
{
    uint64_t x = R64_L;
    int64_t p = cap_words;
    const int prec = tb->precision, bp = tb->bypass_precision;
    uint32_t tok[96];
    for (int64_t i = n - 1; i >= 0; --i) {
        const int32_t c = indexes[i];
        if (c < 0 || c >= tb->T) return ORC_ERR_RANGE;
        const int32_t *cdf = tb->cdfs + (size_t)c * tb->stride;
        const int32_t max_value = tb->sizes[c] - 2;
        int32_t value = symbols[i] - tb->offsets[c];
        uint32_t raw = 0;
        if (tb->bypass) {
            if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
            else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
        } else if (value < 0 || value > max_value) {
            return ORC_ERR_RANGE; /* reference: undefined behaviour (asserts compiled out) */
        }
        const uint32_t start = (uint16_t)cdf[value];
        const uint32_t freq = (uint16_t)(cdf[value + 1] - cdf[value]);
        if (tb->bypass && value == max_value) {
            int k = escape_tokens(raw, bp, tok);
            while (k > 0) { /* reverse order */
                const uint32_t t = tok[--k];
                const uint64_t x_max = ((R64_L >> 16) << 32) * (uint64_t)(1u << (16 - bp));
                if (x >= x_max) { if (p <= 0) return ORC_ERR_CAPACITY; scratch[--p] = (uint32_t)x; x >>= 32; }
                x = (x << bp) | t;
            }
        }
        const uint64_t x_max = ((R64_L >> prec) << 32) * (uint64_t)freq;
        if (x >= x_max) { if (p <= 0) return ORC_ERR_CAPACITY; scratch[--p] = (uint32_t)x; x >>= 32; }
        x = ((x / freq) << prec) + (x % freq) + start;
    }
    if (p < 2) return ORC_ERR_CAPACITY;
    p -= 2;
    scratch[p] = (uint32_t)x;
    scratch[p + 1] = (uint32_t)(x >> 32);
    *first_word = p;
    return ORC_OK;
}

typedef struct { uint64_t x; int64_t pos; } orc_rans64_dstate;

/* rans64.hpp:104-111 set_stream */
void orc_rans64_set_stream(orc_rans64_dstate *st, const uint32_t *words)
{
    st->x = (uint64_t)words[0] | ((uint64_t)words[1] << 32);
    st->pos = 2;
}

static inline uint32_t r64_getbits(orc_rans64_dstate *st, const uint32_t *w, int nb)
{ /* rans64.cpp:49-65 */
    uint64_t x = st->x;
    uint32_t v = (uint32_t)(x & ((1u << nb) - 1));
    x >>= nb;
    if (x < R64_L) x = (x << 32) | w[st->pos++];
    st->x = x;
    return v;
}

/* rans64.cpp:501-598 decode_stream (and :389-499 decode_with_indexes = set_stream + decode_stream) */
int orc_rans64_decode_stream(const orc_rans64_tables *tb, orc_rans64_dstate *st, const uint32_t *w,
                             const int32_t *indexes, int64_t n, int32_t *out)
{
    const int prec = tb->precision, bp = tb->bypass_precision;
    const uint32_t maxb = (1u << bp) - 1;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t c = indexes[i];
        if (c < 0 || c >= tb->T) return ORC_ERR_RANGE;
        const int32_t *cdf = tb->cdfs + (size_t)c * tb->stride;
        const int32_t size = tb->sizes[c], max_value = size - 2;
        const uint32_t cum = (uint32_t)(st->x & ((1u << prec) - 1));
        int s = 0;
        while (s < size && (uint32_t)cdf[s] <= cum) ++s; /* first entry > cum (linear scan, :456-460) */
        s -= 1;
        uint64_t x = st->x;
        x = (uint64_t)(uint32_t)(cdf[s + 1] - cdf[s]) * (x >> prec) + (x & ((1ull << prec) - 1)) - (uint32_t)cdf[s];
        if (x < R64_L) x = (x << 32) | w[st->pos++];
        st->x = x;
        int32_t value = s;
        if (tb->bypass && value == max_value) {
            uint32_t val = r64_getbits(st, w, bp), nb = val;
            while (val == maxb) { val = r64_getbits(st, w, bp); nb += val; }
            uint32_t raw = 0;
            for (uint32_t j = 0; j < nb; ++j) { val = r64_getbits(st, w, bp); raw |= val << (j * bp); }
            value = (int32_t)(raw >> 1);
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + tb->offsets[c];
    }
    return ORC_OK;
}

/* Exact-reciprocal form of the encoder update (rans64.h:167-278); exported so the tests can prove it
 * equals x/freq, x%freq for 63-bit states -- the CUDA lanes=1 encoder uses this form. */
void orc_rans64_rcp(uint32_t start, uint32_t freq, uint32_t prec, uint64_t *rcp, uint32_t *shift,
                    uint32_t *bias, uint32_t *cmpl)
{
    *cmpl = (1u << prec) - freq;
    if (freq < 2) { *rcp = ~0ull; *shift = 0; *bias = start + (1u << prec) - 1; return; }
    uint32_t sh = 0;
    while (freq > (1u << sh)) sh++;
    uint64_t x0 = freq - 1, x1 = 1ull << (sh + 31);
    uint64_t t1 = x1 / freq;
    x0 += (x1 % freq) << 32;
    uint64_t t0 = x0 / freq;
    *rcp = t0 + (t1 << 32);
    *shift = sh - 1;
    *bias = start;
}
uint64_t orc_rans64_put_rcp(uint64_t x, uint64_t rcp, uint32_t shift, uint32_t bias, uint32_t cmpl)
{
    uint64_t q = (uint64_t)(((unsigned __int128)x * rcp) >> 64) >> shift;
    return x + bias + q * cmpl;
}

/* ------------------------------------------------------------------------------------------------
 * 3. tANS (FSE-derived), lanes = 1
 * ---------------------------------------------------------------------------------------------- */
static inline unsigned highbit32(uint32_t v) { return 31 - (unsigned)__builtin_clz(v); }

/* tans.cpp:26-100 Tans_normalizeM2 */
static int tans_normalize_m2(int16_t *norm, uint32_t tableLog, const uint32_t *count, uint64_t total,
                             uint32_t maxSV)
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
        norm[maxV] += (int16_t)toDistribute;
        return ORC_OK;
    }
    {
        const uint64_t vStepLog = 62 - tableLog, mid = (1ull << (vStepLog - 1)) - 1;
        const uint64_t rStep = ((((uint64_t)1 << vStepLog) * toDistribute) + mid) / total;
        uint64_t tmpTotal = mid;
        for (uint32_t s = 0; s <= maxSV; s++) {
            if (norm[s] == -2) {
                uint64_t end = tmpTotal + (count[s] * rStep);
                uint32_t weight = (uint32_t)(end >> vStepLog) - (uint32_t)(tmpTotal >> vStepLog);
                if (weight < 1) return ORC_ERR_GENERIC;
                norm[s] = (int16_t)weight;
                tmpTotal = end;
            }
        }
    }
    return ORC_OK;
}

/* tans.cpp:17-23 FSE_minTableLog + :102-148 Tans_normalizeCount.  count has nsym entries. */
int orc_tans_normalize(int16_t *norm, uint32_t tableLog, const uint32_t *count, int nsym)
{
    static const uint32_t rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const uint32_t maxSV = (uint32_t)nsym - 1;
    uint64_t total = 0;
    for (int i = 0; i < nsym; ++i) total += count[i];
    if (tableLog == 0) tableLog = 11;
    {
        uint32_t minBitsSrc = highbit32((uint32_t)(total - 1)) + 1;
        uint32_t minBitsSym = highbit32(maxSV) + 2;
        uint32_t minBits = minBitsSrc < minBitsSym ? minBitsSrc : minBitsSym;
        if (tableLog < minBits) return ORC_ERR_GENERIC;
    }
    const uint64_t scale = 62 - tableLog, step = ((uint64_t)1 << 62) / total, vStep = 1ull << (scale - 20);
    int still = 1 << tableLog;
    uint32_t largest = 0;
    int16_t largestP = 0;
    const uint32_t lowThreshold = (uint32_t)(total >> tableLog);
    for (uint32_t s = 0; s <= maxSV; s++) {
        if (count[s] == total) return ORC_ERR_GENERIC; /* reference: "rle special case" returns 0 with norm
                                                         left uninitialised; we refuse instead */
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = -1; still--; }
        else {
            int16_t proba = (int16_t)((count[s] * step) >> scale);
            if (proba < 8) {
                uint64_t restToBeat = vStep * rtb[proba];
                proba += (count[s] * step) - ((uint64_t)proba << scale) > restToBeat;
            }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba;
            still -= proba;
        }
    }
    if (-still >= (norm[largest] >> 1)) return tans_normalize_m2(norm, tableLog, count, total, maxSV);
    norm[largest] += (int16_t)still;
    return ORC_OK;
}

#define TANS_STEP(ts) (((ts) >> 1) + ((ts) >> 3) + 3) /* fse.h:618 */

/* Encoder table (tans.cpp:150-228 Tans_buildCTable): stateTable[1<<tableLog] (u16), deltaNbBits[nsym],
 * deltaFindState[nsym]. */
int orc_tans_build_ctable(const int16_t *norm, int nsym, uint32_t tableLog, uint16_t *stateTable,
                          uint32_t *deltaNbBits, int32_t *deltaFindState)
{
    const uint32_t tableSize = 1u << tableLog, tableMask = tableSize - 1, step = TANS_STEP(tableSize);
    uint32_t *cumul = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nsym + 2));
    uint16_t *tableSymbol = (uint16_t *)malloc(sizeof(uint16_t) * tableSize);
    uint32_t high = tableSize - 1;
    cumul[0] = 0;
    for (int u = 1; u <= nsym; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; tableSymbol[high--] = (uint16_t)(u - 1); }
        else cumul[u] = cumul[u - 1] + (uint32_t)norm[u - 1];
    }
    cumul[nsym] = tableSize + 1;
    uint32_t pos = 0;
    for (int s = 0; s < nsym; s++)
        for (int k = 0; k < norm[s]; k++) {
            tableSymbol[pos] = (uint16_t)s;
            pos = (pos + step) & tableMask;
            while (pos > high) pos = (pos + step) & tableMask;
        }
    int rc = ORC_OK;
    if (pos != 0) rc = ORC_ERR_GENERIC;
    if (!rc) {
        for (uint32_t u = 0; u < tableSize; u++) { uint16_t s = tableSymbol[u]; stateTable[cumul[s]++] = (uint16_t)(tableSize + u); }
        uint32_t total = 0;
        for (int s = 0; s < nsym; s++) {
            switch (norm[s]) {
            case 0: deltaNbBits[s] = 0; deltaFindState[s] = 0; break;
            case -1: case 1:
                deltaNbBits[s] = (tableLog << 16) - (1u << tableLog);
                deltaFindState[s] = (int32_t)total - 1; total++; break;
            default: {
                uint32_t maxBitsOut = tableLog - highbit32((uint32_t)(norm[s] - 1));
                uint32_t minStatePlus = (uint32_t)norm[s] << maxBitsOut;
                deltaNbBits[s] = (maxBitsOut << 16) - minStatePlus;
                deltaFindState[s] = (int32_t)total - norm[s];
                total += (uint32_t)norm[s];
            } }
        }
    }
    free(cumul); free(tableSymbol);
    return rc;
}

/* Decoder table (tans.cpp:262-318 Tans_buildDTable, tans.hpp:101-106): entries {u32 newState; u16 symbol;
 * u16 nbBits}; *fast = FSE fastMode. */
typedef struct { uint32_t newState; uint16_t symbol; uint16_t nbBits; } orc_tans_dentry;

int orc_tans_build_dtable(const int16_t *norm, int nsym, uint32_t tableLog, orc_tans_dentry *table, int *fast)
{
    const uint32_t tableSize = 1u << tableLog, tableMask = tableSize - 1, step = TANS_STEP(tableSize);
    if ((uint32_t)(nsym - 1) > 65535u - 1) return ORC_ERR_MAXSYM_LARGE;
    if (tableLog > 12) return ORC_ERR_TABLELOG_LARGE; /* TANS_MAX_TABLELOG, tans.hpp:15 */
    uint16_t *symbolNext = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)nsym);
    uint32_t high = tableSize - 1;
    *fast = 1;
    const int16_t largeLimit = (int16_t)(1 << (tableLog - 1));
    for (int s = 0; s < nsym; s++) {
        if (norm[s] == -1) { table[high--].symbol = (uint16_t)s; symbolNext[s] = 1; }
        else { if (norm[s] >= largeLimit) *fast = 0; symbolNext[s] = (uint16_t)norm[s]; }
    }
    uint32_t pos = 0;
    for (int s = 0; s < nsym; s++)
        for (int k = 0; k < norm[s]; k++) {
            table[pos].symbol = (uint16_t)s;
            pos = (pos + step) & tableMask;
            while (pos > high) pos = (pos + step) & tableMask;
        }
    int rc = pos != 0 ? ORC_ERR_GENERIC : ORC_OK;
    if (!rc)
        for (uint32_t u = 0; u < tableSize; u++) {
            uint16_t sym = table[u].symbol;
            uint16_t next = symbolNext[sym]++;
            table[u].nbBits = (uint8_t)(tableLog - highbit32(next));
            table[u].newState = (uint32_t)((next << table[u].nbBits) - tableSize);
        }
    free(symbolNext);
    return rc;
}

typedef struct {
    int T, tableLog, bypass, bypass_precision, max_nsym;
    const int32_t *nsym, *offsets;
    /* per table t: ctables at [t * (1<<tableLog)], ct_nb/ct_fs at [t * max_nsym]; dtables at [t * (1<<tableLog)] */
    const uint16_t *ct_state; const uint32_t *ct_nb; const int32_t *ct_fs;
    const orc_tans_dentry *dt; const int32_t *dt_fast;
    /* bypass (uniform 2^bp symbols, tans.cpp:430-456 / :505-531) */
    const uint16_t *bct_state; const uint32_t *bct_nb; const int32_t *bct_fs;
    const orc_tans_dentry *bdt; int bdt_fast;
} orc_tans_tables;

/* FSE bit output (csrc/FSE/bitstream.h:185-248), 64-bit container, little-endian bytes. */
typedef struct { uint64_t c; int pos; uint8_t *start, *ptr, *end; } bitc_t;
static void bitc_flush(bitc_t *b)
{
    size_t nb = (size_t)(b->pos >> 3);
    memcpy(b->ptr, &b->c, 8);
    b->ptr += nb;
    if (b->ptr > b->end) b->ptr = b->end;
    b->pos &= 7;
    b->c >>= nb * 8;
}
static inline void bitc_add(bitc_t *b, uint64_t v, unsigned n) { b->c |= (v & ((1ull << n) - 1)) << b->pos; b->pos += (int)n; }

static inline void tans_enc_sym(bitc_t *b, uint64_t *state, const uint16_t *st, uint32_t dnb, int32_t dfs)
{ /* tans.cpp:247-254 Tans_encodeSymbol + BIT_flushBits */
    uint32_t nbOut = (uint32_t)((*state + dnb) >> 16);
    bitc_add(b, *state, nbOut);
    *state = st[(int64_t)(*state >> nbOut) + dfs];
    bitc_flush(b);
}

/* tans.cpp:534-690 TansEncoder::encode_with_indexes.  out capacity follows the reference:
 * n * tableLog / 8 bytes (:573); returns 0 bytes when the stream does not fit (BIT_closeCStream). */
int orc_tans_encode(const orc_tans_tables *tb, const int32_t *symbols, const int32_t *indexes, int64_t n,
                    uint8_t *out, int64_t cap, int64_t *out_len)
{
    const int64_t want = n * tb->tableLog / 8;
    if (cap < want) return ORC_ERR_CAPACITY;
    if (want <= 8) return ORC_ERR_DST_TOO_SMALL; /* BIT_initCStream */
    bitc_t b = {0, 0, out, out, out + want - 8};
    uint64_t state = 1ull << tb->tableLog;
    const int bp = tb->bypass_precision;
    const size_t tsz = (size_t)1 << tb->tableLog;
    uint32_t tok[96];
    for (int64_t i = n - 1; i >= 0; --i) {
        const int32_t c = indexes[i];
        if (c < 0 || c >= tb->T) return ORC_ERR_RANGE;
        const int32_t max_value = tb->nsym[c] - 1;
        int32_t value = symbols[i] - tb->offsets[c];
        uint32_t raw = 0;
        if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
        else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
        if (tb->bypass && value == max_value) {
            int k = escape_tokens(raw, bp, tok);
            while (k > 0) { uint32_t t = tok[--k]; tans_enc_sym(&b, &state, tb->bct_state, tb->bct_nb[t], tb->bct_fs[t]); }
        }
        tans_enc_sym(&b, &state, tb->ct_state + (size_t)c * tsz, tb->ct_nb[(size_t)c * tb->max_nsym + value],
                     tb->ct_fs[(size_t)c * tb->max_nsym + value]);
    }
    bitc_add(&b, state, (unsigned)tb->tableLog); /* Tans_flushCState */
    bitc_flush(&b);
    b.c |= 1ull << b.pos; b.pos += 1;            /* BIT_closeCStream end mark */
    bitc_flush(&b);
    *out_len = (b.ptr >= b.end) ? 0 : (int64_t)(b.ptr - b.start) + (b.pos > 0);
    return ORC_OK;
}

/* FSE bit input (bitstream.h:260-400), reads backward from the end mark. */
typedef struct { uint64_t c; unsigned consumed; const uint8_t *ptr, *start; } bitd_t;

static int bitd_init(bitd_t *d, const uint8_t *src, size_t n)
{ /* bitstream.h:260-294 BIT_initDStream */
    if (n < 1) return ORC_ERR_SRC_SIZE;
    const uint8_t last = src[n - 1];
    if (last == 0) return ORC_ERR_GENERIC; /* end mark not present */
    d->start = src;
    if (n >= 8) {
        d->ptr = src + n - 8;
        memcpy(&d->c, d->ptr, 8);
        d->consumed = 8 - highbit32(last);
    } else {
        d->ptr = src;
        d->c = 0;
        for (size_t k = 0; k < n; ++k) d->c += (uint64_t)src[k] << (8 * k);
        d->consumed = 8 - highbit32(last) + (unsigned)(8 - n) * 8;
    }
    return ORC_OK;
}
static void bitd_reload(bitd_t *d)
{ /* bitstream.h:361-389 BIT_reloadDStream */
    if (d->consumed > 64) return;
    if (d->ptr >= d->start + 8) {
        d->ptr -= d->consumed >> 3; d->consumed &= 7; memcpy(&d->c, d->ptr, 8); return;
    }
    if (d->ptr == d->start) return;
    uint32_t nb = d->consumed >> 3;
    if (d->ptr - nb < d->start) nb = (uint32_t)(d->ptr - d->start);
    d->ptr -= nb; d->consumed -= nb * 8; memcpy(&d->c, d->ptr, 8);
}
static inline uint64_t bitd_read(bitd_t *d, unsigned nb)
{ /* BIT_lookBits (non-BMI path, bitstream.h:321-330) + BIT_skipBits */
    uint64_t v = ((d->c << (d->consumed & 63)) >> 1) >> ((63 - nb) & 63);
    d->consumed += nb;
    return v;
}
static inline uint64_t bitd_read_fast(bitd_t *d, unsigned nb)
{ /* BIT_lookBitsFast (bitstream.h:334-338): only valid for nb >= 1 */
    uint64_t v = (d->c << (d->consumed & 63)) >> ((64 - nb) & 63);
    d->consumed += nb;
    return v;
}
static inline uint32_t tans_dec_sym(bitd_t *d, uint64_t *state, const orc_tans_dentry *t, int fast)
{ /* tans.cpp:337-364 */
    const orc_tans_dentry e = t[*state];
    uint64_t low = fast ? bitd_read_fast(d, e.nbBits) : bitd_read(d, e.nbBits);
    *state = e.newState + low;
    return e.symbol;
}

/* tans.cpp:722-822 TansDecoder::decode_with_indexes */
int orc_tans_decode(const orc_tans_tables *tb, const uint8_t *enc, int64_t len, const int32_t *indexes,
                    int64_t n, int32_t *out)
{
    bitd_t d;
    int rc = bitd_init(&d, enc, (size_t)len);
    if (rc) return rc;
    uint64_t state = bitd_read(&d, (unsigned)tb->tableLog); /* Tans_initDState */
    bitd_reload(&d);
    const int bp = tb->bypass_precision;
    const uint32_t maxb = (1u << bp) - 1;
    const size_t tsz = (size_t)1 << tb->tableLog;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t c = indexes[i];
        if (c < 0 || c >= tb->T) return ORC_ERR_RANGE;
        const int32_t max_value = tb->nsym[c] - 1;
        bitd_reload(&d);
        int32_t value = (int32_t)tans_dec_sym(&d, &state, tb->dt + (size_t)c * tsz, tb->dt_fast[c]);
        if (tb->bypass && value == max_value) {
            uint32_t val = tans_dec_sym(&d, &state, tb->bdt, tb->bdt_fast), nb = val;
            while (val == maxb) { val = tans_dec_sym(&d, &state, tb->bdt, tb->bdt_fast); nb += val; }
            uint32_t raw = 0;
            for (uint32_t j = 0; j < nb; ++j) { val = tans_dec_sym(&d, &state, tb->bdt, tb->bdt_fast); raw |= val << (j * bp); }
            value = (int32_t)(raw >> 1);
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + tb->offsets[c];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * 4. Multi-lane rANS segment ("BLS1") -- CPU specification of THIS repo's format (include/basic_b200.h,
 *    DESIGN.md section "Multi-lane stream").  Not a reference format: the reference coder is single-lane.
 *    The CUDA kernels must reproduce these bytes exactly; lossless round trip and the bpp bound are
 *    checked against the lanes=1 stream above.
 *
 *    32-bit states, L = 2^16, 16-bit renormalisation words, probabilities at the table precision (<= 16).
 *    A segment of n symbols is cut into chunks of `chunk_syms` (multiple of 128) symbols; one chunk = 32
 *    interleaved lanes sharing one word stream.  Within a chunk, local symbol j belongs to lane
 *    (j % 128) / 4 and is coded at step (j / 128) * 4 + (j % 4).
 *
 *    Escapes: the reference's token list (count in unary base 2^bp - 1, then the bp-bit digits LSB first,
 *    rans64.cpp:293-335) is coded through the lane's own state, but in UNITS of up to floor(16 / bp) consecutive tokens:
 *    unit k holds tokens [k * tpu, min((k + 1) * tpu, ntok)), first token in the low bits, and is pushed / popped as one
 *    (bp * count)-bit quantity with one renormalisation check -- a typical escape (count + up to three digits) is one
 *    warp-wide event instead of up to four.  A state is >= 2^16 at every unit boundary, so the decoder can parse the
 *    tokens of a unit (and with them the unit's width) from the low 16 bits before it pops them.
 *
 *    Segment layout (little endian): u32 n_chunks | u32 n_slices | u32 chunk_syms[n_slices] | u32 end_word[n_chunks]
 *    (cumulative word count up to and including chunk k) | u32 state[n_chunks][32] | u16 words of chunk 0, chunk 1, ...
 *    | zero pad to 4 bytes.  Slices: see orc_bls_encode_slices.
 * ---------------------------------------------------------------------------------------------- */
#define BLS_LANES 32
#define BLS_L (1u << 16)

typedef struct { int m; uint32_t tok[96]; } bls_esc;  /* m = UNITS after bls_pack_units: tok[k] = unit value, wid[k] its bits */
typedef struct { int wid[96]; } bls_wid;

/* token list -> units of up to 16 / bp tokens (first token in the low bits); returns the number of units */
static int bls_pack_units(int ntok, int bp, uint32_t *tok, int *wid)
{
    const int tpu = 16 / bp;
    int nu = 0;
    for (int t0 = 0; t0 < ntok; t0 += tpu, ++nu) {
        const int cnt = ntok - t0 < tpu ? ntok - t0 : tpu;
        uint32_t v = 0;
        for (int i = 0; i < cnt; ++i) v |= tok[t0 + i] << (bp * i);
        tok[nu] = v; wid[nu] = bp * cnt;   /* in place: nu <= t0 */
    }
    return nu;
}

static inline int64_t bls_local_index(int64_t step, int lane) { return (step >> 2) * 128 + lane * 4 + (step & 3); }

int64_t orc_bls_num_chunks(int64_t n, int64_t chunk_syms) { return n <= 0 ? 0 : (n + chunk_syms - 1) / chunk_syms; }

/* Encode the m symbols of one slice of one chunk on top of the running lane states x[] ; words are written
 * back-to-front into wbuf, *p is the index of the first word written so far. */
static int bls_encode_slice(const orc_rans64_tables *tb, const int32_t *sym, const int32_t *idx, int64_t m,
                            uint16_t *wbuf, int64_t *pp, uint32_t *x)
{
    bls_esc *esc = (bls_esc *)malloc(sizeof(bls_esc) * BLS_LANES);
    bls_wid *ew = (bls_wid *)malloc(sizeof(bls_wid) * BLS_LANES);
    uint32_t start[BLS_LANES], freq[BLS_LANES];
    int active[BLS_LANES];
    int64_t p = *pp;
    const int prec = tb->precision, bp = tb->bypass_precision;
    const int64_t nsteps = ((m + 127) / 128) * 4;
    int rc = ORC_OK;
    for (int64_t t = nsteps - 1; t >= 0 && !rc; --t) {
        int maxm = 0;
        for (int l = 0; l < BLS_LANES; ++l) {
            const int64_t j = bls_local_index(t, l);
            active[l] = j < m; esc[l].m = 0;
            if (!active[l]) continue;
            const int32_t c = idx[j];
            if (c < 0 || c >= tb->T) { rc = ORC_ERR_RANGE; break; }
            const int32_t *cdf = tb->cdfs + (size_t)c * tb->stride;
            const int32_t max_value = tb->sizes[c] - 2;
            int32_t value = sym[j] - tb->offsets[c];
            uint32_t raw = 0;
            if (tb->bypass) {
                if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
                else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
                if (value == max_value) esc[l].m = bls_pack_units(escape_tokens(raw, bp, esc[l].tok), bp, esc[l].tok, ew[l].wid);
            } else if (value < 0 || value > max_value) { rc = ORC_ERR_RANGE; break; }
            start[l] = (uint32_t)cdf[value]; freq[l] = (uint32_t)(cdf[value + 1] - cdf[value]);
            if (esc[l].m > maxm) maxm = esc[l].m;
        }
        if (rc) break;
        /* escape units, last to first; inside one event words are laid out in ascending lane order */
        for (int u = maxm - 1; u >= 0; --u) {
            int cnt = 0;
            for (int l = 0; l < BLS_LANES; ++l) if (esc[l].m > u && x[l] >= (BLS_L << (16 - ew[l].wid[u]))) cnt++;
            if (p < cnt) { rc = ORC_ERR_CAPACITY; break; }
            p -= cnt; int r = 0;
            for (int l = 0; l < BLS_LANES; ++l) {
                if (esc[l].m <= u) continue;
                if (x[l] >= (BLS_L << (16 - ew[l].wid[u]))) { wbuf[p + r++] = (uint16_t)x[l]; x[l] >>= 16; }
                x[l] = (x[l] << ew[l].wid[u]) | esc[l].tok[u];
            }
        }
        if (rc) break;
        int cnt = 0;
        for (int l = 0; l < BLS_LANES; ++l)
            if (active[l] && (uint64_t)x[l] >= ((uint64_t)(BLS_L >> prec) << 16) * freq[l]) cnt++;
        if (p < cnt) { rc = ORC_ERR_CAPACITY; break; }
        p -= cnt; int r = 0;
        for (int l = 0; l < BLS_LANES; ++l) {
            if (!active[l]) continue;
            if ((uint64_t)x[l] >= ((uint64_t)(BLS_L >> prec) << 16) * freq[l]) { wbuf[p + r++] = (uint16_t)x[l]; x[l] >>= 16; }
            x[l] = ((x[l] / freq[l]) << prec) + (x[l] % freq[l]) + start[l];
        }
    }
    free(esc);
    free(ew);
    *pp = p;
    return rc;
}

/* Encode a whole segment.  out must hold orc_bls_bound(n, chunk_syms) bytes. */
int64_t orc_bls_bound(int64_t n, int64_t chunk_syms)
{
    int64_t nc = orc_bls_num_chunks(n, chunk_syms);
    return 12 + nc * (4 + 128) + n * 24 + 64;
}

/* General form: the segment's symbols are n_slices consecutive runs (slice g has slice_n[g] symbols and follows
 * slice g - 1 in sym / idx).  Chunk k owns symbols [k * slice_cs[g], min(slice_n[g], (k + 1) * slice_cs[g])) of EVERY
 * slice (possibly none); its 32 lanes code slice 0 first, then slice 1, ... with the lane states carried over, so a
 * decoder can stop after any slice, learn more (the next group's parameters) and continue.  Each slice starts on
 * a fresh 128-symbol block (the tail of its last block is idle).  One state flush per lane for the whole segment. */
int orc_bls_encode_slices(const orc_rans64_tables *tb, const int32_t *sym, const int32_t *idx, int n_slices,
                          const int64_t *slice_n, const int64_t *slice_cs, int64_t nc, uint8_t *out, int64_t cap, int64_t *out_len)
{
    int64_t wcap = 64, total = 0;
    for (int g = 0; g < n_slices; ++g) {
        if (slice_cs[g] <= 0 || slice_cs[g] % 128 || orc_bls_num_chunks(slice_n[g], slice_cs[g]) > nc) return ORC_ERR_GENERIC;
        /* callers pass nc = max over slices of ceil(n_g / cs_g): chunks that own nothing are not stored */
        wcap += slice_cs[g] * 12;
        total += slice_n[g];
    }
    int64_t hdr = 8 + 4 * (int64_t)n_slices + nc * 4 + nc * 128;
    if (cap < hdr) return ORC_ERR_CAPACITY;
    uint32_t *h32 = (uint32_t *)out;
    h32[0] = (uint32_t)nc;
    h32[1] = (uint32_t)n_slices;
    for (int g = 0; g < n_slices; ++g) h32[2 + g] = (uint32_t)slice_cs[g];
    uint32_t *nwords = h32 + 2 + n_slices, *states = nwords + nc;
    int64_t pos = hdr, cum_words = 0;
    uint16_t *wbuf = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)wcap);
    int rc = ORC_OK;
    for (int64_t k = 0; k < nc && !rc; ++k) {
        uint32_t x[BLS_LANES];
        for (int l = 0; l < BLS_LANES; ++l) x[l] = BLS_L;
        int64_t first = wcap, off = total;
        for (int g = n_slices - 1; g >= 0 && !rc; --g) {   /* the encoder walks the chunk's symbols backwards */
            off -= slice_n[g];
            const int64_t b = k * slice_cs[g];
            int64_t m = slice_n[g] - b;
            if (m > slice_cs[g]) m = slice_cs[g];
            if (m <= 0) continue;
            rc = bls_encode_slice(tb, sym + off + b, idx + off + b, m, wbuf, &first, x);
        }
        if (rc) break;
        memcpy(states + k * 32, x, sizeof(x));
        const int64_t nw = wcap - first;
        if (pos + nw * 2 + 4 > cap) { rc = ORC_ERR_CAPACITY; break; }
        cum_words += nw;
        nwords[k] = (uint32_t)cum_words;
        memcpy(out + pos, wbuf + first, (size_t)nw * 2);
        pos += nw * 2;
    }
    free(wbuf);
    if (rc) return rc;
    while (pos & 3) out[pos++] = 0;
    *out_len = pos;
    return ORC_OK;
}

int orc_bls_encode(const orc_rans64_tables *tb, const int32_t *sym, const int32_t *idx, int64_t n,
                   int64_t chunk_syms, uint8_t *out, int64_t cap, int64_t *out_len)
{
    if (chunk_syms <= 0 || chunk_syms % 128) return ORC_ERR_GENERIC;
    return orc_bls_encode_slices(tb, sym, idx, 1, &n, &chunk_syms, orc_bls_num_chunks(n, chunk_syms), out, cap, out_len);
}

/* Decode the m symbols of one slice of one chunk, continuing from lane states x[] and word position *wpp. */
static int bls_decode_slice(const orc_rans64_tables *tb, const uint16_t *w, int64_t *wpp, uint32_t *x, const int32_t *idx,
                            int64_t m, int32_t *out)
{
    const int prec = tb->precision, bp = tb->bypass_precision;
    const uint32_t maxb = (1u << bp) - 1, pmask = (1u << prec) - 1;
    int64_t wp = *wpp;
    const int64_t nsteps = ((m + 127) / 128) * 4;
    for (int64_t t = 0; t < nsteps; ++t) {
        int32_t value[BLS_LANES], maxv[BLS_LANES], cc[BLS_LANES];
        int escl[BLS_LANES], phase[BLS_LANES]; uint32_t nb[BLS_LANES], raw[BLS_LANES], jj[BLS_LANES];
        int any = 0;
        /* main event: decode + advance for all active lanes, then renormalise in lane order */
        for (int l = 0; l < BLS_LANES; ++l) {
            const int64_t j = bls_local_index(t, l);
            escl[l] = 0; cc[l] = -1;
            if (j >= m) continue;
            const int32_t c = idx[j];
            if (c < 0 || c >= tb->T) return ORC_ERR_RANGE;
            cc[l] = c;
            const int32_t *cdf = tb->cdfs + (size_t)c * tb->stride;
            const int32_t size = tb->sizes[c];
            maxv[l] = size - 2;
            const uint32_t cum = x[l] & pmask;
            int s = 0;
            while (s < size && (uint32_t)cdf[s] <= cum) ++s;
            s -= 1;
            x[l] = (uint32_t)(cdf[s + 1] - cdf[s]) * (x[l] >> prec) + cum - (uint32_t)cdf[s];
            value[l] = s;
            if (tb->bypass && s == maxv[l]) { escl[l] = 1; phase[l] = 0; nb[l] = 0; raw[l] = 0; jj[l] = 0; any = 1; }
        }
        for (int l = 0; l < BLS_LANES; ++l)
            if (cc[l] >= 0 && x[l] < BLS_L) x[l] = (x[l] << 16) | w[wp++];
        /* escape units: every escaping lane parses up to 16 / bp tokens from the low 16 bits of its state (>= 2^16 here),
         * pops exactly those, then the lanes renormalise in lane order */
        while (any) {
            int was[BLS_LANES];
            for (int l = 0; l < BLS_LANES; ++l) {
                was[l] = escl[l];
                if (!escl[l]) continue;
                int used = 0;
                for (int i = 0; i < 16 / bp && escl[l]; ++i, ++used) {
                    const uint32_t val = (x[l] >> (bp * i)) & maxb;
                    if (phase[l] == 0) { nb[l] += val; if (val != maxb) { phase[l] = 1; if (nb[l] == 0) escl[l] = 0; } }
                    else { if (jj[l] * bp < 32) raw[l] |= val << (jj[l] * bp); if (++jj[l] == nb[l]) escl[l] = 0; }
                }
                x[l] >>= bp * used;
            }
            for (int l = 0; l < BLS_LANES; ++l) if (was[l] && x[l] < BLS_L) x[l] = (x[l] << 16) | w[wp++];
            any = 0;
            for (int l = 0; l < BLS_LANES; ++l) {
                if (!was[l]) continue;
                if (escl[l]) any = 1;
                else { int32_t v = (int32_t)(raw[l] >> 1); value[l] = (raw[l] & 1) ? -v - 1 : v + maxv[l]; }
            }
        }
        for (int l = 0; l < BLS_LANES; ++l)
            if (cc[l] >= 0) out[bls_local_index(t, l)] = value[l] + tb->offsets[cc[l]];
    }
    *wpp = wp;
    return ORC_OK;
}

/* Decode a segment of n_slices slices (indexes / out laid out slice after slice); *consumed = its bytes. */
int orc_bls_decode_slices(const orc_rans64_tables *tb, const uint8_t *enc, int64_t len, const int32_t *idx, int n_slices,
                          const int64_t *slice_n, int32_t *out, int64_t *consumed)
{
    if (len < 8) return ORC_ERR_SRC_SIZE;
    const uint32_t *h32 = (const uint32_t *)enc;
    const int64_t nc = h32[0];
    if ((int64_t)h32[1] != n_slices || len < 8 + 4 * (int64_t)n_slices) return ORC_ERR_SRC_SIZE;
    const uint32_t *cs = h32 + 2;
    int64_t nc_need = 0;
    for (int g = 0; g < n_slices; ++g) {
        if (cs[g] == 0 || cs[g] % 128) return ORC_ERR_SRC_SIZE;
        const int64_t need = orc_bls_num_chunks(slice_n[g], cs[g]);
        if (need > nc_need) nc_need = need;
    }
    if (nc_need != nc) return ORC_ERR_SRC_SIZE;
    const uint32_t *nwords = h32 + 2 + n_slices, *states = nwords + nc;
    int64_t pos = 8 + 4 * (int64_t)n_slices + nc * 4 + nc * 128;
    if (len < pos) return ORC_ERR_SRC_SIZE;
    for (int64_t k = 0; k < nc; ++k) {
        const int64_t nw_k = (int64_t)nwords[k] - (k ? (int64_t)nwords[k - 1] : 0);
        if (nw_k < 0 || pos + nw_k * 2 > len) return ORC_ERR_SRC_SIZE;
        const uint16_t *w = (const uint16_t *)(enc + pos);
        int64_t wp = 0, off = 0;
        uint32_t x[BLS_LANES];
        memcpy(x, states + k * 32, sizeof(x));
        for (int g = 0; g < n_slices; ++g) {
            const int64_t b = k * (int64_t)cs[g];
            int64_t m = slice_n[g] - b;
            if (m > (int64_t)cs[g]) m = cs[g];
            if (m > 0) {
                const int rc = bls_decode_slice(tb, w, &wp, x, idx + off + b, m, out + off + b);
                if (rc) return rc;
            }
            off += slice_n[g];
        }
        if (wp != nw_k) return ORC_ERR_SRC_SIZE;
        pos += nw_k * 2;
    }
    while (pos & 3) pos++;
    *consumed = pos;
    return ORC_OK;
}

int orc_bls_decode(const orc_rans64_tables *tb, const uint8_t *enc, int64_t len, const int32_t *idx, int64_t n,
                   int64_t chunk_syms, int32_t *out, int64_t *consumed)
{
    if (len < 12) return ORC_ERR_SRC_SIZE;
    const uint32_t *h32 = (const uint32_t *)enc;
    if (chunk_syms > 0 && (int64_t)h32[2] != chunk_syms) return ORC_ERR_SRC_SIZE;
    if ((int64_t)h32[0] != orc_bls_num_chunks(n, h32[2] ? h32[2] : 1)) return ORC_ERR_SRC_SIZE;
    return orc_bls_decode_slices(tb, enc, len, idx, 1, &n, out, consumed);
}


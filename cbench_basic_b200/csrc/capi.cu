// extern "C" entry points (include/basic_b200.h) and the host-side orchestration of the kernels.
#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <condition_variable>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define BASIC_HAVE_SSE2 1
#endif
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace basic {

// ---- implemented in the other translation units -----------------------------------------------------
int rans_tables_from_freqs(RansTables &, const int32_t *, int, int, const int32_t *, const int32_t *, int, cudaStream_t);
int rans_tables_from_cdfs(RansTables &, const int32_t *, int, int, const int32_t *, const int32_t *, int, cudaStream_t);
int pmf_to_cdf_device(const float *, int, int, int32_t *);
int launch_rans64_encode(const RansTables &, const int32_t *, const int32_t *, int64_t, int, int, uint32_t *, int64_t,
                         long long *, int *, cudaStream_t, int n_streams = 1);
int launch_rans64_decode(const RansTables &, const uint32_t *, int64_t, void *, int, const int32_t *, int64_t, int, int,
                         int32_t *, int *, cudaStream_t, int n_streams = 1, const long long *d_off = nullptr,
                         const long long *d_nwords = nullptr);
int launch_bls_encode(const RansTables &, int, int, const int32_t *, const int32_t *, int, const void *, int, uint16_t *, int,
                      uint32_t *, uint32_t *, unsigned char *, long long *, int *, int, cudaStream_t);
int launch_bls_decode(const RansTables &, int, int, const unsigned char *, int64_t, const int32_t *, int64_t, int, int, int, int,
                      uint32_t *, uint32_t *, int32_t *, int *, int, cudaStream_t);
int launch_estimate_bits(const RansTables &, int, int, const int32_t *, const int32_t *, int64_t, float *, cudaStream_t);
int launch_quantize_index(const float *, const float *, const int32_t *, int64_t, int, int, int, const float *, int, int32_t *,
                          int32_t *, float *, int, cudaStream_t, int params_cl = 0, const int32_t *perm = nullptr);
int launch_dequantize(const int32_t *, const float *, const int32_t *, int64_t, int, int, int, float *, int, cudaStream_t,
                      int params_cl = 0, const int32_t *perm = nullptr);

int launch_ar_effective_indexes(const int32_t *, int, int, int, int, const int32_t *, const int32_t *, const int32_t *, const int32_t *,
                                const int32_t *, int64_t, int32_t *, int *, cudaStream_t);
int launch_rans64_decode_ar(const RansTables &, const uint32_t *, int64_t, const int32_t *, int, int, int, int, const int32_t *,
                            const int32_t *, const int32_t *, const int32_t *, int64_t, int, int, int32_t *, int *, cudaStream_t);

struct CtxModel;
CtxModel *ctx_new(int C, int G, int k, int device, int sm_count);
void ctx_delete(CtxModel *);
int ctx_set_weights(CtxModel &, const float *, const float *, const float *, const float *, const float *, const float *,
                    const float *, const float *);
int ctx_set_weights_internal(CtxModel &, const float *, const float *, const float *, const float *, const float *, const float *,
                             const float *, const float *, const float *, const float *, const float *, const float *, int);
int ctx_set_map(CtxModel &, const int32_t *, int, int);
int ctx_stage_params(CtxModel &, int, const float *, const float *, int, float *, cudaStream_t, const float *buf_cl = nullptr,
                     const float *prior_cl = nullptr, bool params_cl = false);
bool ctx_uses_tc(const CtxModel &, int B);
int ctx_to_cl(CtxModel &, const float *src, float *dst, int B, int channels, cudaStream_t);
size_t ctx_cl_elems(int B, int channels, int HW);
int ctx_num_stages(const CtxModel &);
int ctx_set_precision(CtxModel &, int, int);
int ctx_precision(const CtxModel &);
int ctx_run_precision(const CtxModel &);
const int32_t *ctx_perm(const CtxModel &);
int mma_bench(int mode, int ts, int n_cols, int iters, int same_acc, long long *cycles);
void ctx_set_run_precision(CtxModel &, int);
int ctx_range_flag_clear(CtxModel &, cudaStream_t);
int ctx_range_flag_read(CtxModel &, cudaStream_t, int *);
int ctx_range_flag_copy(CtxModel &, cudaStream_t, int *);
int ctx_stage_positions(const CtxModel &, int, const int32_t **, int64_t *);
int ctx_dims(const CtxModel &, int *C, int *G, int *H, int *W);
bool ctx_scan_supported(const CtxModel &, int B);
struct ScanDecodeHost {   // (ctx.cuh)
    const RansTables *tables;
    int bypass;
    const unsigned char *seg;
    long long seg_cap;
    int n_chunks;
    const int32_t *chunk_syms;
    int *status;
};
int ctx_scan_run(CtxModel &, int g0, int g1, float *buf, const float *prior, int B, float *params, const float *y, int32_t *sym,
                 int32_t *idx, const float *d_scale_table, int n_scales, cudaStream_t, const int32_t *dq_sym = nullptr,
                 const ScanDecodeHost *dec = nullptr);
bool ctx_scan_decode_supported(const CtxModel &, int B, int n_chunks, int bypass_precision, int freq_precision);

struct TansTables;
TansTables *tans_new();
void tans_delete(TansTables *);
int tans_init(TansTables &, const int32_t *, int, int, const int32_t *, const int32_t *, unsigned, int, unsigned, int role,
              cudaStream_t);
int tans_encode(TansTables &, const int32_t *d_sym, const int32_t *d_idx, int64_t n, uint8_t *d_out, int64_t cap,
                long long *d_len, int *d_status, cudaStream_t);
int tans_decode(TansTables &, const uint8_t *d_enc, int64_t len, const int32_t *d_idx, int64_t n, int32_t *d_out, int *d_status,
                cudaStream_t);

// ---- optional in-library profiling: CUDA events on the launching stream around the phases of the y path
enum { PROF_CTX = 0, PROF_GAUSS = 1, PROF_ENCODE = 2, PROF_DECODE = 3, PROF_N = 8 };
struct ProfSpan { int cat; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static double g_host_ms[PROF_N] = {0};  // host wall time of a few sections, slots 4.. (read with the device phases)
struct HostScope {
    int slot;
    std::chrono::steady_clock::time_point t0;
    explicit HostScope(int s) : slot(s), t0(std::chrono::steady_clock::now()) {}
    ~HostScope() { if (g_prof_on) g_host_ms[slot] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
static std::vector<ProfSpan> g_prof_spans;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    cudaStream_t s;
    ProfSpan sp{};
    bool on;
    ProfScope(int cat, cudaStream_t stream) : s(stream), on(g_prof_on)
    {
        if (!on) return;
        sp.cat = cat; sp.e0 = prof_event(); sp.e1 = prof_event();
        cudaEventRecord(sp.e0, s);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(sp.e1, s);
        g_prof_spans.push_back(sp);
    }
};

// Debug timeline of the y-path calls (BASIC_TRACE=1): host time and device time (an event on the launching stream) at a few
// points of a call, printed to stderr when the call ends.
struct Trace {
    bool on = getenv("BASIC_TRACE") != nullptr;
    struct Pt { const char *label; double host_ms; cudaEvent_t ev; };
    std::vector<Pt> pts;
    std::chrono::steady_clock::time_point t0;
    void mark(const char *label, cudaStream_t s)
    {
        if (!on) return;
        if (pts.empty()) t0 = std::chrono::steady_clock::now();
        cudaEvent_t e = nullptr;
        if (s != (cudaStream_t)-1) { cudaEventCreate(&e); cudaEventRecord(e, s); }
        pts.push_back({label, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), e});
    }
    void dump(const char *what, cudaStream_t s)
    {
        if (!on || pts.empty()) return;
        cudaStreamSynchronize(s);
        fprintf(stderr, "[trace %s]", what);
        for (auto &p : pts) {
            float d = -1.f;
            if (p.ev && pts[0].ev) cudaEventElapsedTime(&d, pts[0].ev, p.ev);
            fprintf(stderr, " %s h=%.3f d=%.3f |", p.label, p.host_ms, d);
            }
        fprintf(stderr, "\n");
        for (auto &p : pts) if (p.ev) cudaEventDestroy(p.ev);
        pts.clear();
    }
};
static Trace g_trace;

static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }
int64_t g_launches = 0;

static constexpr double kAutoBudget = 0.0043;  // target container overhead in auto mode (bar: 0.5 %; measured total = budget + 0.00 .. 0.03 %)
static constexpr int kChunkOverhead = 132;    // 32 states + directory entry

struct SliceDesc {  // mirrors rans_lanes.cu
    long long off, n;
    int cs, pad;
};

}  // namespace basic

using namespace basic;

struct basic_coder {
    int kind = 0, role = 0;
    unsigned precision = 16, max_symbol_value = 255, bypass_precision = 4;
    int bypass = 1, device = 0, sm_count = 148;
    RansTables rt;
    TansTables *tt = nullptr;
    bool initialized = false;
    // Gaussian conditional
    std::vector<float> h_scale;
    DevBuf d_scale;
    // scratch
    DevBuf in_a, in_b, out_i32, words, first, states, segs, small, stream_dev, y_dev, prior_dev, buf, params, sym_all, idx_all,
        yhat_stage, slices_dev, carry_x, carry_wp, buf_cl, prior_cl, batch_first, batch_meta, batch_state;
    DevBuf ar_table, ar_a, ar_b, ar_c, ar_eff;   // in-coder AR lookup: the table, staged ar_indexes / ar_offsets, effective indexes
    int ar_A = 0, ar_I = 0, ar_D1 = 0, ar_D2 = 0;
    void *pinned = nullptr;  // 256 B of pinned host memory for status / length read-back
    // pinned host staging (grow-only): encoded output kept for basic_coder_last_output, and the stream being decoded
    uint8_t *host_out = nullptr, *host_in = nullptr;
    const uint8_t *host_src = nullptr;  // where the current stream's bytes can be read on the host: host_in, or the caller's page-locked buffer
    size_t host_out_cap = 0, host_in_cap = 0;
    int64_t last_len = 0;
    std::vector<cudaEvent_t> out_events;  // one per chunk of the last device-to-host delivery into host_out
    int out_chunks = 0;                   // chunks of that delivery still to be awaited (0 = host_out is complete)
    cudaEvent_t in_event = nullptr;  // completion of the last upload out of host_in
    cudaStream_t copy_stream = nullptr;  // the upload runs beside whatever is already queued on the caller's stream
    cudaEvent_t ev_start = nullptr, ev_prior = nullptr, ev_y = nullptr;  // host inputs of the y path uploaded on copy_stream
    cudaEvent_t ev_sub[4] = {nullptr, nullptr, nullptr, nullptr};       // ... in image sub-batches: one event per sub-batch
    // cache for cache=1 / flush()
    std::vector<int32_t> cache_sym, cache_idx;         // lanes = 1: concatenated operands (device copies made at flush)
    std::vector<std::vector<uint8_t>> cache_segments;  // multi-lane: encoded segments
    // streaming decode state
    int stream_lanes = 1;
    bool stream_fp16 = false;  // the container was written with the context model in 3xFP16 ("BLS2")
    int64_t stream_len = 0;  // bytes of the stream in host_in / stream_dev
    int64_t stream_pos = 0;
    bool stream_set = false;
};

struct basic_ctx {
    CtxModel *m;
    int device;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Returns a device pointer for `p` (copying host data into `staging` when needed).
template <class T>
int to_device(const T *p, size_t count, DevBuf &staging, cudaStream_t s, const T **out)
{
    if (!p || count == 0) { *out = p; return BASIC_OK; }
    if (is_device_ptr(p)) { *out = p; return BASIC_OK; }
    BASIC_TRY(staging.reserve(count * sizeof(T) + 16));
    BASIC_CUDA(cudaMemcpyAsync(staging.p, p, count * sizeof(T), cudaMemcpyHostToDevice, s));
    *out = staging.as<T>();
    return BASIC_OK;
}

// Host inputs of the y path: uploaded on the coder's copy stream (prior first, then y), so that the caller's stream only
// waits for what the next kernel needs -- the first group's context model runs while y is still on the bus, and in the
// decoder the stream's staging copy runs while the prior is.  Device inputs pass through.
// Images are independent, so a large batch is uploaded in sub-batches (plan->n > 1: prior and y of sub-batch 0, then of
// sub-batch 1, ... with one event each) and the caller runs the groups of a sub-batch while the next one is on the bus.
struct SubPlan {
    int n = 1;
    int b0[5] = {0, 0, 0, 0, 0};   // sub-batch i = images [b0[i], b0[i + 1])
    bool host = false;             // events ev_sub[i] were recorded: the caller's stream waits for them itself
};

SubPlan plan_subs(int B, bool host_inputs, bool eligible)
{
    static const int forced = [] { const char *e = getenv("BASIC_SUB_BATCHES"); return e ? atoi(e) : 0; }();   // A/B switch (1 = off)
    SubPlan p;
    p.n = !host_inputs || !eligible || B < 8 ? 1 : (B >= 16 ? 4 : 2);
    if (forced > 0 && host_inputs && eligible) p.n = std::max(1, std::min(std::min(forced, 4), B));
    for (int i = 0; i <= p.n; ++i) p.b0[i] = (int)((long long)i * B / p.n);
    return p;
}

int upload_inputs(basic_coder *c, const float *y, size_t n_y, const float *prior, size_t n_prior, cudaStream_t s,
                  const float **d_y, const float **d_prior, bool *y_pending, SubPlan *plan = nullptr, int B = 1)
{
    *y_pending = false;
    const bool y_host = y && !is_device_ptr(y), p_host = prior && !is_device_ptr(prior);
    *d_y = y;
    *d_prior = prior;
    if (!y_host && !p_host) return BASIC_OK;
    if (!c->copy_stream) BASIC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&c->ev_start, &c->ev_prior, &c->ev_y, &c->ev_sub[0], &c->ev_sub[1], &c->ev_sub[2], &c->ev_sub[3]})
        if (!*e) BASIC_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (p_host) BASIC_TRY(c->prior_dev.reserve(n_prior * sizeof(float) + 16));
    if (y_host) BASIC_TRY(c->y_dev.reserve(n_y * sizeof(float) + 16));
    // the staging buffers may still be read by what an earlier call queued on `s`
    BASIC_CUDA(cudaEventRecord(c->ev_start, s));
    BASIC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_start, 0));
    if (p_host) *d_prior = c->prior_dev.as<float>();
    if (y_host) *d_y = c->y_dev.as<float>();
    if (plan && plan->n > 1) {
        const size_t yi = n_y / (size_t)B, pi = n_prior / (size_t)B;   // floats per image
        for (int i = 0; i < plan->n; ++i) {
            const size_t b0 = (size_t)plan->b0[i], nb = (size_t)(plan->b0[i + 1] - plan->b0[i]);
            if (p_host)
                BASIC_CUDA(cudaMemcpyAsync(c->prior_dev.as<float>() + b0 * pi, prior + b0 * pi, nb * pi * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            if (y_host)
                BASIC_CUDA(cudaMemcpyAsync(c->y_dev.as<float>() + b0 * yi, y + b0 * yi, nb * yi * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            BASIC_CUDA(cudaEventRecord(c->ev_sub[i], c->copy_stream));
        }
        plan->host = true;
        return BASIC_OK;
    }
    if (p_host) {
        BASIC_CUDA(cudaMemcpyAsync(c->prior_dev.p, prior, n_prior * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
        BASIC_CUDA(cudaEventRecord(c->ev_prior, c->copy_stream));
        BASIC_CUDA(cudaStreamWaitEvent(s, c->ev_prior, 0));
    }
    if (y_host) {
        BASIC_CUDA(cudaMemcpyAsync(c->y_dev.p, y, n_y * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
        BASIC_CUDA(cudaEventRecord(c->ev_y, c->copy_stream));
        *y_pending = true;  // the caller makes `s` wait for ev_y before the first kernel that reads y
    }
    return BASIC_OK;
}

int status_error(int st)
{
    if (st & 1) return value_error("index out of range of the coder tables");
    if (st & 2) return value_error("symbol out of table range and bypass_coding is disabled");
    if (st & 4) { set_error("malformed or truncated stream"); return BASIC_ERR_STREAM; }
    return BASIC_OK;
}

struct Small {  // layout of coder->small (device) and coder->pinned (host mirror)
    long long len[4];
    int status;
    float bits;
};

int need_init(basic_coder *c)
{
    if (!c) return value_error("null coder");
    if (!c->initialized) return value_error("ANS not initialized!");
    return BASIC_OK;
}

// lanes = 1: reference stream of (d_sym, d_idx) into c->segs; returns byte length via *len (host, synchronised).
int encode_compat(basic_coder *c, const int32_t *d_sym, const int32_t *d_idx, int64_t n, cudaStream_t s, const uint8_t **d_bytes,
                  int64_t *len)
{
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int64_t cap_words = attempt == 0 ? n + n / 4 + 64 : 12 * n + 64;
        BASIC_TRY(c->segs.reserve((size_t)cap_words * 4));
        BASIC_TRY(c->small.reserve(sizeof(Small)));
        BASIC_CUDA(cudaMemsetAsync(c->small.p, 0, sizeof(Small), s));
        Small *ds = c->small.as<Small>();
        BASIC_TRY(launch_rans64_encode(c->rt, d_sym, d_idx, n, c->bypass, (int)c->bypass_precision, c->segs.as<uint32_t>(),
                                       cap_words, &ds->len[0], &ds->status, s));
        Small *hs = reinterpret_cast<Small *>(c->pinned);
        BASIC_CUDA(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        if ((hs->status & 4) && !(hs->status & 3) && attempt == 0) continue;  // escape-heavy input: retry, worst-case size
        BASIC_TRY(status_error(hs->status));
        *d_bytes = reinterpret_cast<const uint8_t *>(c->segs.as<uint32_t>() + hs->len[0]);
        *len = (cap_words - hs->len[0]) * 4;
        return BASIC_OK;
    }
    return BASIC_ERR_CUDA;
}

// One multi-lane segment of `n_slices` slices (slice g = slice_n[g] symbols, laid out one after the other in
// d_sym / d_idx) into c->segs at byte offset `at` (4-aligned); *seg_len host value (synchronised).
// Chunk count: lanes / 32, or in auto mode as many as keep the flush overhead (132 B per chunk) inside the budget.
int encode_segment(basic_coder *c, const int32_t *d_sym, const int32_t *d_idx, int n_slices, const int64_t *slice_n, int lanes,
                   size_t at, cudaStream_t s, int64_t *seg_len)
{
    Small *ds = c->small.as<Small>();
    Small *hs = reinterpret_cast<Small *>(c->pinned);
    int64_t n = 0, n_max = 0;
    for (int g = 0; g < n_slices; ++g) { n += slice_n[g]; n_max = std::max(n_max, slice_n[g]); }
    int64_t want_chunks;
    if (lanes == BASIC_LANES_AUTO) {
        BASIC_CUDA(cudaMemsetAsync(&ds->bits, 0, sizeof(float), s));
        BASIC_TRY(launch_estimate_bits(c->rt, c->bypass, (int)c->bypass_precision, d_sym, d_idx, n, &ds->bits, s));
        BASIC_CUDA(cudaMemcpyAsync(&hs->bits, &ds->bits, sizeof(float), cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        const double est_bytes = (double)hs->bits / 8.0;
        want_chunks = (int64_t)(kAutoBudget * est_bytes / kChunkOverhead);
    } else {
        want_chunks = (lanes + kLanes - 1) / kLanes;
    }
    if (want_chunks < 1) want_chunks = 1;
    // chunk_syms[g] = ceil(n_g / want) rounded up to whole 128-symbol blocks; chunks that own nothing are dropped
    std::vector<SliceDesc> sl((size_t)n_slices);
    int64_t off = 0, n_chunks = 0, cs_sum = 0;
    for (int g = 0; g < n_slices; ++g) {
        int64_t cs = (slice_n[g] + want_chunks - 1) / want_chunks;
        cs = std::max<int64_t>(128, ((cs + 127) / 128) * 128);
        sl[g].off = off; sl[g].n = slice_n[g]; sl[g].cs = (int)cs; sl[g].pad = 0;
        off += slice_n[g];
        n_chunks = std::max(n_chunks, (slice_n[g] + cs - 1) / cs);
        cs_sum += cs;
    }
    BASIC_TRY(c->slices_dev.reserve(sizeof(SliceDesc) * (size_t)std::max(n_slices, 1)));
    if (n_slices) BASIC_CUDA(cudaMemcpyAsync(c->slices_dev.p, sl.data(), sizeof(SliceDesc) * (size_t)n_slices, cudaMemcpyHostToDevice, s));
    const size_t hdr = 8 + 4 * (size_t)n_slices + (size_t)n_chunks * 132;
    for (int attempt = 0; attempt < 2; ++attempt) {
        // (+ 512: the pair kernel checks the room for a whole block of four steps at once)
        int64_t cap_words64 = (attempt == 0 ? cs_sum + cs_sum / 4 + 64 : 12 * cs_sum + 64) + 512;
        cap_words64 = (cap_words64 + 7) & ~(int64_t)7;
        if (cap_words64 > 0x7fffffff) return value_error("chunk too large: use more lanes");
        const int cap_words = (int)cap_words64;
        const size_t seg_bound = hdr + (size_t)n_chunks * cap_words * 2 + 8;
        BASIC_TRY(c->words.reserve((size_t)n_chunks * cap_words * 2 + 16));
        BASIC_TRY(c->first.reserve((size_t)n_chunks * 4 + 16));
        BASIC_TRY(c->states.reserve((size_t)n_chunks * 128 + 16));
        if (c->segs.cap < at + seg_bound) {  // grow, keeping what earlier segments wrote
            DevBuf bigger;
            BASIC_TRY(bigger.reserve((at + seg_bound) * 2));
            if (at) BASIC_CUDA(cudaMemcpyAsync(bigger.p, c->segs.p, at, cudaMemcpyDeviceToDevice, s));
            BASIC_CUDA(cudaStreamSynchronize(s));
            c->segs.release();
            c->segs = bigger;
        }
        BASIC_CUDA(cudaMemsetAsync(c->small.p, 0, sizeof(Small), s));
        BASIC_TRY(launch_bls_encode(c->rt, c->bypass, (int)c->bypass_precision, d_sym, d_idx, n_slices, c->slices_dev.p, (int)n_chunks,
                                    c->words.as<uint16_t>(), cap_words, c->first.as<uint32_t>(), c->states.as<uint32_t>(),
                                    c->segs.as<unsigned char>() + at, &ds->len[0], &ds->status, c->sm_count, s));
        BASIC_CUDA(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        if ((hs->status & 4) && !(hs->status & 3) && attempt == 0) continue;
        BASIC_TRY(status_error(hs->status));
        *seg_len = hs->len[0];
        return BASIC_OK;
    }
    return BASIC_ERR_CUDA;
}

// Host copy INTO pinned staging memory with non-temporal stores.  A plain memcpy leaves the chunk dirty in the copying
// core's cache, and the DMA engine then pulls every line out of that cache: measured 16 - 20 GB/s on the bus for a
// freshly staged stream against 45 GB/s for the same buffer read from DRAM (tools/pcie_probe.py, BASIC_TRACE).
static void copy_streaming(uint8_t *dst, const uint8_t *src, size_t n)
{
    size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
#ifndef BASIC_HAVE_SSE2
    memcpy(dst, src, n);   // (other host architectures: plain copy)
    return;
#else
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 0);
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 1);
        const __m128i c2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 2);
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 3);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 0, a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 1, b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 2, c2);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 3, d);
        src += 64;
        dst += 64;
    }
    if (n & 63) memcpy(dst, src, n & 63);
    _mm_sfence();  // the stores are globally visible before the upload of the chunk is queued
#endif
}

int reserve_pinned(uint8_t **buf, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return BASIC_OK;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    BASIC_CUDA(cudaHostAlloc(reinterpret_cast<void **>(buf), want, cudaHostAllocDefault));
    *cap = want;
    return BASIC_OK;
}

// Delivers `len` encoded bytes that live on the device: into the caller's buffer, or (out == NULL) into the
// coder's pinned staging buffer, from where basic_coder_last_output hands them out without another device trip.
int finish_out(basic_coder *c);
// granularity of the pipelined host copies (BASIC_HOST_CHUNK_KB overrides)
static const int64_t kHostChunk = [] {
    const char *e = getenv("BASIC_HOST_CHUNK_KB");
    const int64_t kb = e ? atoll(e) : 1024;
    return (kb < 64 ? 64 : kb > 16384 ? 16384 : kb) << 10;
}();
constexpr int kHostThreadsMax = 8;
// host threads of the staging copies: BASIC_HOST_THREADS, else up to 4 but no more than this rank's share of the cores
// (LOCAL_WORLD_SIZE ranks per node under torchrun); 1 = the calling thread only
static const int kHostThreads = [] {
    const char *e = getenv("BASIC_HOST_THREADS");
    int v = 4;
    if (e) v = atoi(e);
    else {
        // a 16 MB stream staged by 4 threads takes 0.33 ms, its upload 0.30 ms: up to 8 threads where the rank has the cores
        const char *l = getenv("LOCAL_WORLD_SIZE");
        const int ranks = l ? std::max(1, atoi(l)) : 1;
        const int cores = (int)std::thread::hardware_concurrency();
        if (cores > 0) v = std::min(kHostThreadsMax, std::max(1, cores / ranks));
    }
    return v < 1 ? 1 : v > kHostThreadsMax ? kHostThreadsMax : v;
}();

// A few persistent host threads for the staging copies (a 16 MB memcpy on one core costs more than its bus transfer).
// run(nt, fn) executes fn(t) for t = 0 .. nt - 1 (t = 0 on the caller) and returns when all are done; one job at a time.
class HostPool {
public:
    static HostPool &get() { static HostPool *p = new HostPool(); return *p; }  // leaked on purpose: no join at exit
    void run(int nt, const std::function<void(int)> &fn)
    {
        std::lock_guard<std::mutex> job(job_mu_);
        if (nt > kHostThreads) nt = kHostThreads;
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn; want_ = nt - 1; done_ = 0; ++gen_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> l(mu_);
        cv_done_.wait(l, [&] { return done_ == want_; });
        fn_ = nullptr;
    }
private:
    HostPool()
    {
        for (int t = 1; t < kHostThreads; ++t) std::thread([this, t] { loop(t); }).detach();
    }
    void loop(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn = nullptr;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (t <= want_) fn = fn_;
            }
            if (!fn) continue;
            (*fn)(t);
            {
                std::lock_guard<std::mutex> l(mu_);
                ++done_;
            }
            cv_done_.notify_one();
        }
    }
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int want_ = 0, done_ = 0;
    unsigned long long gen_ = 0;
};

int copy_out(basic_coder *c, const void *d_src, int64_t len, uint8_t *out, int64_t cap, cudaStream_t s)
{
    if (out) {
        if (len > cap) { set_error("output buffer too small"); return BASIC_ERR_CAPACITY; }
        if (len > 0) BASIC_CUDA(cudaMemcpyAsync(out, d_src, (size_t)len, cudaMemcpyDefault, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        return BASIC_OK;
    }
    // into the pinned staging buffer, chunk by chunk with an event each: basic_coder_take_output copies chunk k to
    // the caller while chunk k + 1 is still on the bus (the call returns without waiting for the copies)
    BASIC_TRY(finish_out(c));  // (an unread earlier delivery; the buffer may move)
    BASIC_TRY(reserve_pinned(&c->host_out, &c->host_out_cap, (size_t)len));
    c->last_len = len;
    const int chunks = (int)((len + kHostChunk - 1) / kHostChunk);
    while ((int)c->out_events.size() < chunks) {
        cudaEvent_t e;
        BASIC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->out_events.push_back(e);
    }
    for (int k = 0; k < chunks; ++k) {
        const int64_t at = (int64_t)k * kHostChunk, nb = std::min(kHostChunk, len - at);
        BASIC_CUDA(cudaMemcpyAsync(c->host_out + at, static_cast<const uint8_t *>(d_src) + at, (size_t)nb, cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaEventRecord(c->out_events[k], s));
    }
    c->out_chunks = chunks;
    return BASIC_OK;
}

// Waits for the pending chunks of host_out (all of them) -- for the pointer-returning accessor.
int finish_out(basic_coder *c)
{
    for (int k = 0; k < c->out_chunks; ++k) BASIC_CUDA(cudaEventSynchronize(c->out_events[k]));
    c->out_chunks = 0;
    return BASIC_OK;
}

// Parses a segment header that lives in host memory: n_slices slices of slice_n[g] symbols are expected.
struct SegInfo {
    int n_chunks = 0, n_slices = 0;
    std::vector<int> cs;  // chunk_syms per slice
    int64_t len = 0;      // bytes of the segment
};

int parse_segment(const uint8_t *p, int64_t avail, int n_slices, const int64_t *slice_n, SegInfo *out)
{
    if (avail < 8) { set_error("truncated multi-lane stream"); return BASIC_ERR_STREAM; }
    uint32_t nc, ns;
    memcpy(&nc, p, 4);
    memcpy(&ns, p + 4, 4);
    if ((int64_t)ns != n_slices || avail < 8 + 4 * (int64_t)ns) { set_error("multi-lane segment does not match the coding groups"); return BASIC_ERR_STREAM; }
    out->cs.resize(ns);
    int64_t nc_need = 0;  // the encoder drops chunks that own nothing: n_chunks = max over slices of ceil(n_g / cs_g)
    for (uint32_t g = 0; g < ns; ++g) {
        uint32_t cs;
        memcpy(&cs, p + 8 + 4 * (int64_t)g, 4);
        if (cs == 0 || cs % 128 || cs > 0x7fffff80u || (slice_n[g] + cs - 1) / cs > (int64_t)nc) {
            set_error("multi-lane segment does not match the number of indexes");
            return BASIC_ERR_STREAM;
        }
        out->cs[g] = (int)cs;
        nc_need = std::max(nc_need, (slice_n[g] + cs - 1) / cs);
    }
    if (nc_need != (int64_t)nc) { set_error("multi-lane segment does not match the number of indexes"); return BASIC_ERR_STREAM; }
    const int64_t hdr = 8 + 4 * (int64_t)ns + (int64_t)nc * 132;
    if (nc > 0x3fffffffu || avail < hdr) { set_error("truncated multi-lane stream"); return BASIC_ERR_STREAM; }
    uint32_t total_words = 0;
    if (nc) memcpy(&total_words, p + 8 + 4 * (int64_t)ns + 4 * (int64_t)(nc - 1), 4);
    int64_t len = hdr + 2 * (int64_t)total_words;
    len = (len + 3) & ~(int64_t)3;
    if (len > avail) { set_error("truncated multi-lane stream"); return BASIC_ERR_STREAM; }
    out->n_chunks = (int)nc;
    out->n_slices = (int)ns;
    out->len = len;
    return BASIC_OK;
}

}  // namespace

// ============================================================================================== C ABI
extern "C" {

const char *basic_last_error(void) { return t_error.c_str(); }

int basic_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int basic_profile_enable(int on)
{
    g_prof_on = on != 0;
    return BASIC_OK;
}

int basic_profile_read(double *ms, int64_t *spans)
{
    for (int i = 0; i < PROF_N; ++i) { ms[i] = g_host_ms[i]; g_host_ms[i] = 0.0; spans[i] = 0; }
    for (auto &sp : g_prof_spans) {
        float t = 0.f;
        if (cudaEventSynchronize(sp.e1) == cudaSuccess && cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess) {
            ms[sp.cat] += t;
            spans[sp.cat] += 1;
        }
        g_prof_pool.push_back(sp.e0);
        g_prof_pool.push_back(sp.e1);
    }
    g_prof_spans.clear();
    cudaGetLastError();
    return BASIC_OK;
}

int basic_debug_mma_bench(int mode, int ts, int n_cols, int iters, int same_acc, long long *cycles)
{
    return mma_bench(mode, ts, n_cols, iters, same_acc, cycles);
}

int64_t basic_launch_count(int reset)
{
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int basic_coder_create(int kind, unsigned precision, unsigned max_symbol_value, int bypass_coding, unsigned bypass_precision,
                       int device, basic_coder **out)
{
    if (!out) return value_error("null out pointer");
    const int base_kind = kind & 0xf;
    if (base_kind != BASIC_KIND_RANS64 && base_kind != BASIC_KIND_TANS) return value_error("unknown coder kind");
    if (bypass_precision < 1 || bypass_precision > 8) return value_error("bypass_precision must be in [1, 8]");
    int ndev = 0;
    BASIC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device"); return BASIC_ERR_CUDA; }
    DeviceGuard guard(device);
    basic_coder *c = new basic_coder();
    c->kind = base_kind;
    c->role = kind >> 4;
    c->precision = precision;
    c->max_symbol_value = max_symbol_value;
    c->bypass = bypass_coding ? 1 : 0;
    c->bypass_precision = bypass_precision;
    c->device = device;
    cudaDeviceProp prop;
    BASIC_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    BASIC_CUDA(cudaHostAlloc(&c->pinned, 256, cudaHostAllocDefault));
    BASIC_TRY(c->small.reserve(sizeof(Small)));
    if (c->kind == BASIC_KIND_TANS) c->tt = tans_new();
    *out = c;
    return BASIC_OK;
}

void basic_coder_destroy(basic_coder *c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    DevBuf *bufs[] = {&c->rt.cdf32, &c->rt.blob, &c->rt.enc, &c->rt.blob_d, &c->ar_table, &c->ar_a, &c->ar_b, &c->ar_c, &c->ar_eff, &c->d_scale, &c->in_a, &c->in_b, &c->out_i32, &c->words, &c->first, &c->states,
                      &c->segs, &c->small, &c->stream_dev, &c->y_dev, &c->prior_dev, &c->buf, &c->params, &c->sym_all,
                      &c->idx_all, &c->yhat_stage, &c->slices_dev, &c->carry_x, &c->carry_wp, &c->buf_cl, &c->prior_cl, &c->batch_first,
                      &c->batch_meta, &c->batch_state};
    for (DevBuf *b : bufs) b->release();
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->host_out) cudaFreeHost(c->host_out);
    if (c->host_in) cudaFreeHost(c->host_in);
    if (c->in_event) cudaEventDestroy(c->in_event);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (cudaEvent_t e : {c->ev_start, c->ev_prior, c->ev_y}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->out_events) cudaEventDestroy(e);
    if (c->tt) tans_delete(c->tt);
    delete c;
}

int basic_coder_init_params(basic_coder *c, const int32_t *freqs, int T, int M, const int32_t *num_symbols, const int32_t *offsets)
{
    if (!c) return value_error("null coder");
    DeviceGuard guard(c->device);
    c->initialized = false;
    if (c->kind == BASIC_KIND_RANS64)
        BASIC_TRY(rans_tables_from_freqs(c->rt, freqs, T, M, num_symbols, offsets, (int)c->precision, 0));
    else
        BASIC_TRY(tans_init(*c->tt, freqs, T, M, num_symbols, offsets, c->precision, c->bypass, c->bypass_precision, c->role, 0));
    c->initialized = true;
    return BASIC_OK;
}

int basic_coder_init_cdf_params(basic_coder *c, const int32_t *cdfs, int T, int M, const int32_t *cdf_sizes, const int32_t *offsets)
{
    if (!c) return value_error("null coder");
    if (c->kind != BASIC_KIND_RANS64) return value_error("init_cdf_params is a rANS call");
    DeviceGuard guard(c->device);
    c->initialized = false;
    BASIC_TRY(rans_tables_from_cdfs(c->rt, cdfs, T, M, cdf_sizes, offsets, (int)c->precision, 0));
    c->initialized = true;
    return BASIC_OK;
}

int basic_coder_cdfs_shape(basic_coder *c, int *T, int *M)
{
    if (!c || c->kind != BASIC_KIND_RANS64 || !c->initialized) { *T = 0; *M = 0; return BASIC_OK; }
    *T = c->rt.T;
    *M = *std::max_element(c->rt.h_sizes.begin(), c->rt.h_sizes.end());
    return BASIC_OK;
}

int basic_coder_get_cdfs(basic_coder *c, int32_t *out)
{
    BASIC_TRY(need_init(c));
    DeviceGuard guard(c->device);
    int T, M;
    basic_coder_cdfs_shape(c, &T, &M);
    BASIC_CUDA(cudaMemcpy2D(out, (size_t)M * 4, c->rt.cdf32.p, (size_t)c->rt.stride * 4, (size_t)M * 4, T, cudaMemcpyDeviceToHost));
    return BASIC_OK;
}

int basic_pmf_to_quantized_cdf(const float *pmf, int n, int precision, int device, int32_t *cdf_out)
{
    DeviceGuard guard(device);
    return pmf_to_cdf_device(pmf, n, precision, cdf_out);
}

int64_t basic_coder_encode_bound(basic_coder *c, int64_t n, int lanes)
{
    if (c && c->kind == BASIC_KIND_TANS) return n * (int64_t)c->precision / 8 + 16;
    if (lanes == BASIC_LANES_REFERENCE) return (12 * n + 64) * 4;
    return 20 + ((n + 127) / 128) * 132 + 24 * n + 64;
}

int basic_coder_encode(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, int lanes, int cache,
                       uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream)
{
    BASIC_TRY(need_init(c));
    if (n < 0 || lanes < 0) return value_error("negative size");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (out_len) *out_len = 0;
    const int32_t *d_sym, *d_idx;
    if (cache && (c->kind == BASIC_KIND_TANS || lanes == BASIC_LANES_REFERENCE)) {
        // reference cache mode (rans64.cpp:327-345): operands are kept, the stream is produced by flush()
        std::vector<int32_t> hs((size_t)n), hi((size_t)n);
        BASIC_CUDA(cudaMemcpyAsync(hs.data(), symbols, (size_t)n * 4, cudaMemcpyDefault, s));
        BASIC_CUDA(cudaMemcpyAsync(hi.data(), indexes, (size_t)n * 4, cudaMemcpyDefault, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        c->cache_sym.insert(c->cache_sym.end(), hs.begin(), hs.end());
        c->cache_idx.insert(c->cache_idx.end(), hi.begin(), hi.end());
        return BASIC_OK;
    }
    BASIC_TRY(to_device(symbols, (size_t)n, c->in_a, s, &d_sym));
    BASIC_TRY(to_device(indexes, (size_t)n, c->in_b, s, &d_idx));
    if (c->kind == BASIC_KIND_TANS) {
        const int64_t cap = n * (int64_t)c->precision / 8;
        BASIC_TRY(c->segs.reserve((size_t)cap + 64));
        BASIC_CUDA(cudaMemsetAsync(c->small.p, 0, sizeof(Small), s));
        Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
        BASIC_TRY(tans_encode(*c->tt, d_sym, d_idx, n, c->segs.as<uint8_t>(), cap, &ds->len[0], &ds->status, s));
        BASIC_CUDA(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        if (hs->status & 8) return value_error("Destination buffer is too small");
        BASIC_TRY(status_error(hs->status));
        BASIC_TRY(copy_out(c, c->segs.p, hs->len[0], out, out_cap, s));
        if (out_len) *out_len = hs->len[0];
        return BASIC_OK;
    }
    if (lanes == BASIC_LANES_REFERENCE) {
        const uint8_t *d_bytes;
        int64_t len;
        BASIC_TRY(encode_compat(c, d_sym, d_idx, n, s, &d_bytes, &len));
        BASIC_TRY(copy_out(c, d_bytes, len, out, out_cap, s));
        if (out_len) *out_len = len;
        return BASIC_OK;
    }
    // multi-lane container: magic | segment
    BASIC_TRY(c->segs.reserve(64));
    int64_t seg_len = 0;
    {
        ProfScope ps(PROF_ENCODE, s);
        BASIC_TRY(encode_segment(c, d_sym, d_idx, 1, &n, lanes, 4, s, &seg_len));
    }
    if (cache) {
        std::vector<uint8_t> seg((size_t)seg_len);
        BASIC_CUDA(cudaMemcpyAsync(seg.data(), c->segs.as<unsigned char>() + 4, (size_t)seg_len, cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        c->cache_segments.push_back(std::move(seg));
        return BASIC_OK;
    }
    BASIC_CUDA(cudaMemcpyAsync(c->segs.p, &kMagic, 4, cudaMemcpyHostToDevice, s));
    BASIC_TRY(copy_out(c, c->segs.p, 4 + seg_len, out, out_cap, s));
    if (out_len) *out_len = 4 + seg_len;
    return BASIC_OK;
}

int basic_coder_last_output(basic_coder *c, const uint8_t **ptr, int64_t *len)
{
    if (!c || !ptr || !len) return value_error("null argument");
    DeviceGuard guard(c->device);
    BASIC_TRY(finish_out(c));
    *ptr = c->host_out;
    *len = c->last_len;
    return BASIC_OK;
}

int64_t basic_coder_output_size(basic_coder *c) { return c ? c->last_len : 0; }

// delivery of a stream into the caller's bytes object with non-temporal stores: measured (cfg2 bench step) 3 % slower with one
// rank per node -- the decoder that follows reads the bytes out of DRAM instead of the cache -- and 10 % faster with eight, where
// the ranks compete for host memory bandwidth (no read-for-ownership: a third less traffic).  BASIC_NT_DELIVERY=0 / 1 overrides.
static const bool kNtDelivery = [] {
    const char *e = getenv("BASIC_NT_DELIVERY");
    if (e) return e[0] == '1';
    const char *l = getenv("LOCAL_WORLD_SIZE");
    return l && atoi(l) >= 4;
}();

int basic_coder_take_output(basic_coder *c, uint8_t *dst, int64_t cap)
{
    if (!c || (!dst && c->last_len)) return value_error("null argument");
    if (cap < c->last_len) { set_error("output buffer too small"); return BASIC_ERR_CAPACITY; }
    DeviceGuard guard(c->device);
    const int64_t len = c->last_len;
    const int chunks = (int)((len + kHostChunk - 1) / kHostChunk);
    const int pending = c->out_chunks;  // chunks [0, pending) still have an event to wait for
    const int nt = std::max(1, std::min(kHostThreads, chunks));
    // slices of a chunk (an eighth) are handed out dynamically: every thread works on the chunk that has just arrived, so what is
    // left to copy after the LAST chunk lands is one slice per thread, not one chunk on one thread
    const int64_t slice = std::max<int64_t>(kHostChunk / 8, 64 << 10);
    const int slices = (int)((len + slice - 1) / slice), per_chunk = (int)(kHostChunk / slice);
    std::atomic<int> next{0};
    const std::function<void(int)> work = [&](int t) {
        if (t > 0) cudaSetDevice(c->device);
        for (int i; (i = next.fetch_add(1, std::memory_order_relaxed)) < slices;) {
            const int k = i / per_chunk;
            if (k < pending) cudaEventSynchronize(c->out_events[k]);
            const int64_t at = (int64_t)i * slice;
            if (kNtDelivery) copy_streaming(dst + at, c->host_out + at, (size_t)std::min(slice, len - at));
            else memcpy(dst + at, c->host_out + at, (size_t)std::min(slice, len - at));
        }
    };
    if (nt <= 1) work(0);
    else HostPool::get().run(nt, work);
    c->out_chunks = 0;
    return cudaGetLastError() == cudaSuccess ? BASIC_OK : value_error("device-to-host delivery failed");
}

int basic_coder_flush(basic_coder *c, int lanes, uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream)
{
    BASIC_TRY(need_init(c));
    DeviceGuard guard(c->device);
    if (out_len) *out_len = 0;
    if (c->kind == BASIC_KIND_TANS || lanes == BASIC_LANES_REFERENCE) {
        std::vector<int32_t> sym, idx;
        sym.swap(c->cache_sym);
        idx.swap(c->cache_idx);
        return basic_coder_encode(c, sym.data(), idx.data(), (int64_t)sym.size(), BASIC_LANES_REFERENCE, 0, out, out_cap, out_len,
                                  stream);
    }
    int64_t total = 4;
    for (auto &sg : c->cache_segments) total += (int64_t)sg.size();
    if (!out) {
        BASIC_TRY(finish_out(c));
        BASIC_TRY(reserve_pinned(&c->host_out, &c->host_out_cap, (size_t)total));
        out = c->host_out;
        c->last_len = total;
    } else if (total > out_cap) { set_error("output buffer too small"); return BASIC_ERR_CAPACITY; }
    std::vector<uint8_t> all((size_t)total);
    memcpy(all.data(), &kMagic, 4);
    size_t at = 4;
    for (auto &sg : c->cache_segments) { memcpy(all.data() + at, sg.data(), sg.size()); at += sg.size(); }
    c->cache_segments.clear();
    BASIC_CUDA(cudaMemcpy(out, all.data(), (size_t)total, cudaMemcpyDefault));
    if (out_len) *out_len = total;
    return BASIC_OK;
}

int basic_coder_set_stream(basic_coder *c, const uint8_t *encoded, int64_t len, int lanes, void *stream)
{
    BASIC_TRY(need_init(c));
    if (c->kind == BASIC_KIND_TANS) { c->stream_set = true; return BASIC_OK; }  // tans.cpp:824-836: stores the string only
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (len < 0) return value_error("negative length");
    HostScope hs_set(4);
    // host copy in pinned memory (segment directories are parsed on the host; the upload runs asynchronously)
    if (c->in_event && (size_t)len + 64 > c->host_in_cap) BASIC_CUDA(cudaEventSynchronize(c->in_event));
    BASIC_TRY(reserve_pinned(&c->host_in, &c->host_in_cap, (size_t)len + 64));
    if ((size_t)len + 64 > c->stream_dev.cap) BASIC_CUDA(cudaStreamSynchronize(s));
    BASIC_TRY(c->stream_dev.reserve((size_t)len + 64));
    c->host_src = c->host_in;
    if (len && is_pinned_host_ptr(encoded)) {
        // page-locked input (e.g. the zero-copy view an encoder hands out): no staging copy -- one upload on the copy stream,
        // directories parsed where the bytes are.  The caller keeps the buffer alive while this stream is being decoded.
        if (!c->in_event) BASIC_CUDA(cudaEventCreateWithFlags(&c->in_event, cudaEventDisableTiming));
        if (!c->copy_stream) BASIC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        BASIC_CUDA(cudaMemcpyAsync(c->stream_dev.p, encoded, (size_t)len, cudaMemcpyHostToDevice, c->copy_stream));
        BASIC_CUDA(cudaMemsetAsync(c->stream_dev.as<uint8_t>() + len, 0, 64, c->copy_stream));
        BASIC_CUDA(cudaEventRecord(c->in_event, c->copy_stream));
        BASIC_CUDA(cudaStreamWaitEvent(s, c->in_event, 0));
        c->host_src = encoded;
        c->stream_len = len;
        c->stream_lanes = lanes;
        c->stream_set = true;
        goto staged;
    } else if (len && is_device_ptr(encoded)) {
        BASIC_CUDA(cudaMemcpyAsync(c->host_in, encoded, (size_t)len, cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaMemcpyAsync(c->stream_dev.p, encoded, (size_t)len, cudaMemcpyDeviceToDevice, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
    } else if (len) {
        if (!c->in_event) BASIC_CUDA(cudaEventCreateWithFlags(&c->in_event, cudaEventDisableTiming));
        if (!c->copy_stream) BASIC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        BASIC_CUDA(cudaEventSynchronize(c->in_event));  // an earlier upload out of host_in may still be in flight
        // staged through pinned memory chunk by chunk: every chunk goes on the bus as soon as it is copied, the host
        // copies run on a few threads, and the upload has its own stream: what the caller has already queued on `s`
        // (the y path: the first group's context model) runs beside it; `s` waits for the upload from here on.
        // (every decoding call ends with a synchronisation of `s`, so nothing is still reading stream_dev)
        const int chunks = (int)((len + kHostChunk - 1) / kHostChunk);
        const int nt = std::min(kHostThreads, chunks);
        cudaStream_t up = c->copy_stream;
        // workers only copy; the calling thread issues the uploads in chunk order as they become ready (CUDA calls from
        // several threads on one stream serialise in the driver and cost more than they save)
        // The unit of the host copies is a SLICE (an eighth of a chunk): all threads work on the first chunk first, so its upload
        // starts after one slice time instead of one chunk time (with whole chunks per thread the first eight chunks became
        // ready together, 125 us in, and the bus idled until then).
        const int64_t slice = std::max<int64_t>(kHostChunk / 8, 64 << 10);
        const int slices = (int)((len + slice - 1) / slice), per_chunk = (int)(kHostChunk / slice);
        std::vector<std::atomic<int>> ready((size_t)chunks), left((size_t)chunks);
        for (int k = 0; k < chunks; ++k) {
            ready[(size_t)k].store(0, std::memory_order_relaxed);
            left[(size_t)k].store(std::min(per_chunk, slices - k * per_chunk), std::memory_order_relaxed);
        }
        auto copy_slice = [&](int i) {
            const int64_t at = (int64_t)i * slice, nb = std::min(slice, len - at);
            copy_streaming(c->host_in + at, encoded + at, (size_t)nb);
            const int k = i / per_chunk;
            if (left[(size_t)k].fetch_sub(1, std::memory_order_acq_rel) == 1) ready[(size_t)k].store(1, std::memory_order_release);
        };
        auto copy_chunk = [&](int k) {
            for (int i = k * per_chunk; i < std::min(slices, (k + 1) * per_chunk); ++i) copy_slice(i);
        };
        auto upload_chunk = [&](int k) {
            const int64_t at = (int64_t)k * kHostChunk, nb = std::min(kHostChunk, len - at);
            cudaMemcpyAsync(c->stream_dev.as<uint8_t>() + at, c->host_in + at, (size_t)nb, cudaMemcpyHostToDevice, up);
        };
        HostScope hs_copy(5);
        g_trace.mark("staging starts (copy stream)", up);
        if (nt <= 1) {
            for (int k = 0; k < chunks; ++k) { copy_chunk(k); upload_chunk(k); }
        } else {
            // chunks are handed out dynamically (a worker that wakes up late just takes fewer); the calling thread copies too
            // and, between its own chunks, uploads whatever prefix is ready
            std::atomic<int> next{0};
            const std::function<void(int)> work = [&](int t) {
                if (t > 0) {
                    for (int i; (i = next.fetch_add(1, std::memory_order_relaxed)) < slices;) copy_slice(i);
                    return;
                }
                int issued = 0;
                while (issued < chunks) {
                    const int i = next.load(std::memory_order_relaxed) < slices ? next.fetch_add(1, std::memory_order_relaxed) : slices;
                    if (i < slices) copy_slice(i);
                    while (issued < chunks && ready[(size_t)issued].load(std::memory_order_acquire)) upload_chunk(issued++);
                    if (i >= slices && issued < chunks && !ready[(size_t)issued].load(std::memory_order_acquire)) std::this_thread::yield();
                }
            };
            HostPool::get().run(nt, work);
        }
        BASIC_CUDA(cudaGetLastError());
        g_trace.mark("uploads issued (copy stream)", up);
        BASIC_CUDA(cudaMemsetAsync(c->stream_dev.as<uint8_t>() + len, 0, 64, up));
        BASIC_CUDA(cudaEventRecord(c->in_event, up));
        BASIC_CUDA(cudaStreamWaitEvent(s, c->in_event, 0));
        c->stream_len = len;
        c->stream_lanes = lanes;
        c->stream_set = true;
        goto staged;
    }
    BASIC_CUDA(cudaMemsetAsync(c->stream_dev.as<uint8_t>() + len, 0, 64, s));
    c->stream_len = len;
    c->stream_lanes = lanes;
    c->stream_set = true;
staged:
    if (lanes == BASIC_LANES_REFERENCE) {
        c->stream_pos = -1;  // state is initialised by the first decode_stream launch
    } else {
        uint32_t magic = 0;
        if (len >= 4) memcpy(&magic, c->host_src, 4);
        if (magic != kMagic && magic != kMagic2 && magic != kMagic0) { c->stream_set = false; set_error("not a multi-lane (BLS) container"); return BASIC_ERR_STREAM; }
        c->stream_fp16 = magic == kMagic2;
        c->stream_pos = 4;
    }
    return BASIC_OK;
}

int basic_coder_decode_stream(basic_coder *c, const int32_t *indexes, int64_t n, int32_t *out, void *stream)
{
    BASIC_TRY(need_init(c));
    if (c->kind == BASIC_KIND_TANS) return value_error("decode_stream is not implemented for tANS (reference stub, tans.cpp:838-915)");
    if (!c->stream_set) return value_error("set_stream has not been called");
    if (n < 0) return value_error("negative size");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int32_t *d_idx;
    BASIC_TRY(to_device(indexes, (size_t)n, c->in_b, s, &d_idx));
    int32_t *d_out = out;
    const bool out_dev = is_device_ptr(out);
    if (!out_dev) { BASIC_TRY(c->out_i32.reserve((size_t)n * 4 + 16)); d_out = c->out_i32.as<int32_t>(); }
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    BASIC_CUDA(cudaMemsetAsync(&ds->status, 0, sizeof(int), s));
    if (c->stream_lanes == BASIC_LANES_REFERENCE) {
        const int init = c->stream_pos < 0;
        BASIC_TRY(launch_rans64_decode(c->rt, c->stream_dev.as<uint32_t>(), c->stream_len / 4, &ds->len[2], init, d_idx,
                                       n, c->bypass, (int)c->bypass_precision, d_out, &ds->status, s));
        c->stream_pos = 0;
    } else {
        if (n > 0 || c->stream_pos < c->stream_len) {
            SegInfo si;
            BASIC_TRY(parse_segment(c->host_src + c->stream_pos, c->stream_len - c->stream_pos, 1, &n, &si));
            ProfScope ps(PROF_DECODE, s);
            BASIC_TRY(launch_bls_decode(c->rt, c->bypass, (int)c->bypass_precision, c->stream_dev.as<unsigned char>() + c->stream_pos,
                                        si.len, d_idx, n, si.cs[0], si.n_chunks, 1, 0, nullptr, nullptr, d_out, &ds->status,
                                        c->sm_count, s));
            c->stream_pos += si.len;
        }
    }
    BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!out_dev && n) BASIC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    return status_error(hs->status);
}

int basic_coder_decode(basic_coder *c, const uint8_t *encoded, int64_t len, const int32_t *indexes, int64_t n, int lanes,
                       int32_t *out, void *stream)
{
    BASIC_TRY(need_init(c));
    if (c->kind == BASIC_KIND_TANS) {
        DeviceGuard guard(c->device);
        cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
        if (len < 1) return value_error("Src size incorrect");
        const int32_t *d_idx;
        BASIC_TRY(to_device(indexes, (size_t)n, c->in_b, s, &d_idx));
        BASIC_TRY(c->stream_dev.reserve((size_t)len + 64));
        BASIC_CUDA(cudaMemsetAsync(c->stream_dev.p, 0, (size_t)len + 64, s));
        BASIC_CUDA(cudaMemcpyAsync(c->stream_dev.as<uint8_t>() + 16, encoded, (size_t)len, cudaMemcpyDefault, s));
        int32_t *d_out = out;
        const bool out_dev = is_device_ptr(out);
        if (!out_dev) { BASIC_TRY(c->out_i32.reserve((size_t)n * 4 + 16)); d_out = c->out_i32.as<int32_t>(); }
        Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
        BASIC_CUDA(cudaMemsetAsync(&ds->status, 0, sizeof(int), s));
        BASIC_TRY(tans_decode(*c->tt, c->stream_dev.as<uint8_t>() + 16, len, d_idx, n, d_out, &ds->status, s));
        BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
        if (!out_dev && n) BASIC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        if (hs->status & 8) return value_error("Error (generic)");  // end mark not present
        return status_error(hs->status);
    }
    BASIC_TRY(basic_coder_set_stream(c, encoded, len, lanes, stream));
    return basic_coder_decode_stream(c, indexes, n, out, stream);
}

// ---- in-coder autoregressive table lookup (reference lanes = 1 stream only) ---------------------------------------------
int basic_coder_init_ar_params(basic_coder *c, const int32_t *ar_tables, int A, int I, int D1, int D2)
{
    if (!c) return value_error("null coder");
    if (c->kind != BASIC_KIND_RANS64) return value_error("the AR lookup is built for the rANS coder");
    if (!ar_tables || A < 1 || I < 1 || D1 < 1 || D2 < 0)
        return value_error("ar_tables should be at least 3-dimensional with shape (ar_tables_size, index_dim, *ar_order_dims)");
    DeviceGuard guard(c->device);
    const size_t count = (size_t)A * I * D1 * (D2 ? D2 : 1);
    BASIC_TRY(c->ar_table.reserve(count * 4));
    BASIC_CUDA(cudaMemcpy(c->ar_table.p, ar_tables, count * 4, cudaMemcpyDefault));
    c->ar_A = A; c->ar_I = I; c->ar_D1 = D1; c->ar_D2 = D2;
    return BASIC_OK;
}

static int ar_args(basic_coder *c, const int32_t *ar_indexes, const int32_t *ar_offsets, int order, int64_t n, cudaStream_t s,
                   const int32_t **d_ai, const int32_t **d_o0, const int32_t **d_o1)
{
    if (!c->ar_A) return value_error("init_ar_params has not been called");
    if (!ar_offsets) return value_error("ar_offsets is required for ar coding!");
    if (order != (c->ar_D2 ? 2 : 1)) return value_error("ar_offsets should have one row per AR order of the tables");
    *d_ai = nullptr;
    if (ar_indexes) BASIC_TRY(to_device(ar_indexes, (size_t)n, c->ar_a, s, d_ai));
    const int32_t *d_off;
    BASIC_TRY(to_device(ar_offsets, (size_t)n * order, c->ar_b, s, &d_off));
    *d_o0 = d_off;
    *d_o1 = order == 2 ? d_off + n : nullptr;
    return BASIC_OK;
}

int basic_coder_encode_ar(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *ar_indexes,
                          const int32_t *ar_offsets, int order, uint8_t *out, int64_t out_cap, int64_t *out_len, void *stream)
{
    BASIC_TRY(need_init(c));
    if (n < 0) return value_error("negative size");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (out_len) *out_len = 0;
    const int32_t *d_sym, *d_idx, *d_ai, *d_o0, *d_o1;
    BASIC_TRY(ar_args(c, ar_indexes, ar_offsets, order, n, s, &d_ai, &d_o0, &d_o1));
    BASIC_TRY(to_device(symbols, (size_t)n, c->in_a, s, &d_sym));
    BASIC_TRY(to_device(indexes, (size_t)n, c->in_b, s, &d_idx));
    BASIC_TRY(c->ar_eff.reserve((size_t)n * 4 + 16));
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    BASIC_CUDA(cudaMemsetAsync(c->small.p, 0, sizeof(Small), s));
    // the encoder knows every symbol: effective table indexes in parallel, then the reference stream
    BASIC_TRY(launch_ar_effective_indexes(c->ar_table.as<int32_t>(), c->ar_A, c->ar_I, c->ar_D1, c->ar_D2, d_ai, d_o0, d_o1, d_sym, d_idx,
                                          n, c->ar_eff.as<int32_t>(), &ds->status, s));
    BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    if (hs->status & 1) return value_error("AR lookup out of range of the AR tables");
    const uint8_t *d_bytes;
    int64_t len;
    BASIC_TRY(encode_compat(c, d_sym, c->ar_eff.as<int32_t>(), n, s, &d_bytes, &len));
    BASIC_TRY(copy_out(c, d_bytes, len, out, out_cap, s));
    if (out_len) *out_len = len;
    return BASIC_OK;
}

int basic_coder_decode_ar(basic_coder *c, const uint8_t *encoded, int64_t len, const int32_t *indexes, int64_t n,
                          const int32_t *ar_indexes, const int32_t *ar_offsets, int order, int32_t *out, void *stream)
{
    BASIC_TRY(need_init(c));
    if (n < 0 || len < 0 || (len & 3)) return value_error("a reference rANS64 stream is a whole number of 32-bit words");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int32_t *d_idx, *d_ai, *d_o0, *d_o1;
    BASIC_TRY(ar_args(c, ar_indexes, ar_offsets, order, n, s, &d_ai, &d_o0, &d_o1));
    BASIC_TRY(to_device(indexes, (size_t)n, c->in_b, s, &d_idx));
    BASIC_TRY(c->stream_dev.reserve((size_t)len + 64));
    if (len) BASIC_CUDA(cudaMemcpyAsync(c->stream_dev.p, encoded, (size_t)len, cudaMemcpyDefault, s));
    int32_t *d_out = out;
    const bool out_dev = is_device_ptr(out);
    if (!out_dev) { BASIC_TRY(c->out_i32.reserve((size_t)n * 4 + 16)); d_out = c->out_i32.as<int32_t>(); }
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    BASIC_CUDA(cudaMemsetAsync(&ds->status, 0, sizeof(int), s));
    BASIC_TRY(launch_rans64_decode_ar(c->rt, c->stream_dev.as<uint32_t>(), len / 4, c->ar_table.as<int32_t>(), c->ar_A, c->ar_I, c->ar_D1,
                                      c->ar_D2, d_ai, d_o0, d_o1, d_idx, n, c->bypass, (int)c->bypass_precision, d_out, &ds->status, s));
    BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!out_dev && n) BASIC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    c->stream_set = false;  // stream_dev was reused
    return status_error(hs->status);
}

// ---- batches of independent reference streams (the z node: one lanes=1 stream per image, compressai_coder.py:233,242) --
// n_streams runs of n symbols each ([n_streams, n] row-major, host or device) -> n_streams reference rANS64 streams, coded
// by one CTA each in ONE launch.  The streams are delivered back to back (into `out`, or with out == NULL into the coder's
// pinned buffer: basic_coder_last_output); out_lens (host, [n_streams]) receives their byte lengths.
int basic_coder_encode_batch(basic_coder *c, const int32_t *symbols, const int32_t *indexes, int64_t n, int n_streams, uint8_t *out,
                             int64_t out_cap, int64_t *out_lens, void *stream)
{
    BASIC_TRY(need_init(c));
    if (c->kind != BASIC_KIND_RANS64) return value_error("batched streams are a rANS call");
    if (n < 0 || n_streams < 0 || !out_lens) return value_error("negative size");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    BASIC_TRY(finish_out(c));
    c->last_len = 0;
    if (n_streams == 0) return BASIC_OK;
    const int32_t *d_sym, *d_idx;
    BASIC_TRY(to_device(symbols, (size_t)n * n_streams, c->in_a, s, &d_sym));
    BASIC_TRY(to_device(indexes, (size_t)n * n_streams, c->in_b, s, &d_idx));
    BASIC_TRY(c->small.reserve(sizeof(Small)));
    BASIC_TRY(c->batch_first.reserve(sizeof(long long) * (size_t)n_streams));
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    std::vector<long long> first((size_t)n_streams);
    int64_t cap_words = 0;
    for (int attempt = 0;; ++attempt) {
        cap_words = attempt == 0 ? n + n / 4 + 64 : 12 * n + 64;
        BASIC_TRY(c->segs.reserve((size_t)cap_words * 4 * (size_t)n_streams));
        BASIC_CUDA(cudaMemsetAsync(c->small.p, 0, sizeof(Small), s));
        BASIC_TRY(launch_rans64_encode(c->rt, d_sym, d_idx, n, c->bypass, (int)c->bypass_precision, c->segs.as<uint32_t>(), cap_words,
                                       c->batch_first.as<long long>(), &ds->status, s, n_streams));
        BASIC_CUDA(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaMemcpyAsync(first.data(), c->batch_first.p, sizeof(long long) * (size_t)n_streams, cudaMemcpyDeviceToHost, s));
        BASIC_CUDA(cudaStreamSynchronize(s));
        if ((hs->status & 4) && !(hs->status & 3) && attempt == 0) continue;  // escape-heavy input: retry, worst-case size
        BASIC_TRY(status_error(hs->status));
        break;
    }
    int64_t total = 0;
    for (int b = 0; b < n_streams; ++b) { out_lens[b] = (cap_words - first[(size_t)b]) * 4; total += out_lens[b]; }
    uint8_t *dst = out;
    if (out) {
        if (total > out_cap) { set_error("output buffer too small"); return BASIC_ERR_CAPACITY; }
    } else {
        BASIC_TRY(reserve_pinned(&c->host_out, &c->host_out_cap, (size_t)total));
        dst = c->host_out;
        c->last_len = total;
    }
    int64_t at = 0;
    for (int b = 0; b < n_streams; ++b) {
        const uint32_t *src = c->segs.as<uint32_t>() + (int64_t)b * cap_words + first[(size_t)b];
        if (out_lens[b]) BASIC_CUDA(cudaMemcpyAsync(dst + at, src, (size_t)out_lens[b], cudaMemcpyDefault, s));
        at += out_lens[b];
    }
    BASIC_CUDA(cudaStreamSynchronize(s));
    return BASIC_OK;
}

// The mirror: `encoded` holds n_streams reference streams back to back (host or device), lens (host) their byte lengths;
// every stream decodes n symbols with its row of indexes ([n_streams, n]) into its row of out.
int basic_coder_decode_batch(basic_coder *c, const uint8_t *encoded, const int64_t *lens, int n_streams, const int32_t *indexes,
                             int64_t n, int32_t *out, void *stream)
{
    BASIC_TRY(need_init(c));
    if (c->kind != BASIC_KIND_RANS64) return value_error("batched streams are a rANS call");
    if (n < 0 || n_streams < 0 || (n_streams && !lens)) return value_error("negative size");
    if (n_streams == 0 || n == 0) return BASIC_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    std::vector<long long> meta(2 * (size_t)n_streams);  // word offsets | word counts
    int64_t total = 0;
    for (int b = 0; b < n_streams; ++b) {
        if (lens[b] < 0 || (lens[b] & 3)) return value_error("a reference rANS64 stream is a whole number of 32-bit words");
        meta[(size_t)b] = total / 4;
        meta[(size_t)n_streams + b] = lens[b] / 4;
        total += lens[b];
    }
    const int32_t *d_idx;
    BASIC_TRY(to_device(indexes, (size_t)n * n_streams, c->in_b, s, &d_idx));
    BASIC_TRY(c->stream_dev.reserve((size_t)total + 64));
    if (total) BASIC_CUDA(cudaMemcpyAsync(c->stream_dev.p, encoded, (size_t)total, cudaMemcpyDefault, s));
    BASIC_TRY(c->batch_meta.reserve(sizeof(long long) * meta.size()));
    BASIC_CUDA(cudaMemcpyAsync(c->batch_meta.p, meta.data(), sizeof(long long) * meta.size(), cudaMemcpyHostToDevice, s));
    BASIC_TRY(c->batch_state.reserve(16 * (size_t)n_streams));
    BASIC_TRY(c->small.reserve(sizeof(Small)));
    int32_t *d_out = out;
    const bool out_dev = is_device_ptr(out);
    if (!out_dev) { BASIC_TRY(c->out_i32.reserve((size_t)n * n_streams * 4 + 16)); d_out = c->out_i32.as<int32_t>(); }
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    BASIC_CUDA(cudaMemsetAsync(&ds->status, 0, sizeof(int), s));
    BASIC_TRY(launch_rans64_decode(c->rt, c->stream_dev.as<uint32_t>(), 0, c->batch_state.p, 1, d_idx, n, c->bypass,
                                   (int)c->bypass_precision, d_out, &ds->status, s, n_streams, c->batch_meta.as<long long>(),
                                   c->batch_meta.as<long long>() + n_streams));
    BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!out_dev) BASIC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * n_streams * 4, cudaMemcpyDeviceToHost, s));
    BASIC_CUDA(cudaStreamSynchronize(s));
    c->stream_set = false;  // stream_dev was reused
    return status_error(hs->status);
}

// ------------------------------------------------------------------------------- Gaussian conditional
int basic_coder_set_scale_table(basic_coder *c, const float *scale_table, int n_scales)
{
    if (!c) return value_error("null coder");
    if (n_scales < 1 || n_scales > 256) return value_error("scale table must have 1..256 entries");
    DeviceGuard guard(c->device);
    c->h_scale.assign(scale_table, scale_table + n_scales);
    for (int i = 1; i < n_scales; ++i)
        if (!(c->h_scale[i] > c->h_scale[i - 1])) return value_error("scale table must be strictly increasing");
    BASIC_TRY(c->d_scale.reserve(sizeof(float) * n_scales));
    BASIC_CUDA(cudaMemcpy(c->d_scale.p, c->h_scale.data(), sizeof(float) * n_scales, cudaMemcpyHostToDevice));
    return BASIC_OK;
}

int basic_gauss_quantize_index(basic_coder *c, const float *y, const float *params, const int32_t *positions, int64_t n_pos, int B,
                               int C, int HW, int32_t *symbols, int32_t *indexes, float *yhat_buf, void *stream)
{
    if (!c || c->h_scale.empty()) return value_error("scale table not set");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!is_device_ptr(params) || (y && !is_device_ptr(y)) || !is_device_ptr(indexes) || (positions && !is_device_ptr(positions)) ||
        (y && !is_device_ptr(symbols)) || (yhat_buf && !is_device_ptr(yhat_buf)))
        return value_error("basic_gauss_quantize_index works on device memory");
    return launch_quantize_index(y, params, positions, n_pos, B, C, HW, c->d_scale.as<float>(), (int)c->h_scale.size(), symbols,
                                 indexes, yhat_buf, c->sm_count, s);
}

int basic_gauss_dequantize(basic_coder *c, const int32_t *symbols, const float *params, const int32_t *positions, int64_t n_pos,
                           int B, int C, int HW, float *yhat_buf, void *stream)
{
    if (!c) return value_error("null coder");
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!is_device_ptr(params) || !is_device_ptr(symbols) || !is_device_ptr(yhat_buf) || (positions && !is_device_ptr(positions)))
        return value_error("basic_gauss_dequantize works on device memory");
    return launch_dequantize(symbols, params, positions, n_pos, B, C, HW, yhat_buf, c->sm_count, s);
}

// ------------------------------------------------------------------------------------- context model
int basic_ctx_create(int C, int G, int kernel_size, int device, basic_ctx **out)
{
    if (!out) return value_error("null out pointer");
    if (C < 1 || G < 1 || C % G) return value_error("in_channels must be a positive multiple of channel_groups");
    if (kernel_size != 5 && kernel_size != 3 && kernel_size != 1) return value_error("kernel_size must be 1, 3 or 5");
    int ndev = 0;
    BASIC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device"); return BASIC_ERR_CUDA; }
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    BASIC_CUDA(cudaGetDeviceProperties(&prop, device));
    basic_ctx *m = new basic_ctx();
    m->m = ctx_new(C, G, kernel_size, device, prop.multiProcessorCount);
    m->device = device;
    *out = m;
    return BASIC_OK;
}

void basic_ctx_destroy(basic_ctx *m)
{
    if (!m) return;
    DeviceGuard guard(m->device);
    ctx_delete(m->m);
    delete m;
}

int basic_ctx_set_weights(basic_ctx *m, const float *ctx_w, const float *ctx_b, const float *m1_w, const float *m1_b, const float *m2_w,
                          const float *m2_b, const float *m3_w, const float *m3_b)
{
    if (!m) return value_error("null model");
    DeviceGuard guard(m->device);
    return ctx_set_weights(*m->m, ctx_w, ctx_b, m1_w, m1_b, m2_w, m2_b, m3_w, m3_b);
}

int basic_ctx_set_weights_internal(basic_ctx *m, const float *ctx_w, const float *ctx_b, const float *m1_w, const float *m1_b,
                                   const float *m2_w, const float *m2_b, const float *m3_w, const float *m3_b, const float *p1_w,
                                   const float *p1_b, const float *p2_w, const float *p2_b, int half)
{
    if (!m) return value_error("null model");
    DeviceGuard guard(m->device);
    return ctx_set_weights_internal(*m->m, ctx_w, ctx_b, m1_w, m1_b, m2_w, m2_b, m3_w, m3_b, p1_w, p1_b, p2_w, p2_b, half);
}

int basic_ctx_set_map(basic_ctx *m, const int32_t *tg, int H, int W)
{
    if (!m) return value_error("null model");
    if (H < 1 || W < 1) return value_error("empty map");
    DeviceGuard guard(m->device);
    return ctx_set_map(*m->m, tg, H, W);
}

int basic_ctx_num_stages(basic_ctx *m) { return m ? ctx_num_stages(*m->m) : 0; }

int basic_ctx_set_precision(basic_ctx *m, int precision, int nacc)
{
    if (!m) return value_error("null model");
    return ctx_set_precision(*m->m, precision, nacc);
}

int basic_ctx_stage_positions(basic_ctx *m, int g, const int32_t **positions_dev, int64_t *n_pos)
{
    if (!m) return value_error("null model");
    return ctx_stage_positions(*m->m, g, positions_dev, n_pos);
}

int basic_ctx_stage_params(basic_ctx *m, int g, const float *buf, const float *prior, int B, float *params, void *stream)
{
    if (!m) return value_error("null model");
    DeviceGuard guard(m->device);
    if (!is_device_ptr(buf) || !is_device_ptr(prior) || !is_device_ptr(params)) return value_error("basic_ctx_stage_params works on device memory");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int prec = ctx_precision(*m->m);
    ctx_set_run_precision(*m->m, prec);
    if (prec != BASIC_CTX_FP16X3) return ctx_stage_params(*m->m, g, buf, prior, B, params, s);
    // 3xFP16 needs |activation| < 4000: the kernels raise a flag otherwise and the stage is repeated in 3xTF32
    BASIC_TRY(ctx_range_flag_clear(*m->m, s));
    BASIC_TRY(ctx_stage_params(*m->m, g, buf, prior, B, params, s));
    int flag = 0;
    BASIC_TRY(ctx_range_flag_read(*m->m, s, &flag));
    if (!flag) return BASIC_OK;
    ctx_set_run_precision(*m->m, BASIC_CTX_TF32X3);
    const int rc = ctx_stage_params(*m->m, g, buf, prior, B, params, s);
    ctx_set_run_precision(*m->m, prec);
    return rc;
}

// -------------------------------------------------------------------------------------- whole y path
int64_t basic_ypath_encode_bound(basic_coder *c, int B, int C, int H, int W, int lanes)
{
    const int64_t n = (int64_t)B * C * H * W;
    return basic_coder_encode_bound(c, n, lanes) + 64 + 4 * (int64_t)C * H * W;  // + one chunk_syms word per coding group
}

static int ypath_setup(basic_coder *c, basic_ctx *model, int B, int C, int H, int W, int *S)
{
    BASIC_TRY(need_init(c));
    if (c->kind != BASIC_KIND_RANS64) return value_error("the y path codes with rANS");
    if (c->h_scale.empty()) return value_error("scale table not set");
    if (B < 1 || C < 1 || H < 1 || W < 1) return value_error("empty input");
    *S = 1;
    if (model) {
        int mc, mg, mh, mw;
        ctx_dims(*model->m, &mc, &mg, &mh, &mw);
        if (mc != C) return value_error("context model channel count does not match the input");
        if (mh != H || mw != W) return value_error("group map not set for this spatial size (basic_ctx_set_map)");
        if (model->device != c->device) return value_error("coder and context model live on different devices");
        *S = ctx_num_stages(*model->m);
    }
    const size_t n = (size_t)B * C * H * W;
    BASIC_TRY(c->buf.reserve(n * 4));
    BASIC_TRY(c->params.reserve(std::max(n * 8, ctx_cl_elems(B, 2 * C, H * W) * 4)));
    BASIC_TRY(c->sym_all.reserve(n * 4 + 16));
    BASIC_TRY(c->idx_all.reserve(n * 4 + 16));
    return BASIC_OK;
}

int basic_ypath_encode(basic_coder *c, basic_ctx *model, const float *y, const float *prior, int B, int C, int H, int W, int lanes,
                       uint8_t *out, int64_t out_cap, int64_t *out_len, float *yhat_out, void *stream)
{
    int S;
    BASIC_TRY(ypath_setup(c, model, B, C, H, W, &S));
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int HW = H * W;
    const size_t n = (size_t)B * C * HW;
    const float *d_y, *d_prior;
    bool y_pending = false;
    g_trace.mark("enter", s);
    // tensor-core context model: it reads channels-last copies of the prior (made once) and of the y_hat buffer
    // (refreshed after every group's write-back)
    if (model) ctx_set_run_precision(*model->m, ctx_precision(*model->m));
    // few-row stages (scanline-like maps) take the persistent stage kernel -- exact FP32, whatever precision is configured
    const bool scan = model && ctx_scan_supported(*model->m, B);
    const bool tc = model && !scan && ctx_uses_tc(*model->m, B);
    // host inputs of a large batch arrive in image sub-batches; the groups of a sub-batch run while the next one is on the bus
    SubPlan subs = plan_subs(B, (y && !is_device_ptr(y)) || (prior && !is_device_ptr(prior)), model && !scan);
    BASIC_TRY(upload_inputs(c, y, n, prior, 2 * n, s, &d_y, &d_prior, &y_pending, &subs, B));
    g_trace.mark("uploaded", s);
    float *buf = c->buf.as<float>(), *params = c->params.as<float>();
    int32_t *sym = c->sym_all.as<int32_t>(), *idx = c->idx_all.as<int32_t>();
    const float *params_src = params;
    float *buf_cl = nullptr, *prior_cl = nullptr;
    if (tc) {
        BASIC_TRY(c->buf_cl.reserve(ctx_cl_elems(B, C, HW) * 4));
        BASIC_TRY(c->prior_cl.reserve(ctx_cl_elems(B, 2 * C, HW) * 4));
        buf_cl = c->buf_cl.as<float>();
        prior_cl = c->prior_cl.as<float>();
    }
    // the encoder knows y: all groups' symbols and indexes first (group g's context = the reconstructions of groups
    // < g, written back by the quantiser), then ONE coding pass over all of them
    std::vector<int64_t> slice_n;
    size_t done = 0;
    auto run_groups = [&]() -> int {
        slice_n.clear();
        done = 0;
        BASIC_CUDA(cudaMemsetAsync(buf, 0, n * 4, s));
        if (tc) {  // (in the operand format of the current mode: redone by the 3xTF32 fallback)
            BASIC_CUDA(cudaMemsetAsync(buf_cl, 0, ctx_cl_elems(B, C, HW) * 4, s));
            if (subs.n == 1) BASIC_TRY(ctx_to_cl(*model->m, d_prior, prior_cl, B, 2 * C, s));
        }
        if (scan) {
            // many-stage map (scanline, the serial JointAR coder): ONE persistent launch walks every stage -- context model,
            // scale indexes, quantisation and the y_hat write-back (ctx_scan.cu k_scan_stages)
            if (y_pending) {
                BASIC_CUDA(cudaStreamWaitEvent(s, c->ev_y, 0));
                y_pending = false;
            }
            for (int g = 0; g < S; ++g) {
                const int32_t *pos = nullptr;
                int64_t n_pos = 0;
                BASIC_TRY(ctx_stage_positions(*model->m, g, &pos, &n_pos));
                slice_n.push_back((int64_t)B * n_pos);
                done += (size_t)B * n_pos;
            }
            ProfScope ps(PROF_CTX, s);
            return ctx_scan_run(*model->m, 0, S, buf, d_prior, B, params, d_y, sym, idx, c->d_scale.as<float>(), (int)c->h_scale.size(), s);
        }
        if (subs.n > 1) {
            // sub-batch by sub-batch (images are independent; the stream order -- group-major, image-major inside a group -- is
            // kept by where the symbols are written)
            const size_t img = (size_t)C * HW, pimg = tc ? ctx_cl_elems(1, 2 * C, HW) : 2 * img;
            const size_t cl_y = tc ? ctx_cl_elems(1, C, HW) : 0, cl_p = tc ? ctx_cl_elems(1, 2 * C, HW) : 0;
            std::vector<int64_t> npos((size_t)S);
            std::vector<const int32_t *> posv((size_t)S);
            for (int g = 0; g < S; ++g) {
                BASIC_TRY(ctx_stage_positions(*model->m, g, &posv[(size_t)g], &npos[(size_t)g]));
                slice_n.push_back((int64_t)B * npos[(size_t)g]);
            }
            for (int i = 0; i < subs.n; ++i) {
                const int b0 = subs.b0[i], nb = subs.b0[i + 1] - b0;
                if (nb <= 0) continue;
                if (subs.host) BASIC_CUDA(cudaStreamWaitEvent(s, c->ev_sub[i], 0));
                if (tc) BASIC_TRY(ctx_to_cl(*model->m, d_prior + b0 * 2 * img, prior_cl + b0 * cl_p, nb, 2 * C, s));
                size_t at = 0;
                for (int g = 0; g < S; ++g) {
                    {
                        ProfScope ps(PROF_CTX, s);
                        BASIC_TRY(ctx_stage_params(*model->m, g, buf + b0 * img, d_prior + b0 * 2 * img, nb, params + b0 * pimg, s,
                                                   tc ? buf_cl + b0 * cl_y : nullptr, tc ? prior_cl + b0 * cl_p : nullptr, tc));
                    }
                    const int64_t n_pos = npos[(size_t)g];
                    if (n_pos > 0) {
                        ProfScope ps(PROF_GAUSS, s);
                        const size_t o = at + (size_t)b0 * n_pos;
                        BASIC_TRY(launch_quantize_index(d_y + b0 * img, params_src + b0 * pimg, posv[(size_t)g], n_pos, nb, C, HW, c->d_scale.as<float>(),
                                                        (int)c->h_scale.size(), sym + o, idx + o, buf + b0 * img, c->sm_count, s, tc ? 1 : 0,
                                                        tc ? ctx_perm(*model->m) : nullptr));
                        if (tc && g + 1 < S) BASIC_TRY(ctx_to_cl(*model->m, buf + b0 * img, buf_cl + b0 * cl_y, nb, C, s));
                    }
                    at += (size_t)B * n_pos;
                }
                done = at;
            }
            return BASIC_OK;
        }
        for (int g = 0; g < S; ++g) {
            const int32_t *pos = nullptr;
            int64_t n_pos = (int64_t)C * HW;
            if (model) {
                ProfScope ps(PROF_CTX, s);
                BASIC_TRY(ctx_stage_params(*model->m, g, buf, d_prior, B, params, s, buf_cl, prior_cl, tc));
                BASIC_TRY(ctx_stage_positions(*model->m, g, &pos, &n_pos));
            } else {
                params_src = d_prior;
            }
            const int64_t cnt = (int64_t)B * n_pos;
            slice_n.push_back(cnt);
            if (cnt == 0) continue;
            if (y_pending) {  // y was uploaded beside the first group's context model
                BASIC_CUDA(cudaStreamWaitEvent(s, c->ev_y, 0));
                y_pending = false;
            }
            ProfScope ps(PROF_GAUSS, s);
            BASIC_TRY(launch_quantize_index(d_y, params_src, pos, n_pos, B, C, HW, c->d_scale.as<float>(), (int)c->h_scale.size(),
                                            sym + done, idx + done, buf, c->sm_count, s, tc ? 1 : 0, tc ? ctx_perm(*model->m) : nullptr));
            if (tc && g + 1 < S) BASIC_TRY(ctx_to_cl(*model->m, buf, buf_cl, B, C, s));
            done += (size_t)cnt;
        }
        return BASIC_OK;
    };
    // 3xFP16 context model: valid while every activation stays below 4000 in magnitude; the kernels raise a flag
    // otherwise and the pass is repeated in 3xTF32.  The container's magic tells the decoder which mode was used.
    bool fp16 = tc && ctx_precision(*model->m) == BASIC_CTX_FP16X3;
    if (fp16) BASIC_TRY(ctx_range_flag_clear(*model->m, s));
    BASIC_TRY(run_groups());
    g_trace.mark("groups queued", s);
    auto fallback_tf32 = [&]() -> int {
        fp16 = false;
        ctx_set_run_precision(*model->m, BASIC_CTX_TF32X3);
        const int rc = run_groups();
        ctx_set_run_precision(*model->m, ctx_precision(*model->m));
        return rc;
    };
    // the flag is read with the coder's own first synchronisation (multi-lane: the size estimate) instead of a separate one;
    // a raised flag -- rare -- then costs the coding pass that was already queued
    int *h_flag = reinterpret_cast<int *>(static_cast<char *>(c->pinned) + 128);
    bool flag_pending = false;
    if (fp16) {
        if (lanes == BASIC_LANES_REFERENCE) {
            int flag = 0;
            BASIC_TRY(ctx_range_flag_read(*model->m, s, &flag));
            if (flag)
                return value_error("context-model activation outside the 3xFP16 range and a lanes=1 stream cannot record the "
                                   "fallback: use ctx_precision tf32x3 or fp32");
        } else {
            BASIC_TRY(ctx_range_flag_copy(*model->m, s, h_flag));
            flag_pending = true;
        }
    }
    if (yhat_out) BASIC_CUDA(cudaMemcpyAsync(yhat_out, buf, n * 4, cudaMemcpyDefault, s));
    if (lanes == BASIC_LANES_REFERENCE) {
        const uint8_t *d_bytes;
        int64_t len;
        BASIC_TRY(encode_compat(c, sym, idx, (int64_t)done, s, &d_bytes, &len));
        BASIC_TRY(copy_out(c, d_bytes, len, out, out_cap, s));
        if (out_len) *out_len = len;
        return BASIC_OK;
    }
    // multi-lane container: magic | one segment with one slice per coding group (lane states carried across groups)
    BASIC_TRY(c->segs.reserve(64));
    int64_t seg_len = 0;
    for (int pass = 0; pass < 2; ++pass) {
        {
            ProfScope ps(PROF_ENCODE, s);
            BASIC_TRY(encode_segment(c, sym, idx, S, slice_n.data(), lanes, 4, s, &seg_len));  // synchronises `s`
        }
        if (!(flag_pending && *h_flag)) break;
        flag_pending = false;  // outside the 3xFP16 range: everything again in 3xTF32, container magic "BLS1"
        BASIC_TRY(fallback_tf32());
        if (yhat_out) BASIC_CUDA(cudaMemcpyAsync(yhat_out, buf, n * 4, cudaMemcpyDefault, s));
    }
    g_trace.mark("segment coded", s);
    // the magic records the arithmetic of the context model: the decoder follows it (a stream written with one mode and read
    // with another would differ in the last bits of the parameters -- garbage symbols at rounding ties, silently)
    const uint32_t *magic = fp16 ? &kMagic2 : tc ? &kMagic : model ? &kMagic0 : &kMagic;
    BASIC_CUDA(cudaMemcpyAsync(c->segs.p, magic, 4, cudaMemcpyHostToDevice, s));
    BASIC_TRY(copy_out(c, c->segs.p, 4 + seg_len, out, out_cap, s));
    if (out_len) *out_len = 4 + seg_len;
    g_trace.mark("copy queued", s);
    g_trace.dump("encode", s);
    return BASIC_OK;
}

int basic_ypath_decode(basic_coder *c, basic_ctx *model, const uint8_t *encoded, int64_t len, const float *prior, int B, int C, int H,
                       int W, int lanes, float *yhat_out, void *stream)
{
    int S;
    BASIC_TRY(ypath_setup(c, model, B, C, H, W, &S));
    DeviceGuard guard(c->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int HW = H * W;
    const size_t n = (size_t)B * C * HW;
    const float *d_prior, *d_none;
    bool none_pending = false;
    g_trace.mark("enter", s);
    // the reconstruction is built in place when the caller's output lives on the device
    const bool direct = is_device_ptr(yhat_out) && (reinterpret_cast<uintptr_t>(yhat_out) & 255) == 0;
    float *buf = direct ? yhat_out : c->buf.as<float>(), *params = c->params.as<float>();
    int32_t *sym = c->sym_all.as<int32_t>(), *idx = c->idx_all.as<int32_t>();
    BASIC_CUDA(cudaMemsetAsync(buf, 0, n * 4, s));
    const float *params_src = params;
    Small *ds = c->small.as<Small>(), *hs = reinterpret_cast<Small *>(c->pinned);
    BASIC_CUDA(cudaMemsetAsync(&ds->status, 0, sizeof(int), s));
    // the context model runs in the mode the stream was written with: "BLS2" = 3xFP16, otherwise the configured mode
    // (a 3xFP16 configuration reads "BLS1" -- the encoder's range fallback -- in 3xTF32)
    if (model) {
        int prec = ctx_precision(*model->m);
        if (lanes != BASIC_LANES_REFERENCE) {
            uint32_t magic = 0;
            if (len >= 4) {
                if (is_device_ptr(encoded)) BASIC_CUDA(cudaMemcpy(&magic, encoded, 4, cudaMemcpyDeviceToHost));
                else memcpy(&magic, encoded, 4);
            }
            if (magic == kMagic2) prec = BASIC_CTX_FP16X3;
            else if (magic == kMagic0) prec = BASIC_CTX_FP32;
            else prec = BASIC_CTX_TF32X3;   // "BLS1" out of the y path with a context model: 3xTF32 (also the 3xFP16 range fallback)
        }
        ctx_set_run_precision(*model->m, prec);
    }
    // many-stage map: one fused launch per stage (ctx_scan.cu k_scan_stages) -- what the encoder used when the stream says exact FP32
    const bool scan = model && ctx_scan_supported(*model->m, B) &&
                      (lanes == BASIC_LANES_REFERENCE || ctx_run_precision(*model->m) == BASIC_CTX_FP32);
    const bool tc = model && !scan && ctx_uses_tc(*model->m, B);
    // a host prior of a large batch arrives in image sub-batches: the first group's context model of a sub-batch runs while
    // the next one is on the bus
    // (one channel group only: with several, later stages read activations an earlier stage cached per image of the WHOLE batch)
    int m_c = 0, m_g = 1, m_h = 0, m_w = 0;
    if (model) ctx_dims(*model->m, &m_c, &m_g, &m_h, &m_w);
    SubPlan subs = plan_subs(B, prior && !is_device_ptr(prior), model && !scan && m_g == 1);
    BASIC_TRY(upload_inputs(c, nullptr, 0, prior, 2 * n, s, &d_none, &d_prior, &none_pending, &subs, B));
    float *buf_cl = nullptr, *prior_cl = nullptr;
    if (tc) {  // channels-last views for the tensor-core context model (see basic_ypath_encode)
        BASIC_TRY(c->buf_cl.reserve(ctx_cl_elems(B, C, HW) * 4));
        BASIC_TRY(c->prior_cl.reserve(ctx_cl_elems(B, 2 * C, HW) * 4));
        buf_cl = c->buf_cl.as<float>();
        prior_cl = c->prior_cl.as<float>();
        BASIC_CUDA(cudaMemsetAsync(buf_cl, 0, ctx_cl_elems(B, C, HW) * 4, s));
        if (subs.n == 1) BASIC_TRY(ctx_to_cl(*model->m, d_prior, prior_cl, B, 2 * C, s));
    }
    // the first group's parameters do not depend on the stream: queue them, then stage the stream into pinned
    // memory and upload it while the GPU is busy
    // (a multi-lane stream on the stage kernel may be decoded inside ONE launch: decided once the segment is parsed)
    bool g0_done = false;
    if (model && subs.n > 1) {
        const size_t img = (size_t)C * HW, pimg = tc ? ctx_cl_elems(1, 2 * C, HW) : 2 * img;
        const size_t cl_y = tc ? ctx_cl_elems(1, C, HW) : 0, cl_p = tc ? ctx_cl_elems(1, 2 * C, HW) : 0;
        ProfScope ps(PROF_CTX, s);
        for (int i = 0; i < subs.n; ++i) {
            const int b0 = subs.b0[i], nb = subs.b0[i + 1] - b0;
            if (nb <= 0) continue;
            if (subs.host) BASIC_CUDA(cudaStreamWaitEvent(s, c->ev_sub[i], 0));
            if (tc) BASIC_TRY(ctx_to_cl(*model->m, d_prior + b0 * 2 * img, prior_cl + b0 * cl_p, nb, 2 * C, s));
            BASIC_TRY(ctx_stage_params(*model->m, 0, buf + b0 * img, d_prior + b0 * 2 * img, nb, params + b0 * pimg, s,
                                       tc ? buf_cl + b0 * cl_y : nullptr, tc ? prior_cl + b0 * cl_p : nullptr, tc));
        }
        g0_done = true;
    } else if (model && !(scan && lanes != BASIC_LANES_REFERENCE)) {
        ProfScope ps(PROF_CTX, s);
        if (scan) BASIC_TRY(ctx_scan_run(*model->m, 0, 1, buf, d_prior, B, params, nullptr, nullptr, idx, c->d_scale.as<float>(), (int)c->h_scale.size(), s));
        else BASIC_TRY(ctx_stage_params(*model->m, 0, buf, d_prior, B, params, s, buf_cl, prior_cl, tc));
        g0_done = true;
    }
    g_trace.mark("g0 queued", s);
    BASIC_TRY(basic_coder_set_stream(c, encoded, len, lanes, stream));
    g_trace.mark("stream set", s);
    // per-group symbol counts = the slices of the segment
    std::vector<int64_t> slice_n((size_t)S);
    for (int g = 0; g < S; ++g) {
        const int32_t *pos = nullptr;
        int64_t n_pos = (int64_t)C * HW;
        if (model) BASIC_TRY(ctx_stage_positions(*model->m, g, &pos, &n_pos));
        slice_n[g] = (int64_t)B * n_pos;
    }
    SegInfo si;
    if (lanes != BASIC_LANES_REFERENCE) {
        BASIC_TRY(parse_segment(c->host_src + c->stream_pos, c->stream_len - c->stream_pos, S, slice_n.data(), &si));
        BASIC_TRY(c->carry_x.reserve((size_t)si.n_chunks * 128 + 16));
        BASIC_TRY(c->carry_wp.reserve((size_t)si.n_chunks * 4 + 16));
    }
    bool fused = false;
    if (scan && lanes != BASIC_LANES_REFERENCE && ctx_scan_decode_supported(*model->m, B, si.n_chunks, (int)c->bypass_precision, c->rt.precision)) {
        // the whole decode in one launch: context model, scale indexes, the coder's chunk warps and the write-back (ctx_scan.cu)
        ProfScope ps(PROF_CTX, s);
        ScanDecodeHost dh = {&c->rt, c->bypass, c->stream_dev.as<unsigned char>() + c->stream_pos, si.len, si.n_chunks, si.cs.data(), &ds->status};
        BASIC_TRY(ctx_scan_run(*model->m, 0, S, buf, d_prior, B, params, nullptr, nullptr, idx, c->d_scale.as<float>(), (int)c->h_scale.size(), s,
                               nullptr, &dh));
        fused = true;
    }
    for (int g = 0; g < S && !fused; ++g) {
        const int32_t *pos = nullptr;
        int64_t n_pos = (int64_t)C * HW;
        if (model) {
            if (g > 0 || !g0_done) {
                ProfScope ps(PROF_CTX, s);
                // (the fused stage launch starts by turning the previous stage's symbols into y_hat)
                if (scan) BASIC_TRY(ctx_scan_run(*model->m, g, g + 1, buf, d_prior, B, params, nullptr, nullptr, idx, c->d_scale.as<float>(), (int)c->h_scale.size(), s,
                                                 g > 0 && slice_n[g - 1] > 0 ? sym : nullptr));
                else BASIC_TRY(ctx_stage_params(*model->m, g, buf, d_prior, B, params, s, buf_cl, prior_cl, tc));
            }
            BASIC_TRY(ctx_stage_positions(*model->m, g, &pos, &n_pos));
        } else {
            params_src = d_prior;
        }
        const int64_t cnt = slice_n[g];
        if (cnt > 0 && !scan) {   // (the fused stage launch has written the indexes already)
            ProfScope ps(PROF_GAUSS, s);
            BASIC_TRY(launch_quantize_index(nullptr, params_src, pos, n_pos, B, C, HW, c->d_scale.as<float>(), (int)c->h_scale.size(),
                                            nullptr, idx, nullptr, c->sm_count, s, tc ? 1 : 0, tc ? ctx_perm(*model->m) : nullptr));
        }
        if (lanes == BASIC_LANES_REFERENCE) {
            if (cnt == 0) continue;
            const int init = c->stream_pos < 0;
            BASIC_TRY(launch_rans64_decode(c->rt, c->stream_dev.as<uint32_t>(), c->stream_len / 4, &ds->len[2], init, idx,
                                           cnt, c->bypass, (int)c->bypass_precision, sym, &ds->status, s));
            c->stream_pos = 0;
        } else {
            // every slice is launched, empty ones too: the lane states must travel from slice to slice
            ProfScope ps(PROF_DECODE, s);
            BASIC_TRY(launch_bls_decode(c->rt, c->bypass, (int)c->bypass_precision, c->stream_dev.as<unsigned char>() + c->stream_pos,
                                        si.len, idx, cnt, si.cs[g], si.n_chunks, S, g, c->carry_x.as<uint32_t>(),
                                        c->carry_wp.as<uint32_t>(), sym, &ds->status, c->sm_count, s));
        }
        if (cnt > 0 && !(scan && g + 1 < S)) {
            ProfScope ps(PROF_GAUSS, s);
            BASIC_TRY(launch_dequantize(sym, params_src, pos, n_pos, B, C, HW, buf, c->sm_count, s, tc ? 1 : 0,
                                        tc ? ctx_perm(*model->m) : nullptr));
            if (tc && g + 1 < S) BASIC_TRY(ctx_to_cl(*model->m, buf, buf_cl, B, C, s));
        }
    }
    if (lanes != BASIC_LANES_REFERENCE) c->stream_pos += si.len;
    g_trace.mark("all queued", s);
    BASIC_CUDA(cudaMemcpyAsync(&hs->status, &ds->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!direct) BASIC_CUDA(cudaMemcpyAsync(yhat_out, buf, n * 4, cudaMemcpyDefault, s));
    g_trace.mark("yhat copied", s);
    BASIC_CUDA(cudaStreamSynchronize(s));
    g_trace.mark("synced", (cudaStream_t)-1);
    g_trace.dump("decode", s);
    if (model) ctx_set_run_precision(*model->m, ctx_precision(*model->m));
    return status_error(hs->status);
}

}  // extern "C"

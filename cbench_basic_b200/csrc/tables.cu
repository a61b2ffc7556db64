// Quantised-CDF table construction on the device.
//   k_build_cdfs : Rans64Base::init_params + pmf_to_quantized_cdf (reference cbench/csrc/ans/rans64.cpp:69-159),
//                  bit-exact, one CTA per table (the normalise / "steal" loop is serial per table).
//   k_pack_tables: packs the int32 CDFs into the shared-memory friendly image used by the coder kernels
//                  (u16 CDF entries + per-table bucket LUT for O(1) symbol lookup instead of the
//                  reference's linear scan, rans64.cpp:456-460).
#include "common.cuh"

namespace basic {

namespace {

constexpr int kBuildThreads = 256;

// Block-wide min of a 64-bit key; result broadcast to all threads.
__device__ unsigned long long block_min_u64(unsigned long long v, unsigned long long *sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    unsigned long long r = sh[0];
    for (int w = 1; w < kBuildThreads / 32; ++w) r = sh[w] < r ? sh[w] : r;
    return r;
}

__device__ long long block_sum_i64(long long v, long long *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    long long r = 0;
    for (int w = 0; w < kBuildThreads / 32; ++w) r += sh[w];
    return r;
}

// The serial part of pmf_to_quantized_cdf on an int32 array `cdf` of n+1 entries living in shared memory,
// whose entries 1..n already hold round(pmf * 2^precision) (rans64.cpp:78-123).
__device__ void quantize_cdf_in_smem(int32_t *cdf, int n, int precision, int *err, void *red)
{
    const int tid = threadIdx.x;
    long long part = 0;
    for (int i = tid; i <= n; i += kBuildThreads) part += cdf[i];
    // std::accumulate(int) then conversion to uint32_t
    const uint32_t total = (uint32_t)(int32_t)block_sum_i64(part, (long long *)red);
    if (total == 0) {
        if (tid == 0) *err = 1;
        return;
    }
    for (int i = tid; i <= n; i += kBuildThreads)
        cdf[i] = (int32_t)(((uint64_t)(1 << precision) * (uint64_t)(int64_t)cdf[i]) / total);
    __syncthreads();
    if (tid == 0) {  // std::partial_sum: serial, a few thousand shared-memory adds, once per model load
        int32_t run = cdf[0];
        for (int i = 1; i <= n; ++i) {
            run += cdf[i];
            cdf[i] = run;
        }
        cdf[n] = 1 << precision;
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;  // block-uniform (shared memory, synchronised below)
        // steal from the lowest-frequency symbol with freq > 1 (first such index on ties)
        unsigned long long best = ~0ull;
        for (int j = tid; j < n; j += kBuildThreads) {
            const uint32_t f = (uint32_t)(cdf[j + 1] - cdf[j]);
            if (f > 1) {
                const unsigned long long key = ((unsigned long long)f << 32) | (uint32_t)j;
                best = key < best ? key : best;
            }
        }
        best = block_min_u64(best, (unsigned long long *)red);
        if (best == ~0ull) {
            if (tid == 0) *err = 1;
            return;
        }
        const int steal = (int)(uint32_t)best;
        if (steal < i) {
            for (int j = steal + 1 + tid; j <= i; j += kBuildThreads) cdf[j]--;
        } else {
            for (int j = i + 1 + tid; j <= steal; j += kBuildThreads) cdf[j]++;
        }
        __syncthreads();
    }
}

// One CTA per table.  dynamic smem: int32 [M + 2].
__global__ void __launch_bounds__(kBuildThreads)
k_build_cdfs(const int32_t *__restrict__ freqs, int M, const int32_t *__restrict__ nsym, int precision,
             int32_t *__restrict__ cdfs, int stride, int *err)
{
    extern __shared__ int32_t sm[];
    __shared__ unsigned long long red[kBuildThreads / 32];
    __shared__ float s_total;
    const int t = blockIdx.x, tid = threadIdx.x, n = nsym[t];
    const int32_t *f = freqs + (size_t)t * M;
    for (int i = tid; i < n; i += kBuildThreads) sm[i + 1] = f[i];
    __syncthreads();
    if (tid == 0) {  // std::accumulate(freq.begin(), freq.end(), 0.0f): float32, strictly left to right
        float s = 0.0f;
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, (float)sm[i + 1]);
        s_total = __fadd_rn(s, 1.0f);  // + tail_mass
    }
    __syncthreads();
    const float total = s_total, scale = (float)(1 << precision);
    for (int i = tid; i < n; i += kBuildThreads) {
        const float p = __fdiv_rn((float)sm[i + 1], total);
        sm[i + 1] = (int32_t)roundf(__fmul_rn(p, scale));
    }
    if (tid == 0) {
        sm[0] = 0;
        sm[n + 1] = (int32_t)roundf(__fmul_rn(__fdiv_rn(1.0f, total), scale));  // the tail / escape symbol
    }
    __syncthreads();
    quantize_cdf_in_smem(sm, n + 1, precision, err, red);
    __syncthreads();
    int32_t *row = cdfs + (size_t)t * stride;
    for (int i = tid; i < stride; i += kBuildThreads) row[i] = i <= n + 1 ? sm[i] : 0;
}

// pmf_to_quantized_cdf as a module-level helper: one table, pmf given as float.
__global__ void __launch_bounds__(kBuildThreads)
k_pmf_to_cdf(const float *__restrict__ pmf, int n, int precision, int32_t *__restrict__ cdf, int *err)
{
    extern __shared__ int32_t sm[];
    __shared__ unsigned long long red[kBuildThreads / 32];
    const int tid = threadIdx.x;
    const float scale = (float)(1 << precision);
    for (int i = tid; i < n; i += kBuildThreads) sm[i + 1] = (int32_t)roundf(__fmul_rn(pmf[i], scale));
    if (tid == 0) sm[0] = 0;
    __syncthreads();
    quantize_cdf_in_smem(sm, n, precision, err, red);
    __syncthreads();
    for (int i = tid; i <= n; i += kBuildThreads) cdf[i] = sm[i];
}

// One CTA per table: u16 CDF + bucket LUT + metadata.
__global__ void __launch_bounds__(256)
k_pack_tables(const int32_t *__restrict__ cdfs, int stride, const TableMeta *__restrict__ meta_in,
              TableMeta *__restrict__ meta_out, uint16_t *__restrict__ cdf16, uint16_t *__restrict__ lut, int precision,
              uint4 *__restrict__ enc, uint16_t *__restrict__ dcdf, uint16_t *__restrict__ lut2)
{
    const int t = blockIdx.x;
    const TableMeta m = meta_in[t];
    const int32_t *row = cdfs + (size_t)t * stride;
    if (threadIdx.x == 0) meta_out[t] = m;
    // the pair decoder's image: d[i] = cdf[i] - 1 (mod 2^16: d[0] = 0xffff, the table's total 2^16 -> 0xffff), four 0xffff
    // sentinels behind every table: "cdf[i] <= cum" is "d[i] < cum" and is false for every sentinel, so the search needs no
    // end-of-table guards; start = (d + 1) & 0xffff, freq = (d[i + 1] - d[i]) & 0xffff
    const uint32_t dbase = m.cdf_base + 4u * (uint32_t)t;
    if (dcdf) {
        for (int i = threadIdx.x; i < m.cdf_size + 4; i += blockDim.x)
            dcdf[dbase + i] = i < m.cdf_size ? (uint16_t)(row[i] - 1) : (uint16_t)0xffffu;
    }
    for (int i = threadIdx.x; i < m.cdf_size; i += blockDim.x) {
        cdf16[m.cdf_base + i] = (uint16_t)row[i];
        // the multi-lane encoder's operands of symbol i (rans_pair.cu): the state update x -> (x / f << 16) + x % f + start as
        // two integer multiplies with m = ceil(2^32 / f); f = 1: m = 2^32 - 1 yields x - 1 with remainder 1, corrected by start'
        uint32_t start = (uint32_t)row[i], f = i + 1 < m.cdf_size ? (uint32_t)(row[i + 1] - row[i]) : 1u;
        if (f == 0 || f > 65536u) f = 1;
        enc[m.cdf_base + i] = make_uint4(f == 1 ? start + 65535u : start, f == 1 ? 0xffffffffu : 0xffffffffu / f + 1u, f, 0u);
    }
    const int nb = 1 << (precision - m.lut_shift);
    const int nsyms = m.cdf_size - 1;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int32_t v = b << m.lut_shift;
        int lo = 0, hi = nsyms - 1;  // largest s with row[s] <= v
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (row[mid] <= v) lo = mid; else hi = mid - 1;
        }
        lut[m.lut_base + b] = (uint16_t)lo;
        if (lut2) lut2[m.lut_base + b] = (uint16_t)(2u * (dbase + (uint32_t)lo));
    }
}

int ceil_log2(int v)
{
    int b = 0;
    while ((1 << b) < v) ++b;
    return b;
}

}  // namespace

// Lays out and fills RansTables::blob from cdf32 (already on the device) and the host-known sizes/offsets.
int rans_tables_pack(RansTables &tb, cudaStream_t stream)
{
    const int T = tb.T;
    std::vector<TableMeta> meta(T);
    uint32_t cdf_total = 0, lut_total = 0;
    for (int t = 0; t < T; ++t) {
        const int size = tb.h_sizes[t];
        if (size < 2 || size > 65535) return value_error("cdf size out of range");
        int lb = ceil_log2(size - 1) + 1;
        const int hi = tb.precision < 12 ? tb.precision : 12;
        if (lb < 4) lb = 4;
        if (lb > hi) lb = hi;
        meta[t].cdf_base = cdf_total;
        meta[t].lut_base = lut_total;
        meta[t].cdf_size = (uint16_t)size;
        meta[t].lut_shift = (uint8_t)(tb.precision - lb);
        meta[t].pad = 0;
        meta[t].offset = tb.h_offsets[t];
        cdf_total += (uint32_t)size;
        lut_total += 1u << lb;
    }
    auto pad16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    tb.meta_bytes = pad16(sizeof(TableMeta) * T);
    tb.cdf16_bytes = pad16((size_t)cdf_total * 2);
    tb.lut_bytes = pad16((size_t)lut_total * 2);
    tb.blob_bytes = tb.meta_bytes + tb.cdf16_bytes + tb.lut_bytes;
    tb.total_cdf = cdf_total;
    tb.total_lut = lut_total;
    BASIC_TRY(tb.blob.reserve(tb.blob_bytes));
    BASIC_TRY(tb.enc.reserve((size_t)cdf_total * sizeof(uint4)));
    tb.dcdf_bytes = pad16((size_t)(cdf_total + 4u * (uint32_t)T) * 2 + 16);
    tb.blob_d_bytes = 2 * (size_t)(cdf_total + 4u * (uint32_t)T) < 65536 ? tb.dcdf_bytes + tb.lut_bytes : 0;   // offsets are u16
    if (tb.blob_d_bytes) {
        BASIC_TRY(tb.blob_d.reserve(tb.blob_d_bytes));
        BASIC_CUDA(cudaMemsetAsync(tb.blob_d.p, 0xff, tb.blob_d_bytes, stream));
    }
    BASIC_CUDA(cudaMemsetAsync(tb.blob.p, 0, tb.blob_bytes, stream));
    DevBuf tmp;
    BASIC_TRY(tmp.reserve(sizeof(TableMeta) * T));
    BASIC_CUDA(cudaMemcpyAsync(tmp.p, meta.data(), sizeof(TableMeta) * T, cudaMemcpyHostToDevice, stream));
    char *b = tb.blob.as<char>();
    k_pack_tables<<<T, 256, 0, stream>>>(tb.cdf32.as<int32_t>(), tb.stride, tmp.as<TableMeta>(),
                                         reinterpret_cast<TableMeta *>(b),
                                         reinterpret_cast<uint16_t *>(b + tb.meta_bytes),
                                         reinterpret_cast<uint16_t *>(b + tb.meta_bytes + tb.cdf16_bytes), tb.precision,
                                         tb.enc.as<uint4>(), tb.blob_d_bytes ? tb.blob_d.as<uint16_t>() : nullptr,
                                         tb.blob_d_bytes ? reinterpret_cast<uint16_t *>(tb.blob_d.as<char>() + tb.dcdf_bytes) : nullptr);
    BASIC_LAUNCHED();
    BASIC_CUDA(cudaStreamSynchronize(stream));
    tmp.release();
    tb.ready = true;
    return BASIC_OK;
}

int rans_tables_from_freqs(RansTables &tb, const int32_t *freqs, int T, int M, const int32_t *nsym, const int32_t *offsets,
                           int precision, cudaStream_t stream)
{
    if (T <= 0 || M <= 0) return value_error("freqs should be 2-dimensional with shape (num_symbols.size(), >num_symbols.max())");
    if (precision < 1 || precision > 16) return value_error("freq_precision must be in [1, 16]");
    int maxn = 0;
    for (int t = 0; t < T; ++t) {
        if (nsym[t] < 1 || nsym[t] > M) return value_error("num_symbols out of range of freqs");
        maxn = nsym[t] > maxn ? nsym[t] : maxn;
    }
    tb.ready = false;
    tb.T = T;
    tb.stride = maxn + 2;
    tb.precision = precision;
    tb.h_sizes.resize(T);
    tb.h_offsets.assign(offsets, offsets + T);
    for (int t = 0; t < T; ++t) tb.h_sizes[t] = nsym[t] + 2;
    DevBuf d_freqs, d_nsym, d_err;
    BASIC_TRY(d_freqs.reserve(sizeof(int32_t) * (size_t)T * M));
    BASIC_TRY(d_nsym.reserve(sizeof(int32_t) * T));
    BASIC_TRY(d_err.reserve(sizeof(int)));
    BASIC_TRY(tb.cdf32.reserve(sizeof(int32_t) * (size_t)T * tb.stride));
    BASIC_CUDA(cudaMemcpyAsync(d_freqs.p, freqs, sizeof(int32_t) * (size_t)T * M, cudaMemcpyHostToDevice, stream));
    BASIC_CUDA(cudaMemcpyAsync(d_nsym.p, nsym, sizeof(int32_t) * T, cudaMemcpyHostToDevice, stream));
    BASIC_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
    const size_t smem = sizeof(int32_t) * (size_t)(maxn + 2);
    if (smem > 48 * 1024) BASIC_CUDA(cudaFuncSetAttribute(k_build_cdfs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_build_cdfs<<<T, kBuildThreads, smem, stream>>>(d_freqs.as<int32_t>(), M, d_nsym.as<int32_t>(), precision,
                                                     tb.cdf32.as<int32_t>(), tb.stride, d_err.as<int>());
    BASIC_LAUNCHED();
    int err = 0;
    BASIC_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BASIC_CUDA(cudaStreamSynchronize(stream));
    d_freqs.release();
    d_nsym.release();
    d_err.release();
    if (err) return value_error("pmf_to_quantized_cdf: cannot normalise table (no frequency to steal)");
    return rans_tables_pack(tb, stream);
}

int rans_tables_from_cdfs(RansTables &tb, const int32_t *cdfs, int T, int M, const int32_t *sizes, const int32_t *offsets,
                          int precision, cudaStream_t stream)
{
    if (T <= 0 || M <= 0) return value_error("cdfs should be 2-dimensional with shape (cdfs_sizes.size(), >cdfs_sizes.max())");
    int maxs = 0;
    for (int t = 0; t < T; ++t) {
        if (sizes[t] < 2 || sizes[t] > M) return value_error("cdfs_sizes out of range of cdfs");
        maxs = sizes[t] > maxs ? sizes[t] : maxs;
    }
    tb.ready = false;
    tb.T = T;
    tb.stride = maxs;
    tb.precision = precision;
    tb.h_sizes.assign(sizes, sizes + T);
    tb.h_offsets.assign(offsets, offsets + T);
    std::vector<int32_t> packed((size_t)T * maxs, 0);
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < sizes[t]; ++i) packed[(size_t)t * maxs + i] = cdfs[(size_t)t * M + i];
    BASIC_TRY(tb.cdf32.reserve(packed.size() * sizeof(int32_t)));
    BASIC_CUDA(cudaMemcpyAsync(tb.cdf32.p, packed.data(), packed.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    BASIC_CUDA(cudaStreamSynchronize(stream));
    return rans_tables_pack(tb, stream);
}

int pmf_to_cdf_device(const float *pmf, int n, int precision, int32_t *cdf_out)
{
    if (n < 1) return value_error("empty pmf");
    // grow-only scratch kept per thread and device (the z node converts one pmf per channel: a cudaMalloc / cudaFree
    // triple per call made its update_state() take seconds)
    static thread_local DevBuf d_pmf, d_cdf, d_err;
    static thread_local int scratch_device = -1;
    int dev = 0;
    BASIC_CUDA(cudaGetDevice(&dev));
    if (dev != scratch_device) {
        d_pmf = DevBuf(); d_cdf = DevBuf(); d_err = DevBuf();  // (buffers of another device are left to that context)
        scratch_device = dev;
    }
    BASIC_TRY(d_pmf.reserve(sizeof(float) * n));
    BASIC_TRY(d_cdf.reserve(sizeof(int32_t) * (n + 1)));
    BASIC_TRY(d_err.reserve(sizeof(int)));
    BASIC_CUDA(cudaMemcpy(d_pmf.p, pmf, sizeof(float) * n, cudaMemcpyHostToDevice));
    BASIC_CUDA(cudaMemset(d_err.p, 0, sizeof(int)));
    const size_t smem = sizeof(int32_t) * (size_t)(n + 1);
    if (smem > 48 * 1024) BASIC_CUDA(cudaFuncSetAttribute(k_pmf_to_cdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pmf_to_cdf<<<1, kBuildThreads, smem>>>(d_pmf.as<float>(), n, precision, d_cdf.as<int32_t>(), d_err.as<int>());
    BASIC_LAUNCHED();
    int err = 0;
    BASIC_CUDA(cudaMemcpy(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    BASIC_CUDA(cudaMemcpy(cdf_out, d_cdf.p, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (err) return value_error("pmf_to_quantized_cdf: cannot normalise table");
    return BASIC_OK;
}

}  // namespace basic

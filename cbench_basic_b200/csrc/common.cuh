// Shared definitions of the sm_100a entropy-coding library (see include/basic_b200.h for the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/basic_b200.h"

namespace basic {

void set_error(const std::string &msg);
extern int64_t g_launches;

#define BASIC_CUDA(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            basic::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr); \
            return BASIC_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

#define BASIC_TRY(expr)          \
    do {                         \
        int _rc = (expr);        \
        if (_rc != BASIC_OK) return _rc; \
    } while (0)

#define BASIC_LAUNCHED()                         \
    do {                                         \
        ++basic::g_launches;                     \
        BASIC_CUDA(cudaGetLastError());          \
    } while (0)

inline int value_error(const std::string &msg)
{
    set_error(msg);
    return BASIC_ERR_VALUE;
}

// Grow-only device buffer (scratch that lives as long as the coder object: no cudaMalloc on the hot path).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return BASIC_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        BASIC_CUDA(cudaMalloc(&p, want));
        cap = want;
        return BASIC_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) are per device: one process may hold coders
// on several GPUs.  first() is true once per (call site, current device).
struct PerDeviceOnce {
    unsigned long long done = 0;  // bit d = device d has been set up (the library is called from one host thread at a time)
    bool first()
    {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};

inline bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Page-locked (cudaHostAlloc / cudaHostRegister) host memory: the DMA engines read it directly, no staging copy needed.
inline bool is_pinned_host_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Per-table metadata, one 16-byte record (a single LDS.128 / LDG.128 in the kernels).
struct __align__(16) TableMeta {
    uint32_t cdf_base;  // first entry of this table inside the packed u16 CDF array
    uint32_t lut_base;  // first entry of this table's bucket LUT
    uint16_t cdf_size;  // number of CDF entries = coded symbols + 1 (reference _cdfs_sizes)
    uint8_t lut_shift;  // bucket = cum >> lut_shift
    uint8_t pad;
    int32_t offset;     // reference _offsets[t]
};

// Device-resident rANS tables (built by tables.cu).
struct RansTables {
    int T = 0, stride = 0, precision = 16;
    std::vector<int32_t> h_sizes, h_offsets;  // host mirrors (sizes are known without a device round trip)
    DevBuf cdf32;                              // int32 [T, stride], the reference's _cdfs (zero padded)
    DevBuf blob;                               // packed: TableMeta[T] | u16 cdf[...] | u16 lut[...]
    DevBuf enc;                                // uint4 per CDF entry (same numbering as the u16 CDFs): the encoder's operands of that symbol
                                               // start' | m = ceil(2^32 / f) | f | 0  (rans_pair.cu); read through L2
    size_t blob_bytes = 0, meta_bytes = 0, cdf16_bytes = 0, lut_bytes = 0;
    DevBuf blob_d;                             // the pair decoder's shared-memory image (rans_pair.cu): u16 d[i] = cdf[i] - 1 of every table followed by
                                               // four 0xffff sentinels | u16 lut2 = BYTE offset of the bucket's first candidate inside the d region
    size_t blob_d_bytes = 0, dcdf_bytes = 0;   // (blob_d_bytes = 0: the tables do not fit the 16-bit offsets)
    uint32_t total_cdf = 0, total_lut = 0;
    bool ready = false;
};

// Views into `blob` (either the global copy or its shared-memory image).
struct TableView {
    const TableMeta *meta;
    const uint16_t *cdf;
    const uint16_t *lut;
};

__host__ __device__ inline TableView make_view(const void *blob, size_t meta_bytes, size_t cdf16_bytes)
{
    const char *b = reinterpret_cast<const char *>(blob);
    TableView v;
    v.meta = reinterpret_cast<const TableMeta *>(b);
    v.cdf = reinterpret_cast<const uint16_t *>(b + meta_bytes);
    v.lut = reinterpret_cast<const uint16_t *>(b + meta_bytes + cdf16_bytes);
    return v;
}

#ifdef __CUDACC__
// Table access of the coding kernels.  SM = true: the blob was staged into shared memory and is read with ld.shared
// through 32-bit window addresses (a generic pointer that may be global or shared compiles to generic loads plus 64-bit
// address arithmetic on the state's dependency chain -- it was 40 % of the decoder's step); SM = false: tables larger
// than the shared memory of an SM stay in HBM / L2 and are read through the read-only path.
// The loads are plain (non-volatile) asm: the tables never change after staging, so the compiler may hoist them off the
// chain; init() orders them behind the staging barrier.
template <bool SM> struct Tab;
template <> struct Tab<true> {
    typedef uint32_t addr_t;
    uint32_t meta, cdf, lut;
    __device__ void init(const void *, unsigned char *smem, size_t meta_bytes, size_t cdf16_bytes)
    {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(smem);
        asm volatile("" : "+r"(b)::"memory");
        meta = b;
        cdf = b + (uint32_t)meta_bytes;
        lut = cdf + (uint32_t)cdf16_bytes;
    }
    __device__ uint4 meta_at(int c) const
    {
        uint4 v;
        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(meta + (uint32_t)c * 16u));
        return v;
    }
    __device__ addr_t cdf_at(uint32_t entry) const { return cdf + 2u * entry; }
    __device__ addr_t lut_at(uint32_t entry) const { return lut + 2u * entry; }
    template <int OFF> static __device__ uint32_t ld16(addr_t a)
    {
        uint16_t v;
        asm("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(a), "n"(OFF));
        return v;
    }
};
template <> struct Tab<false> {
    typedef const unsigned char *addr_t;
    const unsigned char *meta, *cdf, *lut;
    __device__ void init(const void *blob, unsigned char *, size_t meta_bytes, size_t cdf16_bytes)
    {
        meta = reinterpret_cast<const unsigned char *>(blob);
        cdf = meta + meta_bytes;
        lut = cdf + cdf16_bytes;
    }
    __device__ uint4 meta_at(int c) const { return __ldg(reinterpret_cast<const uint4 *>(meta) + c); }
    __device__ addr_t cdf_at(uint32_t entry) const { return cdf + 2ull * entry; }
    __device__ addr_t lut_at(uint32_t entry) const { return lut + 2ull * entry; }
    template <int OFF> static __device__ uint32_t ld16(addr_t a) { return __ldg(reinterpret_cast<const uint16_t *>(a + OFF)); }
};

#endif  // __CUDACC__

static constexpr int kLanes = 32;             // lanes of one chunk = one warp
static constexpr uint32_t kRansL = 1u << 16;  // multi-lane state lower bound
static constexpr uint32_t kMagic0 = 0x30534C42u;  // "BLS0": written by the y path when its context model ran on the exact FP32 kernels
static constexpr uint32_t kMagic = 0x31534C42u;   // "BLS1": no context model involved, or 3xTF32
static constexpr uint32_t kMagic2 = 0x32534C42u;  // "BLS2": same container, written by the y path when its context model ran in 3xFP16
static constexpr int kMaxSmemTables = 200 * 1024;

}  // namespace basic

// Definitions shared by the multi-lane coder kernels (rans_lanes.cu: one warp per chunk, every table size and bypass precision;
// rans_pair.cu: main / helper warp pairs, the default configuration).
#pragma once
#include "common.cuh"

namespace basic {

struct SliceDesc {  // one slice of a segment: `n` symbols starting at `off` in the operand arrays
    long long off, n;
    int cs, pad;
};

struct LaneParams {
    const void *blob;
    size_t blob_bytes, meta_bytes, cdf16_bytes;
    int tables_in_smem;
    int T, precision, bypass, bypass_precision;
    long long n;            // symbols in the slice being coded (decoder) / unused (encoder)
    int chunk_syms;         // chunk_syms of that slice, multiple of 128
    const int *n_chunks_dev;  // device scalar (the auto mode decides it on the device); nullptr -> n_chunks
    int n_chunks;
    int n_slices;           // encoder: slices of the segment, walked last to first
    const SliceDesc *slices;
};


namespace {

constexpr int kWarps = 8;      // warps per CTA when the chunks fit 8 per SM
constexpr int kMaxWarps = 16;  // ... and when there are more (throughput mode: 16 x 148 chunks resident)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kSegHdr = 8;  // u32 n_chunks | u32 n_slices, followed by u32 chunk_syms[n_slices]

template <bool SM> __device__ inline Tab<SM> stage_tables(const void *blob, size_t blob_bytes, size_t meta_bytes, size_t cdf16_bytes,
                                                          unsigned char *smem)
{
    if (SM) {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (size_t i = threadIdx.x; i < blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    Tab<SM> t;
    t.init(blob, smem, meta_bytes, cdf16_bytes);
    return t;
}

__device__ inline void cp_async4(uint32_t smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace

}  // namespace basic

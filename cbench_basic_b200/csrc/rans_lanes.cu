// Multi-lane interleaved rANS ("BLS1" segments) -- the throughput coder.
//
// Format (CPU specification: oracle/ans_oracle.c section 4; DESIGN.md "Multi-lane stream"):
//   a segment of n symbols is cut into chunks of `chunk_syms` (multiple of 128) symbols.  One chunk = one
//   warp = 32 interleaved lanes (32-bit states, L = 2^16, 16-bit renormalisation words) sharing one word
//   stream; local symbol j belongs to lane (j % 128) / 4 and is coded at step (j / 128) * 4 + (j % 4), so
//   every lane moves its operands with one 128-bit load / store per 128-symbol block.  Within one event
//   (a coding step, or one escape sub-step) the lanes that renormalise take consecutive words in lane
//   order: a warp ballot + popc prefix gives each lane its word, no atomics, no divergence.
//   A segment may consist of several SLICES (the y path: one per coding group): chunk k owns symbols
//   [k * chunk_syms[g], (k + 1) * chunk_syms[g]) of every slice g and codes them slice after slice with the lane states
//   carried over -- one state flush per lane for the whole segment, and the decoder can stop after any slice, compute
//   the next group's parameters, and continue (states parked in HBM between launches).
//   Segment bytes: u32 n_chunks | u32 n_slices | u32 chunk_syms[n_slices] | u32 end_word[n_chunks] (cumulative) |
//   u32 state[n_chunks][32] | u16 words | zero pad to 4 bytes.
//
// Tables (u16 CDFs + bucket LUTs, tables.cu) are staged once per CTA into shared memory and read with ld.shared (common.cuh
// Tab<SM>; images larger than an SM's shared memory are read through L2).  With the chunk count fixed by the 0.5 % bpp bar
// a call lasts as long as ONE warp's dependency chain, so everything that does not depend on the state is fetched off the
// chain: table records and encoder operands per 128-symbol block, four CDF probes at once in the decoder, escape payloads of
// bypass_precision 4 as one 36-bit token string.  Kernels are instantiated per <tables in shared
// memory, bypass_precision == 4>.
#include "rans_lanes.cuh"

namespace basic {

namespace {

// ------------------------------------------------------------------------------------------------ encode
// scratch layout: chunk k owns words [k * cap_words, (k + 1) * cap_words), filled back to front.
// Outputs per chunk: first_word[k] (index inside the chunk's scratch), states[k * 32 + lane].
template <bool SM, bool BP4, int MAXW = kMaxWarps>   // (MAXW = 32: see k_bls_decode)
__global__ void __launch_bounds__(MAXW * 32, 1)
k_bls_encode(LaneParams P, const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes,
             uint16_t *__restrict__ scratch, int cap_words, uint32_t *__restrict__ first_word,
             uint32_t *__restrict__ states, int *status)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const Tab<SM> tb = stage_tables<SM>(P.blob, P.blob_bytes, P.meta_bytes, P.cdf16_bytes, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    const int n_chunks = P.n_chunks_dev ? *P.n_chunks_dev : P.n_chunks;
    const int prec = P.precision;
    const uint32_t bp = (uint32_t)P.bypass_precision, maxb = (1u << bp) - 1;
    const int tpu = 16 / (int)bp;  // escape tokens per unit
    // n / bp, n / maxb, n / tpu for n < 160 as multiply + shift (exact there, checked for every bp; a runtime integer division is ~25 instructions)
    const uint32_t r_bp = (65536u + bp - 1) / bp, r_maxb = (65536u + maxb - 1) / maxb, r_tpu = (65536u + tpu - 1) / tpu;
    const bool ptr_ok = ((reinterpret_cast<uintptr_t>(symbols) | reinterpret_cast<uintptr_t>(indexes)) & 15) == 0;
    int st = 0;
    // chunks are dealt round-robin over CTAs first so that few chunks still spread over all SMs
    const int nwarps = blockDim.x >> 5;
    for (int k = blockIdx.x + warp * gridDim.x; k < n_chunks; k += gridDim.x * nwarps) {
        uint16_t *wbuf = scratch + (size_t)k * cap_words;
        int pos = cap_words;  // warp-uniform
        uint32_t x = kRansL;
        for (int g = P.n_slices - 1; g >= 0; --g) {  // the encoder walks the chunk's symbols backwards
        const SliceDesc sd = P.slices[g];
        const long long rem = sd.n - (long long)k * sd.cs;
        if (rem <= 0) continue;
        const long long base = sd.off + (long long)k * sd.cs;
        const int m = (int)(rem < sd.cs ? rem : sd.cs);
        const bool vec_ok = ptr_ok && (sd.off & 3) == 0;
        const int nblocks = (m + 127) >> 7;
        auto load_ops = [&](int blk, int32_t(&sy)[4], int32_t(&ix)[4]) {
            const int j0 = blk * 128 + lane * 4;
            if (vec_ok && j0 + 3 < m) {
                const int4 a = __ldg(reinterpret_cast<const int4 *>(symbols + base + j0));
                const int4 b = __ldg(reinterpret_cast<const int4 *>(indexes + base + j0));
                sy[0] = a.x; sy[1] = a.y; sy[2] = a.z; sy[3] = a.w;
                ix[0] = b.x; ix[1] = b.y; ix[2] = b.z; ix[3] = b.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool ok = j0 + q < m;
                    sy[q] = ok ? symbols[base + j0 + q] : 0;
                    ix[q] = ok ? indexes[base + j0 + q] : 0;
                }
            }
        };
        int32_t sy[4], ix[4], syn[4] = {0, 0, 0, 0}, ixn[4] = {0, 0, 0, 0};
        if (nblocks > 0) load_ops(nblocks - 1, syn, ixn);
        for (int blk = nblocks - 1; blk >= 0; --blk) {
            const int j0 = blk * 128 + lane * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) { sy[q] = syn[q]; ix[q] = ixn[q]; }
            if (blk > 0) load_ops(blk - 1, syn, ixn);  // one block ahead: off the dependency chain
            // --- everything that does not depend on the state, for the block's four symbols at once: table lookups,
            // escape classification, the reciprocal of the frequency.  What is left per symbol is the state's own chain.
            uint32_t start[4], freq[4], raw[4];
            float rcp[4];
            bool esc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool active = j0 + q < m;
                int32_t c = ix[q];
                if ((uint32_t)c >= (uint32_t)P.T) { if (active) st |= 1; c = 0; }
                const uint4 mt = tb.meta_at(c);  // cdf_base | lut_base | cdf_size, lut_shift | offset
                const int32_t maxv = (int32_t)(mt.z & 0xffffu) - 2;
                int32_t v = sy[q] - (int32_t)mt.w;
                raw[q] = 0;
                esc[q] = false;
                if (P.bypass) {
                    if (v < 0) { raw[q] = (uint32_t)(-2 * v - 1); v = maxv; }
                    else if (v >= maxv) { raw[q] = (uint32_t)(2 * (v - maxv)); v = maxv; }
                    esc[q] = active && v == maxv;
                } else if (v < 0 || v > maxv) { if (active) st |= 2; v = 0; }
                const typename Tab<SM>::addr_t at = tb.cdf_at(mt.x + (uint32_t)v);
                start[q] = Tab<SM>::template ld16<0>(at);
                freq[q] = (uint16_t)(Tab<SM>::template ld16<2>(at) - start[q]);
                rcp[q] = __frcp_rn(__uint2float_rz(freq[q]));
            }
#pragma unroll
            for (int q = 3; q >= 0; --q) {
                const bool active = j0 + q < m;
                // --- escape units, last to first (oracle: bls_encode_slice): the token list (count tokens, then the
                // digits) goes through the state in units of up to tpu = 16 / bp tokens, one renormalisation check each
                if (BP4) {
                    // bypass_precision 4 (the reference's default): a 32-bit payload has at most 8 digits, so there is one
                    // count token and the whole token list is the 36-bit string  nd | raw << 4 ; unit u = its bits [16u, 16u + 16)
                    if (__any_sync(kFull, esc[q])) {
                        const int nd = esc[q] ? (35 - __clz(raw[q])) >> 2 : 0;
                        const int ntok = esc[q] ? nd + 1 : 0;
                        const int nunits = (ntok + 3) >> 2;
                        const unsigned long long toks = (unsigned long long)nd | ((unsigned long long)raw[q] << 4);
                        const int maxunits = (int)__reduce_max_sync(kFull, (unsigned)nunits);
                        for (int u = maxunits - 1; u >= 0; --u) {
                            const bool part = nunits > u;
                            const int wbits = part ? 4 * min(4, ntok - 4 * u) : 0;
                            const uint32_t unit = (uint32_t)(toks >> (16 * u)) & 0xffffu;
                            const bool emit = part && x >= (1u << (32 - wbits));
                            const unsigned em = __ballot_sync(kFull, emit);
                            pos -= __popc(em);
                            if (emit) {
                                const int at = pos + __popc(em & lt_mask);
                                if (at >= 0) wbuf[at] = (uint16_t)x; else st |= 4;
                                x >>= 16;
                            }
                            if (part) x = (x << wbits) | unit;
                        }
                    }
                } else
                if (__any_sync(kFull, esc[q])) {
                    int nd = 0, ncnt = 0, ntok = 0;
                    if (esc[q]) {
                        nd = (int)(((32 - __clz(raw[q]) + bp - 1) * r_bp) >> 16);  // digits of raw (0 for raw = 0)
                        ncnt = (int)(((uint32_t)nd * r_maxb) >> 16) + 1;
                        ntok = ncnt + nd;
                    }
                    const int nunits = (int)(((uint32_t)(ntok + tpu - 1) * r_tpu) >> 16);
                    const int maxunits = (int)__reduce_max_sync(kFull, (unsigned)nunits);
                    for (int u = maxunits - 1; u >= 0; --u) {
                        const bool part = nunits > u;
                        // the unit: tokens [u * tpu, min((u + 1) * tpu, ntok)), first token in the low bits
                        uint32_t unit = 0;
                        int wbits = 0;
                        if (part) {
                            const int t0 = u * tpu, cnt = min(tpu, ntok - t0);
                            for (int i = 0; i < cnt; ++i) {
                                const int t = t0 + i;
                                uint32_t tok;
                                if (t >= ncnt) tok = (raw[q] >> ((t - ncnt) * bp)) & maxb;
                                else tok = t < ncnt - 1 ? maxb : (uint32_t)(nd - (ncnt - 1) * (int)maxb);
                                unit |= tok << (bp * i);
                            }
                            wbits = (int)bp * cnt;
                        }
                        const bool emit = part && x >= (1u << (32 - wbits));
                        const unsigned em = __ballot_sync(kFull, emit);
                        pos -= __popc(em);
                        if (emit) {
                            const int at = pos + __popc(em & lt_mask);
                            if (at >= 0) wbuf[at] = (uint16_t)x; else st |= 4;
                            x >>= 16;
                        }
                        if (part) x = (x << wbits) | unit;
                    }
                }
                // --- the symbol itself
                // x >= ((L >> prec) << 16) * freq = freq << (32 - prec), without the 64-bit product
                const bool emit = active && (x >> (32 - prec)) >= freq[q];
                const unsigned em = __ballot_sync(kFull, emit);
                pos -= __popc(em);
                if (emit) {
                    const int at = pos + __popc(em & lt_mask);
                    if (at >= 0) wbuf[at] = (uint16_t)x; else st |= 4;
                    x >>= 16;
                }
                if (active) {
                    // x < freq << 16 after the renormalisation above, so qt < 2^16 and a float estimate is within
                    // one of it (exact general division when the invariant does not hold: prec < 16)
                    uint32_t qt, rem;
                    if (prec == 16) {
                        qt = __float2uint_rz(__uint2float_rz(x) * rcp[q]);
                        rem = x - qt * freq[q];
                        if ((int32_t)rem < 0) { --qt; rem += freq[q]; }
                        else if (rem >= freq[q]) { ++qt; rem -= freq[q]; }
                    } else {
                        qt = x / freq[q];
                        rem = x - qt * freq[q];
                    }
                    x = (qt << prec) + rem + start[q];
                }
            }
        }
        }
        states[(size_t)k * 32 + lane] = x;
        if (lane == 0) first_word[k] = (uint32_t)(pos < 0 ? 0 : pos);
    }
    if (st) atomicOr(status, st);
}

// One CTA: cumulative word counts -> directory, plus the total segment size in bytes.
__global__ void __launch_bounds__(1024)
k_bls_scan(const int *n_chunks_dev, int n_chunks_arg, int n_slices, const SliceDesc *__restrict__ slices,
           const uint32_t *__restrict__ first_word, int cap_words, uint32_t *__restrict__ seg /* segment header in the output buffer */,
           long long *seg_bytes)
{
    __shared__ uint32_t part[1024];
    const int n_chunks = n_chunks_dev ? *n_chunks_dev : n_chunks_arg;
    const int tid = threadIdx.x;
    const int per = (n_chunks + 1023) / 1024;
    uint32_t s = 0;
    for (int i = 0; i < per; ++i) {
        const int k = tid * per + i;
        if (k < n_chunks) s += (uint32_t)cap_words - first_word[k];
    }
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
        seg[0] = (uint32_t)n_chunks;
        seg[1] = (uint32_t)n_slices;
        for (int g = 0; g < n_slices; ++g) seg[2 + g] = (uint32_t)slices[g].cs;
        const long long words_at = kSegHdr + 4ll * n_slices + 4ll * n_chunks + 128ll * n_chunks;
        long long total = words_at + 2ll * run;
        total = (total + 3) & ~3ll;
        *seg_bytes = total;
    }
    __syncthreads();
    uint32_t run = part[tid];
    for (int i = 0; i < per; ++i) {
        const int k = tid * per + i;
        if (k < n_chunks) {
            run += (uint32_t)cap_words - first_word[k];
            seg[2 + n_slices + k] = run;  // cumulative END of chunk k, in words
        }
    }
}

// Gather: states + words of every chunk into the contiguous segment.
__global__ void __launch_bounds__(256)
k_bls_gather(const int *n_chunks_dev, int n_chunks_arg, int n_slices, const uint16_t *__restrict__ scratch, int cap_words,
             const uint32_t *__restrict__ first_word, const uint32_t *__restrict__ states, unsigned char *__restrict__ seg)
{
    const int n_chunks = n_chunks_dev ? *n_chunks_dev : n_chunks_arg;
    const uint32_t *end_word = reinterpret_cast<const uint32_t *>(seg) + 2 + n_slices;
    uint32_t *out_states = reinterpret_cast<uint32_t *>(seg) + 2 + n_slices + n_chunks;
    uint16_t *out_words = reinterpret_cast<uint16_t *>(seg + kSegHdr + 4ll * n_slices + 4ll * n_chunks + 128ll * n_chunks);
    for (int k = blockIdx.x; k < n_chunks; k += gridDim.x) {
        if (threadIdx.x < 32) out_states[(size_t)k * 32 + threadIdx.x] = states[(size_t)k * 32 + threadIdx.x];
        const uint32_t end = end_word[k], beg = k ? end_word[k - 1] : 0;
        const uint16_t *src = scratch + (size_t)k * cap_words + first_word[k];
        // 32-bit stores (the destination is 4-byte aligned from its first even word on; the source has whatever parity the
        // chunk's first word had): half the store instructions of a word-by-word copy
        uint16_t *dst = out_words + beg;
        const uint32_t n = end - beg;
        const uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(dst) >> 1) & 1u) < n ? (uint32_t)((reinterpret_cast<uintptr_t>(dst) >> 1) & 1u) : n;
        if (threadIdx.x == 0 && head) dst[0] = src[0];
        const uint32_t pairs = (n - head) >> 1;
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst + head);
        const uint16_t *s2 = src + head;
        if ((reinterpret_cast<uintptr_t>(s2) & 3) == 0) {
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s2);
            for (uint32_t i = threadIdx.x; i < pairs; i += blockDim.x) dst32[i] = s32[i];
        } else {
            for (uint32_t i = threadIdx.x; i < pairs; i += blockDim.x) dst32[i] = (uint32_t)s2[2 * i] | ((uint32_t)s2[2 * i + 1] << 16);
        }
        if (threadIdx.x == 32 && ((n - head) & 1)) dst[n - 1] = src[n - 1];
        if (k == n_chunks - 1 && threadIdx.x == 0 && (end & 1)) out_words[end] = 0;  // pad to 4 bytes
    }
}

// ------------------------------------------------------------------------------------------------ decode
// The renormalisation words of a chunk are consumed strictly in order (wp is warp-uniform), so each warp streams
// them through a private shared-memory ring filled by cp.async two groups ahead: the state update chain never
// waits on a global load.  Ring = kRingUnits u32 units (2 words each), filled in groups of kGroupUnits.
constexpr int kRingUnits = 512, kGroupUnits = 128;

// MAXW: warps per CTA the instantiation is compiled for (registers per thread follow: 32 warps = 64 registers).  Thirty-two
// state-carrying warps per SM hide twice the chain latency of sixteen -- for streams with more than 16 chunks per SM.
constexpr int kHugeWarps = 32;

template <bool SM, bool BP4, int MAXW = kMaxWarps>
__global__ void __launch_bounds__(MAXW * 32, 1)
k_bls_decode(LaneParams P, const unsigned char *__restrict__ seg, long long seg_cap, const int32_t *__restrict__ indexes,
             int32_t *__restrict__ out, int seg_slices, int first_slice, int last_slice, uint32_t *__restrict__ carry_x,
             uint32_t *__restrict__ carry_wp, int *status)
{
    extern __shared__ __align__(16) unsigned char smem[];
    // ring first (fixed size), tables behind it
    uint32_t *ring32 = reinterpret_cast<uint32_t *>(smem) + (threadIdx.x >> 5) * kRingUnits;
    const uint16_t *ring16 = reinterpret_cast<const uint16_t *>(ring32);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring32);
    const Tab<SM> tb = stage_tables<SM>(P.blob, P.blob_bytes, P.meta_bytes, P.cdf16_bytes, smem + (blockDim.x >> 5) * kRingUnits * 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    const int n_chunks = P.n_chunks;
    const int prec = P.precision;
    const uint32_t pmask = (1u << prec) - 1;
    const uint32_t bp = (uint32_t)P.bypass_precision, maxb = (1u << bp) - 1;
    const int tpu = 16 / (int)bp;  // escape tokens per unit
    const uint32_t *end_word = reinterpret_cast<const uint32_t *>(seg) + 2 + seg_slices;
    const uint32_t *states = end_word + n_chunks;
    const long long words_at = kSegHdr + 4ll * seg_slices + 4ll * n_chunks + 128ll * n_chunks;
    const uint32_t *units = reinterpret_cast<const uint32_t *>(seg + words_at);  // 2 words per unit, 4-byte aligned
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(indexes)) & 15) == 0;
    int st = 0;
    // chunks are dealt round-robin over CTAs first so that few chunks still spread over all SMs
    const int nwarps = blockDim.x >> 5;
    for (int k = blockIdx.x + warp * gridDim.x; k < n_chunks; k += gridDim.x * nwarps) {
        const long long base = (long long)k * P.chunk_syms;
        const long long rem = P.n - base;
        const int m = (int)(rem <= 0 ? 0 : rem < P.chunk_syms ? rem : P.chunk_syms);
        uint32_t wend = end_word[k], wbeg = k ? end_word[k - 1] : 0;
        if (wend < wbeg || words_at + 2ll * wend > seg_cap) { st |= 4; wend = wbeg = 0; }  // corrupt directory
        const uint32_t u_lim = (wend + 1) >> 1;       // units holding words of this chunk end here
        // absolute word index, warp-uniform; a later slice continues where the previous launch stopped
        uint32_t wp = first_slice ? wbeg : carry_wp[k];
        if (wp < wbeg || wp > wend) { st |= 4; wp = wend; }
        uint32_t fill_u = wp >> 1, ready_u = fill_u;
        auto issue_group = [&]() {
#pragma unroll
            for (int i = 0; i < kGroupUnits / 32; ++i) {
                const uint32_t u = fill_u + i * 32 + lane;
                if (u < u_lim) cp_async4(ring_s + (u & (kRingUnits - 1)) * 4, units + u);
            }
            cp_async_commit();
            fill_u += kGroupUnits;
        };
        __syncwarp();  // the previous chunk's readers are done with the ring
        issue_group();
        issue_group();
        cp_async_wait<1>();
        __syncwarp();
        ready_u = fill_u - kGroupUnits;
        // Words must have landed before the events that consume them.  The check runs once per block of four coding
        // steps (up to 128 words) and once per escape sub-step (up to 32 words); both keep 160 words ahead, so that the
        // steps of a block are still covered after any number of escape sub-steps in between.  Unread words never
        // exceed 3 groups = 384 of the ring's 512 units.
        auto ensure = [&]() {
            if (((wp + 160) >> 1) + 1 > ready_u) {
                cp_async_wait<0>();
                __syncwarp();
                ready_u = fill_u;
                issue_group();
            }
        };
        uint32_t x = first_slice ? states[(size_t)k * 32 + lane] : carry_x[(size_t)k * 32 + lane];
        const int nblocks = (m + 127) >> 7;
        auto load_ix = [&](int blk) -> int4 {
            const int j0 = blk * 128 + lane * 4;
            if (vec_ok && j0 + 3 < m) return __ldg(reinterpret_cast<const int4 *>(indexes + base + j0));
            int4 b;
            b.x = j0 + 0 < m ? indexes[base + j0 + 0] : 0;
            b.y = j0 + 1 < m ? indexes[base + j0 + 1] : 0;
            b.z = j0 + 2 < m ? indexes[base + j0 + 2] : 0;
            b.w = j0 + 3 < m ? indexes[base + j0 + 3] : 0;
            return b;
        };
        int4 ixn = make_int4(0, 0, 0, 0);
        if (nblocks > 0) ixn = load_ix(0);
        for (int blk = 0; blk < nblocks; ++blk) {
            const int j0 = blk * 128 + lane * 4;
            int32_t res[4];
            const bool full4 = vec_ok && j0 + 3 < m;
            const int32_t ix[4] = {ixn.x, ixn.y, ixn.z, ixn.w};
            if (blk + 1 < nblocks) ixn = load_ix(blk + 1);  // one block ahead: off the dependency chain
            // table records of the block's four symbols: known from the indexes alone, fetched before the chain starts
            uint4 mt[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int32_t c = ix[q];
                if ((uint32_t)c >= (uint32_t)P.T) { if (j0 + q < m) st |= 1; c = 0; }
                mt[q] = tb.meta_at(c);  // cdf_base | lut_base | cdf_size, lut_shift | offset
            }
            ensure();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool active = j0 + q < m;
                const typename Tab<SM>::addr_t cd = tb.cdf_at(mt[q].x);
                const int nsyms = (int)(mt[q].z & 0xffffu) - 1, maxv = nsyms - 1;
                const uint32_t cum = x & pmask;
                // chain: bucket LUT -> four CDF entries at once (the answer is among the first three candidates unless
                // the bucket sits in a tail of width-1 symbols) -> multiply -> ballot -> word
                int s = (int)Tab<SM>::template ld16<0>(tb.lut_at(mt[q].y + (cum >> ((mt[q].z >> 16) & 0xffu))));
                const typename Tab<SM>::addr_t e = cd + 2 * s;
                const uint32_t c0 = Tab<SM>::template ld16<0>(e), c1 = Tab<SM>::template ld16<2>(e);
                const uint32_t c2 = Tab<SM>::template ld16<4>(e), c3 = Tab<SM>::template ld16<6>(e);
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
                {
                    const bool need = active && x < kRansL;
                    const unsigned nm = __ballot_sync(kFull, need);
                    if (need) {
                        const uint32_t at = wp + __popc(nm & lt_mask);
                        uint32_t word = 0;
                        if (at < wend) word = ring16[at & (2 * kRingUnits - 1)]; else st |= 4;
                        x = (x << 16) | word;
                    }
                    wp += __popc(nm);
                }
                int32_t value = s;
                const bool esc = active && P.bypass && s == maxv;
                if (BP4) {
                    // bypass_precision 4: the first unit starts with the digit count nb (<= 8 for a 32-bit payload, one count
                    // token), followed by the digits, least significant first, four tokens per unit
                    if (__any_sync(kFull, esc)) {
                        bool in = esc, first = true;
                        uint32_t nb = 0, raw = 0, jj = 0;
                        while (__any_sync(kFull, in)) {
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
                            ensure();
                            const bool need = was && x < kRansL;
                            const unsigned nm = __ballot_sync(kFull, need);
                            if (need) {
                                const uint32_t at = wp + __popc(nm & lt_mask);
                                uint32_t word = 0;
                                if (at < wend) word = ring16[at & (2 * kRingUnits - 1)]; else st |= 4;
                                x = (x << 16) | word;
                            }
                            wp += __popc(nm);
                        }
                        if (esc) {
                            const int32_t v2 = (int32_t)(raw >> 1);
                            value = (raw & 1) ? -v2 - 1 : v2 + maxv;
                        }
                    }
                } else
                if (__any_sync(kFull, esc)) {
                    bool in = esc;
                    int phase = 0;
                    uint32_t nb = 0, raw = 0, jj = 0;
                    while (__any_sync(kFull, in)) {
                        // one unit: parse up to tpu tokens out of the low 16 bits (x >= 2^16 here), pop exactly those
                        const bool was = in;
                        if (in) {
                            int used = 0;
                            for (int i = 0; i < tpu && in; ++i, ++used) {
                                const uint32_t val = (x >> (bp * i)) & maxb;
                                if (phase == 0) {
                                    nb += val;
                                    if (val != maxb) { phase = 1; if (nb == 0) in = false; }
                                    else if (nb > 64) { in = false; st |= 4; }
                                } else {
                                    if (jj * bp < 32) raw |= val << (jj * bp);
                                    if (++jj == nb) in = false;
                                }
                            }
                            x >>= bp * used;
                        }
                        ensure();
                        const bool need = was && x < kRansL;
                        const unsigned nm = __ballot_sync(kFull, need);
                        if (need) {
                            const uint32_t at = wp + __popc(nm & lt_mask);
                            uint32_t word = 0;
                            if (at < wend) word = ring16[at & (2 * kRingUnits - 1)]; else st |= 4;
                            x = (x << 16) | word;
                        }
                        wp += __popc(nm);
                    }
                    if (esc) {
                        const int32_t v2 = (int32_t)(raw >> 1);
                        value = (raw & 1) ? -v2 - 1 : v2 + maxv;
                    }
                }
                res[q] = value + (int32_t)mt[q].w;
            }
            if (full4) {
                *reinterpret_cast<int4 *>(out + base + j0) = make_int4(res[0], res[1], res[2], res[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (j0 + q < m) out[base + j0 + q] = res[q];
            }
        }
        cp_async_wait<0>();
        if (last_slice) {
            if (wp != wend && lane == 0) st |= 4;
        } else {
            carry_x[(size_t)k * 32 + lane] = x;
            if (lane == 0) carry_wp[k] = wp;
        }
    }
    if (st) atomicOr(status, st);
}


// Sampled size estimate for the auto lane count: mean cost in bits of up to `max_samples` (symbol, index)
// pairs taken at a regular stride; *out_bits = estimated bits of the whole segment (float accumulation is fine:
// the estimate only picks the chunk count, which is stored in the stream).
__global__ void __launch_bounds__(256)
k_estimate_bits(LaneParams P, const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, long long stride,
                long long samples, float *out_bits)
{
    __shared__ float red[8];
    const TableView tv = make_view(P.blob, P.meta_bytes, P.cdf16_bytes);
    const uint32_t bp = (uint32_t)P.bypass_precision, maxb = (1u << bp) - 1;
    float bits = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < samples; i += (long long)gridDim.x * blockDim.x) {
        const long long j = i * stride;
        int32_t c = indexes[j];
        if ((uint32_t)c >= (uint32_t)P.T) c = 0;
        const TableMeta mt = tv.meta[c];
        const int32_t maxv = (int32_t)mt.cdf_size - 2;
        int32_t v = symbols[j] - mt.offset;
        uint32_t raw = 0;
        float extra = 0.f;
        if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = maxv; }
        else if (v >= maxv) { raw = (uint32_t)(2 * (v - maxv)); v = maxv; }
        if (P.bypass && v == maxv) {
            const int nd = raw ? (int)((32 - __clz(raw) + bp - 1) / bp) : 0;
            extra = (float)((nd / (int)maxb + 1 + nd) * (int)bp);
        }
        const uint32_t start = tv.cdf[mt.cdf_base + v];
        const uint32_t freq = (uint16_t)(tv.cdf[mt.cdf_base + v + 1] - start);
        bits += (float)P.precision - __log2f((float)(freq ? freq : 1)) + extra;
    }
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(kFull, bits, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out_bits, t * (float)((double)P.n / (double)samples));
    }
}

}  // namespace

// rans_pair.cu
bool pair_kernels_apply(const RansTables &tb, int bypass_precision);
int launch_pair_encode(const RansTables &tb, const LaneParams &P, const int32_t *d_sym, const int32_t *d_idx, uint16_t *d_scratch,
                       int cap_words, uint32_t *d_first, uint32_t *d_states, int *d_status, int sm_count, cudaStream_t stream);
int launch_pair_decode(const RansTables &tb, const LaneParams &P, const unsigned char *d_seg, int64_t seg_cap, const int32_t *d_idx,
                       int seg_slices, int slice, uint32_t *d_carry_x, uint32_t *d_carry_wp, int32_t *d_out, int *d_status,
                       int sm_count, cudaStream_t stream);

static int smem_for(const RansTables &tb) { return tb.blob_bytes <= (size_t)kMaxSmemTables ? (int)tb.blob_bytes : 0; }
static constexpr int kRingBytes = kMaxWarps * kRingUnits * 4;
// warps per CTA: 8 while every chunk gets its own resident warp, 16 beyond that (more lanes in flight per SM)
static constexpr int kSmemLimit = 232448;  // opt-in dynamic shared memory of one CTA on sm_100
static int warps_for(int n_chunks, int sm_count, int table_smem)
{
    if (n_chunks <= kWarps * sm_count) return kWarps;
    return table_smem + kMaxWarps * kRingUnits * 4 <= kSmemLimit ? kMaxWarps : kWarps;
}

static LaneParams make_params(const RansTables &tb, int bypass, int bypass_precision, int64_t n, int chunk_syms, int n_chunks,
                              int n_slices, const SliceDesc *slices)
{
    LaneParams P;
    P.blob = tb.blob.p;
    P.blob_bytes = tb.blob_bytes;
    P.meta_bytes = tb.meta_bytes;
    P.cdf16_bytes = tb.cdf16_bytes;
    P.tables_in_smem = smem_for(tb) > 0;
    P.T = tb.T;
    P.precision = tb.precision;
    P.bypass = bypass;
    P.bypass_precision = bypass_precision;
    P.n = n;
    P.chunk_syms = chunk_syms;
    P.n_chunks = n_chunks;
    P.n_chunks_dev = nullptr;
    P.n_slices = n_slices;
    P.slices = slices;
    return P;
}

static int grid_for(int n_chunks, int sm_count)
{
    int g = n_chunks < sm_count ? n_chunks : sm_count;  // one resident CTA per SM (tables fill its shared memory)
    return g < 1 ? 1 : g;
}

static int set_attrs()
{
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        const int dec_smem = kMaxSmemTables + kRingBytes < kSmemLimit ? kMaxSmemTables + kRingBytes : kSmemLimit;
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_encode<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTables));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_encode<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTables));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_encode<true, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTables));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_decode<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_decode<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_decode<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_decode<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem));
        BASIC_CUDA(cudaFuncSetAttribute(k_bls_decode<true, true, kHugeWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    }
    return BASIC_OK;
}

// Encodes one segment of `n_slices` slices (device array `d_slices`: offsets into d_sym / d_idx, symbol counts and
// chunk_syms) into `seg_out` (device).  *d_seg_bytes (device) receives its size.
int launch_bls_encode(const RansTables &tb, int bypass, int bypass_precision, const int32_t *d_sym, const int32_t *d_idx,
                      int n_slices, const void *d_slices, int n_chunks, uint16_t *d_scratch, int cap_words, uint32_t *d_first,
                      uint32_t *d_states, unsigned char *d_seg_out, long long *d_seg_bytes, int *d_status, int sm_count,
                      cudaStream_t stream)
{
    const SliceDesc *sl = reinterpret_cast<const SliceDesc *>(d_slices);
    const LaneParams P = make_params(tb, bypass, bypass_precision, 0, 128, n_chunks, n_slices, sl);
    const int smem = smem_for(tb);
    BASIC_TRY(set_attrs());
    static const bool huge_off = [] { const char *e = getenv("BASIC_CODER_WARPS32"); return e && e[0] == '0'; }();   // A/B switch
    if (!huge_off && n_chunks > kMaxWarps * sm_count && smem > 0 && bypass_precision == 4) {
        // more than 16 chunks per SM (lanes chosen freely): 32 state-carrying warps per SM beat eight main / helper groups
        k_bls_encode<true, true, 32><<<grid_for(n_chunks, sm_count), 32 * 32, smem, stream>>>(P, d_sym, d_idx, d_scratch, cap_words, d_first,
                                                                                                d_states, d_status);
        BASIC_LAUNCHED();
    } else if (n_chunks > 0 && smem > 0 && pair_kernels_apply(tb, bypass_precision)) {
        BASIC_TRY(launch_pair_encode(tb, P, d_sym, d_idx, d_scratch, cap_words, d_first, d_states, d_status, sm_count, stream));
    } else if (n_chunks > 0) {
        const dim3 grid(grid_for(n_chunks, sm_count)), block(warps_for(n_chunks, sm_count, smem) * 32);
        // instantiations: tables in shared memory or not x bypass_precision 4 (the reference's default; its escape code is a
        // fraction of the general one -- the kernels are instruction-fetch sensitive with one warp per scheduler)
        auto kern = smem > 0 ? (bypass_precision == 4 ? k_bls_encode<true, true> : k_bls_encode<true, false>)
                             : (bypass_precision == 4 ? k_bls_encode<false, true> : k_bls_encode<false, false>);
        kern<<<grid, block, smem, stream>>>(P, d_sym, d_idx, d_scratch, cap_words, d_first, d_states, d_status);
        BASIC_LAUNCHED();
    }
    k_bls_scan<<<1, 1024, 0, stream>>>(nullptr, n_chunks, n_slices, sl, d_first, cap_words, reinterpret_cast<uint32_t *>(d_seg_out),
                                       d_seg_bytes);
    BASIC_LAUNCHED();
    if (n_chunks > 0) {
        int g = n_chunks < 4 * sm_count ? n_chunks : 4 * sm_count;
        k_bls_gather<<<g, 256, 0, stream>>>(nullptr, n_chunks, n_slices, d_scratch, cap_words, d_first, d_states, d_seg_out);
        BASIC_LAUNCHED();
    }
    return BASIC_OK;
}

// Decodes slice `slice` (n symbols, chunk_syms) of a segment with seg_slices slices; lane states and word positions
// are parked in d_carry_x [n_chunks * 32] / d_carry_wp [n_chunks] between the slices of one segment.
int launch_bls_decode(const RansTables &tb, int bypass, int bypass_precision, const unsigned char *d_seg, int64_t seg_cap,
                      const int32_t *d_idx, int64_t n, int chunk_syms, int n_chunks, int seg_slices, int slice, uint32_t *d_carry_x,
                      uint32_t *d_carry_wp, int32_t *d_out, int *d_status, int sm_count, cudaStream_t stream)
{
    if (n_chunks <= 0) return BASIC_OK;
    const LaneParams P = make_params(tb, bypass, bypass_precision, n, chunk_syms, n_chunks, 0, nullptr);
    const int smem = smem_for(tb);
    // pairs while every chunk gets its own resident main warp (the lane count the size bar allows); beyond that -- lanes chosen
    // freely, throughput mode -- sixteen state-carrying warps per SM do better than eight pairs
    if (smem > 0 && n_chunks <= 8 * sm_count && pair_kernels_apply(tb, bypass_precision))
        return launch_pair_decode(tb, P, d_seg, seg_cap, d_idx, seg_slices, slice, d_carry_x, d_carry_wp, d_out, d_status, sm_count, stream);
    BASIC_TRY(set_attrs());
    int nw = warps_for(n_chunks, sm_count, smem);
    const dim3 grid(grid_for(n_chunks, sm_count));
    auto kern = smem > 0 ? (bypass_precision == 4 ? k_bls_decode<true, true> : k_bls_decode<true, false>)
                         : (bypass_precision == 4 ? k_bls_decode<false, true> : k_bls_decode<false, false>);
    static const bool huge_off = [] { const char *e = getenv("BASIC_CODER_WARPS32"); return e && e[0] == '0'; }();   // A/B switch
    if (!huge_off && smem > 0 && bypass_precision == 4 && n_chunks > kMaxWarps * sm_count &&
        smem + kHugeWarps * kRingUnits * 4 <= kSmemLimit) {   // more than 16 chunks per SM: 32 warps per CTA
        nw = kHugeWarps;
        kern = k_bls_decode<true, true, kHugeWarps>;
    }
    const dim3 block(nw * 32);
    kern<<<grid, block, smem + nw * kRingUnits * 4, stream>>>(P, d_seg, seg_cap, d_idx, d_out, seg_slices, slice == 0, slice == seg_slices - 1,
                                                              d_carry_x, d_carry_wp, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// Adds the estimated size in bits of the segment to *d_bits (device float, zeroed by the caller).
int launch_estimate_bits(const RansTables &tb, int bypass, int bypass_precision, const int32_t *d_sym, const int32_t *d_idx,
                         int64_t n, float *d_bits, cudaStream_t stream)
{
    if (n <= 0) return BASIC_OK;
    const LaneParams P = make_params(tb, bypass, bypass_precision, n, 128, 0, 0, nullptr);
    const long long max_samples = 1 << 16;
    const long long stride = n > max_samples ? n / max_samples : 1;
    const long long samples = (n + stride - 1) / stride;
    int blocks = (int)((samples + 255) / 256);
    if (blocks > 64) blocks = 64;
    k_estimate_bits<<<blocks, 256, 0, stream>>>(P, d_sym, d_idx, stride, samples, d_bits);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

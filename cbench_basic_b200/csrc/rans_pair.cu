// Multi-lane interleaved rANS as MAIN / HELPER warp pairs -- the kernels of the default configuration (tables resident in
// shared memory, bypass_precision 4).  Same container, byte for byte, as rans_lanes.cu and the CPU specification
// (oracle/ans_oracle.c section 4).
//
// Why pairs: the 0.5 % size bar fixes the number of chunks (= warps that carry a state), so a coding call lasts as long as
// ONE warp needs for its chunk: steps x the length of the state's dependency chain, with the instructions the same warp has
// to issue around the chain added on top (one warp per scheduler: nothing else hides them).  Here the warp that owns the
// states (main) issues the chain and nothing else; a second warp (helper) does everything that does not depend on the state:
//   decoder helper: loads the indexes, fetches the table records, writes per-symbol lookup operands into a shared-memory
//                   ring one block (4 steps) ahead, streams the renormalisation words global -> shared (cp.async), and
//                   moves the decoded symbols shared -> global with 128-bit stores;
//   encoder helper: loads symbols and indexes, classifies escapes, looks up (start, freq), computes 1 / freq, and writes
//                   one 16-byte operand record per symbol into the ring.
// Hand-over is by monotonically increasing block counters in shared memory (st.release / ld.acquire, CTA scope).
#include <cstdio>
#include <cstdlib>

#include "rans_lanes.cuh"

namespace basic {

namespace {

constexpr int kPairs = 8;          // chunk slots per CTA: warps [0, kPairs) are the main warps, [kPairs, 2 kPairs) their helpers
constexpr int kWordUnits = 512;    // word ring per slot: u32 units of two 16-bit words (2 KB)
constexpr int kWordGroup = 128;    // ... filled in groups of this many units
constexpr int kDecParBlocks = 2;   // decoder operand ring: blocks of 4 steps x 32 lanes x 16 B (2 KB each)
constexpr int kEncParBlocks = 6;   // encoder operand ring (two blocks per helper)
constexpr int kEncHelpers = 3;     // encoder: helper warps per main warp (a helper is as latency-bound as a main warp); helper h prepares the blocks with (block number) % kEncHelpers == h
constexpr int kOutBlocks = 2;      // decoder output ring: blocks of 4 steps x 32 lanes x 4 B
constexpr int kCtrlBytes = 64;     // per slot: params_ready | out_done | stored | wp_pub | ready_w | w_epoch | pad | pad | flags of the ring's blocks [8]

constexpr int kDecSlotBytes = kWordUnits * 4 + kDecParBlocks * 2048 + kOutBlocks * 512;
constexpr int kEncSlotBytes = kEncParBlocks * 2048;

__device__ inline void st_release(uint32_t addr, uint32_t v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// Per-block hand-over flags.  st.release.cta compiles to MEMBAR.ALL.CTA + ST, and the membar waits for EVERY memory
// operation of the warp still in flight -- the helper's operand prefetch from global memory, the main warp's word stores --
// which put a full global-memory round trip into every block (measured: 1500 of a helper's 1650 cycles per block).  The
// data handed over lives in shared memory only and is written by the same warp that raises the flag: shared-memory stores of
// one warp are performed in program order, so __syncwarp() (all lanes' data stores issued) followed by a plain store of the
// flag is sufficient; the asm is volatile with a memory clobber, so the compiler keeps the order too.
__device__ inline void st_flag(uint32_t addr, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ inline uint32_t ld_acquire(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ inline void st_relaxed(uint32_t addr, uint32_t v)
{
    asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ inline uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ inline void sts128(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ inline uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ inline void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ inline uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");  // zero-extended
    return v;
}
// The helper keeps ONE block of operands in registers ahead of its work; what hides the DRAM latency (a block of four steps
// is shorter than a trip to HBM) is a prefetch into L2 several blocks further ahead.
constexpr int kPrefetchBlocks = 12;
__device__ inline void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// spin until the counter at `addr` reaches `want` (warp-uniform: every lane polls the same word)
__device__ inline uint32_t wait_ge(uint32_t addr, uint32_t want)
{
    uint32_t v = ld_acquire(addr);
    while ((int32_t)(v - want) < 0) {
        v = ld_acquire(addr);   // (no nanosleep: its granularity is far above a block's duration)
    }
    return v;
}

enum { C_PARAMS = 0, C_DONE = 4, C_STORED = 8, C_WP = 12, C_READYW = 16, C_EPOCH = 20, C_FLAGS = 32 };

// ------------------------------------------------------------------------------------------------ decode
// One coding step of the main warp.  Operand record Q (written by the helper): lut address | cdf address |
// lut shift, symbols << 8, active << 31 | offset.  Everything the common case needs is straight-line predicated code: the
// single warp of a scheduler pays the full latency of every dependent instruction and of every branch.
__global__ void __launch_bounds__(kPairs * 64, 1)
k_pair_decode(LaneParams P, const uint4 *__restrict__ blob_d, int blob_d_bytes, int dcdf_bytes,
              const unsigned char *__restrict__ seg, long long seg_cap, const int32_t *__restrict__ indexes,
              int32_t *__restrict__ out, int seg_slices, int first_slice, int last_slice, uint32_t *__restrict__ carry_x,
              uint32_t *__restrict__ carry_wp, int *status)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp % kPairs;
    const bool is_main = warp < kPairs;
    unsigned char *slot_mem = smem + kPairs * kCtrlBytes + slot * kDecSlotBytes;
    const uint32_t ctrl = (uint32_t)__cvta_generic_to_shared(smem + slot * kCtrlBytes);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(slot_mem);             // words
    const uint32_t par_s = ring_s + kWordUnits * 4;                                   // operands
    const uint32_t out_s = par_s + kDecParBlocks * 2048;                              // decoded symbols
    if (threadIdx.x < kPairs * kCtrlBytes / 4) reinterpret_cast<uint32_t *>(smem)[threadIdx.x] = 0;
    // the decoder's table image (tables.cu): d[i] = cdf[i] - 1 with sentinels | lut2 = byte offsets into the d region
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem + kPairs * (kCtrlBytes + kDecSlotBytes));
        for (int i = threadIdx.x; i < blob_d_bytes / 16; i += blockDim.x) dst[i] = blob_d[i];
        __syncthreads();
    }
    uint32_t d_s = (uint32_t)__cvta_generic_to_shared(smem + kPairs * (kCtrlBytes + kDecSlotBytes));   // d region (16-byte aligned)
    asm volatile("" : "+r"(d_s)::"memory");
    const uint32_t lut_s = d_s + (uint32_t)dcdf_bytes;
    const uint4 *meta_tab = reinterpret_cast<const uint4 *>(P.blob);   // 16-byte table records, read by the helper (L1 resident)
    const unsigned lt_mask = (1u << lane) - 1;
    const int n_chunks = P.n_chunks;
    const bool bypass = P.bypass != 0;
    const uint32_t *end_word = reinterpret_cast<const uint32_t *>(seg) + 2 + seg_slices;
    const uint32_t *states = end_word + n_chunks;
    const long long words_at = kSegHdr + 4ll * seg_slices + 4ll * n_chunks + 128ll * n_chunks;
    const uint32_t *units = reinterpret_cast<const uint32_t *>(seg + words_at);  // 2 words per unit, 4-byte aligned
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(indexes)) & 15) == 0;
    int st = 0;
#ifdef PAIR_TIMING
    long long td_wait = 0, td_work = 0, td_words = 0;
    const long long td_start = clock64();
#endif
    uint32_t gb = 0;     // blocks of this slot so far (all its chunks): the counters in `ctrl` count in these
    uint32_t seq = 0;    // chunks of this slot so far
    for (int k = blockIdx.x + slot * gridDim.x; k < n_chunks; k += gridDim.x * kPairs) {
        ++seq;
        const long long base = (long long)k * P.chunk_syms;
        const long long rem = P.n - base;
        const int m = (int)(rem <= 0 ? 0 : rem < P.chunk_syms ? rem : P.chunk_syms);
        const int nblocks = (m + 127) >> 7;
        uint32_t wend = end_word[k], wbeg = k ? end_word[k - 1] : 0;
        if (wend < wbeg || words_at + 2ll * wend > seg_cap) { st |= 4; wend = wbeg = 0; }  // corrupt directory
        uint32_t wp = first_slice ? wbeg : carry_wp[k];  // absolute word index; a later slice continues where the previous stopped
        if (wp < wbeg || wp > wend) { st |= 4; wp = wend; }
        const uint32_t gb0 = gb;
        if (is_main) {
            // ================================================================================= main: the state's chain
            uint32_t x = first_slice ? states[(size_t)k * 32 + lane] : carry_x[(size_t)k * 32 + lane];
            wait_ge(ctrl + C_EPOCH, seq);               // the helper serves this chunk's words from now on
            uint32_t ready_w = ld_acquire(ctrl + C_READYW);
            // words must have landed before the events that consume them: 160 ahead at every block and escape sub-step
            // (a block's four steps take up to 128, a sub-step up to 32)
            auto ensure = [&]() {
                const uint32_t need = wp + 160 < wend ? wp + 160 : wend;
                if ((int32_t)(ready_w - need) < 0) {
#ifdef PAIR_TIMING
                    const long long t_e0 = clock64();
#endif
                    if (lane == 0) st_relaxed(ctrl + C_WP, wp);
                    ready_w = wait_ge(ctrl + C_READYW, need);
#ifdef PAIR_TIMING
                    td_words += clock64() - t_e0;
#endif
                }
            };
            // one renormalisation event: the lanes below L take consecutive words in lane order
            auto refill = [&](bool need) {
                const unsigned nm = __ballot_sync(kFull, need);
                const uint32_t at = wp + __popc(nm & lt_mask);
                const uint32_t word = lds16(ring_s + ((at & (2 * kWordUnits - 1)) << 1));
                x = need ? (x << 16) | word : x;
                wp += __popc(nm);
            };
            for (int blk = 0; blk < nblocks; ++blk, ++gb) {
#ifdef PAIR_TIMING
                const long long t_w0 = clock64();
#endif
                wait_ge(ctrl + C_PARAMS, gb + 1);
                if (gb >= kOutBlocks) wait_ge(ctrl + C_STORED, gb + 1 - kOutBlocks);
#ifdef PAIR_TIMING
                const long long t_w1 = clock64();
                td_wait += t_w1 - t_w0;
#endif
                const uint32_t pbase = par_s + (gb % kDecParBlocks) * 2048 + lane * 16;
                const uint32_t obase = out_s + (gb % kOutBlocks) * 512 + lane * 4;
                uint4 Pq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) Pq[q] = lds128(pbase + q * 512);
                ensure();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // operands: lut2 address of the table | address of its escape symbol's d entry | value offset (the decoded value
                    // is (address of the symbol's d entry >> 1) + this) | lut shift, escape symbol << 8, active << 31
                    const uint4 Q = Pq[q];
                    const bool active = (int32_t)Q.w < 0;
                    const uint32_t cum = x & 0xffffu;
                    const uint32_t e = d_s + lds16(Q.x + ((cum >> (Q.w & 31u)) << 1));   // first candidate of the bucket
                    const uint32_t e0 = lds16(e), e1 = lds16(e + 2), e2 = lds16(e + 4), e3 = lds16(e + 6);
                    const bool a1 = e1 < cum, a2 = e2 < cum, a3 = e3 < cum;               // (monotone; false at every sentinel)
                    uint32_t dsel = a2 ? e2 : a1 ? e1 : e0, dnext = a2 ? e3 : a1 ? e2 : e1;
                    uint32_t es = e + (a1 ? 2u : 0u) + (a2 ? 2u : 0u);                    // address of the decoded symbol's d entry
                    const bool rare = active && (a3 || (bypass && es == Q.y));
                    const uint32_t xc = ((dnext - dsel) & 0xffffu) * (x >> 16) + ((cum - 1u - dsel) & 0xffffu);   // (before the vote: off its latency)
                    if (!__any_sync(kFull, rare)) {
                        // ---- the common step
                        x = active ? xc : x;
                        refill(active && x < kRansL);
                        sts32(obase + q * 128, (es >> 1) + Q.z);
                        continue;
                    }
                    // ---- a lane sits in a tail of narrow symbols (eight more entries per round) and / or decoded an escape
                    if (__any_sync(kFull, a3)) {
                        bool more = a3;
                        uint32_t ec = e + 6;       // d[ec] < cum is known for the lanes still searching
                        uint32_t lo = e3;
                        while (__any_sync(kFull, more)) {
                            uint32_t cc[9];
#pragma unroll
                            for (int i = 1; i <= 8; ++i) cc[i] = lds16(ec + 2 * i);
                            cc[0] = lo;
                            int adv = 0;
                            bool run = more;
#pragma unroll
                            for (int i = 1; i <= 8; ++i) {
                                run = run && cc[i] < cum;
                                adv += (int)run;
                            }
                            uint32_t st_ = cc[0], nx_ = cc[1];
#pragma unroll
                            for (int i = 1; i <= 7; ++i) if (adv >= i) { st_ = cc[i]; nx_ = cc[i + 1]; }
                            if (more) {
                                ec += 2u * (uint32_t)adv;
                                lo = cc[8];
                                if (adv < 8) { es = ec; dsel = st_; dnext = nx_; more = false; }
                            }
                        }
                    }
                    {
                        const uint32_t freq = (dnext - dsel) & 0xffffu;
                        const uint32_t xn = freq * (x >> 16) + ((cum - 1u - dsel) & 0xffffu);
                        x = active ? xn : x;
                        refill(active && x < kRansL);
                    }
                    uint32_t value = (es >> 1) + Q.z;
                    const bool esc = active && bypass && es == Q.y;
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
                            refill(was && x < kRansL);
                            if (wp > wend) { st |= 4; wp = wend; in = false; }   // truncated / corrupt stream
                        }
                        if (esc) {   // value = maxv + offset at this point
                            const int32_t v2 = (int32_t)(raw >> 1);
                            const int32_t maxv = (int32_t)((Q.w >> 8) & 0xffffu);
                            value = (raw & 1) ? (uint32_t)((int32_t)value - maxv - v2 - 1) : (uint32_t)((int32_t)value + v2);
                        }
                    }
                    sts32(obase + q * 128, value);
                }
                if (wp > wend) { st |= 4; wp = wend; }   // reads past the chunk's words: truncated / corrupt stream
                __syncwarp();
                if (lane == 0) {
                    st_relaxed(ctrl + C_WP, wp);
                    st_flag(ctrl + C_DONE, gb + 1);
                }
#ifdef PAIR_TIMING
                td_work += clock64() - t_w1;
#endif
            }
            if (last_slice) {
                if (wp != wend && lane == 0) st |= 4;
            } else {
                carry_x[(size_t)k * 32 + lane] = x;
                if (lane == 0) carry_wp[k] = wp;
            }
        } else {
            // ================================================================================= helper: everything else
            const uint32_t u_lim = (wend + 1) >> 1;  // units holding words of this chunk end here
            const uint32_t wp0 = wp;
            uint32_t fill_u = wp0 >> 1, wp_known = wp0;
            bool serving = false;                    // the word ring belongs to this chunk (the main warp left the previous one)
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
            int blk_par = 0, blk_out = 0;
            int4 ixn = make_int4(0, 0, 0, 0);
            if (nblocks > 0) ixn = load_ix(0);
            for (int b = 1 + (lane >> 2); b < kPrefetchBlocks && b < nblocks; b += 8) prefetch_l2(indexes + base + b * 128 + (lane & 3) * 32);
            for (uint32_t u = fill_u + lane * 32; u < u_lim && u < fill_u + 3 * kWordGroup; u += 32 * 32) prefetch_l2(units + u);
            while (blk_out < nblocks || !serving) {
                const uint32_t done = ld_acquire(ctrl + C_DONE);
                bool progress = false;
                // --- lookup operands of the next block (its ring slot is free once the main warp left block - kDecParBlocks)
                if (blk_par < nblocks && (int32_t)(gb0 + blk_par - done) < kDecParBlocks) {
                    const int j0 = blk_par * 128 + lane * 4;
                    const int32_t ix[4] = {ixn.x, ixn.y, ixn.z, ixn.w};
                    if (blk_par + 1 < nblocks) ixn = load_ix(blk_par + 1);
                    if (lane < 4 && blk_par + kPrefetchBlocks < nblocks)   // four 128-byte lines per block
                        prefetch_l2(indexes + base + (blk_par + kPrefetchBlocks) * 128 + lane * 32);
                    const uint32_t pbase = par_s + ((gb0 + blk_par) % kDecParBlocks) * 2048 + lane * 16;
                    uint4 rec[4];   // (lookups first, stores after: see the encoder)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool active = j0 + q < m;
                        int32_t c = ix[q];
                        if ((uint32_t)c >= (uint32_t)P.T) { if (active) st |= 1; c = 0; }
                        const uint4 mt = __ldg(meta_tab + c);  // cdf_base | lut_base | cdf_size, lut_shift | offset
                        const uint32_t maxv = (mt.z & 0xffffu) - 2u;                  // the escape symbol = last coded symbol
                        const uint32_t tbase = d_s + 2u * (mt.x + 4u * (uint32_t)c);  // address of the table's d[0]
                        uint4 o;
                        o.x = lut_s + 2u * mt.y;
                        o.y = tbase + 2u * maxv;
                        o.z = mt.w - (tbase >> 1);
                        o.w = ((mt.z >> 16) & 0xffu) | (maxv << 8) | (active ? 0x80000000u : 0u);
                        rec[q] = o;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts128(pbase + q * 512, rec[q]);
                    __syncwarp();
                    ++blk_par;
                    if (lane == 0) st_flag(ctrl + C_PARAMS, gb0 + blk_par);
                    progress = true;
                }
                // --- renormalisation words: global -> ring, as far ahead as the ring allows
                if ((int32_t)(done - gb0) >= 0) {
                    if (!serving) {
                        serving = true;
                        if (lane == 0) {
                            st_relaxed(ctrl + C_READYW, fill_u >= u_lim ? 0x7fffffffu : wp0);
                            st_release(ctrl + C_EPOCH, seq);
                        }
                        progress = true;
                    }
                    const uint32_t pub = ld_acquire(ctrl + C_WP);   // (a value left by the previous chunk lies below wp0)
                    if ((int32_t)(pub - wp_known) > 0 && pub <= wend) wp_known = pub;
                    bool filled = false;
                    while (fill_u < u_lim && fill_u + kWordGroup - (wp_known >> 1) <= (uint32_t)kWordUnits) {
#pragma unroll
                        for (int i = 0; i < kWordGroup / 32; ++i) {
                            const uint32_t u = fill_u + i * 32 + lane;
                            if (u < u_lim) cp_async4(ring_s + (u & (kWordUnits - 1)) * 4, units + u);
                        }
                        fill_u += kWordGroup;
                        filled = true;
                        if (lane < kWordGroup / 32 && fill_u + kWordGroup + lane * 32 < u_lim)   // the group after the next: into L2
                            prefetch_l2(units + fill_u + kWordGroup + lane * 32);
                    }
                    if (filled) {
                        cp_async_commit();
                        cp_async_wait<0>();
                        __syncwarp();
                        if (lane == 0) st_release(ctrl + C_READYW, fill_u >= u_lim ? 0x7fffffffu : fill_u * 2);
                        progress = true;
                    }
                }
                // --- decoded symbols of a finished block: ring -> global
                if (blk_out < nblocks && (int32_t)(done - (gb0 + blk_out)) > 0) {
                    const int j0 = blk_out * 128 + lane * 4;
                    const uint32_t obase = out_s + ((gb0 + blk_out) % kOutBlocks) * 512 + lane * 4;
                    int4 r;
                    r.x = (int32_t)lds32(obase);
                    r.y = (int32_t)lds32(obase + 128);
                    r.z = (int32_t)lds32(obase + 256);
                    r.w = (int32_t)lds32(obase + 384);
                    if (vec_ok && j0 + 3 < m) {
                        *reinterpret_cast<int4 *>(out + base + j0) = r;
                    } else {
                        if (j0 + 0 < m) out[base + j0 + 0] = r.x;
                        if (j0 + 1 < m) out[base + j0 + 1] = r.y;
                        if (j0 + 2 < m) out[base + j0 + 2] = r.z;
                        if (j0 + 3 < m) out[base + j0 + 3] = r.w;
                    }
                    __syncwarp();
                    ++blk_out;
                    if (lane == 0) st_flag(ctrl + C_STORED, gb0 + blk_out);
                    progress = true;
                }
            }
            gb += nblocks;
        }
    }
#ifdef PAIR_TIMING
    if (blockIdx.x == 0 && slot == 0 && lane == 0 && gb && is_main)
        printf("[pair decode] main: blocks %u, total %lld cycles, waiting for operands / output ring %lld, for words %lld, working %lld = %lld per block\n",
               gb, clock64() - td_start, td_wait, td_words, td_work, td_work / gb);
#endif
    if (st) atomicOr(status, st);
}

// ------------------------------------------------------------------------------------------------ encode
// scratch layout as in rans_lanes.cu: chunk k owns words [k * cap_words, (k + 1) * cap_words), filled back to front;
// outputs per chunk: first_word[k], states[k * 32 + lane].
//
// Operand record of one symbol (helper -> main): start' | m | f, escape << 31 | escape payload, with
//   m = ceil(2^32 / f): x < f << 16 after the renormalisation, so umulhi(x, m) is the quotient or one more (fixed up with the
//       sign of the remainder): two integer multiplies instead of int -> float -> int conversions on the chain;
//   f = 1: m = 2^32 - 1 gives x - 1 with remainder 1, and start' = start + 65535 makes the new state come out right;
//   a position past the end of the chunk is the identity symbol f = 2^16, start = 0 (never emits, leaves x unchanged).
__device__ inline void stg16_if(uint16_t *ptr, uint32_t v, bool p)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u16 [%0], %1;\n\t}" ::"l"(ptr), "h"((uint16_t)v), "r"((uint32_t)p) : "memory");
}

__global__ void __launch_bounds__(kPairs * 32 * (1 + kEncHelpers), 1)
k_pair_encode(LaneParams P, const uint4 *__restrict__ enc_tab, const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes,
              uint16_t *__restrict__ scratch, int cap_words, uint32_t *__restrict__ first_word, uint32_t *__restrict__ states,
              int *status)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp % kPairs;
    const bool is_main = warp < kPairs;
    const uint32_t hid = is_main ? 0u : (uint32_t)(warp / kPairs - 1);   // which of the slot's helpers
    const uint32_t ctrl = (uint32_t)__cvta_generic_to_shared(smem + slot * kCtrlBytes);
    const uint32_t par_s = (uint32_t)__cvta_generic_to_shared(smem + kPairs * kCtrlBytes + slot * kEncSlotBytes);
    if (threadIdx.x < kPairs * kCtrlBytes / 4) reinterpret_cast<uint32_t *>(smem)[threadIdx.x] = 0;
    // (no table image in shared memory here: the helpers need the 16-byte table records only -- 1 KB, L1 resident -- and the
    // precomputed operand table, both read through the read-only path)
    const uint4 *meta_tab = reinterpret_cast<const uint4 *>(P.blob);
    __syncthreads();
    const unsigned lt_mask = (1u << lane) - 1;
    const int n_chunks = P.n_chunks_dev ? *P.n_chunks_dev : P.n_chunks;
    const bool ptr_ok = ((reinterpret_cast<uintptr_t>(symbols) | reinterpret_cast<uintptr_t>(indexes)) & 15) == 0;
    int st = 0;
    uint32_t gb = 0;
#ifdef PAIR_TIMING
    long long t_wait = 0, t_work = 0;
    const long long t_start = clock64();
#endif
    for (int k = blockIdx.x + slot * gridDim.x; k < n_chunks; k += gridDim.x * kPairs) {
        uint16_t *wbuf = scratch + (size_t)k * cap_words;
        int pos = cap_words;  // warp-uniform
        uint32_t x = kRansL;
        for (int g = P.n_slices - 1; g >= 0; --g) {  // the encoder walks the chunk's symbols backwards
            const SliceDesc sd = P.slices[g];
            const long long rem = sd.n - (long long)k * sd.cs;
            if (rem <= 0) continue;
            const long long base = sd.off + (long long)k * sd.cs;
            const int m = (int)(rem < sd.cs ? rem : sd.cs);
            const int nblocks = (m + 127) >> 7;
            if (is_main) {
                // ============================================================================= main: the state's chain
                for (int blk = nblocks - 1; blk >= 0; --blk, ++gb) {
#ifdef PAIR_TIMING
                    const long long t_w0 = clock64();
#endif
                    // the ring slot carries its own sequence word: (block number + 1) | a symbol of the block escapes << 31
                    uint32_t flag = ld_acquire(ctrl + C_FLAGS + (gb % kEncParBlocks) * 4);
                    while ((flag & 0x7fffffffu) != gb + 1) {
                        flag = ld_acquire(ctrl + C_FLAGS + (gb % kEncParBlocks) * 4);
                    }
#ifdef PAIR_TIMING
                    const long long t_w1 = clock64();
                    t_wait += t_w1 - t_w0;
#endif
                    const uint32_t pbase = par_s + (gb % kEncParBlocks) * 2048 + lane * 16;
                    const uint32_t has_esc = flag >> 31;
                    uint4 Pq[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) Pq[q] = lds128(pbase + q * 512);
                    // room for everything a block can emit (4 steps x (1 symbol + 3 escape units) x 32 words): checked here,
                    // not per word; on overflow the words are dropped and the call is repeated with a worst-case buffer
                    const bool room = pos >= 4 * 4 * 32;
                    if (!room) st |= 4;
                    // the symbol itself: a word goes out when x >= f << 16, then x = (x / f << 16) + x % f + start
                    auto put = [&](const uint4 &Q, uint32_t f) {
                        const uint32_t t = x >> 16;
                        const bool emit = t >= f;
                        const unsigned em = __ballot_sync(kFull, emit);
                        pos -= room ? __popc(em) : 0;
                        stg16_if(wbuf + pos + __popc(em & lt_mask), x, emit && room);
                        x = emit ? t : x;
                        uint32_t qt = __umulhi(x, Q.y);
                        const int32_t rm = (int32_t)(x - qt * f);
                        const int32_t neg = rm >> 31;      // -1: the estimate was one too large
                        qt += (uint32_t)neg;
                        x = (qt << 16) + (uint32_t)(rm + (neg & (int32_t)f)) + Q.x;
                    };
                    if (!has_esc) {
#pragma unroll
                        for (int q = 3; q >= 0; --q) put(Pq[q], Pq[q].z);
                    } else {
#pragma unroll
                        for (int q = 3; q >= 0; --q) {
                            const bool esc = (int32_t)Pq[q].z < 0;
                            const uint32_t raw = Pq[q].w;
                            // escape units, last to first: the token list is the 36-bit string  nd | raw << 4 ; unit u = its
                            // bits [16u, 16u + 16) (a 32-bit payload has at most 8 digits: one count token)
                            if (__any_sync(kFull, esc)) {
                                const int nd = esc ? (35 - __clz(raw)) >> 2 : 0;
                                const int ntok = esc ? nd + 1 : 0;
                                const int nunits = (ntok + 3) >> 2;
                                const unsigned long long toks = (unsigned long long)nd | ((unsigned long long)raw << 4);
                                const int maxunits = (int)__reduce_max_sync(kFull, (unsigned)nunits);
                                for (int u = maxunits - 1; u >= 0; --u) {
                                    const bool part = nunits > u;
                                    const int wbits = part ? 4 * min(4, ntok - 4 * u) : 0;
                                    const uint32_t unit = (uint32_t)(toks >> (16 * u)) & 0xffffu;
                                    const bool emit = part && x >= (1u << (32 - wbits));
                                    const unsigned em = __ballot_sync(kFull, emit);
                                    pos -= room ? __popc(em) : 0;
                                    stg16_if(wbuf + pos + __popc(em & lt_mask), x, emit && room);
                                    x = emit ? x >> 16 : x;
                                    x = part ? (x << wbits) | unit : x;
                                }
                            }
                            put(Pq[q], Pq[q].z & 0x1ffffu);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) st_flag(ctrl + C_DONE, gb + 1);
#ifdef PAIR_TIMING
                    t_work += clock64() - t_w1;
#endif
                }
            } else {
                // ============================================================================= helper: operands
                const bool vec_ok = ptr_ok && (sd.off & 3) == 0;
                auto load_ops = [&](int blk, int4 &a, int4 &b) {
                    const int j0 = blk * 128 + lane * 4;
                    if (vec_ok && j0 + 3 < m) {
                        a = __ldg(reinterpret_cast<const int4 *>(symbols + base + j0));
                        b = __ldg(reinterpret_cast<const int4 *>(indexes + base + j0));
                    } else {
                        int32_t sy[4], ix[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool ok = j0 + q < m;
                            sy[q] = ok ? symbols[base + j0 + q] : 0;
                            ix[q] = ok ? indexes[base + j0 + q] : 0;
                        }
                        a = make_int4(sy[0], sy[1], sy[2], sy[3]);
                        b = make_int4(ix[0], ix[1], ix[2], ix[3]);
                    }
                };
                // this helper's blocks of the slice: every kEncHelpers-th, starting with the first whose number matches
                int4 syn = make_int4(0, 0, 0, 0), ixn = syn;
                const int first = nblocks - 1 - (int)((hid + kEncHelpers - gb % kEncHelpers) % kEncHelpers);
                if (first >= 0) load_ops(first, syn, ixn);
                for (int b = nblocks - 2 - (lane >> 3); b >= 0 && b > nblocks - 1 - kPrefetchBlocks; b -= 4)
                    prefetch_l2(((lane & 4) ? indexes : symbols) + base + b * 128 + (lane & 3) * 32);
                for (int blk = nblocks - 1; blk >= 0; --blk, ++gb) {
                    if (gb % kEncHelpers != hid) continue;
                    const int j0 = blk * 128 + lane * 4;
                    const int32_t sy[4] = {syn.x, syn.y, syn.z, syn.w}, ix[4] = {ixn.x, ixn.y, ixn.z, ixn.w};
                    if (blk >= kEncHelpers) load_ops(blk - kEncHelpers, syn, ixn);
                    if (lane < 8 && blk >= kPrefetchBlocks)   // four 128-byte lines of symbols, four of indexes
                        prefetch_l2((lane < 4 ? symbols : indexes) + base + (blk - kPrefetchBlocks) * 128 + (lane & 3) * 32);
#ifdef PAIR_TIMING
                    const long long t_h0 = clock64();
#endif
                    if (gb >= kEncParBlocks) wait_ge(ctrl + C_DONE, gb + 1 - kEncParBlocks);  // the ring slot is free again
#ifdef PAIR_TIMING
                    const long long t_h1 = clock64();
                    t_wait += t_h1 - t_h0;
#endif
                    const uint32_t pbase = par_s + (gb % kEncParBlocks) * 2048 + lane * 16;
                    bool any_esc = false;
                    uint4 rec[4];   // all four lookups first, the stores after them: a store (a compiler barrier) between two
                                    // lookups would serialise their L2 round trips
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool active = j0 + q < m;
                        int32_t c = ix[q];
                        if ((uint32_t)c >= (uint32_t)P.T) { if (active) st |= 1; c = 0; }
                        const uint4 mt = __ldg(meta_tab + c);  // cdf_base | lut_base | cdf_size, lut_shift | offset
                        const int32_t maxv = (int32_t)(mt.z & 0xffffu) - 2;
                        int32_t v = sy[q] - (int32_t)mt.w;
                        uint32_t raw = 0;
                        bool esc = false;
                        if (P.bypass) {
                            if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = maxv; }
                            else if (v >= maxv) { raw = (uint32_t)(2 * (v - maxv)); v = maxv; }
                            esc = active && v == maxv;
                        } else if (v < 0 || v > maxv) { if (active) st |= 2; v = 0; }
                        // the symbol's precomputed operands (tables.cu; 16 bytes through L2 -- the four loads of a block overlap)
                        uint4 o = __ldg(enc_tab + mt.x + (uint32_t)v);
                        o.z |= esc ? 0x80000000u : 0u;
                        o.w = raw;
                        if (!active) { o.x = 0; o.y = 65536u; o.z = 65536u; o.w = 0; }
                        any_esc |= esc;
                        rec[q] = o;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts128(pbase + q * 512, rec[q]);
                    const bool blk_esc = __any_sync(kFull, any_esc);
                    __syncwarp();
                    if (lane == 0) st_flag(ctrl + C_FLAGS + (gb % kEncParBlocks) * 4, (gb + 1) | (blk_esc ? 0x80000000u : 0u));
#ifdef PAIR_TIMING
                    t_work += clock64() - t_h1;
#endif
                }
            }
        }
        if (is_main) {
            states[(size_t)k * 32 + lane] = x;
            if (lane == 0) first_word[k] = (uint32_t)(pos < 0 ? 0 : pos);
        }
    }
#ifdef PAIR_TIMING
    if (blockIdx.x == 0 && slot == 0 && lane == 0 && gb)
        printf("[pair encode] warp %d (%s): blocks %u, total %lld cycles, waiting %lld, working %lld = %lld per own block\n", warp,
               is_main ? "main" : "helper", gb, clock64() - t_start, t_wait, t_work, t_work / (gb / (is_main ? 1 : kEncHelpers) + 1));
#endif
    if (st) atomicOr(status, st);
}

}  // namespace

static constexpr int kSmemLimit = 232448;  // opt-in dynamic shared memory of one CTA on sm_100

static int pair_smem(const RansTables &tb, int slot_bytes) { return kPairs * (kCtrlBytes + slot_bytes) + (int)tb.blob_bytes; }
static int enc_smem() { return kPairs * (kCtrlBytes + kEncSlotBytes); }
static int dec_smem(const RansTables &tb) { return kPairs * (kCtrlBytes + kDecSlotBytes) + (int)tb.blob_d_bytes; }

// The pair kernels serve the default configuration: bypass_precision 4 and a table image that fits beside the rings.
bool pair_kernels_apply(const RansTables &tb, int bypass_precision)
{
    static const bool off = [] { const char *e = getenv("BASIC_CODER_PAIRS"); return e && e[0] == '0'; }();  // A/B switch
    return !off && bypass_precision == 4 && tb.precision == 16 && tb.blob_d_bytes > 0 && dec_smem(tb) <= kSmemLimit;
}

static int pair_attrs()
{
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        BASIC_CUDA(cudaFuncSetAttribute(k_pair_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        BASIC_CUDA(cudaFuncSetAttribute(k_pair_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    }
    return BASIC_OK;
}

static int pair_grid(int n_chunks, int sm_count)
{
    int g = n_chunks < sm_count ? n_chunks : sm_count;  // chunks are dealt over CTAs first: few chunks still spread over all SMs
    return g < 1 ? 1 : g;
}

int launch_pair_encode(const RansTables &tb, const LaneParams &P, const int32_t *d_sym, const int32_t *d_idx, uint16_t *d_scratch,
                       int cap_words, uint32_t *d_first, uint32_t *d_states, int *d_status, int sm_count, cudaStream_t stream)
{
    BASIC_TRY(pair_attrs());
    k_pair_encode<<<pair_grid(P.n_chunks, sm_count), kPairs * 32 * (1 + kEncHelpers), enc_smem(), stream>>>(
        P, tb.enc.as<uint4>(), d_sym, d_idx, d_scratch, cap_words, d_first, d_states, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int launch_pair_decode(const RansTables &tb, const LaneParams &P, const unsigned char *d_seg, int64_t seg_cap, const int32_t *d_idx,
                       int seg_slices, int slice, uint32_t *d_carry_x, uint32_t *d_carry_wp, int32_t *d_out, int *d_status,
                       int sm_count, cudaStream_t stream)
{
    BASIC_TRY(pair_attrs());
    k_pair_decode<<<pair_grid(P.n_chunks, sm_count), kPairs * 64, dec_smem(tb), stream>>>(
        P, tb.blob_d.as<uint4>(), (int)tb.blob_d_bytes, (int)tb.dcdf_bytes, d_seg, seg_cap, d_idx, d_out, seg_slices, slice == 0, slice == seg_slices - 1, d_carry_x, d_carry_wp, d_status);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

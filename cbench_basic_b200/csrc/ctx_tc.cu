// Context model on the 5th-generation tensor cores: one gathered-A GEMM per layer per stage, 3xTF32.
//
// Same semantics as k_layer (ctx.cu; reference cbench/nn/layers/masked_conv.py:102-228,287-305): for the rows
// (= batch x cells of one stage and out-group) of a 128-row tile,
//      D[row, n] = sum_k A[row, k] * W[n, k]      A = masked im2col of y_hat (5x5 conv) or the visible input
//                                                 channels of the previous layer (+ prior) (1x1 layers)
// Precision: the reference is FP32 and north_star asks for means / scales within 1e-5, which plain TF32 (1e-3)
// cannot give.  Both operands are split on the fly into hi = tf32(x) and lo = x - hi and three MMAs are issued
// per k-step, hi*hi + hi*lo + lo*hi (error ~2^-21 of the product magnitude).  The tensor core's FP32 accumulation
// is measurably not round-to-nearest (error grows linearly with the chain: 1.4e-5 after 864 MMAs), so the K loop is
// cut into segments of `seg_kb` k-blocks, each accumulated in its own TMEM slot (4 slots, round-robin) and drained
// by dedicated warps into an FP32 register accumulator (round-to-nearest adds) while the next segment runs.
//
// Structure (persistent: one CTA per SM walks 128 x 128 output tiles, n-tile fastest; 320 threads):
//   warp 0 lane 0 : weight producer -- one cp.async.bulk (UBLKCP) of a pre-packed, pre-swizzled 32 KB [hi|lo]
//                   tile per k-block, completing on the stage's mbarrier
//   warp 1        : allocates TMEM; lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8),
//                   tcgen05.commit releases the stage / publishes the accumulators
//   warps 2..5    : A producers -- thread = row; gather 32 k per k-block from the NCHW activations (coalesced
//                   across the warp for a fixed channel), mask, split hi/lo, st.shared.v4 into the SWIZZLE_128B
//                   K-major image
//   warps 6..9    : drain + epilogue -- thread = row (a warp reads its own TMEM lane quarter): tcgen05.ld of every
//                   finished segment into registers, at the end bias / + prior / LeakyReLU -> NCHW store
// K-blocks no row of the launch can see (masked taps, invisible channel groups) are skipped by every role.  The
// roles only meet through mbarriers, so tile t's epilogue overlaps tile t + 1's main loop.
#include <cstdio>
#include <cstdlib>

#include "ctx.cuh"

namespace basic {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, NTHREADS = 320;
constexpr int TILE_BYTES = BM * BK * 4;      // 16 KB: one of A_hi, A_lo, B_hi, B_lo
// Operand staging.  TS = true: the A operand (hi and lo images of the gathered rows) is written by the producers
// straight into tensor memory (tcgen05.st, thread = row = TMEM lane) and tcgen05.mma reads it from there; shared
// memory only holds the weights.  That halves the shared-memory traffic per MMA, which is what bounds 3xTF32 with
// both operands in shared memory (24 KB of operand reads per 128x128x8 k-step against 128 B/cycle).
// TS = false: both operands in shared memory (SWIZZLE_128B K-major), kept as the reference variant.
template <bool TS> struct Cfg {
    static constexpr int STAGES = TS ? 4 : 3;
    static constexpr int SLOTS = TS ? 2 : 4;                          // TMEM accumulator slots of BN columns
    static constexpr int STAGE_BYTES = (TS ? 2 : 4) * TILE_BYTES;     // [A_hi | A_lo |] B_hi | B_lo
    static constexpr int B_OFF = TS ? 0 : 2 * TILE_BYTES;
    static constexpr int A_COL0 = SLOTS * BN;                         // TS: stage s keeps A_hi at A_COL0 + 64 s, A_lo 32 further
};
constexpr int MAX_STAGES = 4, MAX_SLOTS = 4;
constexpr int MAX_KB = 512;                  // k-blocks a CTA may walk (conv: 25 taps x ceil(C / 32))
constexpr int MAX_G = 8;
constexpr float kSlope = 0.01f;              // nn.LeakyReLU default negative_slope

// shared memory map (offsets from the 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_BARS = 3 * 4 * TILE_BYTES;            // stages: 192 KB (SS) / 128 KB (TS); then full[], empty[], slot_full[], slot_empty[]
constexpr int OFF_TMEM = OFF_BARS + 8 * (2 * MAX_STAGES + 2 * MAX_SLOTS);
constexpr int OFF_NKB = OFF_TMEM + 4;
constexpr int OFF_LIST = ((OFF_NKB + 4 + 15) / 16) * 16;         // uint4 [MAX_KB]
constexpr int OFF_MASK = OFF_LIST + 16 * MAX_KB;        // uint32 [MAX_G][BM]
constexpr int SMEM_BYTES = OFF_MASK + 4 * MAX_G * BM + 1024 /* alignment slack */;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ inline uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ inline void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ inline void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ inline void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ inline void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ inline void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ inline void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ inline void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ inline void tmem_alloc(uint32_t slot_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ inline void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ inline void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ inline void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ inline void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ inline void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ inline void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ inline void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO (unused
// for swizzled K-major, 1) | SBO = 1024 B between 8-row groups | version 1 | layout type 2 (SWIZZLE_128B).
__device__ inline uint64_t smem_desc(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ inline uint32_t tf32_hi(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// byte offset of 16-byte chunk j (k = 4j .. 4j+3) of row r inside a [128 rows x 128 B] SWIZZLE_128B tile
__device__ __host__ inline int swz(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }

// ------------------------------------------------------------------------------------------------- the kernel
template <bool TS>
__global__ void __launch_bounds__(NTHREADS, 1)
k_layer_tc(LayerArgs a)
{
    constexpr int STAGES = Cfg<TS>::STAGES, SLOTS = Cfg<TS>::SLOTS, STAGE_BYTES = Cfg<TS>::STAGE_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
    int *s_nkb = reinterpret_cast<int *>(smem + OFF_NKB);
    // k-block list entry: x = weight k-block index; y = element offset of the block's first channel from the row's
    // base pointer (c0 * HW + tap shift); z = source (bit 0) | tap << 1 | valid 4-channel chunks << 8 | grouped << 16;
    // w = visibility group of each of the 8 chunks, 4 bits each
    uint4 *s_list = reinterpret_cast<uint4 *>(smem + OFF_LIST);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + OFF_MASK);  // [MAX_G][BM]: a producer thread's private row masks
    auto bar_full = [&](int s) { return s_base + OFF_BARS + 8 * s; };
    auto bar_empty = [&](int s) { return s_base + OFF_BARS + 8 * (MAX_STAGES + s); };
    auto bar_slot_full = [&](int q) { return s_base + OFF_BARS + 8 * (2 * MAX_STAGES + q); };
    auto bar_slot_empty = [&](int q) { return s_base + OFF_BARS + 8 * (2 * MAX_STAGES + MAX_SLOTS + q); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = a.B * a.ncells;
    const int G = a.G, HW = a.HW;
    const int seg_kb = a.nacc;                         // k-blocks per TMEM accumulation segment
    constexpr uint32_t tmem_cols = 512;                // all of this SM's tensor memory (1 CTA / SM)
    // persistent: tiles are dealt round-robin, n-tile fastest, so the CTAs gathering the same rows run together (L2)
    const int n_ntiles = (a.n_count + BN - 1) / BN;
    const int n_tiles = n_ntiles * ((rows + BM - 1) / BM);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 128 + 1);  // 128 A-producer threads + the weight producer's expect_tx arrive
            mbar_init(bar_empty(s), 1);       // one tcgen05.commit
        }
        for (int q = 0; q < SLOTS; ++q) {
            mbar_init(bar_slot_full(q), 1);     // one tcgen05.commit
            mbar_init(bar_slot_empty(q), 128);  // the 128 drain threads
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), tmem_cols);
    // ---- warp 0: the k-block list every role walks, from the launch-level visibility (a.vis_or): k-blocks no row
    // of this (stage, out-group) can see -- masked taps, invisible channel groups -- are skipped by everybody
    if (warp == 0) {
        int count = 0;
        const int total = a.is_conv ? a.ksize * a.ksize * a.kb_src0 : a.kb_total;
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            bool vis = false;
            uint4 e = make_uint4(0, 0, 0, 0);
            if (i < total) {
                int src = 0, tap = 0, c0, nch, groups, shift = 0;
                if (a.is_conv) {
                    const int kbt = a.kb_src0 /* k-blocks per tap */, pad = a.ksize / 2;
                    tap = i / kbt;
                    c0 = (i - tap * kbt) * BK;
                    nch = a.Cin;
                    groups = G;
                    shift = (tap / a.ksize - pad) * a.W_img + (tap % a.ksize - pad);
                } else {
                    src = i >= a.kb_src0;
                    c0 = (src ? i - a.kb_src0 : i) * BK;
                    const Source sc = src ? a.src1 : a.src0;
                    nch = sc.channels;
                    groups = sc.groups;
                }
                const int nvalid = min(BK / 4, (nch - c0) / 4);
                uint32_t gids = 0;
                if (groups == 0) vis = true;
                else {
                    const int cpg = nch / groups;
                    for (int j = 0; j < nvalid; ++j) {
                        const int g = (c0 + 4 * j) / cpg;
                        gids |= (uint32_t)g << (4 * j);
                        vis |= a.is_conv ? ((a.vis_or[g] >> tap) & 1u) : ((a.vis_or[0] >> g) & 1u);
                    }
                }
                const bool src_cl = a.is_conv ? a.src0.cl : (src ? a.src1.cl : a.src0.cl);
                const int rel = src_cl ? shift * nch + c0 : c0 * HW + shift;   // from the row's base pointer, in elements
                e = make_uint4((uint32_t)i, (uint32_t)rel, (uint32_t)src | ((uint32_t)tap << 1) | ((uint32_t)nvalid << 8) |
                               (groups ? 1u << 16 : 0u) | (src_cl ? 1u << 17 : 0u), gids);
            }
            const unsigned m = __ballot_sync(0xffffffffu, vis);
            if (vis) s_list[count + __popc(m & ((1u << lane) - 1))] = e;
            count += __popc(m);
        }
        if (lane == 0) *s_nkb = count;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_kb = *s_nkb;
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ================================================================================ weight producer
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int nt = t % n_ntiles;
                const unsigned char *wsrc = a.wpack + (size_t)(a.ntile_base + nt) * a.kb_total * (2 * TILE_BYTES);
                for (int i = 0; i < n_kb; ++i, ++it) {
                    const int s = it % STAGES, round = it / STAGES;
                    mbar_wait(bar_empty(s), (round & 1) ^ 1);
                    if (a.debug & 2) { mbar_arrive(bar_full(s)); continue; }
                    mbar_arrive_expect_tx(bar_full(s), 2 * TILE_BYTES);
                    bulk_g2s(s_base + OFF_STAGES + s * STAGE_BYTES + Cfg<TS>::B_OFF, wsrc + (size_t)s_list[i].x * (2 * TILE_BYTES),
                             2 * TILE_BYTES, bar_full(s));
                }
            }
        }
    } else if (warp == 1) {
        // ==================================================================================== MMA issuer
        if (lane == 0) {
            int it = 0, segg = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int i = 0; i < n_kb; ++i, ++it) {
                    const int s = it % STAGES, round = it / STAGES;
                    const int slot = segg % SLOTS;
                    const bool seg_first = i % seg_kb == 0, seg_last = (i % seg_kb == seg_kb - 1) || i == n_kb - 1;
                    if (seg_first) {  // the slot must have been drained (fresh barrier: passes)
                        mbar_wait(bar_slot_empty(slot), ((segg / SLOTS) & 1) ^ 1);
                        tc_fence_after();
                    }
                    mbar_wait(bar_full(s), round & 1);
                    tc_fence_after();
                    const uint32_t st = s_base + OFF_STAGES + s * STAGE_BYTES;
                    const uint32_t d = tmem_base + (uint32_t)(slot * BN);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t bh = smem_desc(st + Cfg<TS>::B_OFF + ks * 32), bl = smem_desc(st + Cfg<TS>::B_OFF + TILE_BYTES + ks * 32);
                        const uint32_t acc0 = (seg_first && ks == 0) ? 0u : 1u;
                        if constexpr (TS) {
                            const uint32_t ah = tmem_base + (uint32_t)(Cfg<TS>::A_COL0 + s * 64 + ks * 8), al = ah + 32;
                            umma_tf32_ts(d, ah, bh, kIdesc, acc0);
                            umma_tf32_ts(d, ah, bl, kIdesc, 1u);
                            umma_tf32_ts(d, al, bh, kIdesc, 1u);
                        } else {
                            const uint64_t ah = smem_desc(st + ks * 32), al = smem_desc(st + TILE_BYTES + ks * 32);
                            umma_tf32(d, ah, bh, kIdesc, acc0);
                            umma_tf32(d, ah, bl, kIdesc, 1u);
                            umma_tf32(d, al, bh, kIdesc, 1u);
                        }
                    }
                    umma_commit(bar_empty(s));  // frees the stage once the MMAs above have read it
                    if (seg_last) {
                        umma_commit(bar_slot_full(slot));  // segment complete -> drain warps
                        ++segg;
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ================================================================================== A producers
        // this thread's row of the tile; with A in tensor memory a warp can only write TMEM lanes 32 * (warp % 4) .. + 31
        const int r = TS ? (warp & 3) * 32 + lane : tid - 64;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        // shared-space addresses of this row's eight 16-byte chunks inside a SWIZZLE_128B tile
        uint32_t chunk_off[BK / 4];
#pragma unroll
        for (int j = 0; j < BK / 4; ++j) chunk_off[j] = (uint32_t)swz(r, j);
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (n_kb == 0) break;
            // ---- per-row facts of this tile
            const int row = (t / n_ntiles) * BM + r;
            int rb = -1, rhw = 0;
            uint32_t rgrp = 0, row_tap0 = 0;
            if (row < rows) {
                rb = row / a.ncells;
                const int cell = a.cell_base + (row - rb * a.ncells);
                rhw = a.cell_hw[cell];
                if (a.is_conv) {
                    if (G == 1) row_tap0 = a.cell_tap[cell];
                    else for (int j = 0; j < G; ++j) s_mask[j * BM + r] = a.cell_tap[(size_t)cell * G + j];  // private to this thread
                } else {
                    rgrp = a.cell_grp[cell];
                }
            }
            // per-row base pointers (an invalid row never loads: its visibility mask is forced to 0)
            const long long rbz = rb < 0 ? 0 : rb;
            const int nch0 = a.is_conv ? a.Cin : a.src0.channels;
            const float *row0p = a.src0.cl ? a.src0.ptr + (rbz * HW + rhw) * nch0 : a.src0.ptr + rbz * nch0 * HW + rhw;
            const float *row1p = !a.src1.ptr ? nullptr
                                 : a.src1.cl ? a.src1.ptr + (rbz * HW + rhw) * a.src1.channels
                                             : a.src1.ptr + rbz * a.src1.channels * HW + rhw;
            // gathers the 32 k of k-block i for this thread's row into registers (masked elements = 0)
            auto gather = [&](int i, float(&v)[BK]) {
                const uint4 e = s_list[i];
                const float *p = (e.z & 1u) ? row1p : row0p;
                const int tap = (int)((e.z >> 1) & 31u), nvalid = (int)((e.z >> 8) & 15u);
                uint32_t vis;  // bit j: chunk j (channels c0 + 4j .. + 3) is loaded
                if (!((e.z >> 16) & 1u)) vis = 0xffu;                                  // ungrouped source: always visible
                else if (G == 1) vis = a.is_conv ? (((row_tap0 >> tap) & 1u) ? 0xffu : 0u) : ((rgrp & 1u) ? 0xffu : 0u);
                else {
                    vis = 0;
#pragma unroll
                    for (int j = 0; j < BK / 4; ++j) {
                        const uint32_t g = (e.w >> (4 * j)) & 15u;
                        const uint32_t bit = a.is_conv ? ((s_mask[g * BM + r] >> tap) & 1u) : ((rgrp >> g) & 1u);
                        vis |= bit << j;
                    }
                }
                vis &= (1u << nvalid) - 1u;
                if (rb < 0 || (a.debug & 1)) vis = 0;
                int idx = (int)e.y;
                if ((e.z >> 17) & 1u) {  // channels-last source: the 32 channels of the block are contiguous
                    const float4 *p4 = reinterpret_cast<const float4 *>(p + idx);
#pragma unroll
                    for (int j = 0; j < BK / 4; ++j) {
                        const float4 t = ((vis >> j) & 1u) ? __ldg(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < BK / 4; ++j) {
                        const bool ok = (vis >> j) & 1u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            v[4 * j + q] = ok ? __ldg(p + idx) : 0.f;
                            idx += HW;
                        }
                    }
                }
            };
            // splits into hi = tf32(v), lo = v - hi and stores both SWIZZLE_128B images of the stage
            auto publish = [&](const float(&v)[BK]) {
                const int s = it % STAGES, round = it / STAGES;
                ++it;
                mbar_wait(bar_empty(s), (round & 1) ^ 1);
                if constexpr (TS) {
                    tc_fence_after();  // the MMAs that read this stage's TMEM columns have completed (tcgen05.commit)
                    uint32_t h[BK];
#pragma unroll
                    for (int q = 0; q < BK; ++q) h[q] = tf32_hi(v[q]);
                    tmem_st32(lane_base + (uint32_t)(Cfg<TS>::A_COL0 + s * 64), h);
#pragma unroll
                    for (int q = 0; q < BK; ++q) h[q] = __float_as_uint(v[q] - __uint_as_float(h[q]));
                    tmem_st32(lane_base + (uint32_t)(Cfg<TS>::A_COL0 + s * 64 + 32), h);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(bar_full(s));
                    return;
                }
                const uint32_t st = s_base + OFF_STAGES + s * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < BK / 4; ++j) {
                    const uint32_t h0 = tf32_hi(v[4 * j + 0]), h1 = tf32_hi(v[4 * j + 1]), h2 = tf32_hi(v[4 * j + 2]), h3 = tf32_hi(v[4 * j + 3]);
                    const float l0 = v[4 * j + 0] - __uint_as_float(h0), l1 = v[4 * j + 1] - __uint_as_float(h1);
                    const float l2 = v[4 * j + 2] - __uint_as_float(h2), l3 = v[4 * j + 3] - __uint_as_float(h3);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + chunk_off[j]), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st + TILE_BYTES + chunk_off[j]), "f"(l0), "f"(l1), "f"(l2), "f"(l3) : "memory");
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                mbar_arrive(bar_full(s));
            };
            // software pipeline: the loads of k-block i + 1 are in flight while k-block i is split and stored
            float va[BK], vb[BK];
            gather(0, va);
            for (int i = 0; i < n_kb; i += 2) {
                if (i + 1 < n_kb) gather(i + 1, vb);
                publish(va);
                if (i + 1 < n_kb) {
                    if (i + 2 < n_kb) gather(i + 2, va);
                    publish(vb);
                }
            }
        }
    } else {
        // ============================================================================= drain + epilogue
        const int r = (warp & 3) * 32 + lane;  // a warp may only read TMEM lanes 32 * (warp % 4) .. + 31
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int n_seg = (n_kb + seg_kb - 1) / seg_kb;
        const long long ohw = (long long)a.Ntot * HW;
        int segg = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int nt = t % n_ntiles;
            const int row = (t / n_ntiles) * BM + r;
            int rb = -1, rhw = 0;
            if (row < rows) {
                rb = row / a.ncells;
                rhw = a.cell_hw[a.cell_base + (row - rb * a.ncells)];
            }
            float acc[BN];
#pragma unroll
            for (int q = 0; q < BN; ++q) acc[q] = 0.f;
            for (int seg = 0; seg < n_seg; ++seg, ++segg) {
                const int slot = segg % SLOTS;
                mbar_wait(bar_slot_full(slot), (segg / SLOTS) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < BN / 32; ++cc) {
                    uint32_t tt[32];
                    tmem_ld32(lane_base + (uint32_t)(slot * BN + cc * 32), tt);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[cc * 32 + q] += __uint_as_float(tt[q]);
                }
                tc_fence_before();
                mbar_arrive(bar_slot_empty(slot));
            }
            if (rb >= 0 && a.out_cl) {
                // channels-last: this row's 128 channels are contiguous -> 128-bit stores, full sectors
                const long long o0 = ((long long)rb * HW + rhw) * a.Ntot + a.n_begin + nt * BN;
#pragma unroll
                for (int q = 0; q < BN; q += 4) {
                    if (nt * BN + q < a.n_count) {  // n_count % 4 == 0 on this path
                        float4 val = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
                        if (a.bias) {
                            const float4 bv = __ldg(reinterpret_cast<const float4 *>(a.bias + a.n_begin + nt * BN + q));
                            val.x += bv.x; val.y += bv.y; val.z += bv.z; val.w += bv.w;
                        }
                        if (a.add) {
                            const float4 av = __ldg(reinterpret_cast<const float4 *>(a.add + o0 + q));
                            val.x += av.x; val.y += av.y; val.z += av.z; val.w += av.w;
                        }
                        if (a.lrelu) {
                            val.x = val.x > 0.f ? val.x : val.x * kSlope; val.y = val.y > 0.f ? val.y : val.y * kSlope;
                            val.z = val.z > 0.f ? val.z : val.z * kSlope; val.w = val.w > 0.f ? val.w : val.w * kSlope;
                        }
                        *reinterpret_cast<float4 *>(a.out + o0 + q) = val;
                    }
                }
            } else if (rb >= 0) {
                const long long obase = (long long)rb * ohw + rhw;
#pragma unroll
                for (int q = 0; q < BN; ++q) {
                    const int n = nt * BN + q;
                    if (n < a.n_count) {
                        const int ch = a.n_begin + n;
                        const long long o = obase + (long long)ch * HW;
                        float val = acc[q] + (a.bias ? __ldg(a.bias + ch) : 0.f);
                        if (a.add) val += __ldg(a.add + o);
                        if (a.lrelu) val = val > 0.f ? val : val * kSlope;
                        a.out[o] = val;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();  // lane 0 ran the issue loop alone: reconverge before the .sync.aligned dealloc
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Packs weights into per-(out-group, n-tile, k-block) [hi | lo] SWIZZLE_128B images.
//   conv : w [N][Cin][k2]            K' = tap * Cpad + c            (Cpad = Cin rounded up to 32)
//   dense: w [N][c_src0 + c_src1]    K' = [src0 padded to 32 | src1 padded to 32]
__global__ void __launch_bounds__(256)
k_pack_w_tc(const float *__restrict__ w, unsigned char *__restrict__ out, int N, int G, int ntiles, int kb_total, int is_conv,
            int Cin, int k2, int c_src0, int c_src1, int kb_src0)
{
    const int n_count = N / G;
    const long long total = (long long)G * ntiles * kb_total * BN * BK;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(e % BK);
        const int nl = (int)((e / BK) % BN);
        const long long tkb = e / (BK * BN);
        const int wkb = (int)(tkb % kb_total);
        const int tile = (int)(tkb / kb_total);
        const int og = tile / ntiles, ntile = tile - og * ntiles;
        const int n_in = ntile * BN + nl;
        float v = 0.f;
        if (n_in < n_count) {
            const int n = og * n_count + n_in;
            if (is_conv) {
                const int tap = wkb / kb_src0, c = (wkb - tap * kb_src0) * BK + kk;
                if (c < Cin) v = w[((size_t)n * Cin + c) * k2 + tap];
            } else {
                const int Korig = c_src0 + c_src1;
                if (wkb < kb_src0) {
                    const int c = wkb * BK + kk;
                    if (c < c_src0) v = w[(size_t)n * Korig + c];
                } else {
                    const int c = (wkb - kb_src0) * BK + kk;
                    if (c < c_src1) v = w[(size_t)n * Korig + c_src0 + c];
                }
            }
        }
        const uint32_t hi = tf32_hi(v);
        const float lo = v - __uint_as_float(hi);
        unsigned char *t = out + (size_t)tkb * (2 * TILE_BYTES);
        const int o = swz(nl, kk >> 2) + (kk & 3) * 4;
        *reinterpret_cast<uint32_t *>(t + o) = hi;
        *reinterpret_cast<float *>(t + TILE_BYTES + o) = lo;
    }
}

}  // namespace

int pack_weights_tc(PackedW &dst, const float *w_dev, int N, int G, int is_conv, int Cin, int k2, int c_src0, int c_src1,
                    cudaStream_t stream)
{
    const int n_count = N / G;
    dst.ntiles_per_group = (n_count + BN - 1) / BN;
    if (is_conv) {
        dst.kb_src0 = (Cin + BK - 1) / BK;
        dst.kb_total = k2 * dst.kb_src0;
    } else {
        dst.kb_src0 = (c_src0 + BK - 1) / BK;
        dst.kb_total = dst.kb_src0 + (c_src1 + BK - 1) / BK;
    }
    const size_t bytes = (size_t)G * dst.ntiles_per_group * dst.kb_total * (2 * TILE_BYTES);
    BASIC_TRY(dst.buf.reserve(bytes));
    k_pack_w_tc<<<1024, 256, 0, stream>>>(w_dev, dst.buf.as<unsigned char>(), N, G, dst.ntiles_per_group, dst.kb_total, is_conv,
                                          Cin, k2, c_src0, c_src1, dst.kb_src0);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// The tensor path takes every layer of every stage of a model or none (its activations are channels-last).  It
// needs 4-channel chunks that never straddle a visibility group, at most MAX_G groups, a k-block list that fits,
// and stages big enough for 128-row MMA tiles to make sense (scanline-like maps stay on the exact-FP32 kernel).
bool tc_model_eligible(const CtxModel &m, int B)
{
    if (m.precision != BASIC_CTX_TF32X3 || !m.has_conv || m.S < 1) return false;
    if (m.G > MAX_G || m.k * m.k * ((m.C + BK - 1) / BK) > MAX_KB || m.k * m.k > 31) return false;
    auto ok4 = [&](int channels) { return channels % m.G == 0 && (channels / m.G) % 4 == 0; };
    if (!ok4(m.C) || !ok4(m.c_ctx)) return false;
    if (m.has_merger && (!ok4(m.c_m1) || !ok4(m.c_m2))) return false;
    return (long long)B * m.G * m.H * m.W / m.S >= 64;
}

// [B, channels, HW] -> [B, HW, channels], 32 x 32 tiles through shared memory (both sides coalesced)
__global__ void __launch_bounds__(256)
k_nchw_to_cl(const float *__restrict__ src, float *__restrict__ dst, int channels, int HW)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *s = src + (size_t)b * channels * HW;
    float *d = dst + (size_t)b * channels * HW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, hw = hw0 + tx;
        tile[ty + 8 * i][tx] = (c < channels && hw < HW) ? s[(size_t)c * HW + hw] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hw = hw0 + ty + 8 * i, c = c0 + tx;
        if (c < channels && hw < HW) d[(size_t)hw * channels + c] = tile[tx][ty + 8 * i];
    }
}

int launch_nchw_to_cl(const float *src, float *dst, int B, int channels, int HW, cudaStream_t stream)
{
    if (B == 0) return BASIC_OK;
    dim3 grid((HW + 31) / 32, (channels + 31) / 32, B);
    k_nchw_to_cl<<<grid, 256, 0, stream>>>(src, dst, channels, HW);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int launch_layer_tc(const CtxModel &m, const LayerArgs &a, cudaStream_t stream)
{
    const int rows = a.B * a.ncells;
    if (rows == 0 || a.n_count == 0) return BASIC_OK;
    static bool attr_done = false;
    static const bool ts = !(getenv("BASIC_TC_SS") && atoi(getenv("BASIC_TC_SS")));
    if (!attr_done) {
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_done = true;
    }
    const long long tiles = (long long)((a.n_count + BN - 1) / BN) * ((rows + BM - 1) / BM);
    dim3 grid((unsigned)(tiles < m.sm_count ? tiles : m.sm_count));  // persistent: one CTA per SM
    LayerArgs b = a;
    static const int dbg = getenv("BASIC_TC_DEBUG") ? atoi(getenv("BASIC_TC_DEBUG")) : 0;
    b.debug = dbg;
    if (ts) k_layer_tc<true><<<grid, NTHREADS, SMEM_BYTES, stream>>>(b);
    else k_layer_tc<false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(b);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

}  // namespace basic

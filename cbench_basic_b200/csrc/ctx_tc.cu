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
// Structure (persistent: one CTA per SM walks 128 x 128 output tiles, n-tile fastest; 512 threads = 4 warpgroups,
// registers moved between them with setmaxnreg):
//   WG0 warp 0 lane 0 : weight producer -- one cp.async.bulk (UBLKCP) of a pre-packed, pre-swizzled 32 KB [hi|lo]
//                       tile per k-block, completing on the stage's mbarrier
//       warp 1        : allocates TMEM; lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8, A from
//                       TMEM), tcgen05.commit releases the stage / publishes the accumulators
//   WG1, WG2          : A producers, alternating k-blocks -- thread = row = TMEM lane; gather 32 k per k-block from
//                       the channels-last activations (128 contiguous bytes), mask, split hi/lo, tcgen05.st.  One
//                       k-block costs a producer ~1500 cycles of dependent latency (barrier, split, TMEM store round
//                       trip, next gather) against 768 cycles of MMA, hence two groups
//   WG3               : drain + epilogue -- thread = row (a warp reads its own TMEM lane quarter): tcgen05.ld of every
//                       finished segment into registers; at the end bias / LeakyReLU -> padded shared-memory row ->
//                       one cp.async.bulk store per row (channels-last), or strided stores (+ prior) for NCHW
// K-blocks no row of the launch can see (masked taps, invisible channel groups) are skipped by every role.  The
// roles only meet through mbarriers, so tile t's epilogue overlaps tile t + 1's main loop.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ctx.cuh"

namespace basic {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, NTHREADS = 512;
// Operand staging: the A operand (hi and lo images of the gathered rows) is written by the producers straight into
// tensor memory (tcgen05.st, thread = row = TMEM lane) and tcgen05.mma reads it from there; shared memory only
// holds the weights.  With both operands in shared memory 3xTF32 is shared-memory bound (24 KB of operand reads
// per 128x128x8 k-step against 128 B/cycle = 2x the MMA time).
//
// Two arithmetic modes, same structure (every product = hi*hi + hi*lo + lo*hi, ~22 significant bits):
//   MODE 0  3xTF32: hi = tf32(x), lo = x - hi; kind::tf32, K = 8 per MMA, 4-byte operands.
//   MODE 1  3xFP16: operands scaled by powers of two (activations x 16, weights so that max |w| lands in [2^12, 2^13)),
//           hi = fp16(X), lo = fp16(X - hi); kind::f16, K = 16 per MMA, 2-byte operands: half the MMA instructions,
//           half the operand bytes, twice the stages.  |activation| >= 4000 would leave the fp16 range: the producers
//           raise a flag and the y-path driver repeats the call in MODE 0 (the container records the mode).
template <int MODE> struct Cfg {
    static constexpr int ESIZE = MODE ? 2 : 4;                 // operand bytes
    static constexpr int TILE_BYTES = BM * BK * ESIZE;         // one of B_hi, B_lo: 16 KB / 8 KB
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;         // B_hi | B_lo
    static constexpr int STAGES = 4;                           // operand ring: B in shared memory, A in TMEM columns
    static constexpr int SLOTS = MODE ? 3 : 2;                 // TMEM accumulator slots of BN columns (512 columns in all)
    static constexpr int A_COL0 = SLOTS * BN;                  // stage s keeps A_hi at column A_COL0 + 2 A_COLS s, A_lo A_COLS further
    static constexpr int A_COLS = BK * ESIZE / 4;              // TMEM columns of one A image (hi or lo) of a k-block
    static constexpr int KSTEPS = MODE ? BK / 16 : BK / 8;     // MMAs (x 3) per k-block
    // cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2) / F16 (0) at bits 7-9, 10-12, K-major both,
    // N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t IDESC = (1u << 4) | ((MODE ? 0u : 2u) << 7) | ((MODE ? 0u : 2u) << 10) | ((uint32_t)(BN >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
};
constexpr int MAX_STAGES = 8;
constexpr int STAGE_REGION = 128 * 1024;     // >= STAGES * STAGE_BYTES of either mode
constexpr int MAX_SLOTS = 3;
constexpr float kActScale = 16.f;            // MODE 1: activations are multiplied by this before the fp16 split
constexpr float kActLimit = 4000.f;          // ... and must stay below this in magnitude
constexpr int NGROUPS = 2;                   // A-producer warpgroups, alternating k-blocks
constexpr int MAX_KB = 512;                  // k-blocks a CTA may walk (conv: 25 taps x ceil(C / 32))
constexpr int MAX_G = 8;
constexpr float kSlope = 0.01f;              // nn.LeakyReLU default negative_slope
constexpr int OUT_ROW = BN * 4 + 16;         // bytes between rows of the epilogue staging tile (16 B pad: conflict-free v4 stores)

// shared memory map (offsets from the 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_OUT = STAGE_REGION;                    // epilogue staging: BM rows of OUT_ROW bytes
constexpr int OFF_BARS = OFF_OUT + BM * OUT_ROW;         // full[MAX_STAGES], empty[MAX_STAGES], slot_full[SLOTS], slot_empty[SLOTS]
constexpr int OFF_TMEM = OFF_BARS + 8 * (2 * MAX_STAGES + 2 * MAX_SLOTS);
constexpr int OFF_NKB = OFF_TMEM + 4;
constexpr int OFF_LIST = ((OFF_NKB + 4 + 15) / 16) * 16;         // uint4 [MAX_KB]
constexpr int OFF_MASK = OFF_LIST + 16 * MAX_KB;        // uint32 [NGROUPS][MAX_G][BM]: a producer thread's private row masks
constexpr int MAX_BIAS = 2048;                          // output channels of one launch (n_count) the bias cache holds
constexpr int OFF_BIAS = OFF_MASK + 4 * NGROUPS * MAX_G * BM;     // float [MAX_BIAS]: bias[n_begin ..] (zeros without a bias)
constexpr int SMEM_BYTES = OFF_BIAS + 4 * MAX_BIAS + 1024 /* alignment slack */;
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");

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
template <int MODE>
__device__ inline void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (MODE == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
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
__device__ inline void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
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
__device__ inline bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int R> __device__ inline void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ inline void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ inline void bulk_s2g(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ inline void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ inline void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ inline void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO (unused for swizzled
// K-major, 1) | SBO = bytes between 8-row groups >> 4 | version 1 | layout type.  MODE 0: 128-byte rows, SWIZZLE_128B
// (type 2), SBO 1024; MODE 1: 64-byte rows, SWIZZLE_64B (type 4), SBO 512.
template <int MODE> __device__ inline uint64_t smem_desc(uint32_t addr)
{
    constexpr uint64_t sbo = MODE ? 512 : 1024, type = MODE ? 4 : 2;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (type << 61);
}

__device__ inline uint32_t tf32_hi(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// byte offset of 16-byte chunk j of row r inside a [128 rows x 128 B] SWIZZLE_128B / [128 rows x 64 B] SWIZZLE_64B tile
__device__ __host__ inline int swz(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }
__device__ __host__ inline int swz64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }
// two floats -> packed fp16 pair (lo in the low half), and back
__device__ inline uint32_t pack_h2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// Eight consecutive channels -> the two 16-byte chunks of the 3xFP16 activation format (split16): X = 16 x,
// hi = fp16(X), lo = fp16(X - hi); chunk 2m = {hi(c0,c1), hi(c2,c3), hi(c4,c5), hi(c6,c7)}, chunk 2m + 1 = the same of lo.
// The words are the TMEM columns the MMA reads and a k-block's eight chunk loads land in column order (hi 0..15, lo 0..15),
// so a producer moves them from global memory to tensor memory with no instruction in between.  amax: max |x| (range flag).
__device__ inline float h_lo(uint32_t p);
__device__ inline float h_hi(uint32_t p);
__device__ inline void split16(float4 x0, float4 x1, float &amax, uint4 &hi, uint4 &lo)
{
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x0.x), fabsf(x0.y)), fmaxf(fabsf(x0.z), fabsf(x0.w))));
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x1.x), fabsf(x1.y)), fmaxf(fabsf(x1.z), fabsf(x1.w))));
    const float a = x0.x * kActScale, b = x0.y * kActScale, c = x0.z * kActScale, d = x0.w * kActScale;
    const float e = x1.x * kActScale, f = x1.y * kActScale, g = x1.z * kActScale, h = x1.w * kActScale;
    hi.x = pack_h2(a, b); hi.y = pack_h2(c, d); hi.z = pack_h2(e, f); hi.w = pack_h2(g, h);
    lo.x = pack_h2(a - h_lo(hi.x), b - h_hi(hi.x));
    lo.y = pack_h2(c - h_lo(hi.y), d - h_hi(hi.y));
    lo.z = pack_h2(e - h_lo(hi.z), f - h_hi(hi.z));
    lo.w = pack_h2(g - h_lo(hi.w), h - h_hi(hi.w));
}
__device__ inline float h_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xffffu))); }
__device__ inline float h_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }

// ------------------------------------------------------------------------------------------------- the kernel
// DBG = 1: the instrumented build (BASIC_TC_DEBUG switches, BASIC_TC_TIMELINE stamps); the production build has none of it
template <int MODE, int DBG>
__global__ void __launch_bounds__(NTHREADS, 1)
k_layer_tc(LayerArgs a)
{
    const int dbg = DBG ? a.debug : 0;                    // (constant 0 / NULL in the production build: every debug and
    long long *const tline = DBG ? a.timeline : nullptr;  //  timeline branch folds away)
    constexpr int STAGES = Cfg<MODE>::STAGES, STAGE_BYTES = Cfg<MODE>::STAGE_BYTES, TILE_BYTES = Cfg<MODE>::TILE_BYTES;
    constexpr int A_COLS = Cfg<MODE>::A_COLS, SLOTS = Cfg<MODE>::SLOTS, A_COL0 = Cfg<MODE>::A_COL0;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
    // k-block list entry: x = weight k-block index; y = element offset of the block's first channel from the row's
    // base pointer (c0 * HW + tap shift); z = source (bit 0) | tap << 1 | valid 4-channel chunks << 8 | grouped << 16;
    // w = visibility group of each of the 8 chunks, 4 bits each
    uint4 *s_list = reinterpret_cast<uint4 *>(smem + OFF_LIST);
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    auto bar_full = [&](int s) { return s_base + OFF_BARS + 8 * s; };
    auto bar_empty = [&](int s) { return s_base + OFF_BARS + 8 * (MAX_STAGES + s); };
    auto bar_slot_full = [&](int q) { return s_base + OFF_BARS + 8 * (2 * MAX_STAGES + q); };
    auto bar_slot_empty = [&](int q) { return s_base + OFF_BARS + 8 * (2 * MAX_STAGES + MAX_SLOTS + q); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const int rows = a.B * a.ncells;
    const int G = a.G, HW = a.HW;
    const int seg_kb = a.nacc;                         // k-blocks per TMEM accumulation segment
    constexpr uint32_t tmem_cols = 512;                // all of this SM's tensor memory (1 CTA / SM)
    // persistent: tiles are dealt round-robin, n-tile fastest, so the CTAs gathering the same rows run together (L2)
    const int n_ntiles = (a.n_count + BN - 1) / BN;
    const int n_tiles = n_ntiles * ((rows + BM - 1) / BM);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 128 + 1);  // the 128 threads of one A-producer group + the weight producer's expect_tx arrive
            mbar_init(bar_empty(s), 1);       // one tcgen05.commit
        }
        for (int q = 0; q < SLOTS; ++q) {
            mbar_init(bar_slot_full(q), 1);     // one tcgen05.commit
            mbar_init(bar_slot_empty(q), 128);  // the 128 drain threads
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), tmem_cols);
    // the epilogue reads its bias from shared memory (a dependent global load per stored vector serialised it)
    if (wg == 3)
        for (int i = tid - 384; i < ((a.n_count + 3) & ~3); i += 128) s_bias[i] = (a.bias && i < a.n_count) ? __ldg(a.bias + a.n_begin + i) : 0.f;
    // ---- the k-block list every role walks (built on the host, launch_layer_tc): k-blocks no row of this
    // (stage, out-group) can see -- masked taps, invisible channel groups -- are not in it
    for (int i = tid; i < a.n_kb; i += NTHREADS) s_list[i] = a.kb_list[i];
    // programmatic dependent launch: the next layer's kernel may start its own prologue as soon as SMs free up, and this one
    // waits here -- after its prologue (barriers, TMEM, list, bias: nothing a previous kernel writes) -- for the kernels
    // before it to have completed (their activations are read / overwritten from here on)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_kb = a.n_kb;
    const uint32_t tmem_base = *s_tmem;
    // debug timeline: slot j of local tile lt of this CTA
    auto stamp_v = [&](int lt, int j, long long v) {
        if (tline && lt < 16) tline[((size_t)blockIdx.x * 16 + lt) * 16 + j] = v;
    };
    auto stamp = [&](int lt, int j) {
        if (tline && lt < 16) tline[((size_t)blockIdx.x * 16 + lt) * 16 + j] = clock64();
    };

    if (wg == 0) {
        reg_dec<32>();
        // both loops run warp-wide with one elected lane issuing, so that addresses and descriptors stay in uniform
        // registers (a lane-0 branch makes the compiler wrap every tcgen05.mma in a divergence loop of register moves)
        if (warp == 0) {
            // ============================================================================ weight producer
            int s = 0, par = 1;  // stage and the parity that means "this round's buffer is free"
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int nt = t % n_ntiles;
                const unsigned char *wsrc = a.wpack + (size_t)(a.ntile_base + nt) * a.kb_total * (2 * TILE_BYTES);
                for (int i = 0; i < n_kb; ++i) {
                    mbar_wait(bar_empty(s), par);
                    const unsigned char *src = wsrc + (size_t)s_list[i].x * (2 * TILE_BYTES);
                    if (elect_one()) {
                        if (dbg & 2) mbar_arrive(bar_full(s));
                        else {
                            mbar_arrive_expect_tx(bar_full(s), 2 * TILE_BYTES);
                            bulk_g2s(s_base + OFF_STAGES + s * STAGE_BYTES, src, 2 * TILE_BYTES, bar_full(s));
                        }
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; par ^= 1; }
                }
            }
        } else if (warp == 1) {
            // ================================================================================ MMA issuer
            const uint64_t desc0 = smem_desc<MODE>(s_base + OFF_STAGES);  // stage 0, k-step 0, hi image; addresses advance the low field
            int s = 0, par = 0, slot = 0, slot_par = 1, in_seg = 0, lt = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt) {
                if (lane == 0) stamp(lt, 0);
                long long w_full = 0, w_slot = 0, w_issue = 0;
                for (int i = 0; i < n_kb; ++i) {
                    const bool seg_first = in_seg == 0, seg_last = in_seg == seg_kb - 1 || i == n_kb - 1;
                    const long long c0 = tline ? clock64() : 0;
                    if (seg_first) {  // the slot must have been drained (fresh barrier: passes)
                        mbar_wait(bar_slot_empty(slot), slot_par);
                    }
                    const long long c1 = tline ? clock64() : 0;
                    mbar_wait(bar_full(s), par);
                    tc_fence_after();
                    const long long c2 = tline ? clock64() : 0;
                    if (elect_one()) {
                        const uint64_t bh0 = desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
                        const uint32_t d = tmem_base + (uint32_t)(slot * BN), a0 = tmem_base + (uint32_t)(A_COL0 + s * 2 * A_COLS);
                        if (!(dbg & 4)) {
#pragma unroll
                            for (int ks = 0; ks < Cfg<MODE>::KSTEPS; ++ks) {  // one k-step = 32 bytes of K in either mode
                                const uint64_t bh = bh0 + (uint64_t)(ks * 2), bl = bh + (uint64_t)(TILE_BYTES >> 4);
                                const uint32_t ah = a0 + ks * 8, al = ah + A_COLS;
                                umma_ts<MODE>(d, ah, bh, Cfg<MODE>::IDESC, (seg_first && ks == 0) ? 0u : 1u);
                                umma_ts<MODE>(d, ah, bl, Cfg<MODE>::IDESC, 1u);
                                umma_ts<MODE>(d, al, bh, Cfg<MODE>::IDESC, 1u);
                            }
                        }
                        umma_commit(bar_empty(s));                       // frees the stage once the MMAs above have read it
                        if (seg_last) umma_commit(bar_slot_full(slot));  // segment complete -> drain warps
                    }
                    __syncwarp();
                    if (tline) { w_slot += c1 - c0; w_full += c2 - c1; w_issue += clock64() - c2; }
                    if (i == 0 && lane == 0) stamp(lt, 1);
                    if (++s == STAGES) { s = 0; par ^= 1; }
                    if (seg_last) {
                        in_seg = 0;
                        if (++slot == SLOTS) { slot = 0; slot_par ^= 1; }
                    } else ++in_seg;
                }
                if (lane == 0) { stamp(lt, 2); stamp_v(lt, 8, w_full); stamp_v(lt, 9, w_slot); stamp_v(lt, 15, w_issue); }
            }
        }
    } else if (wg <= NGROUPS) {
        // ================================================================================== A producers
        // group `grp` takes the k-blocks whose running number is = grp (mod NGROUPS)
        reg_inc<144>();
        const int grp = wg - 1;
        const int r = (warp & 3) * 32 + lane;  // this thread's row of the tile = TMEM lane (a warp reaches lanes 32 * (warp % 4) ..)
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        int it0 = 0, lt = 0;  // running k-block number of the tile's first k-block
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt, it0 += n_kb) {
            if (n_kb == 0) break;
            if (tid == 128) stamp(lt, 3);
            // ---- per-row facts of this tile
            const int row = (t / n_ntiles) * BM + r;
            int rb = -1, rhw = 0, rslot = 0;
            uint32_t rgrp = 0;
            uint32_t *my_mask = reinterpret_cast<uint32_t *>(smem + OFF_MASK) + grp * (MAX_G * BM) + r;  // [g * BM]: private to this thread
            if (row < rows) {
                rb = row / a.ncells;
                const int cell = a.cell_base + (row - rb * a.ncells);
                rhw = a.cell_hw[cell];
                rslot = __ldg(a.perm + rhw);
                if (a.is_conv) for (int j = 0; j < G; ++j) my_mask[j * BM] = a.cell_tap[(size_t)cell * G + j];
                else rgrp = a.cell_grp[cell];
            }
            // sources are blocked channels-last (ctx.cuh): chunk c4 of slot s of image b starts at
            // ((b * NB + s / 32) * (channels / 4) + c4) * 128 + (s % 32) * 4.  Per tile and source the row's own base
            // pointer; the conv recomputes it when the tap (and with it the slot) changes
            const long long rbz = rb < 0 ? 0 : rb;
            const int NB = (HW + 31) >> 5;
            const int nq0 = (a.is_conv ? a.Cin : a.src0.channels) >> 2, nq1 = a.src1.channels >> 2;
            auto slot_base = [&](const float *ptr, int nq, int slot) {
                return ptr + ((rbz * NB + (slot >> 5)) * nq) * 128 + (slot & 31) * 4;
            };
            const float *base0 = slot_base(a.src0.ptr, nq0, rslot);
            const float *base1 = a.src1.ptr ? slot_base(a.src1.ptr, nq1, rslot) : nullptr;
            int last_shift = 0;              // conv: tap whose slot was looked up last (shift 0 = the row's own position)
            const float *tap_base = base0;   // ... and the row's base pointer at that tap
            // gathers the 32 k of k-block i for this thread's row into registers (masked elements = 0)
            auto gather = [&](int i, float(&v)[BK]) {
                const uint4 e = s_list[i];
                const bool s1 = e.z & 1u;
                const int tap = (int)((e.z >> 1) & 31u), nvalid = (int)((e.z >> 8) & 15u);
                uint32_t vis = 0;  // bit j: chunk j (channels c0 + 4j .. + 3) is loaded
                if (!((e.z >> 16) & 1u)) vis = 0xffu;  // ungrouped source: always visible
                else if (G == 1) vis = (a.is_conv ? ((my_mask[0] >> tap) & 1u) : (rgrp & 1u)) ? 0xffu : 0u;  // one group: all or nothing
                else {
#pragma unroll
                    for (int j = 0; j < BK / 4; ++j) {
                        const uint32_t g = (e.w >> (4 * j)) & 15u;
                        const uint32_t bit = a.is_conv ? ((my_mask[g * BM] >> tap) & 1u) : ((rgrp >> g) & 1u);
                        vis |= bit << j;
                    }
                }
                vis &= (1u << nvalid) - 1u;
                if (rb < 0 || (dbg & 1)) vis = 0;
                // row base at the tap's position (inside the image whenever the tap is visible; 1x1 layers: the row's own slot)
                const int shift = (int)(short)(e.y & 0xffffu);
                if ((shift != last_shift || (dbg & 64)) && vis != 0) {  // consecutive k-blocks of one tap share the lookup
                    tap_base = slot_base(a.src0.ptr, nq0, __ldg(a.perm + rhw + shift));
                    last_shift = shift;
                }
                const float *p = (s1 ? base1 : (shift ? tap_base : base0)) + (int)(e.y >> 16) * 128;
                // eight 4-channel chunks, 512 bytes apart; the warp's lanes (rows of one coding group) are neighbouring slots
#pragma unroll
                for (int j = 0; j < BK / 4; ++j) {
                    const float4 t4 = ((vis >> j) & 1u) ? __ldg(reinterpret_cast<const float4 *>(p + j * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    // MODE 1: even chunks are hi columns, odd chunks lo columns -> registers in TMEM column order
                    constexpr int kDummy = 0;
                    const int d = MODE ? ((j & 1) ? 16 + 4 * (j >> 1) : 4 * (j >> 1)) + kDummy : 4 * j;
                    v[d + 0] = t4.x; v[d + 1] = t4.y; v[d + 2] = t4.z; v[d + 3] = t4.w;
                }
            };
            // splits k-block i into hi = tf32(v), lo = v - hi and stores both into the stage's TMEM columns
            long long w_empty = 0, w_st = 0, w_g = 0;
            auto publish = [&](int i, const float(&v)[BK]) {
                const int it = it0 + i, s = it % STAGES, round = it / STAGES;
                const long long c0 = tline ? clock64() : 0;
                mbar_wait(bar_empty(s), (round & 1) ^ 1);
                const long long c1 = tline ? clock64() : 0;
                w_empty += c1 - c0;
                if (dbg & 16) { mbar_arrive(bar_full(s)); return; }
                tc_fence_after();  // the MMAs that read this stage's TMEM columns have completed (tcgen05.commit)
                const uint32_t col = lane_base + (uint32_t)(A_COL0 + s * 2 * A_COLS);
                if constexpr (MODE == 0) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {  // 16 columns at a time keeps the temporaries small
                        uint32_t h[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) h[q] = tf32_hi(v[half * 16 + q]);
                        tmem_st16(col + half * 16, h);
#pragma unroll
                        for (int q = 0; q < 16; ++q) h[q] = __float_as_uint(v[half * 16 + q] - __uint_as_float(h[q]));
                        tmem_st16(col + A_COLS + half * 16, h);
                    }
                } else {
                    // the activations arrive already split (split16) and in column order: hi columns 0..15, lo columns 0..15
                    uint32_t h[BK];
#pragma unroll
                    for (int q = 0; q < BK; ++q) h[q] = __float_as_uint(v[q]);
                    tmem_st32(col, h);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_full(s));
                if (tline) w_st += clock64() - c1;
            };
            // this group's k-blocks of the tile: i = i_first, i_first + NGROUPS, ...; the loads of the next one are in
            // flight while the current one is split and stored
            const int i_first = ((grp - it0) % NGROUPS + NGROUPS) % NGROUPS;
            float v0[BK], v1[BK], v2[BK];
            auto gather_t = [&](int i, float(&v)[BK]) {
                const long long c0 = tline ? clock64() : 0;
                gather(i, v);
                if (tline) w_g += clock64() - c0;
            };
            // three register buffers: the loads of this group's next two k-blocks are in flight while one is published
            constexpr int NG = NGROUPS;
            if (i_first < n_kb) gather_t(i_first, v0);
            if (i_first + NG < n_kb) gather_t(i_first + NG, v1);
            for (int i = i_first; i < n_kb; i += 3 * NG) {
                if (i + 2 * NG < n_kb) gather_t(i + 2 * NG, v2);
                publish(i, v0);
                if (i + NG < n_kb) {
                    if (i + 3 * NG < n_kb) gather_t(i + 3 * NG, v0);
                    publish(i + NG, v1);
                }
                if (i + 2 * NG < n_kb) {
                    if (i + 4 * NG < n_kb) gather_t(i + 4 * NG, v1);
                    publish(i + 2 * NG, v2);
                }
            }
            if (tid == 128) { stamp(lt, 4); stamp_v(lt, 10, w_empty); stamp_v(lt, 11, w_st); stamp_v(lt, 12, w_g); }
        }
    } else {
        // ============================================================================= drain + epilogue
        reg_inc<192>();
        const int r = (warp & 3) * 32 + lane;  // a warp may only read TMEM lanes 32 * (warp % 4) .. + 31
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int n_seg = (n_kb + seg_kb - 1) / seg_kb;
        const float osc = MODE ? a.out_scale : 1.f;  // MODE 1: undoes the power-of-two operand scales (exact)
        const long long ohw = (long long)a.Ntot * HW;
        const uint32_t out_row = s_base + OFF_OUT + (uint32_t)r * OUT_ROW;  // this thread's private staging row
        int segg = 0, lt = 0;
        float amax = 0.f;  // MODE 1: largest activation magnitude written (range flag)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt) {
            if (tid == 384) stamp(lt, 5);
            const int nt = t % n_ntiles;
            const int row = (t / n_ntiles) * BM + r;
            int rb = -1, rhw = 0, rslot = 0;
            if (row < rows) {
                rb = row / a.ncells;
                rhw = a.cell_hw[a.cell_base + (row - rb * a.ncells)];
                rslot = __ldg(a.perm + rhw);
            }
            float acc[BN];
#pragma unroll
            for (int q = 0; q < BN; ++q) acc[q] = 0.f;
            long long w_sf = 0;
            for (int seg = 0; seg < n_seg; ++seg, ++segg) {
                const int slot = segg % SLOTS;
                const long long c0 = tline ? clock64() : 0;
                mbar_wait(bar_slot_full(slot), (segg / SLOTS) & 1);
                if (tline) w_sf += clock64() - c0;
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < BN / 32; ++cc) {
                    if (dbg & 8) break;
                    uint32_t tt[32];
                    tmem_ld32(lane_base + (uint32_t)(slot * BN + cc * 32), tt);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[cc * 32 + q] += __uint_as_float(tt[q]);
                }
                tc_fence_before();
                mbar_arrive(bar_slot_empty(slot));
            }
            if (tid == 384) { stamp(lt, 6); stamp_v(lt, 14, w_sf); }
            // phase 1 (unrolled: the accumulators are registers): raw sums -> this thread's padded staging row
#pragma unroll
            for (int q = 0; q < BN; q += 4)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_row + q * 4), "f"(acc[q] * osc), "f"(acc[q + 1] * osc), "f"(acc[q + 2] * osc), "f"(acc[q + 3] * osc) : "memory");
            __syncwarp();
            const int n_left = min(BN, a.n_count - nt * BN);
            // phase 2 (compact loops: the instruction cache has to hold every role's code)
            if (a.out_cl) {
                // blocked channels-last: chunk by chunk, the warp's rows (neighbouring positions) fill 16-byte slots of
                // one or two 512-byte blocks (n_count % 4 == 0 on this path)
                if (rb >= 0) {
                    const int NB = (HW + 31) >> 5;
                    float *op = a.out + (((long long)rb * NB + (rslot >> 5)) * (a.Ntot >> 2) + ((a.n_begin + nt * BN) >> 2)) * 128 + (rslot & 31) * 4;
                    const float *bp = s_bias + nt * BN;
                    const float *srow_c = reinterpret_cast<const float *>(smem + OFF_OUT + r * OUT_ROW);
                    const float *ap = a.add ? a.add + (long long)rb * ohw + rhw + (long long)(a.n_begin + nt * BN) * HW : nullptr;
                    auto finish4 = [&](int q) {
                        float4 val = *reinterpret_cast<const float4 *>(srow_c + q);
                        const float4 bv = *reinterpret_cast<const float4 *>(bp + q);
                        val.x += bv.x; val.y += bv.y; val.z += bv.z; val.w += bv.w;
                        if (ap) {  // merger-less: + prior (NCHW)
                            val.x += __ldg(ap + (long long)(q + 0) * HW); val.y += __ldg(ap + (long long)(q + 1) * HW);
                            val.z += __ldg(ap + (long long)(q + 2) * HW); val.w += __ldg(ap + (long long)(q + 3) * HW);
                        }
                        if (a.lrelu) {
                            val.x = val.x > 0.f ? val.x : val.x * kSlope; val.y = val.y > 0.f ? val.y : val.y * kSlope;
                            val.z = val.z > 0.f ? val.z : val.z * kSlope; val.w = val.w > 0.f ? val.w : val.w * kSlope;
                        }
                        return val;
                    };
                    if (MODE == 1 && !a.out_f32) {  // the next layer's operand format: eight channels -> a hi and a lo chunk
#pragma unroll 4
                        for (int q = 0; q < n_left; q += 8) {
                            uint4 hi, lo;
                            split16(finish4(q), finish4(q + 4), amax, hi, lo);
                            *reinterpret_cast<uint4 *>(op + q * 32) = hi;
                            *reinterpret_cast<uint4 *>(op + (q + 4) * 32) = lo;
                        }
                    } else {
#pragma unroll 8
                        for (int q = 0; q < n_left; q += 4) *reinterpret_cast<float4 *>(op + q * 32) = finish4(q);
                    }
                }
            } else if (rb >= 0) {
                // NCHW (+ prior): for a fixed channel the warp's rows are neighbouring cells -> a strided sector run per store
                const float *srow = reinterpret_cast<const float *>(smem + OFF_OUT + r * OUT_ROW);
                float *op = a.out + (long long)rb * ohw + rhw + (long long)(a.n_begin + nt * BN) * HW;
                const float *ap = a.add ? a.add + (long long)rb * ohw + rhw + (long long)(a.n_begin + nt * BN) * HW : nullptr;
#pragma unroll 2
                for (int q = 0; q < n_left; ++q) {
                    float val = srow[q] + s_bias[nt * BN + q];
                    if (ap) val += __ldg(ap + (long long)q * HW);
                    if (a.lrelu) val = val > 0.f ? val : val * kSlope;
                    op[(long long)q * HW] = val;
                }
            }
            if (tid == 384) stamp(lt, 7);
        }
        if (MODE == 1 && !(amax < kActLimit) && a.range_flag) atomicOr(a.range_flag, 1);  // (NaN / inf raise it too)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();  // lane 0 ran the issue loop alone: reconverge before the .sync.aligned dealloc
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Packs weights into per-(out-group, n-tile, k-block) [hi | lo] K-major swizzled images (MODE 0: tf32 pairs in
// SWIZZLE_128B rows of 32 floats; MODE 1: fp16 pairs of w * scale in SWIZZLE_64B rows of 32 halves).
//   conv : w [N][Cin][k2]            K' = tap * Cpad + c            (Cpad = Cin rounded up to 32)
//   dense: w [N][c_src0 + c_src1]    K' = [src0 padded to 32 | src1 padded to 32]
template <int MODE>
__global__ void __launch_bounds__(256)
k_pack_w_tc(const float *__restrict__ w, unsigned char *__restrict__ out, int N, int G, int ntiles, int kb_total, int is_conv,
            int Cin, int k2, int c_src0, int c_src1, int kb_src0, float scale)
{
    constexpr int TILE_BYTES = Cfg<MODE>::TILE_BYTES;
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
        unsigned char *t = out + (size_t)tkb * (2 * TILE_BYTES);
        if constexpr (MODE == 0) {
            const uint32_t hi = tf32_hi(v);
            const float lo = v - __uint_as_float(hi);
            const int o = swz(nl, kk >> 2) + (kk & 3) * 4;
            *reinterpret_cast<uint32_t *>(t + o) = hi;
            *reinterpret_cast<float *>(t + TILE_BYTES + o) = lo;
        } else {
            const float x = v * scale;
            const __half hi = __float2half_rn(x);
            const __half lo = __float2half_rn(x - __half2float(hi));
            const int o = swz64(nl, kk >> 3) + (kk & 7) * 2;
            *reinterpret_cast<__half *>(t + o) = hi;
            *reinterpret_cast<__half *>(t + TILE_BYTES + o) = lo;
        }
    }
}

__global__ void __launch_bounds__(256)
k_absmax(const float *__restrict__ w, long long n, unsigned int *out)
{
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bit patterns
}


// ---------------------------------------------------------------------------------------------------------------
// Micro-benchmark (tools/mma_bench.py, basic_debug_mma_bench): back-to-back tcgen05.mma of one shape from one warp,
// operands = whatever the shared / tensor memory holds; cycles per MMA between the first issue and the commit's arrival.
template <int MODE, bool TS>
__global__ void __launch_bounds__(128, 1)
k_mma_bench(int n_cols, int iters, int same_acc, long long *out)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | ((MODE ? 0u : 2u) << 7) | ((MODE ? 0u : 2u) << 10) | ((uint32_t)(n_cols >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
        const uint64_t adesc = smem_desc<MODE>(s_base), bdesc = smem_desc<MODE>(s_base + 16384);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                const uint32_t d = tmem_base + (same_acc ? 0u : (uint32_t)((i & 1) * n_cols));
                if constexpr (TS) umma_ts<MODE>(d, tmem_base + 448, bdesc, idesc, i ? 1u : 0u);
                else {
                    if constexpr (MODE == 0)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int mma_bench(int mode, int ts, int n_cols, int iters, int same_acc, long long *cycles)
{
    long long *d = nullptr;
    BASIC_CUDA(cudaMalloc(&d, 16));
    const int smem = 96 * 1024;
#define BASIC_MB(M, T)                                                                                         \
    do {                                                                                                       \
        BASIC_CUDA(cudaFuncSetAttribute(k_mma_bench<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        k_mma_bench<M, T><<<1, 128, smem>>>(n_cols, iters, same_acc, d);                                       \
    } while (0)
    if (mode == 0 && ts) BASIC_MB(0, true);
    else if (mode == 0) BASIC_MB(0, false);
    else if (ts) BASIC_MB(1, true);
    else BASIC_MB(1, false);
#undef BASIC_MB
    BASIC_CUDA(cudaDeviceSynchronize());
    BASIC_CUDA(cudaMemcpy(cycles, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return BASIC_OK;
}

int pack_weights_tc(PackedW &dst, const float *w_dev, int N, int G, int is_conv, int Cin, int k2, int c_src0, int c_src1, int mode,
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
    const int tile_bytes = mode ? Cfg<1>::TILE_BYTES : Cfg<0>::TILE_BYTES;
    const size_t bytes = (size_t)G * dst.ntiles_per_group * dst.kb_total * (2 * tile_bytes);
    BASIC_TRY(dst.buf.reserve(bytes + 16));
    dst.scale = 1.f;
    if (mode) {
        // power-of-two scale that puts the largest weight magnitude into [2^12, 2^13): fp16 keeps 11 bits of it and the
        // residual (2^-11 of the value) stays a normal fp16 number for every weight down to 2^-15 of the largest
        unsigned int *d_max = reinterpret_cast<unsigned int *>(dst.buf.as<unsigned char>() + bytes);
        BASIC_CUDA(cudaMemsetAsync(d_max, 0, 4, stream));
        const long long nw = (long long)N * (is_conv ? (long long)Cin * k2 : (long long)(c_src0 + c_src1));
        k_absmax<<<256, 256, 0, stream>>>(w_dev, nw, d_max);
        BASIC_LAUNCHED();
        float wmax = 0.f;
        BASIC_CUDA(cudaMemcpyAsync(&wmax, d_max, 4, cudaMemcpyDeviceToHost, stream));
        BASIC_CUDA(cudaStreamSynchronize(stream));
        if (!(wmax < 3.0e38f)) return value_error("context-model weights are not finite");
        int e = 0;
        if (wmax > 0.f) frexpf(wmax, &e);  // wmax = f * 2^e, f in [0.5, 1)
        dst.scale = ldexpf(1.f, 13 - e);
        k_pack_w_tc<1><<<1024, 256, 0, stream>>>(w_dev, dst.buf.as<unsigned char>(), N, G, dst.ntiles_per_group, dst.kb_total, is_conv,
                                                 Cin, k2, c_src0, c_src1, dst.kb_src0, dst.scale);
    } else {
        k_pack_w_tc<0><<<1024, 256, 0, stream>>>(w_dev, dst.buf.as<unsigned char>(), N, G, dst.ntiles_per_group, dst.kb_total, is_conv,
                                                 Cin, k2, c_src0, c_src1, dst.kb_src0, 1.f);
    }
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// The tensor path takes every layer of every stage of a model or none (its activations are channels-last).  It
// needs 4-channel chunks that never straddle a visibility group, at most MAX_G groups, a k-block list that fits,
// and stages big enough for 128-row MMA tiles to make sense (scanline-like maps stay on the exact-FP32 kernel).
bool tc_model_eligible(const CtxModel &m, int B)
{
    if ((m.run_precision != BASIC_CTX_TF32X3 && m.run_precision != BASIC_CTX_FP16X3) || !m.has_conv || m.S < 1) return false;
    if (m.internal) return false;  // the internal 2G-group merger runs on the exact FP32 kernels
    if (!m.has_merger && m.S == 1 && m.stages[0].tap_or == 0) return false;  // nothing to multiply: params = prior + bias
    if (m.G > MAX_G || m.k * m.k * ((m.C + BK - 1) / BK) > MAX_KB || m.k * m.k > 31) return false;
    if (std::max(m.c_ctx, std::max(m.c_m1, m.c_m2)) / m.G > MAX_BIAS) return false;
    auto ok4 = [&](int channels) { return channels % m.G == 0 && (channels / m.G) % 4 == 0; };
    if (!ok4(m.C) || !ok4(m.c_ctx)) return false;
    if (m.has_merger && (!ok4(m.c_m1) || !ok4(m.c_m2))) return false;
    return (long long)B * m.G * m.H * m.W / m.S >= 64;
}

// 3xFP16 keeps activations in 8-channel chunk pairs: every channel count (per visibility group) must be a multiple of 8
bool tc_fp16_ok(const CtxModel &m)
{
    auto ok8 = [&](int channels) { return channels % m.G == 0 && (channels / m.G) % 8 == 0; };
    return ok8(m.C) && ok8(m.c_ctx) && (!m.has_merger || (ok8(m.c_m1) && ok8(m.c_m2)));
}

// [B, channels, HW] -> blocked channels-last (ctx.cuh), 32 positions x 32 channels per CTA through shared memory
// (both sides coalesced); positions past HW in the last block are written as zeros.  split = 1: 3xFP16 operand
// format (split16) instead of the floats themselves.
__global__ void __launch_bounds__(256)
k_nchw_to_cl(const float *__restrict__ src, float *__restrict__ dst, int channels, int HW, int split, int *range_flag,
             const int32_t *__restrict__ iperm)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, blk = blockIdx.x, hw0 = blk * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *s = src + (size_t)b * channels * HW;
    const int hw = hw0 + tx < HW ? iperm[hw0 + tx] : -1;  // the position stored in slot hw0 + tx
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i;
        tile[ty + 8 * i][tx] = (c < channels && hw >= 0) ? s[(size_t)c * HW + hw] : 0.f;
    }
    __syncthreads();
    // thread = (chunk ty of 4 channels, slot tx): one 16-byte chunk each, a warp writes 512 contiguous bytes
    if (c0 + ty * 4 < channels) {
        float *d = dst + (((size_t)b * gridDim.x + blk) * (channels >> 2) + (c0 >> 2) + ty) * 128 + tx * 4;
        if (split) {  // chunk pair m = ty / 2 holds channels 8m .. 8m + 7: the even chunk their hi halves, the odd one the lo halves
            const int m8 = (ty >> 1) * 8;
            const float4 x0 = make_float4(tile[m8 + 0][tx], tile[m8 + 1][tx], tile[m8 + 2][tx], tile[m8 + 3][tx]);
            const float4 x1 = make_float4(tile[m8 + 4][tx], tile[m8 + 5][tx], tile[m8 + 6][tx], tile[m8 + 7][tx]);
            float amax = 0.f;
            uint4 hi, lo;
            split16(x0, x1, amax, hi, lo);
            *reinterpret_cast<uint4 *>(d) = (ty & 1) ? lo : hi;
            if (!(amax < kActLimit) && range_flag) atomicOr(range_flag, 1);
        } else {
            *reinterpret_cast<float4 *>(d) = make_float4(tile[ty * 4 + 0][tx], tile[ty * 4 + 1][tx], tile[ty * 4 + 2][tx], tile[ty * 4 + 3][tx]);
        }
    }
}

// blocked channels-last (floats) -> [B, channels, HW]: the inverse of k_nchw_to_cl, for the public stage API
__global__ void __launch_bounds__(256)
k_cl_to_nchw(const float *__restrict__ src, float *__restrict__ dst, int channels, int HW, const int32_t *__restrict__ iperm)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, blk = blockIdx.x, hw0 = blk * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (c0 + ty * 4 < channels) {
        const float4 x = *reinterpret_cast<const float4 *>(src + (((size_t)b * gridDim.x + blk) * (channels >> 2) + (c0 >> 2) + ty) * 128 + tx * 4);
        tile[ty * 4 + 0][tx] = x.x; tile[ty * 4 + 1][tx] = x.y; tile[ty * 4 + 2][tx] = x.z; tile[ty * 4 + 3][tx] = x.w;
    }
    __syncthreads();
    float *d = dst + (size_t)b * channels * HW;
    const int hw = hw0 + tx < HW ? iperm[hw0 + tx] : -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i;
        if (c < channels && hw >= 0) d[(size_t)c * HW + hw] = tile[ty + 8 * i][tx];
    }
}

int launch_cl_to_nchw(const float *src, float *dst, int B, int channels, int HW, const int32_t *iperm, cudaStream_t stream)
{
    if (B == 0) return BASIC_OK;
    dim3 grid((HW + 31) / 32, (channels + 31) / 32, B);
    k_cl_to_nchw<<<grid, 256, 0, stream>>>(src, dst, channels, HW, iperm);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

int launch_nchw_to_cl(const float *src, float *dst, int B, int channels, int HW, const int32_t *iperm, cudaStream_t stream, int split,
                      int *range_flag)
{
    if (B == 0) return BASIC_OK;
    dim3 grid((HW + 31) / 32, (channels + 31) / 32, B);
    k_nchw_to_cl<<<grid, 256, 0, stream>>>(src, dst, channels, HW, split, range_flag, iperm);
    BASIC_LAUNCHED();
    return BASIC_OK;
}

// The k-block list of one (stage, out-group, layer): every 32-k block at least one row of the launch can see, in
// weight order.  Entry: x = weight k-block index; y = position shift of the tap (int16) | first 4-channel chunk << 16; z = source (bit 0) | tap << 1 | valid 4-channel chunks << 8 |
// grouped << 16; w = visibility group of each of the 8 chunks, 4 bits each.
static void build_kb_list(const LayerArgs &a, std::vector<uint4> &out)
{
    out.clear();
    const int total = a.is_conv ? a.ksize * a.ksize * a.kb_src0 : a.kb_total;
    for (int i = 0; i < total; ++i) {
        int src = 0, tap = 0, c0, nch, groups, shift = 0;
        if (a.is_conv) {
            const int kbt = a.kb_src0 /* k-blocks per tap */, pad = a.ksize / 2;
            tap = i / kbt;
            c0 = (i - tap * kbt) * BK;
            nch = a.Cin;
            groups = a.G;
            shift = (tap / a.ksize - pad) * a.W_img + (tap % a.ksize - pad);
        } else {
            src = i >= a.kb_src0;
            if (a.fold_src0 && !src) continue;   // constant source folded into the bias
            c0 = (src ? i - a.kb_src0 : i) * BK;
            const Source sc = src ? a.src1 : a.src0;
            nch = sc.channels;
            groups = sc.groups;
        }
        const int nvalid = std::min(BK / 4, (nch - c0) / 4);
        uint32_t gids = 0;
        bool vis = groups == 0;
        if (groups) {
            const int cpg = nch / groups;
            for (int j = 0; j < nvalid; ++j) {
                const int g = (c0 + 4 * j) / cpg;
                gids |= (uint32_t)g << (4 * j);
                vis |= a.is_conv ? ((a.vis_or[g] >> tap) & 1u) : ((a.vis_or[0] >> g) & 1u);
            }
        }
        if (!vis) continue;
        out.push_back(make_uint4((uint32_t)i, ((uint32_t)shift & 0xffffu) | ((uint32_t)(c0 / 4) << 16),
                                 (uint32_t)src | ((uint32_t)tap << 1) | ((uint32_t)nvalid << 8) | (groups ? 1u << 16 : 0u), gids));
    }
}

int launch_layer_tc(CtxModel &m, const LayerArgs &a, cudaStream_t stream)
{
    const int rows = a.B * a.ncells;
    if (rows == 0 || a.n_count == 0) return BASIC_OK;
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BASIC_CUDA(cudaFuncSetAttribute(k_layer_tc<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    const long long tiles = (long long)((a.n_count + BN - 1) / BN) * ((rows + BM - 1) / BM);
    dim3 grid((unsigned)(tiles < m.sm_count ? tiles : m.sm_count));  // persistent: one CTA per SM
    LayerArgs b = a;
    // k-block list of this (stage, out-group, layer): built once per map, kept in device memory
    const int n_keys = m.S * m.G * 4;
    if ((int)m.kb_count.size() != n_keys) {
        BASIC_CUDA(cudaStreamSynchronize(stream));  // launches in flight may still read the old pool
        m.kb_count.assign((size_t)n_keys, -1);
        BASIC_TRY(m.kb_pool.reserve((size_t)n_keys * MAX_KB * sizeof(uint4)));
    }
    if (a.list_key < 0 || a.list_key >= n_keys) return value_error("k-block list key out of range");
    uint4 *slot = m.kb_pool.as<uint4>() + (size_t)a.list_key * MAX_KB;
    if (m.kb_count[a.list_key] < 0) {
        std::vector<uint4> list;
        build_kb_list(a, list);
        if ((int)list.size() > MAX_KB) return value_error("k-block list too long for the tensor-core context kernel");
        if (!list.empty()) BASIC_CUDA(cudaMemcpy(slot, list.data(), list.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        m.kb_count[a.list_key] = (int)list.size();
    }
    b.kb_list = slot;
    b.n_kb = m.kb_count[a.list_key];
    static const int dbg = getenv("BASIC_TC_DEBUG") ? atoi(getenv("BASIC_TC_DEBUG")) : 0;
    b.debug = dbg;
    static const int nacc_env = getenv("BASIC_TC_NACC") ? atoi(getenv("BASIC_TC_NACC")) : 0;
    if (nacc_env > 0) b.nacc = nacc_env;
    // BASIC_TC_TIMELINE=<launch index>: clock64 stamps of the roles of that launch, printed for a few CTAs
    static const int tl_at = getenv("BASIC_TC_TIMELINE") ? atoi(getenv("BASIC_TC_TIMELINE")) : -1;
    static int tl_count = 0;
    static long long *tl_buf = nullptr;
    const bool tl = tl_at >= 0 && tl_count++ == tl_at;
    if (tl) {
        cudaMalloc(&tl_buf, (size_t)grid.x * 16 * 16 * sizeof(long long));
        cudaMemset(tl_buf, 0, (size_t)grid.x * 16 * 16 * sizeof(long long));
        b.timeline = tl_buf;
    }
    const bool instrumented = b.debug != 0 || b.timeline != nullptr;
    // launched with programmatic stream serialisation: the kernel's griddepcontrol.wait orders it after its predecessors
    static const bool pdl = !(getenv("BASIC_TC_NOPDL") && atoi(getenv("BASIC_TC_NOPDL")));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto kfn = instrumented ? (b.mode ? k_layer_tc<1, 1> : k_layer_tc<0, 1>) : (b.mode ? k_layer_tc<1, 0> : k_layer_tc<0, 0>);
    BASIC_CUDA(cudaLaunchKernelEx(&cfg, kfn, b));
    BASIC_LAUNCHED();
    if (tl) {
        std::vector<long long> h((size_t)grid.x * 16 * 16);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), tl_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        fprintf(stderr, "timeline: launch %d rows %d n_count %d conv %d tiles %lld (cycles from the CTA's first stamp)\n", tl_at, rows, a.n_count, a.is_conv, tiles);
        for (unsigned cta : {0u, grid.x / 2, grid.x - 1}) {
            long long t0 = h[(size_t)cta * 256 + 0];
            for (int lt = 0; lt < 16; ++lt) {
                const long long *e = &h[((size_t)cta * 16 + lt) * 16];
                if (!e[0]) break;
                fprintf(stderr, "  cta %3u tile %d | mma: start %7lld first-issue %7lld last-commit %7lld | producer: start %7lld done %7lld | drain: start %7lld drained %7lld stored %7lld || waits: mma full %lld slot %lld | prodA empty %lld st %lld gather %lld (%lld) | drain slot_full %lld | mma issue %lld\n",
                        cta, lt, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0, e[8], e[9], e[10], e[11], e[12], e[13], e[14], e[15]);
            }
        }
    }
    return BASIC_OK;
}

}  // namespace basic

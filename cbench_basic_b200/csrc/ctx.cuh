// Context-model state shared by the exact-FP32 path (ctx.cu) and the tcgen05 3xTF32 path (ctx_tc.cu).
#pragma once
#include "common.cuh"

namespace basic {

// Packed weights of one layer for the tensor-core path: per (out-group, n-tile, k-block) one 32 KB image
// [hi: 128 rows x 128 B, SWIZZLE_128B K-major | lo: same], ready for a single bulk copy into shared memory.
struct PackedW {
    DevBuf buf;
    int kb_total = 0;         // k-blocks (32 k each) per n-tile
    int ntiles_per_group = 0; // n-tiles (128 output channels each) per out-group
    int kb_src0 = 0;          // dense: k-blocks of the first source; conv: k-blocks per tap
    float scale = 1.f;        // fp16 images: power of two the weights were multiplied by
};

struct CtxModel {
    int C = 0, G = 1, k = 5, device = 0, sm_count = 148;
    bool has_conv = false, has_merger = false;
    int c_ctx = 0, c_m1 = 0, c_m2 = 0;  // 2C, 10C/3, 8C/3
    // weights, K-major ("transposed"): wt[kk][o]
    DevBuf w_ctx, b_ctx;                // conv: kk = tap * C + c
    DevBuf w_m1, b_m1, w_m2, b_m2, w_m3, b_m3;
    DevBuf b_m1_fold;                   // G = 1: b_m1 + W_m1[:, :2C] . b_ctx -- the first merger layer's bias at cells that see no neighbour (ctx == bias)
    // the coder's INTERNAL merger (pgm_coder.py:1207-1239: 2G channel groups, the G "prior" groups with id -1): besides the
    // context branch above (widths c_m1 = c_m2 = bottleneck / 2, layers 2 and 3 also read the prior branch) a prior branch
    // prior -> p1 -> p2 of width c_p that no rule masks.  Exact FP32 kernels only.
    bool internal = false;
    int c_p = 0;
    DevBuf w_p1, b_p1, w_p2, b_p2, a_p1, a_p2;
    // map
    int H = 0, W = 0, S = 0;
    std::vector<int32_t> h_tg;
    struct Stage {
        std::vector<int> cell_off;       // [G + 1] offsets into the stage's cell arrays
        size_t cells_at = 0;             // offset of this stage inside d_cells (in cells)
        size_t pos_at = 0;               // offset inside d_positions
        int64_t n_pos = 0;
        uint32_t tap_or = 0;             // OR of all tap masks of the stage (0 -> conv contributes bias only)
        std::vector<uint32_t> og_tap_or; // [G out-groups][G in-groups] OR of the tap masks over the (stage, out-group) cells
        std::vector<uint32_t> og_grp_or; // [G out-groups] OR of the group-visibility bits
    };
    std::vector<Stage> stages;
    DevBuf d_cell_hw;    // int32 [ncells]            cell -> h*W+w
    DevBuf d_cell_tap;   // uint32 [ncells][G]        visible 5x5 taps per input channel group
    DevBuf d_cell_grp;   // uint32 [ncells]           bit j: in-group j visible through "<="
    DevBuf d_positions;  // int32 [C*H*W]             coded element offsets, stage-major
    DevBuf d_stage_cells;  // int2 [S]                first cell, cells of every stage (persistent stage kernel, G = 1)
    int max_stage_cells = 0;
    DevBuf scan_barrier;   // (SCAN_TIMING builds: cycle counters of the persistent stage kernel)
    uint32_t scan_step = 0, scan_call = 0;   // tags of the stage kernel's exchanged words: (stage, layer) steps and coding calls so far
    DevBuf ws_ctxc;        // ... k_scan_blocks: convolution weights with only the visible taps, N-major [2C][ntaps * C]
    uint32_t ws_ctxc_key = 0xffffffffu;   // the tap set they were packed for
    DevBuf scan_cs;        // ... single-launch decoding: chunk_syms of every slice
    DevBuf scan_ws;        // ... its tagged per-stage layer outputs [row][N], tagged position-major y_hat, position-major prior
    int scan_nctas = 0;
    DevBuf ws_ctx, ws_m1, ws_m2, ws_m3;   // N-major weight copies of the persistent stage kernel (conv: [2C][tap][C])
    DevBuf d_perm;       // int32 [H*W]               position -> slot of the tensor path's activation layout (stage-major)
    DevBuf d_iperm;      // int32 [H*W]               slot -> position
    // activations (grow-only)
    DevBuf a_ctx, a_m1, a_m2;
    int act_B = 0;
    // tensor-core path
    int precision = 0;   // configured: BASIC_CTX_FP32 | BASIC_CTX_TF32X3 | BASIC_CTX_FP16X3
    int run_precision = 0;  // what the next stage calls use (the y-path driver falls back from FP16X3 to TF32X3 when an
                            // activation leaves the fp16 range, and follows the mode recorded in a stream when decoding)
    PackedW q_ctx, q_m1, q_m2, q_m3;  // fp16 images of the weights (FP16X3)
    DevBuf range_flag;   // int: set by the FP16X3 kernels when |activation| >= 4000
    int nacc = 4;        // k-blocks (32 k each) accumulated in TMEM before a segment is drained into FP32 registers
    PackedW p_ctx, p_m1, p_m2, p_m3;
    DevBuf kb_pool;                   // k-block lists of the tensor path, one slot of 512 entries per (stage, out-group, layer)
    std::vector<int> kb_count;        // entries per slot, -1 = not built yet (reset by set_map / set_weights)
    DevBuf cl_ctx, cl_m1, cl_m2;      // channels-last activations of the tensor path
    DevBuf cl_buf, cl_prior;          // channels-last copies made when the caller only has NCHW (public stage API)
    DevBuf cl_params;                 // ... and the blocked channels-last parameters behind its NCHW result
};


// What the stage kernel needs to decode inside its one launch (ctx_scan.cu k_scan_stages / scan_decode_share)
struct ScanDecodeHost {
    const RansTables *tables;
    int bypass;
    const unsigned char *seg;     // device: the segment
    long long seg_cap;
    int n_chunks;
    const int32_t *chunk_syms;    // host: per slice (pageable memory is fine: copied before the call returns)
    int *status;                  // device: coder status flags (atomicOr)
};

struct Source {          // one block of K coming from an NCHW activation tensor
    const float *ptr;    // [B, channels, H, W]
    int channels;        // channels of this tensor
    int groups;          // channel groups subject to the visibility rule (0 = always visible)
    int cl;              // 1: channels-last [B, H*W, channels] (tensor-core path), 0: NCHW
};

struct LayerArgs {
    // rows = B x cells(stage, out-group)
    const int32_t *cell_hw;
    const int32_t *perm;        // tensor path: position -> slot of the blocked channels-last layout
    const uint32_t *cell_tap;   // conv only
    const uint32_t *cell_grp;   // dense only
    int ncells, cell_base;      // cells of this (stage, out-group) start at cell_base
    int B, HW, W_img, H_img, G;
    // K
    int is_conv, ksize, Cin;    // conv: Cin input channels of `src0`
    Source src0, src1;          // dense: K = src0.channels + src1.channels
    const float *wt;            // [K][Ntot] K-major
    const float *bias;          // [Ntot]
    int Ntot, n_begin, n_count; // this out-group's output channels [n_begin, n_begin + n_count)
    // epilogue
    float *out;                 // [B, Ntot, H, W]
    const float *add;           // optional [B, Ntot, H, W] added in the epilogue (merger-less: + prior)
    int lrelu;
    int out_cl;                 // tensor path: `out` is blocked channels-last (`add` stays NCHW)
    int out_f32;                // tensor path, 3xFP16: write plain floats (the parameter tensor), not the split16 operand format
    uint32_t tap_or;            // stage-level OR of the tap masks (conv)
    // tensor-core path
    const unsigned char *wpack; // PackedW image
    int kb_total, kb_src0, ntile_base, nacc;
    int mode;                   // tensor path: 0 = 3xTF32, 1 = 3xFP16
    float out_scale;            // 3xFP16: 1 / (activation scale * weight scale), applied to the accumulators
    int *range_flag;            // 3xFP16: raised when an activation magnitude reaches 4000
    int list_key;               // (stage * G + out-group) * 4 + layer: slot of this launch's k-block list
    const uint4 *kb_list;       // k-block list (built on the host once per key) and its length
    int n_kb;
    int debug;
    long long *timeline;        // optional [grid][16 tiles][8] clock64 stamps (BASIC_TC_TIMELINE), else NULL
    uint32_t vis_or[8];         // conv: OR over the launch's rows of the tap mask per input group; dense: [0] = OR of group bits
    int fold_src0;              // tensor path, dense: src0 is constant over the launch's rows and already inside `bias` (stage 0)
};


// ctx_tc.cu
// Activation layout of the tensor-core path ("blocked channels-last"): element (b, hw, c) of a [B, channels, HW] tensor
// lives at (((b * NB + hw / 32) * (channels / 4) + c / 4) * 32 + hw % 32) * 4 + c % 4, NB = ceil(HW / 32): the 4-channel
// chunks of 32 neighbouring positions are 512 contiguous bytes, so a warp whose lanes are neighbouring positions reads
// or writes whole lines per 128-bit access (plain channels-last costs one line per lane).  `hw` above is not the raster
// position but its SLOT: positions are renumbered stage-major (by the coding group of channel group 0, then raster order;
// CtxModel::d_perm), so the rows of one coding group -- what a warp of the GEMM kernel holds -- are neighbouring slots.
inline size_t cl_elems(int B, int channels, int HW) { return (size_t)B * ((HW + 31) / 32) * 32 * channels; }
bool tc_model_eligible(const CtxModel &m, int B);
bool tc_fp16_ok(const CtxModel &m);
int launch_nchw_to_cl(const float *src, float *dst, int B, int channels, int HW, const int32_t *iperm, cudaStream_t stream,
                      int split = 0, int *range_flag = nullptr);
int launch_layer_tc(CtxModel &m, const LayerArgs &a, cudaStream_t stream);
int launch_cl_to_nchw(const float *src, float *dst, int B, int channels, int HW, const int32_t *iperm, cudaStream_t stream);
int pack_weights_tc(PackedW &dst, const float *w_dev /* [N][Korig] state_dict layout */, int N, int G, int is_conv, int Cin,
                    int k2, int c_src0, int c_src1, int mode, cudaStream_t stream);

}  // namespace basic

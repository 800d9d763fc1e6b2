// libse3tn: context, per-object weight sets, TMA tensor maps and the launch schedule of the
// se(3)-TrackNet hot path behind the C ABI declared in include/se3tn.h.
#include "../../include/se3tn.h"
#include "conv_common.h"
#include "aux_kernels.h"
#include "metrics.h"
#include "render.h"
#include <dlfcn.h>
#include "depth_fill.h"
#include "ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace se3tn;

namespace {

// --------------------------------------------------------------------------------------------
// Network schedule: 14 conv launches cover the reference's 17 convs (se3_tracknet.py:57-78).
// --------------------------------------------------------------------------------------------
enum Buf { B_X0A, B_X0B, B_Y1A, B_Y1B, B_P1A, B_P1B, B_T1, B_T2, B_U, B_CAT, B_F1, B_T4, B_F2, B_H1, B_H2, B_H3, B_COUNT };

constexpr size_t kBufFloats[B_COUNT] = {
    kStemImgFloats, kStemImgFloats,                 // X0A, X0B  (182 x 184 x 4)
    88 * 88 * 64, 88 * 88 * 64,                     // Y1A, Y1B
    44 * 44 * 64, 44 * 44 * 64,                     // P1A, P1B
    44 * 44 * 64, 44 * 44 * 64, 44 * 44 * 64,       // T1, T2, U
    44 * 44 * 128,                                  // CAT
    22 * 22 * 256, 22 * 22 * 256, 22 * 22 * 256,    // F1, T4, F2
    11 * 11 * 1024, 11 * 11 * 1024, 11 * 11 * 1024  // H1, H2, H3
};

enum Kind { K_STEM, K_S1, K_S2 };

struct LayerSpec {
    Kind kind;
    Buf in, out, res;            // res == B_COUNT: none
    int Hin, Win, in_c;          // input spatial + channels per pixel of the input buffer
    int cin, cout, groups;       // per group
    int out_c, out_coff;         // channels per pixel of the output buffer, channel offset
    int act;
    int block_n;
};

constexpr Buf NONE = B_COUNT;
const LayerSpec kLayers[14] = {
    // kind   in     out    res    Hin  Win  in_c  cin  cout groups out_c coff act       BN
    {K_STEM, B_X0A, B_Y1A, NONE,  182, 184,   4,   32,   64, 1,    64,   0, ACT_SELU,  64},   // convA1
    {K_STEM, B_X0B, B_Y1B, NONE,  182, 184,   4,   32,   64, 1,    64,   0, ACT_SELU,  64},   // convB1
    {K_S1,   B_P1A, B_T1,  NONE,   44,  44,  64,   64,   64, 1,    64,   0, ACT_RELU,  64},   // convA2.conv1
    {K_S1,   B_T1,  B_CAT, B_P1A,  44,  44,  64,   64,   64, 1,   128,   0, ACT_RELU,  64},   // convA2.conv2 (+id) -> cat[0:64]
    {K_S1,   B_P1B, B_T2,  NONE,   44,  44,  64,   64,   64, 1,    64,   0, ACT_RELU,  64},   // convB2.conv1
    {K_S1,   B_T2,  B_U,   B_P1B,  44,  44,  64,   64,   64, 1,    64,   0, ACT_RELU,  64},   // convB2.conv2 (+id)
    {K_S1,   B_U,   B_T2,  NONE,   44,  44,  64,   64,   64, 1,    64,   0, ACT_RELU,  64},   // convB3.conv1
    {K_S1,   B_T2,  B_CAT, B_U,    44,  44,  64,   64,   64, 1,   128,  64, ACT_RELU,  64},   // convB3.conv2 (+id) -> cat[64:128]
    {K_S2,   B_CAT, B_F1,  NONE,   44,  44, 128,  128,  256, 1,   256,   0, ACT_SELU, 256},   // convAB1
    {K_S1,   B_F1,  B_T4,  NONE,   22,  22, 256,  256,  256, 1,   256,   0, ACT_RELU, 256},   // convAB2.conv1
    {K_S1,   B_T4,  B_F2,  B_F1,   22,  22, 256,  256,  256, 1,   256,   0, ACT_RELU, 256},   // convAB2.conv2 (+id) = 'feature'
    {K_S2,   B_F2,  B_H1,  NONE,   22,  22, 256,  256, 1024, 1,  1024,   0, ACT_SELU, 256},   // trans_conv1 ++ rot_conv1
    {K_S1,   B_H1,  B_H2,  NONE,   11,  11, 1024, 512,  512, 2,  1024,   0, ACT_RELU, 256},   // {trans,rot}_conv2.conv1
    {K_S1,   B_H2,  B_H3,  B_H1,   11,  11, 1024, 512,  512, 2,  1024,   0, ACT_RELU, 256},   // {trans,rot}_conv2.conv2 (+id)
};

inline int layer_taps(const LayerSpec& L) { return L.kind == K_STEM ? 7 : 9; }
inline int layer_ktot(const LayerSpec& L) { return layer_taps(L) * L.cin; }
inline int layer_rows(const LayerSpec& L) { return L.cout * L.groups; }
inline int layer_Ho(const LayerSpec& L) { return L.kind == K_STEM ? 88 : (L.kind == K_S2 ? L.Hin / 2 : L.Hin); }

constexpr size_t kFcFloats = 6 * 512 + 6;

size_t blob_floats() {
    size_t n = 0;
    for (const LayerSpec& L : kLayers) n += static_cast<size_t>(layer_rows(L)) * layer_ktot(L) + layer_rows(L);
    return n + kFcFloats;
}

struct WeightSet {
    float* dev = nullptr;           // exact fp32 blob
    float* dev_tf32 = nullptr;      // same layout, conv weights rounded to tf32 (biases / fc untouched)
    size_t w_off[14], b_off[14];
    size_t fc_off;
    CUtensorMap bmap[14];           // over dev_tf32
    uint8_t* dev_bf16 = nullptr;    // blob-sized: conv weights as [32 bf16 hi | 32 bf16 lo] per 32-word K chunk
    uint8_t* dev_stem_bf16 = nullptr;  // 2 x [64][448 words]: stem weights, two tiles per filter row (see aux_kernels.cu)
    CUtensorMap bmap_bf16[14];
    uint8_t* dev_stack = nullptr;   // 8 x [128][288 words]: resident-weight layers with hi / lo rows stacked along N (conv_umma2.cu STACK)
    CUtensorMap bmap_stack[8];
    // resident 64-channel layers (li 2..7) for the v2 kernel: weight rows in the epilogue's channel order (aux_kernels.cu
    // permute_rows64_kernel); [li-2] -> tf32 words | bf16 [hi|lo] chunks, 64 x 576 words each.  bmap_res[0..1] = the stems' natural maps.
    float* dev_perm = nullptr; float* dev_perm_tmp = nullptr;
    CUtensorMap bmap_res[8], bmap_res_bf16[8];
    CUtensorMap bmap_pair[8], bmap_bf16_pair[8];   // Cout=64 layers for the CTA-pair kernels: box = 32 weight rows (half per CTA)
    float mean32[8], std32[8];
    double mean64[8], std64[8];
    int stats_f64 = 0;
    bool has_stats = false;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_create_error;

}  // namespace

struct se3tn_ctx {
    int device = 0;
    int max_batch = 0;
    int num_sms = 0;
    bool own_workspace = false;
    uint8_t* workspace = nullptr;
    float* buf[B_COUNT] = {};
    CUtensorMap amap[14][7];
    CUtensorMap amap2[14][4];        // v2 kernel: boxes extended vertically (one per filter column / parity view)
    int conv_version = 2;            // SE3TN_CONV=1 selects the first-generation kernel
    int dual_m = 0;                  // SE3TN_DUAL_M=1: two M tiles per CTA on the BN=256 layers (halves weight fill traffic; measured
                                     // slightly slower than MT=1 once fills stopped being the limiter: the epilogue cannot overlap)
    int pair = 0;                    // SE3TN_PAIR=1: run the Cout=64 layers as cta_group::2 CTA pairs.  Correct, but measured no faster
                                     // (64-ch layers equal, stem 18 % slower): the pair MMA's ~1.3x per-SM advantage at N=64 is eaten by
                                     // the cross-CTA barrier round trips; kept as an experiment
    int pdl = 1;                     // SE3TN_PDL=0 disables programmatic dependent launch between conv kernels
    std::map<int, MeshDev> meshes;   // CAD models of the rasteriser (device copies), keyed by mesh id
    MeshDev* d_meshes = nullptr; int mesh_rows = 0; bool meshes_dirty = false;
    uint8_t* render_proj = nullptr; uint8_t* render_unif = nullptr; int render_max_nv = 0, render_proj_nv = 0;   // rasteriser workspace
    FillScratch fill = {nullptr, nullptr, nullptr, nullptr}; size_t fill_pixels = 0;   // depth hole-filling scratch (grows on demand)
    int fuse_pool = 1;               // SE3TN_FUSE_POOL=0: store the last head activation (debug buffer H3) and pool it in head_kernel
    float* pool_part = nullptr;      // [max_batch][4][1024] column sums from the last conv's epilogue
    int streamk = 0;                 // SE3TN_STREAMK=1: deal (unit, chunk) steps evenly over the CTAs in the BN=256 layers.  Measured no net gain at batch 64
                                     // (the 128 KB partial dump + fix-up per CTA costs what the 12.5 % shorter makespan wins) and results then
                                     // depend in the last ulps on a pair's position in the batch, so off by default
    float* sk_part = nullptr; int* sk_flags = nullptr; int sk_seq = 0;   // stream-K partial slots (one per SM), flags, launch counter
    unsigned long long* trace = nullptr;   // SE3TN_TRACE=1: [14 layers][256 CTAs][8] globaltimer stamps of the last forward (conv_umma2.cu trace_stamp)
    int debug_flags = 0;             // SE3TN_DEBUG_SKIP: timing experiments (bit0 no B fills, bit1 no A fills); results invalid
    int base_off_mode = 0;           // SE3TN_BASE_OFF: UMMA descriptor base_offset convention for row-shifted starts
    EncodeTiledFn encode = nullptr;
    std::map<int, WeightSet> weights;
    // device copies of per-set stats, rebuilt when a set changes: [max_id+1][8]
    float* d_mean32 = nullptr; float* d_std32 = nullptr; double* d_mean64 = nullptr; double* d_std64 = nullptr;
    int stats_rows = 0; bool stats_dirty = true; int stats_f64 = 0;
    // per-weight-set device tables for multi-set launches, rebuilt when a set is (re)loaded: entry [wid*14 + layer]
    CUtensorMap* d_bmaps_tf32 = nullptr; CUtensorMap* d_bmaps_bf16 = nullptr; CUtensorMap* d_bmaps_x3 = nullptr;
    const float** d_bias = nullptr; const float** d_fc = nullptr;   // d_fc[wid] -> [6][512] weights then [6] biases
    int table_rows = 0; bool tables_dirty = true;
    int launches = 0;
    bool profiling = false;
    cudaEvent_t ev0[SE3TN_PROFILE_SLOTS] = {}, ev1[SE3TN_PROFILE_SLOTS] = {};
    bool ev_used[SE3TN_PROFILE_SLOTS] = {};
    int umma_block_n_override = 0;
    std::string err;
};

namespace {

// conv-input storage mode of the packers for a precision: 0 raw fp32, 1 tf32 words, 2 bf16 hi/lo per pixel
inline int store_mode_of(int precision) {
    return precision == SE3TN_PREC_TF32 ? 1 : (precision == SE3TN_PREC_BF16X3 || precision == SE3TN_PREC_BF16) ? 2 : 0;
}

int fail(se3tn_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU_TRY(ctx, expr)                                                                        \
    do { cudaError_t e_ = (expr);                                                                \
         if (e_ != cudaSuccess) return fail((ctx), SE3TN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

struct ProfScope {
    se3tn_ctx* c; int slot; cudaStream_t s;
    ProfScope(se3tn_ctx* c_, int slot_, cudaStream_t s_) : c(c_), slot(slot_), s(s_) {
        if (c->profiling) { cudaEventRecord(c->ev0[slot], s); }
    }
    ~ProfScope() { if (c->profiling) { cudaEventRecord(c->ev1[slot], s); c->ev_used[slot] = true; } }
};

size_t workspace_floats(int max_batch) {
    size_t n = 0;
    for (int b = 0; b < B_COUNT; ++b) {
        size_t f = kBufFloats[b] * static_cast<size_t>(max_batch);
        n += (f + 255) & ~size_t(255);           // keep every buffer 1 KB aligned
    }
    return n;
}

// rank-4 fp32 tensor map, SWIZZLE_128B, box inner = 32 floats
int make_map4(se3tn_ctx* c, CUtensorMap* m, const void* base, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3],
              const cuuint32_t box[4], CUtensorMapL2promotion l2, const char* what) {
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[512];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(%s) failed: CUresult %d dims {%llu,%llu,%llu,%llu} strides {%llu,%llu,%llu} box {%u,%u,%u,%u}",
                 what, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
                 (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1], (unsigned long long)strides_bytes[2],
                 box[0], box[1], box[2], box[3]);
        return fail(c, SE3TN_ERR_CUDA, msg);
    }
    return SE3TN_OK;
}

int make_map2(se3tn_ctx* c, CUtensorMap* m, const void* base, cuuint64_t inner, cuuint64_t rows, cuuint32_t box_rows, const char* what) {
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[256];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(%s) failed: CUresult %d dims {%llu,%llu} box {32,%u}", what, (int)r,
                 (unsigned long long)inner, (unsigned long long)rows, box_rows);
        return fail(c, SE3TN_ERR_CUDA, msg);
    }
    return SE3TN_OK;
}

void tile_box(const LayerSpec& L, int& bw, int& bh, int& bn) {
    if (L.kind == K_STEM) { bw = 8; bh = 8; bn = 2; } else { bw = 11; bh = 11; bn = 1; }
}

int block_n_of(const se3tn_ctx* c, const LayerSpec& L) {
    int bn = L.block_n;
    if (c->umma_block_n_override && L.cout % c->umma_block_n_override == 0 && L.cout >= c->umma_block_n_override)
        bn = c->umma_block_n_override;
    return bn;
}

// Activation-side tensor maps: built once per context (they depend only on the workspace layout).
int build_activation_maps(se3tn_ctx* c) {
    const cuuint64_t N = static_cast<cuuint64_t>(c->max_batch);
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        const float* base = c->buf[L.in];
        int bw, bh, bn; tile_box(L, bw, bh, bn);
        char what[64];
        if (L.kind == K_STEM) {
            // Overlapping-window view of the zero-padded NHWC4 input: coordinate (k, ox, oy, n) ->
            // float offset k + 8*ox + (2*oy + r)*rowpitch + n*imgpitch; one map per filter row r.
            const cuuint64_t rowpitch = static_cast<cuuint64_t>(kStemW) * 4 * sizeof(float);
            const cuuint64_t dims[4] = {32, 88, 88, N};
            const cuuint64_t strides[3] = {8 * sizeof(float), 2 * rowpitch, static_cast<cuuint64_t>(kStemH) * rowpitch};
            const cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
            for (int r = 0; r < 7; ++r) {
                snprintf(what, sizeof what, "layer %d stem row %d", li, r);
                int rc = make_map4(c, &c->amap[li][r], reinterpret_cast<const uint8_t*>(base) + r * rowpitch, dims, strides, box,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                if (rc) return rc;
            }
        } else if (L.kind == K_S1) {
            const cuuint64_t C = L.in_c, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {C, W, H, N};
            const cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
            const cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
            snprintf(what, sizeof what, "layer %d s1", li);
            int rc = make_map4(c, &c->amap[li][0], base, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
            if (rc) return rc;
        } else {
            // stride 2: four parity views (py, px) of the input, each a dense half-resolution tensor
            const cuuint64_t C = L.in_c, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {C, W / 2, H / 2, N};
            const cuuint64_t strides[3] = {2 * C * 4, 2 * W * C * 4, H * W * C * 4};
            const cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    snprintf(what, sizeof what, "layer %d s2 parity %d%d", li, py, px);
                    int rc = make_map4(c, &c->amap[li][py * 2 + px], base + (static_cast<size_t>(py) * W + px) * C, dims, strides, box,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                    if (rc) return rc;
                }
        }
    }
    return SE3TN_OK;
}

// v2 activation maps: same tensors, boxes extended by the vertical filter extent so that the
// vertical taps become descriptor row shifts inside one shared-memory tile (conv_umma2.cu).
int build_activation_maps_v2(se3tn_ctx* c) {
    const cuuint64_t N = static_cast<cuuint64_t>(c->max_batch);
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        const float* base = c->buf[L.in];
        char what[64];
        if (L.kind == K_STEM) {
            // even / odd input-row views of the zero-padded NHWC4 stem input; x is the overlapping
            // 8-pixel window view (stride 2 pixels = 32 B, extent 128 B)
            const cuuint64_t rowpitch = static_cast<cuuint64_t>(kStemW) * 4 * sizeof(float);
            const cuuint64_t strides[3] = {8 * sizeof(float), 2 * rowpitch, static_cast<cuuint64_t>(kStemH) * rowpitch};
            for (int odd = 0; odd < 2; ++odd) {
                const cuuint64_t dims[4] = {32, 88, odd ? 90u : 91u, N};
                const cuuint32_t box[4] = {32, 11, odd ? 13u : 14u, 1};
                snprintf(what, sizeof what, "v2 layer %d stem %s rows", li, odd ? "odd" : "even");
                int rc = make_map4(c, &c->amap2[li][odd], reinterpret_cast<const uint8_t*>(base) + odd * rowpitch, dims, strides, box,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                if (rc) return rc;
            }
        } else if (L.kind == K_S1) {
            const cuuint64_t C = L.in_c, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {C, W, H, N};
            const cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
            const cuuint32_t box[4] = {32, 11, 13, 1};
            snprintf(what, sizeof what, "v2 layer %d s1", li);
            int rc = make_map4(c, &c->amap2[li][0], base, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
            if (rc) return rc;
        } else {
            const cuuint64_t C = L.in_c, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {C, W / 2, H / 2, N};
            const cuuint64_t strides[3] = {2 * C * 4, 2 * W * C * 4, H * W * C * 4};
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    const cuuint32_t box[4] = {32, 11, py ? 12u : 11u, 1};
                    snprintf(what, sizeof what, "v2 layer %d s2 parity %d%d", li, py, px);
                    int rc = make_map4(c, &c->amap2[li][py * 2 + px], base + (static_cast<size_t>(py) * W + px) * C, dims, strides, box,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                    if (rc) return rc;
                }
        }
    }
    return SE3TN_OK;
}

void fill_plan2(const se3tn_ctx* c, const LayerSpec& L, int first, int n, int block_n, Umma2Plan& t) {
    memset(&t, 0, sizeof t);
    t.chunks = L.cin / 32;
    t.bw = 11; t.bh = 11; t.bn = 1;
    t.img_first = first;
    t.base_off_mode = c->base_off_mode;
    t.debug = c->debug_flags;
    t.pdl = c->pdl;
    t.n_tiles = L.cout / block_n;
    const int Ho = layer_Ho(L);
    if (L.kind == K_STEM) {
        // pooled tile: 11x11 conv outputs starting at (2*5*ty - 1, 2*5*tx - 1) -> 5x5 pooled outputs
        t.units_per_chunk = 2;
        Unit& e = t.units[0]; e.map = 0; e.c1 = 0; e.c2 = 0; e.ntaps = 4; e.rows = 11 * 14;
        for (int j = 0; j < 4; ++j) { e.taps[j].row_shift = (int8_t)(j * 11); e.taps[j].w_tap = (int8_t)(2 * j); }
        Unit& o = t.units[1]; o.map = 1; o.c1 = 0; o.c2 = 0; o.ntaps = 3; o.rows = 11 * 13;
        for (int j = 0; j < 3; ++j) { o.taps[j].row_shift = (int8_t)(j * 11); o.taps[j].w_tap = (int8_t)(2 * j + 1); }
        t.step_x = t.step_y = 10; t.off_x = t.off_y = -1;
        t.tiles_x = t.tiles_y = 9;
    } else if (L.kind == K_S1) {
        t.units_per_chunk = 3;
        for (int s2 = 0; s2 < 3; ++s2) {
            Unit& u = t.units[s2]; u.map = 0; u.c1 = (int8_t)(s2 - 1); u.c2 = -1; u.ntaps = 3; u.rows = 11 * 13;
            for (int r = 0; r < 3; ++r) { u.taps[r].row_shift = (int8_t)(r * 11); u.taps[r].w_tap = (int8_t)(r * 3 + s2); }
        }
        t.step_x = t.step_y = 11; t.off_x = t.off_y = 0;
        t.tiles_x = t.tiles_y = Ho / 11;
    } else {
        t.units_per_chunk = 6;
        for (int s2 = 0; s2 < 3; ++s2) {
            const int px = (s2 == 1) ? 0 : 1;
            const int c1 = (s2 == 0) ? -1 : 0;
            Unit& ev = t.units[s2 * 2]; ev.map = (int8_t)(0 * 2 + px); ev.c1 = (int8_t)c1; ev.c2 = 0; ev.ntaps = 1; ev.rows = 11 * 11;
            ev.taps[0].row_shift = 0; ev.taps[0].w_tap = (int8_t)(1 * 3 + s2);
            Unit& od = t.units[s2 * 2 + 1]; od.map = (int8_t)(1 * 2 + px); od.c1 = (int8_t)c1; od.c2 = -1; od.ntaps = 2; od.rows = 11 * 12;
            od.taps[0].row_shift = 0;  od.taps[0].w_tap = (int8_t)(0 * 3 + s2);
            od.taps[1].row_shift = 11; od.taps[1].w_tap = (int8_t)(2 * 3 + s2);
        }
        t.step_x = t.step_y = 11; t.off_x = t.off_y = 0;
        t.tiles_x = t.tiles_y = Ho / 11;
    }
    t.m_tiles = n * t.tiles_x * t.tiles_y;
}

void fill_geom(const LayerSpec& L, int n, bool round_tf32, ConvGeom& g) {
    memset(&g, 0, sizeof g);
    g.Hin = L.Hin; g.Win = L.Win; g.in_cstride = L.in_c; g.in_coff = 0;
    g.Ho = layer_Ho(L); g.Wo = g.Ho;
    g.stride = (L.kind == K_S1) ? 1 : 2;
    g.cin = L.cin; g.cout = L.cout; g.groups = L.groups;
    g.num_taps = layer_taps(L);
    g.n_img = n;
    if (L.kind == K_STEM) {
        // tap r: padded input row 2*oy + r, 32 contiguous floats from padded x = 2*ox
        for (int r = 0; r < 7; ++r) { g.taps[r].dy = (int16_t)r; g.taps[r].dx = 0; g.taps[r].map = (int8_t)r; g.taps[r].c1 = 0; g.taps[r].c2 = 0; }
    } else {
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                Tap& t = g.taps[r * 3 + s];
                t.dy = (int16_t)(r - 1); t.dx = (int16_t)(s - 1);
                if (L.kind == K_S1) { t.map = 0; t.c1 = (int8_t)(s - 1); t.c2 = (int8_t)(r - 1); }
                else {
                    // iy = 2*oy + dy: dy=-1 -> odd row oy-1; dy=0 -> even row oy; dy=+1 -> odd row oy
                    const int py = (r == 1) ? 0 : 1, px = (s == 1) ? 0 : 1;
                    t.map = (int8_t)(py * 2 + px);
                    t.c2 = (int8_t)(r == 0 ? -1 : 0); t.c1 = (int8_t)(s == 0 ? -1 : 0);
                }
            }
    }
    g.out_cstride = L.out_c; g.out_coff = L.out_coff;
    g.res_cstride = (L.res != NONE) ? (L.res == B_H1 ? 1024 : (L.res == B_F1 ? 256 : 64)) : 0;
    g.res_coff = 0;
    g.act = L.act;
    g.round_tf32 = round_tf32 ? 1 : 0;
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) dst[i] = ptx::to_tf32(src[i]);
}

int sync_stats(se3tn_ctx* c, cudaStream_t s) {
    if (!c->stats_dirty) return SE3TN_OK;
    int max_id = -1, f64 = -1;
    for (auto& kv : c->weights) if (kv.second.has_stats) {
        if (kv.first > max_id) max_id = kv.first;
        if (f64 < 0) f64 = kv.second.stats_f64;
        else if (f64 != kv.second.stats_f64) return fail(c, SE3TN_ERR_STATE, "all weight sets must use the same mean/std dtype");
    }
    if (max_id < 0) return fail(c, SE3TN_ERR_STATE, "se3tn_set_stats has not been called");
    const int rows = max_id + 1;
    if (rows > c->stats_rows) {
        cudaFree(c->d_mean32); cudaFree(c->d_std32); cudaFree(c->d_mean64); cudaFree(c->d_std64);
        CU_TRY(c, cudaMalloc(&c->d_mean32, rows * 8 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&c->d_std32, rows * 8 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&c->d_mean64, rows * 8 * sizeof(double)));
        CU_TRY(c, cudaMalloc(&c->d_std64, rows * 8 * sizeof(double)));
        c->stats_rows = rows;
    }
    std::vector<float> m32(rows * 8, 0.f), s32(rows * 8, 1.f);
    std::vector<double> m64(rows * 8, 0.0), s64(rows * 8, 1.0);
    for (auto& kv : c->weights) if (kv.second.has_stats && kv.first >= 0) {
        memcpy(&m32[kv.first * 8], kv.second.mean32, sizeof(float) * 8); memcpy(&s32[kv.first * 8], kv.second.std32, sizeof(float) * 8);
        memcpy(&m64[kv.first * 8], kv.second.mean64, sizeof(double) * 8); memcpy(&s64[kv.first * 8], kv.second.std64, sizeof(double) * 8);
    }
    // synchronous copies from stack-lifetime host vectors (rare: only when stats change)
    CU_TRY(c, cudaStreamSynchronize(s));
    CU_TRY(c, cudaMemcpy(c->d_mean32, m32.data(), rows * 8 * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_std32, s32.data(), rows * 8 * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_mean64, m64.data(), rows * 8 * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_std64, s64.data(), rows * 8 * sizeof(double), cudaMemcpyHostToDevice));
    c->stats_f64 = f64; c->stats_dirty = false;
    return SE3TN_OK;
}

int sync_tables(se3tn_ctx* c, cudaStream_t s) {
    if (!c->tables_dirty) return SE3TN_OK;
    int max_id = -1;
    for (auto& kv : c->weights) if (kv.second.dev && kv.first > max_id) max_id = kv.first;
    if (max_id < 0) return fail(c, SE3TN_ERR_STATE, "no weight set loaded");
    const int rows = max_id + 1;
    CU_TRY(c, cudaStreamSynchronize(s));
    if (rows > c->table_rows) {
        cudaFree(c->d_bmaps_tf32); cudaFree(c->d_bmaps_bf16); cudaFree(c->d_bmaps_x3); cudaFree(c->d_bias); cudaFree(c->d_fc);
        CU_TRY(c, cudaMalloc(&c->d_bmaps_x3, sizeof(CUtensorMap) * rows * kLayersPerSet));
        CU_TRY(c, cudaMalloc(&c->d_bmaps_tf32, sizeof(CUtensorMap) * rows * kLayersPerSet));
        CU_TRY(c, cudaMalloc(&c->d_bmaps_bf16, sizeof(CUtensorMap) * rows * kLayersPerSet));
        CU_TRY(c, cudaMalloc(&c->d_bias, sizeof(float*) * rows * kLayersPerSet));
        CU_TRY(c, cudaMalloc(&c->d_fc, sizeof(float*) * rows));
        c->table_rows = rows;
    }
    std::vector<CUtensorMap> m1(rows * kLayersPerSet), m2(rows * kLayersPerSet), m3(rows * kLayersPerSet);
    memset(m3.data(), 0, m3.size() * sizeof(CUtensorMap));
    std::vector<const float*> bias(rows * kLayersPerSet, nullptr), fc(rows, nullptr);
    memset(m1.data(), 0, m1.size() * sizeof(CUtensorMap)); memset(m2.data(), 0, m2.size() * sizeof(CUtensorMap));
    for (auto& kv : c->weights) {
        if (!kv.second.dev || kv.first < 0) continue;
        for (int li = 0; li < kLayersPerSet; ++li) {
            m1[kv.first * kLayersPerSet + li] = (li < 8) ? kv.second.bmap_res[li] : kv.second.bmap[li];
            // bf16: the stem always runs stacked; bf16x3: every resident-weight layer does
            m2[kv.first * kLayersPerSet + li] = (li < 2) ? kv.second.bmap_stack[li] : (li < 8 ? kv.second.bmap_res_bf16[li] : kv.second.bmap_bf16[li]);
            m3[kv.first * kLayersPerSet + li] = (li < 8) ? kv.second.bmap_stack[li] : kv.second.bmap_bf16[li];
            bias[kv.first * kLayersPerSet + li] = kv.second.dev + kv.second.b_off[li];
        }
        fc[kv.first] = kv.second.dev + kv.second.fc_off;
    }
    CU_TRY(c, cudaMemcpy(c->d_bmaps_tf32, m1.data(), m1.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_bmaps_bf16, m2.data(), m2.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_bmaps_x3, m3.data(), m3.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_bias, bias.data(), bias.size() * sizeof(float*), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(c->d_fc, fc.data(), fc.size() * sizeof(float*), cudaMemcpyHostToDevice));
    c->tables_dirty = false;
    return SE3TN_OK;
}

// The conv stack on images [first, first+n) of the context buffers.  img_wid (device, indexed by absolute image
// index) non-null: every image uses its own weight set in the same launches (tensor-core modes, v2 kernel);
// `weight_id` is then only a representative loaded set.
int run_network(se3tn_ctx* c, int weight_id, int first, int n, int precision,
                float* out_trans, float* out_rot, float* out_feature, cudaStream_t s, const int* img_wid = nullptr) {
    auto it = c->weights.find(weight_id);
    if (it == c->weights.end()) return fail(c, SE3TN_ERR_STATE, "weight set " + std::to_string(weight_id) + " not loaded");
    const WeightSet& ws = it->second;
    const bool bf16 = (precision == SE3TN_PREC_BF16X3 || precision == SE3TN_PREC_BF16);
    const bool tf32 = (precision == SE3TN_PREC_TF32);
    const bool tensor = tf32 || bf16;
    if (!tensor && precision != SE3TN_PREC_FP32) return fail(c, SE3TN_ERR_INVALID, "unknown precision");
    if (bf16 && c->conv_version != 2) return fail(c, SE3TN_ERR_INVALID, "bf16 modes need the v2 conv kernel (unset SE3TN_CONV)");
    const int kprec = tf32 ? 0 : (precision == SE3TN_PREC_BF16X3 ? 1 : 2);
    if (img_wid) {
        if (!tensor || c->conv_version != 2) return fail(c, SE3TN_ERR_INVALID, "multi-weight-set launches need a tensor-core precision and the v2 kernel");
        int rc = sync_tables(c, s); if (rc) return rc;
    }
    const float* wbase = tf32 ? ws.dev_tf32 : ws.dev;
    const bool fused_pool = tensor && c->conv_version == 2 && c->fuse_pool && !c->dual_m;
    auto bufp = [&](Buf b) { return c->buf[b] + kBufFloats[b] * static_cast<size_t>(first); };
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        ConvGeom g; fill_geom(L, n, tf32, g);
        ConvPtrs p;
        p.in = bufp(L.in); p.out = bufp(L.out); p.res = (L.res != NONE) ? bufp(L.res) : nullptr;
        p.w = wbase + ws.w_off[li]; p.bias = ws.dev + ws.b_off[li];
        p.img_wid = img_wid;
        p.gbmaps = img_wid ? (precision == SE3TN_PREC_BF16X3 ? c->d_bmaps_x3 : (bf16 ? c->d_bmaps_bf16 : c->d_bmaps_tf32)) + li : nullptr;
        p.gbias = img_wid ? c->d_bias + li : nullptr;
        p.sk_part = nullptr; p.sk_flags = nullptr; p.pool_part = nullptr;
        p.trace = c->trace ? c->trace + static_cast<size_t>(li) * 256 * 8 : nullptr;
        if (tensor && c->conv_version == 2) {
            UmmaMaps maps;
            const int nmaps = (L.kind == K_STEM) ? 2 : (L.kind == K_S2 ? 4 : 1);
            for (int m = 0; m < 7; ++m) maps.a[m] = c->amap2[li][m < nmaps ? m : 0];
            maps.b = bf16 ? ws.bmap_bf16[li] : ws.bmap[li];
            const int BN = block_n_of(c, L);
            const bool pool = (L.kind == K_STEM);
            const bool resident = (BN == 64 && L.cout == 64);
            if (resident && li < 8) maps.b = bf16 ? ws.bmap_res_bf16[li] : ws.bmap_res[li];
            const bool pair = resident && c->pair && !img_wid && li < 8;
            if (pair) maps.b = bf16 ? ws.bmap_bf16_pair[li] : ws.bmap_pair[li];
            else if (resident && li < 8 && bf16 && (pool || precision == SE3TN_PREC_BF16X3)) maps.b = ws.bmap_stack[li];   // must mirror Cfg2::kStack
            Umma2Plan t; fill_plan2(c, L, first, n, BN, t);
            t.pair = pair ? 1 : 0;
            t.sk_seq = 0;
            if (li == 13 && fused_pool) p.pool_part = c->pool_part;   // indexed by absolute image (tile n0)
            if (BN == 256 && c->streamk && c->sk_part && !(c->dual_m && !img_wid)) {
                p.sk_part = c->sk_part; p.sk_flags = c->sk_flags;
                if (++c->sk_seq == 0x7fffffff) c->sk_seq = 1;
                t.sk_seq = c->sk_seq;
            }
            g.n_img = first + n;                       // absolute image indices (TMA maps address image 0)
            p.out = c->buf[L.out]; p.res = (L.res != NONE) ? c->buf[L.res] : nullptr;
            if (pool) {                                // fused MaxPool2d(3,2,1): write the pooled tensor directly
                g.Ho = g.Wo = 44;
                p.out = c->buf[li == 0 ? B_P1A : B_P1B];
                g.out_cstride = 64; g.out_coff = 0;
            }
            { ProfScope ps(c, li, s); CU_TRY(c, launch_conv_umma2(maps, g, t, p, BN, resident, L.kind == K_STEM ? KIND_STEM : (L.kind == K_S2 ? KIND_S2 : KIND_S1),
                                                                     (BN == 256 && c->dual_m && !img_wid) ? 2 : 1, kprec, c->num_sms, s)); }
            ++c->launches;
            continue;
        }
        if (tf32) {
            UmmaMaps maps;
            const int nmaps = (L.kind == K_STEM) ? 7 : (L.kind == K_S2 ? 4 : 1);
            for (int m = 0; m < 7; ++m) maps.a[m] = c->amap[li][m < nmaps ? m : 0];
            maps.b = ws.bmap[li];
            UmmaTiling t;
            tile_box(L, t.bw, t.bh, t.bn);
            t.tiles_x = g.Wo / t.bw; t.tiles_y = g.Ho / t.bh;
            const int BN = block_n_of(c, L);
            t.n_tiles = L.cout / BN;
            t.chunks_per_tap = L.cin / 32;
            t.m_tiles = ((n + t.bn - 1) / t.bn) * t.tiles_x * t.tiles_y;
            // The TMA maps address images absolutely (image 0 of the buffer), so this path works in
            // absolute image indices: tiles start at image `first`, rows are valid below first + n.
            t.img_first = first;
            g.n_img = first + n;
            p.out = c->buf[L.out]; p.res = (L.res != NONE) ? c->buf[L.res] : nullptr;
            { ProfScope ps(c, li, s); CU_TRY(c, launch_conv_umma(maps, g, t, p, BN, c->num_sms, s)); }
        } else {
            { ProfScope ps(c, li, s); CU_TRY(c, launch_conv_direct(g, p, s)); }
        }
        ++c->launches;
        if (li == 0) { ProfScope ps(c, 14, s); CU_TRY(c, launch_maxpool(bufp(B_Y1A), bufp(B_P1A), n, 88, 88, 64, s)); ++c->launches; }
        if (li == 1) { ProfScope ps(c, 15, s); CU_TRY(c, launch_maxpool(bufp(B_Y1B), bufp(B_P1B), n, 88, 88, 64, s)); ++c->launches; }
    }
    if (fused_pool) {
        ProfScope ps(c, 16, s);
        CU_TRY(c, launch_head_pooled(c->pool_part + static_cast<size_t>(first) * 4 * 1024, ws.dev + ws.fc_off, ws.dev + ws.fc_off + 6 * 512, out_trans, out_rot, n, 121,
                                     img_wid ? img_wid + first : nullptr, img_wid ? c->d_fc : nullptr, s));
    } else { ProfScope ps(c, 16, s); CU_TRY(c, launch_head(bufp(B_H3), ws.dev + ws.fc_off, ws.dev + ws.fc_off + 6 * 512, out_trans, out_rot, n, 121, bf16 ? 1 : 0,
                                                   img_wid ? img_wid + first : nullptr, img_wid ? c->d_fc : nullptr, s)); }
    ++c->launches;
    if (out_feature) { CU_TRY(c, launch_nhwc_to_nchw(bufp(B_F2), out_feature, n, 22 * 22, 256, bf16 ? 1 : 0, s)); ++c->launches; }
    return SE3TN_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

size_t se3tn_workspace_bytes(int max_batch) {
    if (max_batch <= 0) return 0;
    return workspace_floats(max_batch) * sizeof(float);
}

const char* se3tn_last_error(se3tn_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int se3tn_create(int device, int max_batch, void* workspace, se3tn_ctx** out) {
    if (!out || max_batch <= 0) return fail(nullptr, SE3TN_ERR_INVALID, "se3tn_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, SE3TN_ERR_CUDA, std::string("se3tn_create: no such CUDA device: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    CU_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, SE3TN_ERR_UNSUPPORTED, "se3tn_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                    ", this library is sm_100a only (no fallback path)");
    CU_TRY(nullptr, cudaSetDevice(device));
    se3tn_ctx* c = new se3tn_ctx();
    c->device = device; c->max_batch = max_batch; c->num_sms = prop.multiProcessorCount;
    if (const char* ov = getenv("SE3TN_BLOCK_N")) c->umma_block_n_override = atoi(ov);
    if (const char* ov = getenv("SE3TN_CONV")) c->conv_version = atoi(ov) == 1 ? 1 : 2;
    if (const char* ov = getenv("SE3TN_BASE_OFF")) c->base_off_mode = atoi(ov);
    if (const char* ov = getenv("SE3TN_DUAL_M")) c->dual_m = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_DEBUG_SKIP")) c->debug_flags = atoi(ov);
    if (const char* ov = getenv("SE3TN_PDL")) c->pdl = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_PAIR")) c->pair = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_STREAMK")) c->streamk = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_FUSE_POOL")) c->fuse_pool = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_TRACE")) {
        if (atoi(ov) != 0 && cudaMalloc(&c->trace, 14 * 256 * 8 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(c->trace, 0, 14 * 256 * 8 * sizeof(unsigned long long));
    }

    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete c; return fail(nullptr, SE3TN_ERR_CUDA, "se3tn_create: cuTensorMapEncodeTiled entry point unavailable");
    }
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);

    const size_t bytes = se3tn_workspace_bytes(max_batch);
    if (workspace) {
        if (reinterpret_cast<uintptr_t>(workspace) % 1024) { delete c; return fail(nullptr, SE3TN_ERR_INVALID, "se3tn_create: workspace must be 1024-byte aligned"); }
        c->workspace = static_cast<uint8_t*>(workspace);
    } else {
        e = cudaMalloc(&c->workspace, bytes);
        if (e != cudaSuccess) { delete c; return fail(nullptr, SE3TN_ERR_NOMEM, std::string("se3tn_create: cudaMalloc(workspace): ") + cudaGetErrorString(e)); }
        c->own_workspace = true;
    }
    e = cudaMalloc(&c->pool_part, static_cast<size_t>(max_batch) * 4 * 1024 * sizeof(float));
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_NOMEM, "se3tn_create: pool buffer: " + m); }
    if (c->streamk) {
        e = cudaMalloc(&c->sk_part, static_cast<size_t>(c->num_sms) * 128 * 256 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->sk_flags, (c->num_sms + 1) * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(c->sk_flags, 0, (c->num_sms + 1) * sizeof(int));
        if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_NOMEM, "se3tn_create: stream-K workspace: " + m); }
    }
    // zero once: the stem buffers' 3-pixel halo is the conv padding and is never written again
    e = cudaMemset(c->workspace, 0, bytes);
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_CUDA, "se3tn_create: cudaMemset: " + m); }
    float* p = reinterpret_cast<float*>(c->workspace);
    for (int b = 0; b < B_COUNT; ++b) {
        c->buf[b] = p;
        p += (kBufFloats[b] * static_cast<size_t>(max_batch) + 255) & ~size_t(255);
    }
    int rc = build_activation_maps(c);
    if (!rc) rc = build_activation_maps_v2(c);
    if (rc) { g_create_error = c->err; se3tn_destroy(c); return rc; }
    *out = c;
    return SE3TN_OK;
}

void se3tn_destroy(se3tn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& kv : c->weights) { cudaFree(kv.second.dev); cudaFree(kv.second.dev_tf32); cudaFree(kv.second.dev_bf16); cudaFree(kv.second.dev_stem_bf16); cudaFree(kv.second.dev_stack); cudaFree(kv.second.dev_perm); cudaFree(kv.second.dev_perm_tmp); }
    cudaFree(c->d_mean32); cudaFree(c->d_std32); cudaFree(c->d_mean64); cudaFree(c->d_std64);
    cudaFree(c->d_bmaps_tf32); cudaFree(c->d_bmaps_bf16); cudaFree(c->d_bmaps_x3); cudaFree(c->d_bias); cudaFree(c->d_fc);
    for (int i = 0; i < SE3TN_PROFILE_SLOTS; ++i) { if (c->ev0[i]) cudaEventDestroy(c->ev0[i]); if (c->ev1[i]) cudaEventDestroy(c->ev1[i]); }
    cudaFree(c->sk_part); cudaFree(c->sk_flags); cudaFree(c->pool_part); cudaFree(c->trace);
    for (auto& kv : c->meshes) { cudaFree(const_cast<float*>(kv.second.pos)); cudaFree(const_cast<float*>(kv.second.nrm)); cudaFree(const_cast<uint8_t*>(kv.second.col)); cudaFree(const_cast<int*>(kv.second.faces)); }
    cudaFree(c->d_meshes); cudaFree(c->render_proj); cudaFree(c->render_unif);
    cudaFree(c->fill.a); cudaFree(c->fill.b); cudaFree(c->fill.lut); cudaFree(c->fill.minmax);
    if (c->own_workspace) cudaFree(c->workspace);
    delete c;
}

int se3tn_load_weights(se3tn_ctx* c, int weight_id, const float* blob, size_t n_floats) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!blob || weight_id < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_load_weights: bad arguments");
    const size_t expect = blob_floats();
    if (n_floats != expect || expect != SE3TN_WEIGHT_BLOB_FLOATS)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_load_weights: blob has " + std::to_string(n_floats) + " floats, expected " + std::to_string(expect));
    CU_TRY(c, cudaSetDevice(c->device));
    WeightSet& ws = c->weights[weight_id];
    if (!ws.dev) {
        CU_TRY(c, cudaMalloc(&ws.dev, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_tf32, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_bf16, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_stem_bf16, 2 * 64 * 448 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_stack, 8 * 128 * 288 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_perm, 6 * 2 * 64 * 576 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_perm_tmp, 64 * 576 * sizeof(float)));
    }
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemcpy(ws.dev, blob, expect * sizeof(float), cudaMemcpyHostToDevice));
    round_tf32_kernel<<<1024, 256>>>(ws.dev, ws.dev_tf32, expect);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaDeviceSynchronize());
    size_t off = 0;
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        ws.w_off[li] = off; off += static_cast<size_t>(layer_rows(L)) * layer_ktot(L);
        ws.b_off[li] = off; off += layer_rows(L);
        char what[48]; snprintf(what, sizeof what, "layer %d weights", li);
        int rc = make_map2(c, &ws.bmap[li], ws.dev_tf32 + ws.w_off[li], layer_ktot(L), layer_rows(L), block_n_of(c, L), what);
        if (rc) return rc;
        snprintf(what, sizeof what, "layer %d bf16 weights", li);
        if (L.kind == K_STEM) {
            uint8_t* dst = ws.dev_stem_bf16 + static_cast<size_t>(li) * 64 * 448 * sizeof(float);
            CU_TRY(c, launch_split_stem_weights(ws.dev + ws.w_off[li], dst, 0));
            rc = make_map2(c, &ws.bmap_bf16[li], dst, 448, 64, 64, what);
        } else {
            uint8_t* dst = ws.dev_bf16 + ws.w_off[li] * sizeof(float);
            CU_TRY(c, launch_split_weights(ws.dev + ws.w_off[li], dst, static_cast<size_t>(layer_rows(L)) * layer_ktot(L), 0));
            rc = make_map2(c, &ws.bmap_bf16[li], dst, layer_ktot(L), layer_rows(L), block_n_of(c, L), what);
        }
        if (rc) return rc;
        if (L.cout == 64 && li < 8) {
            // resident-weight layers of the v2 kernel.  Stems keep the natural row order; the 64-channel 3x3 layers use the
            // row order of the 16x256b epilogue (all three precisions, single-CTA and pair maps alike).
            const bool stem = (L.kind == K_STEM);
            const float* wsrc = ws.dev + ws.w_off[li];
            const float* w_tf32 = ws.dev_tf32 + ws.w_off[li];
            const uint8_t* w_bf16 = stem ? ws.dev_stem_bf16 + static_cast<size_t>(li) * 64 * 448 * sizeof(float) : ws.dev_bf16 + ws.w_off[li] * sizeof(float);
            if (!stem) {
                if (layer_ktot(L) != 576 || layer_rows(L) != 64) return fail(c, SE3TN_ERR_STATE, "resident layer shape");
                float* pt = ws.dev_perm + static_cast<size_t>(li - 2) * 2 * 64 * 576;
                CU_TRY(c, launch_permute_rows64(wsrc, ws.dev_perm_tmp, 576, 0));
                round_tf32_kernel<<<144, 256>>>(ws.dev_perm_tmp, pt, 64 * 576);
                CU_TRY(c, cudaGetLastError());
                CU_TRY(c, launch_split_weights(ws.dev_perm_tmp, pt + 64 * 576, 64 * 576, 0));
                wsrc = ws.dev_perm_tmp; w_tf32 = pt; w_bf16 = reinterpret_cast<const uint8_t*>(pt + 64 * 576);
            }
            snprintf(what, sizeof what, "layer %d stacked weights", li);
            uint8_t* sdst = ws.dev_stack + static_cast<size_t>(li) * 128 * 288 * sizeof(float);
            CU_TRY(c, launch_split_stack_weights(wsrc, sdst, stem, 0));
            rc = make_map2(c, &ws.bmap_stack[li], sdst, stem ? 224 : 288, 128, 128, what);
            if (rc) return rc;
            snprintf(what, sizeof what, "layer %d resident weights", li);
            rc = make_map2(c, &ws.bmap_res[li], w_tf32, layer_ktot(L), 64, 64, what);
            if (!rc) rc = make_map2(c, &ws.bmap_res_bf16[li], w_bf16, stem ? 448 : layer_ktot(L), 64, 64, what);
            if (rc) return rc;
            snprintf(what, sizeof what, "layer %d pair weights", li);
            rc = make_map2(c, &ws.bmap_pair[li], w_tf32, layer_ktot(L), 64, 32, what);
            if (!rc) rc = make_map2(c, &ws.bmap_bf16_pair[li], w_bf16, stem ? 448 : layer_ktot(L), 64, 32, what);
            if (rc) return rc;
        }
    }
    CU_TRY(c, cudaDeviceSynchronize());
    ws.fc_off = off;
    c->tables_dirty = true;
    return SE3TN_OK;
}

int se3tn_set_stats(se3tn_ctx* c, int weight_id, const void* mean8, const void* std8, int is_f64) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!mean8 || !std8 || weight_id < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_set_stats: bad arguments");
    WeightSet& ws = c->weights[weight_id];
    for (int i = 0; i < 8; ++i) {
        if (is_f64) {
            ws.mean64[i] = static_cast<const double*>(mean8)[i]; ws.std64[i] = static_cast<const double*>(std8)[i];
            ws.mean32[i] = static_cast<float>(ws.mean64[i]); ws.std32[i] = static_cast<float>(ws.std64[i]);
        } else {
            ws.mean32[i] = static_cast<const float*>(mean8)[i]; ws.std32[i] = static_cast<const float*>(std8)[i];
            ws.mean64[i] = ws.mean32[i]; ws.std64[i] = ws.std32[i];
        }
    }
    ws.stats_f64 = is_f64 ? 1 : 0; ws.has_stats = true;
    c->stats_dirty = true;
    return SE3TN_OK;
}

int se3tn_preprocess(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                     const double* K, const double* poses, const double* object_width,
                     const uint8_t* rgbA, const uint16_t* depthA, const int32_t* weight_ids, int n,
                     int precision, float* out_A, float* out_B, uint8_t* crop_rgb, uint16_t* crop_depth, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!frame_rgb || !frame_depth || !K || !poses || !object_width || !rgbA || !depthA || H <= 0 || W <= 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_preprocess: null/invalid argument");
    if (n < 0 || n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_preprocess: n exceeds max_batch");
    if ((out_A == nullptr) != (out_B == nullptr)) return fail(c, SE3TN_ERR_INVALID, "se3tn_preprocess: out_A/out_B must both be given or both NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU_TRY(c, cudaSetDevice(c->device));
    int rc = sync_stats(c, s); if (rc) return rc;
    PreprocessArgs a;
    a.frame_rgb = frame_rgb; a.frame_depth = frame_depth; a.H = H; a.W = W;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
    a.poses = poses; a.object_width = object_width; a.rgbA = rgbA; a.depthA = depthA; a.weight_ids = weight_ids;
    a.mean32 = c->d_mean32; a.std32 = c->d_std32; a.mean64 = c->d_mean64; a.std64 = c->d_std64;
    a.stats_f64 = c->stats_f64; a.round_tf32 = store_mode_of(precision); a.b_precropped = 0;
    a.stemA = c->buf[B_X0A]; a.stemB = c->buf[B_X0B]; a.nchwA = out_A; a.nchwB = out_B;
    a.crop_rgb = crop_rgb; a.crop_depth = crop_depth;
    { ProfScope ps(c, 17, s); CU_TRY(c, launch_preprocess(a, n, s)); }
    ++c->launches;
    return SE3TN_OK;
}

int se3tn_normalize(se3tn_ctx* c, const uint8_t* rgbA, const uint16_t* depthA, const uint8_t* rgbB, const uint16_t* depthB,
                    const double* poses, const int32_t* weight_ids, int n, int precision, float* out_A, float* out_B, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!rgbA || !depthA || !rgbB || !depthB || !poses) return fail(c, SE3TN_ERR_INVALID, "se3tn_normalize: null argument");
    if (n < 0 || n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_normalize: n exceeds max_batch");
    if ((out_A == nullptr) != (out_B == nullptr)) return fail(c, SE3TN_ERR_INVALID, "se3tn_normalize: out_A/out_B must both be given or both NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU_TRY(c, cudaSetDevice(c->device));
    int rc = sync_stats(c, s); if (rc) return rc;
    PreprocessArgs a;
    memset(&a, 0, sizeof a);
    a.frame_rgb = rgbB; a.frame_depth = depthB; a.H = kImg; a.W = kImg; a.b_precropped = 1;
    a.poses = poses; a.rgbA = rgbA; a.depthA = depthA; a.weight_ids = weight_ids;
    a.mean32 = c->d_mean32; a.std32 = c->d_std32; a.mean64 = c->d_mean64; a.std64 = c->d_std64;
    a.stats_f64 = c->stats_f64; a.round_tf32 = store_mode_of(precision);
    a.stemA = c->buf[B_X0A]; a.stemB = c->buf[B_X0B]; a.nchwA = out_A; a.nchwB = out_B;
    { ProfScope ps(c, 17, s); CU_TRY(c, launch_preprocess(a, n, s)); }
    ++c->launches;
    return SE3TN_OK;
}

int se3tn_compute_bbox(se3tn_ctx* c, const double* poses, const double* K, const double* widths, const double* scale,
                       int32_t* out_bbox, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses || !K || !widths || !scale || !out_bbox || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_compute_bbox: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, launch_bbox(poses, K, widths, scale, out_bbox, n, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_crop_bbox(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W, const int32_t* bbox, int n,
                    int out_h, int out_w, uint8_t* crop_rgb, uint16_t* crop_depth, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!frame_rgb || !frame_depth || !bbox || !crop_rgb || !crop_depth || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || n < 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_crop_bbox: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, launch_crop(frame_rgb, frame_depth, H, W, bbox, n, out_h, out_w, crop_rgb, crop_depth, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_forward(se3tn_ctx* c, int weight_id, const float* A, const float* B, int n,
                  float* out_trans, float* out_rot, float* out_feature, int precision, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!A || !B || !out_trans || !out_rot) return fail(c, SE3TN_ERR_INVALID, "se3tn_forward: null argument");
    if (n < 0 || n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_forward: n exceeds max_batch");
    if (n == 0) return SE3TN_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU_TRY(c, cudaSetDevice(c->device));
    c->launches = 0;
    const int round = store_mode_of(precision);
    { ProfScope ps(c, 19, s);
      CU_TRY(c, launch_nchw_to_stem(A, c->buf[B_X0A], n, round, s));
      CU_TRY(c, launch_nchw_to_stem(B, c->buf[B_X0B], n, round, s)); }
    c->launches += 2;
    return run_network(c, weight_id, 0, n, precision, out_trans, out_rot, out_feature, s);
}

int se3tn_forward_preprocessed(se3tn_ctx* c, int weight_id, int first, int n,
                               float* out_trans, float* out_rot, float* out_feature, int precision, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out_trans || !out_rot) return fail(c, SE3TN_ERR_INVALID, "se3tn_forward_preprocessed: null argument");
    if (first < 0 || n < 0 || first + n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_forward_preprocessed: range exceeds max_batch");
    if (n == 0) return SE3TN_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    return run_network(c, weight_id, first, n, precision, out_trans, out_rot, out_feature, static_cast<cudaStream_t>(stream));
}

int se3tn_pose_update(se3tn_ctx* c, const double* poses_in, const float* trans, const float* rot,
                      double tn, double rn, double* poses_out, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses_in || !trans || !rot || !poses_out || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_pose_update: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    { ProfScope ps(c, 18, static_cast<cudaStream_t>(stream)); CU_TRY(c, launch_pose_update(poses_in, trans, rot, static_cast<float>(tn), static_cast<float>(rn), poses_out, n, static_cast<cudaStream_t>(stream))); }
    ++c->launches;
    return SE3TN_OK;
}

int se3tn_so3_log(se3tn_ctx* c, const double* poses_a, const double* poses_b, double tn, double rn,
                  double* trans_label, double* rot_label, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses_a || !poses_b || !trans_label || !rot_label || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_so3_log: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, launch_so3_log(poses_a, poses_b, tn, rn, trans_label, rot_label, n, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_track_batch(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                      const double* K, const double* poses_in, const double* object_width,
                      const uint8_t* rgbA, const uint16_t* depthA,
                      const int32_t* weight_ids_host, const int32_t* weight_ids_dev, int n,
                      double tn, double rn, int precision,
                      float* out_trans, float* out_rot, double* poses_out, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out_trans || !out_rot || !poses_out) return fail(c, SE3TN_ERR_INVALID, "se3tn_track_batch: null output");
    if ((weight_ids_host == nullptr) != (weight_ids_dev == nullptr))
        return fail(c, SE3TN_ERR_INVALID, "se3tn_track_batch: weight_ids_host and weight_ids_dev must both be given or both NULL");
    c->launches = 0;
    int rc = se3tn_preprocess(c, frame_rgb, frame_depth, H, W, K, poses_in, object_width, rgbA, depthA, weight_ids_dev, n,
                              precision, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    bool multi = false;
    if (weight_ids_host) {
        for (int i = 0; i < n; ++i) {
            if (c->weights.find(weight_ids_host[i]) == c->weights.end() || !c->weights[weight_ids_host[i]].dev)
                return fail(c, SE3TN_ERR_STATE, "weight set " + std::to_string(weight_ids_host[i]) + " not loaded");
            if (weight_ids_host[i] != weight_ids_host[0]) multi = true;
        }
    }
    if (multi && precision != SE3TN_PREC_FP32 && c->conv_version == 2) {
        // every track picks its own weight set inside the same 14 conv launches
        rc = run_network(c, weight_ids_host[0], 0, n, precision, out_trans, out_rot, nullptr, static_cast<cudaStream_t>(stream), weight_ids_dev);
        if (rc) return rc;
    } else {
        int first = 0;
        while (first < n) {
            const int wid = weight_ids_host ? weight_ids_host[first] : 0;
            int last = first + 1;
            while (last < n && (weight_ids_host ? weight_ids_host[last] : 0) == wid) ++last;
            rc = run_network(c, wid, first, last - first, precision, out_trans + first * 3, out_rot + first * 3, nullptr,
                             static_cast<cudaStream_t>(stream));
            if (rc) return rc;
            first = last;
        }
    }
    return se3tn_pose_update(c, poses_in, out_trans, out_rot, tn, rn, poses_out, n, stream);
}

int se3tn_add_adi(se3tn_ctx* c, const double* model_pts, int m, const double* pred, const double* gt, int n,
                  double* out_add, double* out_adi, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!model_pts || !pred || !gt || m <= 0 || n < 0 || (!out_add && !out_adi)) return fail(c, SE3TN_ERR_INVALID, "se3tn_add_adi: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, launch_add_adi(model_pts, m, pred, gt, n, out_add, out_adi, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_vocap(se3tn_ctx* c, const double* errs, int n, double* out_ap, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out_ap || n < 0 || (n > 0 && !errs)) return fail(c, SE3TN_ERR_INVALID, "se3tn_vocap: null/invalid argument");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, vocap(errs, n, out_ap, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_allgather_poses(se3tn_ctx* c, void* nccl_comm, const double* local_poses, double* all_poses, int n_local, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!nccl_comm || n_local < 0 || (n_local > 0 && (!local_poses || !all_poses))) return fail(c, SE3TN_ERR_INVALID, "se3tn_allgather_poses: bad arguments");
    if (n_local == 0) return SE3TN_OK;
    // ncclResult_t ncclAllGather(const void* send, void* recv, size_t sendcount, ncclDataType_t, ncclComm_t, cudaStream_t)
    typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
    static AllGatherFn fn = nullptr;
    if (!fn) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy torch (or the host) already loaded
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) fn = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
        if (!fn) return fail(c, SE3TN_ERR_UNSUPPORTED, "se3tn_allgather_poses: libnccl.so.2 / ncclAllGather not found");
    }
    CU_TRY(c, cudaSetDevice(c->device));
    const int kNcclFloat64 = 8;
    const int rc = fn(local_poses, all_poses, static_cast<size_t>(n_local) * 16, kNcclFloat64, nccl_comm, static_cast<cudaStream_t>(stream));
    if (rc != 0) return fail(c, SE3TN_ERR_CUDA, "se3tn_allgather_poses: ncclAllGather returned " + std::to_string(rc));
    return SE3TN_OK;
}

int se3tn_fill_depth(se3tn_ctx* c, const uint16_t* depth_mm, int H, int W, double max_depth,
                     uint16_t* out_mm, float* out_m, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!depth_mm || H <= 0 || W <= 0 || (!out_mm && !out_m)) return fail(c, SE3TN_ERR_INVALID, "se3tn_fill_depth: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t px = static_cast<size_t>(H) * W;
    if (px > c->fill_pixels) {
        CU_TRY(c, cudaStreamSynchronize(s));
        cudaFree(c->fill.a); cudaFree(c->fill.b); c->fill.a = c->fill.b = nullptr; c->fill_pixels = 0;
        CU_TRY(c, cudaMalloc(&c->fill.a, px * sizeof(float)));
        CU_TRY(c, cudaMalloc(&c->fill.b, px * sizeof(float)));
        c->fill_pixels = px;
    }
    if (!c->fill.lut) {
        CU_TRY(c, cudaMalloc(&c->fill.lut, (kFillLutEntries + 1) * sizeof(float)));
        CU_TRY(c, cudaMalloc(&c->fill.minmax, 2 * sizeof(unsigned)));
    }
    CU_TRY(c, launch_fill_depth(depth_mm, H, W, static_cast<float>(max_depth), c->fill, out_mm, out_m, s));
    c->launches += 8;
    return SE3TN_OK;
}

int se3tn_set_mesh(se3tn_ctx* c, int mesh_id, const float* pos, const float* nrm, const uint8_t* col,
                   const int32_t* faces, int nv, int nf) {
    if (!c) return SE3TN_ERR_INVALID;
    if (mesh_id < 0 || mesh_id > 4095 || !pos || !nrm || !col || !faces || nv <= 0 || nf <= 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_set_mesh: bad arguments");
    for (int i = 0; i < 3 * nf; ++i)
        if (faces[i] < 0 || faces[i] >= nv) return fail(c, SE3TN_ERR_INVALID, "se3tn_set_mesh: face index out of range");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    MeshDev& m = c->meshes[mesh_id];
    cudaFree(const_cast<float*>(m.pos)); cudaFree(const_cast<float*>(m.nrm)); cudaFree(const_cast<uint8_t*>(m.col)); cudaFree(const_cast<int*>(m.faces));
    m = MeshDev{};
    float* dpos; float* dnrm; uint8_t* dcol; int* dfaces;
    CU_TRY(c, cudaMalloc(&dpos, sizeof(float) * 3 * nv)); m.pos = dpos;
    CU_TRY(c, cudaMalloc(&dnrm, sizeof(float) * 3 * nv)); m.nrm = dnrm;
    CU_TRY(c, cudaMalloc(&dcol, 3 * static_cast<size_t>(nv))); m.col = dcol;
    CU_TRY(c, cudaMalloc(&dfaces, sizeof(int) * 3 * nf)); m.faces = dfaces;
    CU_TRY(c, cudaMemcpy(dpos, pos, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(dnrm, nrm, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(dcol, col, 3 * static_cast<size_t>(nv), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(dfaces, faces, sizeof(int) * 3 * nf, cudaMemcpyHostToDevice));
    m.nv = nv; m.nf = nf;
    c->meshes_dirty = true;
    return SE3TN_OK;
}

int se3tn_render(se3tn_ctx* c, const double* K, const double* poses, const double* object_width,
                 const int32_t* mesh_ids, int n, uint8_t* rgbA, uint16_t* depthA, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (n < 0 || !K || (n > 0 && (!poses || !object_width || !rgbA || !depthA))) return fail(c, SE3TN_ERR_INVALID, "se3tn_render: bad arguments");
    if (n == 0) return SE3TN_OK;
    if (n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_render: n exceeds the context's max_batch");
    if (c->meshes.empty()) return fail(c, SE3TN_ERR_STATE, "se3tn_render: no mesh loaded (se3tn_set_mesh)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->meshes_dirty) {
        const int rows = c->meshes.rbegin()->first + 1;
        CU_TRY(c, cudaStreamSynchronize(s));
        if (rows > c->mesh_rows) { cudaFree(c->d_meshes); CU_TRY(c, cudaMalloc(&c->d_meshes, sizeof(MeshDev) * rows)); c->mesh_rows = rows; }
        std::vector<MeshDev> tab(rows, c->meshes.begin()->second);      // unused ids alias the first model
        for (auto& kv : c->meshes) tab[kv.first] = kv.second;
        CU_TRY(c, cudaMemcpy(c->d_meshes, tab.data(), sizeof(MeshDev) * rows, cudaMemcpyHostToDevice));
        c->render_max_nv = 0;
        for (auto& kv : c->meshes) c->render_max_nv = std::max(c->render_max_nv, kv.second.nv);
        if (c->render_max_nv > c->render_proj_nv) {          // projected-vertex workspace: max_batch x largest model
            cudaFree(c->render_proj); c->render_proj = nullptr;
            CU_TRY(c, cudaMalloc(&c->render_proj, static_cast<size_t>(c->max_batch) * c->render_max_nv * render_projected_bytes_per_vertex()));
            c->render_proj_nv = c->render_max_nv;
        }
        if (!c->render_unif) CU_TRY(c, cudaMalloc(&c->render_unif, static_cast<size_t>(c->max_batch) * render_uniform_bytes()));
        c->meshes_dirty = false;
    }
    RenderArgs a;
    a.poses = poses; a.object_width = object_width; a.mesh_ids = mesh_ids; a.meshes = c->d_meshes; a.n_meshes = c->mesh_rows;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
    a.rgb = rgbA; a.depth = depthA;
    a.projected = c->render_proj; a.uniforms = c->render_unif; a.max_nv = c->render_max_nv;
    { ProfScope ps(c, 20, s); CU_TRY(c, launch_render(a, n, s)); }
    c->launches += 2;
    return SE3TN_OK;
}

int se3tn_debug_buffer(se3tn_ctx* c, int id, float** ptr, size_t* floats_per_image) {
    if (!c) return SE3TN_ERR_INVALID;
    if (id < 0 || id >= B_COUNT || !ptr || !floats_per_image) return fail(c, SE3TN_ERR_INVALID, "se3tn_debug_buffer: bad id");
    *ptr = c->buf[id]; *floats_per_image = kBufFloats[id];
    return SE3TN_OK;
}

int se3tn_last_launch_count(se3tn_ctx* c) { return c ? c->launches : 0; }

int se3tn_get_trace(se3tn_ctx* c, unsigned long long* out) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out) return fail(c, SE3TN_ERR_INVALID, "se3tn_get_trace: null argument");
    if (!c->trace) return fail(c, SE3TN_ERR_STATE, "se3tn_get_trace: the context was created without SE3TN_TRACE=1");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemcpy(out, c->trace, 14 * 256 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return SE3TN_OK;
}

int se3tn_set_profiling(se3tn_ctx* c, int enable) {
    if (!c) return SE3TN_ERR_INVALID;
    CU_TRY(c, cudaSetDevice(c->device));
    if (enable && !c->ev0[0]) {
        for (int i = 0; i < SE3TN_PROFILE_SLOTS; ++i) { CU_TRY(c, cudaEventCreate(&c->ev0[i])); CU_TRY(c, cudaEventCreate(&c->ev1[i])); }
    }
    c->profiling = enable != 0;
    for (int i = 0; i < SE3TN_PROFILE_SLOTS; ++i) c->ev_used[i] = false;
    return SE3TN_OK;
}

int se3tn_get_profile(se3tn_ctx* c, float* ms) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!ms) return fail(c, SE3TN_ERR_INVALID, "se3tn_get_profile: null argument");
    if (!c->ev0[0]) return fail(c, SE3TN_ERR_STATE, "se3tn_get_profile: profiling was never enabled");
    for (int i = 0; i < SE3TN_PROFILE_SLOTS; ++i) {
        ms[i] = 0.f;
        if (!c->ev_used[i]) continue;
        CU_TRY(c, cudaEventSynchronize(c->ev1[i]));
        CU_TRY(c, cudaEventElapsedTime(&ms[i], c->ev0[i], c->ev1[i]));
        c->ev_used[i] = false;
    }
    return SE3TN_OK;
}

}  // extern "C"

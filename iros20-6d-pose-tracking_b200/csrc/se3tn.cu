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

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace se3tn;

namespace {

// --------------------------------------------------------------------------------------------
// Network schedule: 14 conv layers cover the reference's 17 convs (se3_tracknet.py:57-78): the two heads' first convs
// are one layer with concatenated output channels, their basic blocks one grouped layer each.  Layers 0-7 are one
// conv_resident_kernel launch each; layers 8-13 are ONE conv_trunk_kernel launch.
// --------------------------------------------------------------------------------------------
enum Buf { B_X0A, B_X0B, B_Y1A, B_Y1B, B_P1A, B_P1B, B_T1, B_T2, B_U, B_CAT, B_F1, B_T4, B_F2, B_H1, B_H2, B_H3, B_COUNT };

constexpr size_t kBufFloats[B_COUNT] = {
    kStemImgFloats, kStemImgFloats,                 // X0A, X0B  (182 x 184 x 4)
    88 * 88 * 64, 88 * 88 * 64,                     // Y1A, Y1B  (fp32 mode only: the tensor-core stems pool in their epilogue)
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
constexpr int kFirstTrunkLayer = 8;

inline int layer_taps(const LayerSpec& L) { return L.kind == K_STEM ? 7 : 9; }
inline int layer_ktot(const LayerSpec& L) { return layer_taps(L) * L.cin; }
inline int layer_rows(const LayerSpec& L) { return L.cout * L.groups; }
inline int layer_Ho(const LayerSpec& L) { return L.kind == K_STEM ? 88 : (L.kind == K_S2 ? L.Hin / 2 : L.Hin); }
inline int res_channels(const LayerSpec& L) { return L.res == B_H1 ? 1024 : (L.res == B_F1 ? 256 : 64); }

constexpr size_t kFcFloats = 6 * 512 + 6;

size_t blob_floats() {
    size_t n = 0;
    for (const LayerSpec& L : kLayers) n += static_cast<size_t>(layer_rows(L)) * layer_ktot(L) + layer_rows(L);
    return n + kFcFloats;
}

struct WeightSet {
    float* dev = nullptr;           // exact fp32 blob (biases, fc and the fp32 mode's conv weights)
    float* dev_tf32 = nullptr;      // same layout, conv weights rounded to tf32
    uint8_t* dev_bf16 = nullptr;    // blob-sized: conv weights as [32 bf16 hi | 32 bf16 lo] per 32-word K chunk (PREC_BF16X3 ring layers)
    uint8_t* dev_h = nullptr;       // half-blob-sized: conv weights as plain bf16, K-major (PREC_BF16; layers 2..7 with permuted rows)
    uint8_t* dev_stack = nullptr;   // 8 x [128][288 words]: resident layers with hi / lo rows stacked along N (conv_umma2.cu STACK)
    float* dev_perm = nullptr; float* dev_perm_tmp = nullptr;   // 64-channel layers: rows in the 16x256b epilogue's channel order, tf32 words
    size_t w_off[14], b_off[14];
    size_t fc_off;
    // weight tensor maps per precision: [li] -> the map the kernel of that layer wants
    CUtensorMap bmap_tf32[kLayersPerSet], bmap_x3[kLayersPerSet], bmap_h[kLayersPerSet];   // rows 14..19: trunk layers with 128-row boxes
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
    CUtensorMap amap4[14][4];        // activation views, 4 bytes per channel (TF32 / BF16X3; also the stems' input in every mode)
    CUtensorMap amap2[14][4];        // activation views, 2 bytes per channel (PREC_BF16, layers 2..13)
    int pdl = 1;                     // SE3TN_PDL=0 disables programmatic dependent launch between the kernels of a step
    int stem_ws = 0;                 // SE3TN_STEM_WS=1: weights-stationary stem (conv_stem_t.cu); 3 = timing experiment without the pooling epilogue
    std::map<int, MeshDev> meshes;   // CAD models of the rasteriser (device copies), keyed by mesh id
    MeshDev* d_meshes = nullptr; int mesh_rows = 0; bool meshes_dirty = false;
    uint8_t* render_proj = nullptr; uint8_t* render_unif = nullptr; int render_max_nv = 0, render_proj_nv = 0;   // rasteriser workspace
    FillScratch fill = {nullptr, nullptr, nullptr, nullptr}; size_t fill_pixels = 0;   // depth hole-filling scratch (grows on demand)
    float* pool_part = nullptr;      // [max_batch][4][1024] column sums from the last conv's epilogue
    unsigned* sched = nullptr;       // trunk kernel: next-unit counter + done[6][max_batch] + split-K slice counters; zero between steps (head_pooled_kernel clears it)
    float* partial = nullptr;        // split-K scratch of the latency mode (n <= 4): trunk_partial_floats()
    bool sched_dirty = false;        // a step failed between the trunk launch and the head launch: clear before the next one
    unsigned long long* trace = nullptr;   // SE3TN_TRACE=1: [14 slots][256 CTAs][8] globaltimer stamps of the last forward (conv_umma2.cu trace_stamp)
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
    // CUDA graphs of whole track_batch steps (preprocess -> 8 resident convs -> trunk -> head + pose update), keyed by every baked-in argument
    int use_graphs = 1;              // SE3TN_GRAPH=0: plain stream launches; set to 0 at run time if capture is not possible
    bool last_was_graph = false;
    struct StepGraph { std::vector<unsigned long long> key; cudaGraphExec_t exec; int launches; unsigned long long last_use; };
    std::vector<StepGraph> graphs; unsigned long long graph_clock = 0;
    cudaStream_t cap_stream = nullptr;   // steps are captured on this private stream (the caller's may be the legacy default stream, which cannot be captured) and replayed on the caller's
    cudaEvent_t ev0[SE3TN_PROFILE_SLOTS] = {}, ev1[SE3TN_PROFILE_SLOTS] = {};
    bool ev_used[SE3TN_PROFILE_SLOTS] = {};
    // se3tn_track_host: context-owned pinned staging and device-side inputs / outputs (stable addresses -> the step's graph is reused)
    struct HostIO {
        uint8_t* pin = nullptr; size_t pin_bytes = 0;          // pinned host staging: inputs, then outputs
        uint8_t* dev = nullptr; size_t dev_bytes = 0;          // device: frame rgb | frame depth | poses | widths | rgbA | depthA | ids | out poses | out trans | out rot
        int H = 0, W = 0, n_cap = 0;
    } hio;
    std::string err;
};

namespace {

// conv-input storage mode of the packers for a precision: 0 raw fp32, 1 tf32 words, 2 bf16 hi/lo per pixel
inline int store_mode_of(int precision) {
    return precision == SE3TN_PREC_TF32 ? 1 : (precision == SE3TN_PREC_BF16X3 || precision == SE3TN_PREC_BF16) ? 2 : 0;
}
// kernel-side precision of a public one (-1: the fp32 FFMA mode)
inline int kprec_of(int precision) {
    return precision == SE3TN_PREC_TF32 ? PREC_TF32 : precision == SE3TN_PREC_BF16X3 ? PREC_BF16X3 : precision == SE3TN_PREC_BF16 ? PREC_BF16 : -1;
}

int fail(se3tn_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU_TRY(ctx, expr)                                                                        \
    do { cudaError_t e_ = (expr);                                                                \
         if (e_ != cudaSuccess) return fail((ctx), SE3TN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

// Every entry point works on the context's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

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

// rank-4 tensor map over 32-bit words, SWIZZLE_128B, box inner = 32 words
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

// weight matrix [rows][inner words] -> boxes of 32 words x box_rows
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

// Activation-side tensor maps: built once per context (they depend only on the workspace layout).  Boxes are extended
// by the vertical filter extent so the vertical taps become descriptor row shifts inside one shared-memory tile
// (conv_umma2.cu).  bpc = bytes per channel of the storage format (stems: always their 16-byte-per-pixel input).
int build_activation_maps(se3tn_ctx* c, int bpc, CUtensorMap (*out)[4]) {
    const cuuint64_t N = static_cast<cuuint64_t>(c->max_batch);
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        const uint8_t* base = reinterpret_cast<const uint8_t*>(c->buf[L.in]);
        char what[64];
        if (L.kind == K_STEM) {
            if (bpc != 4) continue;
            // even / odd input-row views of the zero-padded NHWC4 stem input; x is the overlapping
            // 8-pixel window view (stride 2 pixels = 32 B, extent 128 B)
            const cuuint64_t rowpitch = static_cast<cuuint64_t>(kStemW) * 4 * sizeof(float);
            const cuuint64_t strides[3] = {8 * sizeof(float), 2 * rowpitch, static_cast<cuuint64_t>(kStemH) * rowpitch};
            for (int odd = 0; odd < 2; ++odd) {
                const cuuint64_t dims[4] = {32, 88, odd ? 90u : 91u, N};
                const cuuint32_t box[4] = {32, 11, odd ? 13u : 14u, 1};
                snprintf(what, sizeof what, "layer %d stem %s rows", li, odd ? "odd" : "even");
                int rc = make_map4(c, &out[li][odd], base + odd * rowpitch, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                if (rc) return rc;
            }
        } else if (L.kind == K_S1) {
            const cuuint64_t Cb = static_cast<cuuint64_t>(L.in_c) * bpc, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {Cb / 4, W, H, N};
            const cuuint64_t strides[3] = {Cb, W * Cb, H * W * Cb};
            const cuuint32_t box[4] = {32, 11, 13, 1};
            snprintf(what, sizeof what, "layer %d s1 (%d B/ch)", li, bpc);
            int rc = make_map4(c, &out[li][0], base, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
            if (rc) return rc;
        } else {
            // stride 2: four parity views (py, px) of the input, each a dense half-resolution tensor
            const cuuint64_t Cb = static_cast<cuuint64_t>(L.in_c) * bpc, W = L.Win, H = L.Hin;
            const cuuint64_t dims[4] = {Cb / 4, W / 2, H / 2, N};
            const cuuint64_t strides[3] = {2 * Cb, 2 * W * Cb, H * W * Cb};
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    const cuuint32_t box[4] = {32, 11, py ? 12u : 11u, 1};
                    snprintf(what, sizeof what, "layer %d s2 parity %d%d (%d B/ch)", li, py, px, bpc);
                    int rc = make_map4(c, &out[li][py * 2 + px], base + (static_cast<size_t>(py) * W + px) * Cb, dims, strides, box,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
                    if (rc) return rc;
                }
        }
    }
    return SE3TN_OK;
}

// geometry of one conv for the fp32 FFMA kernel
void fill_geom(const LayerSpec& L, int n, ConvGeom& g) {
    memset(&g, 0, sizeof g);
    g.Hin = L.Hin; g.Win = L.Win; g.in_cstride = L.in_c; g.in_coff = 0;
    g.Ho = layer_Ho(L); g.Wo = g.Ho;
    g.stride = (L.kind == K_S1) ? 1 : 2;
    g.cin = L.cin; g.cout = L.cout; g.groups = L.groups;
    g.num_taps = layer_taps(L);
    g.n_img = n;
    if (L.kind == K_STEM) {
        // tap r: padded input row 2*oy + r, 32 contiguous floats from padded x = 2*ox
        for (int r = 0; r < 7; ++r) { g.taps[r].dy = (int16_t)r; g.taps[r].dx = 0; }
    } else {
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) { g.taps[r * 3 + s].dy = (int16_t)(r - 1); g.taps[r * 3 + s].dx = (int16_t)(s - 1); }
    }
    g.out_cstride = L.out_c; g.out_coff = L.out_coff;
    g.res_cstride = (L.res != NONE) ? res_channels(L) : 0;
    g.res_coff = 0;
    g.act = L.act;
}

// one layer as the tcgen05 kernels see it (conv_common.h LayerDesc)
void fill_layer_desc(const se3tn_ctx* c, const WeightSet& ws, int li, int kprec, LayerDesc& d, int block_n = 0) {
    const LayerSpec& L = kLayers[li];
    if (!block_n) block_n = L.block_n;
    const int row = (block_n == L.block_n) ? li : 14 + li - kFirstTrunkLayer;      // table row / weight map with the matching box height
    memset(&d, 0, sizeof d);
    const int bpc = prec_bytes_per_channel(kprec);
    const bool stem = (L.kind == K_STEM);
    const CUtensorMap (*amaps)[4] = (bpc == 2 && !stem) ? c->amap2 : c->amap4;
    for (int m = 0; m < 4; ++m) d.amap[m] = amaps[li][m];
    d.bmap = (kprec == PREC_TF32) ? ws.bmap_tf32[row] : (kprec == PREC_BF16X3 ? ws.bmap_x3[row] : ws.bmap_h[row]);
    d.bias = ws.dev + ws.b_off[li];
    d.kind = stem ? KIND_STEM : (L.kind == K_S2 ? KIND_S2 : KIND_S1);
    if (stem) {                                    // K = 8 pixels x 4 channels x 4 bytes per filter row in every mode
        d.chunks = 1; d.cin_words = 32; d.in_cbase_words = 0; d.in_gstride_words = 0;
        d.out = reinterpret_cast<uint8_t*>(c->buf[li == 0 ? B_P1A : B_P1B]);   // fused MaxPool2d(3,2,1): the pooled tensor is written directly
        d.out_c = 64; d.out_coff = 0; d.Ho = d.Wo = 44;
        d.tiles_x = d.tiles_y = 9;                 // pooled 5x5 blocks
    } else {
        d.chunks = L.cin * bpc / kChunkBytes; d.cin_words = L.cin * bpc / 4; d.in_cbase_words = 0; d.in_gstride_words = L.cin * bpc / 4;
        d.out = reinterpret_cast<uint8_t*>(c->buf[L.out]);
        d.out_c = L.out_c; d.out_coff = L.out_coff; d.Ho = d.Wo = layer_Ho(L);
        d.tiles_x = d.tiles_y = d.Ho / 11;
    }
    d.res = (L.res != NONE) ? reinterpret_cast<const uint8_t*>(c->buf[L.res]) : nullptr;
    d.res_c = (L.res != NONE) ? res_channels(L) : 0;
    d.cout = L.cout; d.groups = L.groups; d.n_tiles = L.cout / block_n;
    d.act = L.act; d.li = row;
    d.units_per_image = d.tiles_x * d.tiles_y * d.n_tiles * d.groups;
    d.dep_layer = -1; d.dep_target = 0; d.unit_base = 0;
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
    memset(m1.data(), 0, m1.size() * sizeof(CUtensorMap)); memset(m2.data(), 0, m2.size() * sizeof(CUtensorMap)); memset(m3.data(), 0, m3.size() * sizeof(CUtensorMap));
    std::vector<const float*> bias(rows * kLayersPerSet, nullptr), fc(rows, nullptr);
    for (auto& kv : c->weights) {
        if (!kv.second.dev || kv.first < 0) continue;
        for (int li = 0; li < kLayersPerSet; ++li) {
            m1[kv.first * kLayersPerSet + li] = kv.second.bmap_tf32[li];
            m2[kv.first * kLayersPerSet + li] = kv.second.bmap_h[li];
            m3[kv.first * kLayersPerSet + li] = kv.second.bmap_x3[li];
            bias[kv.first * kLayersPerSet + li] = kv.second.dev + kv.second.b_off[li < 14 ? li : kFirstTrunkLayer + li - 14];
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

void drop_graphs(se3tn_ctx* c) {
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

// optional pose update fused into the head kernel (tensor-core modes): K6 for the same n tracks
struct PoseArgs { const double* in = nullptr; double* out = nullptr; float tn = 0.f, rn = 0.f; };

// The conv stack on images [first, first+n) of the context buffers.  img_wid (device, indexed by absolute image
// index) non-null: every image uses its own weight set in the same launches (tensor-core modes);
// `weight_id` is then only a representative loaded set.  *pose_done tells the caller whether `pose` was applied
// (the fp32 FFMA mode leaves it to a separate pose_update_kernel launch).
int run_network(se3tn_ctx* c, int weight_id, int first, int n, int precision,
                float* out_trans, float* out_rot, float* out_feature, cudaStream_t s, const int* img_wid = nullptr,
                const PoseArgs* pose = nullptr, bool* pose_done = nullptr) {
    if (pose_done) *pose_done = false;
    auto it = c->weights.find(weight_id);
    if (it == c->weights.end() || !it->second.dev) return fail(c, SE3TN_ERR_STATE, "weight set " + std::to_string(weight_id) + " not loaded");
    const WeightSet& ws = it->second;
    const int kprec = kprec_of(precision);
    const bool tensor = kprec >= 0;
    if (!tensor && precision != SE3TN_PREC_FP32) return fail(c, SE3TN_ERR_INVALID, "unknown precision");
    auto bufp = [&](Buf b) { return c->buf[b] + kBufFloats[b] * static_cast<size_t>(first); };
    const float* fcw = ws.dev + ws.fc_off;
    if (!tensor) {
        // ---- fp32 FFMA cross-check mode: 14 direct convs + 2 max-pools + head ----
        if (img_wid) return fail(c, SE3TN_ERR_INVALID, "multi-weight-set launches need a tensor-core precision");
        for (int li = 0; li < 14; ++li) {
            const LayerSpec& L = kLayers[li];
            ConvGeom g; fill_geom(L, n, g);
            ConvPtrs p;
            p.in = bufp(L.in); p.out = bufp(L.out); p.res = (L.res != NONE) ? bufp(L.res) : nullptr;
            p.w = ws.dev + ws.w_off[li]; p.bias = ws.dev + ws.b_off[li];
            { ProfScope ps(c, li, s); CU_TRY(c, launch_conv_direct(g, p, s)); }
            ++c->launches;
            if (li == 0) { ProfScope ps(c, 14, s); CU_TRY(c, launch_maxpool(bufp(B_Y1A), bufp(B_P1A), n, 88, 88, 64, s)); ++c->launches; }
            if (li == 1) { ProfScope ps(c, 15, s); CU_TRY(c, launch_maxpool(bufp(B_Y1B), bufp(B_P1B), n, 88, 88, 64, s)); ++c->launches; }
        }
        { ProfScope ps(c, 16, s); CU_TRY(c, launch_head(bufp(B_H3), fcw, fcw + 6 * 512, out_trans, out_rot, n, 121, 0, nullptr, nullptr, s)); }
        ++c->launches;
        if (out_feature) { CU_TRY(c, launch_nhwc_to_nchw(bufp(B_F2), out_feature, n, 22 * 22, 256, 0, s)); ++c->launches; }
        return SE3TN_OK;
    }
    // ---- tensor-core modes: 8 resident-weight launches + 1 trunk launch + head ----
    if (img_wid) { int rc = sync_tables(c, s); if (rc) return rc; }
    const CUtensorMap* gbmaps = img_wid ? (kprec == PREC_BF16X3 ? c->d_bmaps_x3 : (kprec == PREC_BF16 ? c->d_bmaps_bf16 : c->d_bmaps_tf32)) : nullptr;
    const float* const* gbias = img_wid ? c->d_bias : nullptr;
    if (c->sched_dirty) { CU_TRY(c, cudaMemsetAsync(c->sched, 0, trunk_sched_words(c->max_batch) * sizeof(unsigned), s)); c->sched_dirty = false; }
    for (int li = 0; li < kFirstTrunkLayer; ++li) {
        ResidentParams rp;
        fill_layer_desc(c, ws, li, kprec, rp.L);
        rp.img_first = first; rp.n_img = n;
        rp.m_tiles = n * rp.L.tiles_x * rp.L.tiles_y;
        if (kLayers[li].kind == K_STEM) { rp.step_x = rp.step_y = 10; rp.off_x = rp.off_y = -1; }   // 11x11 conv outputs from (10*t - 1): the 5x5 pooled block's window
        else { rp.step_x = rp.step_y = 11; rp.off_x = rp.off_y = 0; }
        rp.img_wid = img_wid; rp.gbmaps = gbmaps; rp.gbias = gbias;
        rp.trace = c->trace ? c->trace + static_cast<size_t>(li) * 256 * 8 : nullptr;
        if (kLayers[li].kind == K_STEM && c->stem_ws && !img_wid && kprec != PREC_TF32) {
            ProfScope ps(c, li, s);
            CU_TRY(c, launch_conv_stem_ws(rp, ws.dev_stack + static_cast<size_t>(li) * 128 * 288 * sizeof(float), kprec, c->stem_ws >> 1, c->num_sms, c->pdl != 0, s));
        } else { ProfScope ps(c, li, s); CU_TRY(c, launch_conv_resident(rp, rp.L.kind, kprec, c->num_sms, c->pdl != 0, s)); }
        ++c->launches;
    }
    {
        TrunkParams tp;
        memset(&tp, 0, sizeof tp);
        // work units of 256 output channels; small batches (at most half of the SMs busy per layer otherwise) use 128:
        // twice the units per layer and half the latency of each -- the layers of one image are a serial chain
        const int bn = (n * 4 * 2 <= c->num_sms) ? 128 : 256;
        // latency mode: a handful of tracks keep only 8 CTAs per layer busy, and the six layers of an image are a serial chain:
        // cut every unit's K loop into kSplitK pieces (conv_trunk_kernel).  Its fp32 sums are grouped differently, so results agree
        // with the throughput mode to rounding, not bit for bit; within the mode (n = 1..4) they do not depend on n.
        int ksplit = (n <= kSplitMaxImages) ? kSplitK : 1;
        for (int l = 0; l < 14 - kFirstTrunkLayer; ++l) {
            fill_layer_desc(c, ws, kFirstTrunkLayer + l, kprec, tp.layer[l], bn);
            while (tp.layer[l].chunks % ksplit) ksplit /= 2;       // 2-byte storage: convAB1 has only two 128-byte chunks per pixel
        }
        int base = 0, base0 = 0;
        for (int l = 0; l < 14 - kFirstTrunkLayer; ++l) {
            LayerDesc& d = tp.layer[l];
            d.unit_base = base; base += n * d.units_per_image * ksplit;
            d.base_unit0 = base0; base0 += n * d.units_per_image;
            // completion signals per unit: one per epilogue warp that finishes part of it (8; in 4-piece latency mode 4 warps of each piece finish one block each)
            if (l > 0) { d.dep_layer = l - 1; d.dep_target = (ksplit == 4 ? 16u : 8u) * static_cast<unsigned>(tp.layer[l - 1].units_per_image); }
        }
        if (ksplit > 1 && base0 > kSplitMaxUnits) return fail(c, SE3TN_ERR_STATE, "split-K scratch too small");
        tp.ksplit = ksplit; tp.partial = c->partial;
        tp.slice_cnt = c->sched + 1 + static_cast<size_t>(kTrunkMaxLayers) * c->max_batch;
        tp.layer[5].pool_part = c->pool_part;      // AdaptiveAvgPool2d(1) fused into the last conv's epilogue (indexed by absolute image)
        tp.n_layers = 6; tp.total_units = base;
        tp.img_first = first; tp.n_img = n; tp.max_batch = c->max_batch;
        tp.sched = c->sched; tp.img_wid = img_wid; tp.gbmaps = gbmaps; tp.gbias = gbias;
        tp.trace = c->trace ? c->trace + static_cast<size_t>(kFirstTrunkLayer) * 256 * 8 : nullptr;
        c->sched_dirty = true;                     // cleared again by the head kernel below
        { ProfScope ps(c, kFirstTrunkLayer, s); CU_TRY(c, launch_conv_trunk(tp, kprec, bn, c->num_sms, c->pdl != 0, s)); }
        ++c->launches;
    }
    {
        ProfScope ps(c, 16, s);
        CU_TRY(c, launch_head_pooled(c->pool_part + static_cast<size_t>(first) * 4 * 1024, fcw, fcw + 6 * 512, out_trans, out_rot, n, 121,
                                     img_wid ? img_wid + first : nullptr, img_wid ? c->d_fc : nullptr,
                                     pose ? pose->in : nullptr, pose ? pose->out : nullptr, pose ? pose->tn : 0.f, pose ? pose->rn : 0.f,
                                     c->sched, static_cast<int>(trunk_sched_words(c->max_batch)), s));
        c->sched_dirty = false;
        if (pose && pose_done) *pose_done = true;
    }
    ++c->launches;
    if (out_feature) { CU_TRY(c, launch_nhwc_to_nchw(reinterpret_cast<const uint8_t*>(c->buf[B_F2]) + static_cast<size_t>(first) * 22 * 22 * 256 * prec_bytes_per_channel(kprec), out_feature, n, 22 * 22, 256,
                                                       kprec == PREC_BF16X3 ? 1 : (kprec == PREC_BF16 ? 2 : 0), s)); ++c->launches; }
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
    DeviceGuard guard(device);
    se3tn_ctx* c = new se3tn_ctx();
    c->device = device; c->max_batch = max_batch; c->num_sms = prop.multiProcessorCount;
    if (const char* ov = getenv("SE3TN_PDL")) c->pdl = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_GRAPH")) c->use_graphs = atoi(ov) != 0;
    if (const char* ov = getenv("SE3TN_STEM_WS")) c->stem_ws = atoi(ov);
    if (const char* ov = getenv("SE3TN_TRACE")) {
        if (atoi(ov) != 0 && cudaMalloc(&c->trace, SE3TN_TRACE_WORDS * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(c->trace, 0, SE3TN_TRACE_WORDS * sizeof(unsigned long long));
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
    e = cudaMalloc(&c->partial, trunk_partial_floats() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->sched, trunk_sched_words(max_batch) * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(c->sched, 0, trunk_sched_words(max_batch) * sizeof(unsigned));
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_NOMEM, "se3tn_create: scheduler state: " + m); }
    // zero once: the stem buffers' 3-pixel halo is the conv padding and is never written again
    e = cudaMemset(c->workspace, 0, bytes);
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_CUDA, "se3tn_create: cudaMemset: " + m); }
    float* p = reinterpret_cast<float*>(c->workspace);
    for (int b = 0; b < B_COUNT; ++b) {
        c->buf[b] = p;
        p += (kBufFloats[b] * static_cast<size_t>(max_batch) + 255) & ~size_t(255);
    }
    int rc = build_activation_maps(c, 4, c->amap4);
    if (!rc) rc = build_activation_maps(c, 2, c->amap2);
    if (rc) { g_create_error = c->err; se3tn_destroy(c); return rc; }
    // the memsets above ran on the NULL stream: later launches may use non-blocking streams, which do not wait for it
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); se3tn_destroy(c); return fail(nullptr, SE3TN_ERR_CUDA, "se3tn_create: " + m); }
    *out = c;
    return SE3TN_OK;
}

void se3tn_destroy(se3tn_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (auto& kv : c->weights) { cudaFree(kv.second.dev); cudaFree(kv.second.dev_tf32); cudaFree(kv.second.dev_bf16); cudaFree(kv.second.dev_h); cudaFree(kv.second.dev_stack); cudaFree(kv.second.dev_perm); cudaFree(kv.second.dev_perm_tmp); }
    cudaFree(c->d_mean32); cudaFree(c->d_std32); cudaFree(c->d_mean64); cudaFree(c->d_std64);
    cudaFree(c->d_bmaps_tf32); cudaFree(c->d_bmaps_bf16); cudaFree(c->d_bmaps_x3); cudaFree(c->d_bias); cudaFree(c->d_fc);
    for (int i = 0; i < SE3TN_PROFILE_SLOTS; ++i) { if (c->ev0[i]) cudaEventDestroy(c->ev0[i]); if (c->ev1[i]) cudaEventDestroy(c->ev1[i]); }
    drop_graphs(c);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    cudaFree(c->sched); cudaFree(c->partial); cudaFree(c->pool_part); cudaFree(c->trace);
    for (auto& kv : c->meshes) { cudaFree(const_cast<float*>(kv.second.pos)); cudaFree(const_cast<float*>(kv.second.nrm)); cudaFree(const_cast<uint8_t*>(kv.second.col)); cudaFree(const_cast<int*>(kv.second.faces)); }
    cudaFree(c->d_meshes); cudaFree(c->render_proj); cudaFree(c->render_unif);
    cudaFree(c->fill.a); cudaFree(c->fill.b); cudaFree(c->fill.lut); cudaFree(c->fill.minmax);
    cudaFree(c->hio.dev); if (c->hio.pin) cudaFreeHost(c->hio.pin);
    if (c->own_workspace) cudaFree(c->workspace);
    delete c;
}

int se3tn_load_weights(se3tn_ctx* c, int weight_id, const float* blob, size_t n_floats) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!blob || weight_id < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_load_weights: bad arguments");
    const size_t expect = blob_floats();
    if (n_floats != expect || expect != SE3TN_WEIGHT_BLOB_FLOATS)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_load_weights: blob has " + std::to_string(n_floats) + " floats, expected " + std::to_string(expect));
    DeviceGuard guard(c->device);
    WeightSet& ws = c->weights[weight_id];
    if (!ws.dev) {
        CU_TRY(c, cudaMalloc(&ws.dev, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_tf32, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_bf16, expect * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_h, expect * sizeof(uint16_t)));
        CU_TRY(c, cudaMalloc(&ws.dev_stack, 8 * 128 * 288 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_perm, 6 * 64 * 576 * sizeof(float)));
        CU_TRY(c, cudaMalloc(&ws.dev_perm_tmp, 64 * 576 * sizeof(float)));
    }
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemcpy(ws.dev, blob, expect * sizeof(float), cudaMemcpyHostToDevice));
    round_tf32_kernel<<<1024, 256>>>(ws.dev, ws.dev_tf32, expect);
    CU_TRY(c, cudaGetLastError());
    size_t off = 0;
    for (int li = 0; li < 14; ++li) {
        const LayerSpec& L = kLayers[li];
        const size_t words = static_cast<size_t>(layer_rows(L)) * layer_ktot(L);
        ws.w_off[li] = off; off += words;
        ws.b_off[li] = off; off += layer_rows(L);
        const float* wsrc = ws.dev + ws.w_off[li];
        char what[64];
        int rc = SE3TN_OK;
        if (li >= kFirstTrunkLayer) {
            // trunk layers: natural row order, tiles of {32 words, 256 rows} streamed through the weight ring
            snprintf(what, sizeof what, "layer %d weights", li);
            rc = make_map2(c, &ws.bmap_tf32[li], ws.dev_tf32 + ws.w_off[li], layer_ktot(L), layer_rows(L), 256, what);
            uint8_t* d3 = ws.dev_bf16 + ws.w_off[li] * sizeof(float);
            CU_TRY(c, launch_split_weights(wsrc, d3, words, 0));                          // [32 hi | 32 lo] per 32-word K chunk
            if (!rc) rc = make_map2(c, &ws.bmap_x3[li], d3, layer_ktot(L), layer_rows(L), 256, what);
            uint8_t* dh = ws.dev_h + ws.w_off[li] * sizeof(uint16_t);
            CU_TRY(c, launch_to_bf16(wsrc, dh, words, 0));                                // plain bf16, 64 channels per 128-byte chunk
            if (!rc) rc = make_map2(c, &ws.bmap_h[li], dh, layer_ktot(L) / 2, layer_rows(L), 256, what);
            // the same matrices in 128-row boxes: small batches run the trunk with 128-channel work units (twice the units per layer)
            const int ls = 14 + li - kFirstTrunkLayer;
            if (!rc) rc = make_map2(c, &ws.bmap_tf32[ls], ws.dev_tf32 + ws.w_off[li], layer_ktot(L), layer_rows(L), 128, what);
            if (!rc) rc = make_map2(c, &ws.bmap_x3[ls], d3, layer_ktot(L), layer_rows(L), 128, what);
            if (!rc) rc = make_map2(c, &ws.bmap_h[ls], dh, layer_ktot(L) / 2, layer_rows(L), 128, what);
        } else {
            // resident-weight layers.  Stems keep the natural row order; the 64-channel 3x3 layers use the row order of the
            // 16x256b epilogue (all precisions).  bf16 modes: hi / lo rows stacked along N, except the 64-channel layers in PREC_BF16.
            const bool stem = (L.kind == K_STEM);
            const float* w_tf32 = ws.dev_tf32 + ws.w_off[li];
            if (!stem) {
                if (layer_ktot(L) != 576 || layer_rows(L) != 64) return fail(c, SE3TN_ERR_STATE, "resident layer shape");
                float* pt = ws.dev_perm + static_cast<size_t>(li - 2) * 64 * 576;
                CU_TRY(c, launch_permute_rows64(wsrc, ws.dev_perm_tmp, 576, 0));
                round_tf32_kernel<<<144, 256>>>(ws.dev_perm_tmp, pt, 64 * 576);
                CU_TRY(c, cudaGetLastError());
                wsrc = ws.dev_perm_tmp; w_tf32 = pt;
            }
            snprintf(what, sizeof what, "layer %d resident weights", li);
            rc = make_map2(c, &ws.bmap_tf32[li], w_tf32, layer_ktot(L), 64, 64, what);
            uint8_t* sdst = ws.dev_stack + static_cast<size_t>(li) * 128 * 288 * sizeof(float);
            CU_TRY(c, launch_split_stack_weights(wsrc, sdst, stem, 0));
            if (!rc) rc = make_map2(c, &ws.bmap_x3[li], sdst, stem ? 224 : 288, 128, 128, what);
            if (stem) ws.bmap_h[li] = ws.bmap_x3[li];                                      // the stem input is [4 hi | 4 lo] per pixel in both bf16 modes
            else {
                uint8_t* dh = ws.dev_h + ws.w_off[li] * sizeof(uint16_t);
                CU_TRY(c, launch_to_bf16(wsrc, dh, words, 0));                            // permuted rows, plain bf16
                if (!rc) rc = make_map2(c, &ws.bmap_h[li], dh, 288, 64, 64, what);
            }
            CU_TRY(c, cudaDeviceSynchronize());                                           // dev_perm_tmp is reused by the next layer
        }
        if (rc) return rc;
    }
    CU_TRY(c, cudaDeviceSynchronize());
    ws.fc_off = off;
    c->tables_dirty = true;
    drop_graphs(c);                                // captured steps hold the old tensor maps / table pointers
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
    drop_graphs(c);
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
    DeviceGuard guard(c->device);
    int rc = sync_stats(c, s); if (rc) return rc;
    PreprocessArgs a;
    a.frame_rgb = frame_rgb; a.frame_depth = frame_depth; a.H = H; a.W = W;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
    a.poses = poses; a.object_width = object_width; a.rgbA = rgbA; a.depthA = depthA; a.weight_ids = weight_ids;
    a.mean32 = c->d_mean32; a.std32 = c->d_std32; a.mean64 = c->d_mean64; a.std64 = c->d_std64;
    a.stats_f64 = c->stats_f64; a.stats_rows = c->stats_rows; a.round_tf32 = store_mode_of(precision); a.b_precropped = 0;
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
    DeviceGuard guard(c->device);
    int rc = sync_stats(c, s); if (rc) return rc;
    PreprocessArgs a;
    memset(&a, 0, sizeof a);
    a.frame_rgb = rgbB; a.frame_depth = depthB; a.H = kImg; a.W = kImg; a.b_precropped = 1;
    a.poses = poses; a.rgbA = rgbA; a.depthA = depthA; a.weight_ids = weight_ids;
    a.mean32 = c->d_mean32; a.std32 = c->d_std32; a.mean64 = c->d_mean64; a.std64 = c->d_std64;
    a.stats_f64 = c->stats_f64; a.stats_rows = c->stats_rows; a.round_tf32 = store_mode_of(precision);
    a.stemA = c->buf[B_X0A]; a.stemB = c->buf[B_X0B]; a.nchwA = out_A; a.nchwB = out_B;
    { ProfScope ps(c, 17, s); CU_TRY(c, launch_preprocess(a, n, s)); }
    ++c->launches;
    return SE3TN_OK;
}

int se3tn_compute_bbox(se3tn_ctx* c, const double* poses, const double* K, const double* widths, const double* scale,
                       int32_t* out_bbox, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses || !K || !widths || !scale || !out_bbox || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_compute_bbox: null/invalid argument");
    DeviceGuard guard(c->device);
    CU_TRY(c, launch_bbox(poses, K, widths, scale, out_bbox, n, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_crop_bbox(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W, const int32_t* bbox, int n,
                    int out_h, int out_w, uint8_t* crop_rgb, uint16_t* crop_depth, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!frame_rgb || !frame_depth || !bbox || !crop_rgb || !crop_depth || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || n < 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_crop_bbox: null/invalid argument");
    DeviceGuard guard(c->device);
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
    DeviceGuard guard(c->device);
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
    DeviceGuard guard(c->device);
    return run_network(c, weight_id, first, n, precision, out_trans, out_rot, out_feature, static_cast<cudaStream_t>(stream));
}

int se3tn_pose_update(se3tn_ctx* c, const double* poses_in, const float* trans, const float* rot,
                      double tn, double rn, double* poses_out, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses_in || !trans || !rot || !poses_out || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_pose_update: null/invalid argument");
    DeviceGuard guard(c->device);
    { ProfScope ps(c, 18, static_cast<cudaStream_t>(stream)); CU_TRY(c, launch_pose_update(poses_in, trans, rot, static_cast<float>(tn), static_cast<float>(rn), poses_out, n, static_cast<cudaStream_t>(stream))); }
    ++c->launches;
    return SE3TN_OK;
}

int se3tn_so3_log(se3tn_ctx* c, const double* poses_a, const double* poses_b, double tn, double rn,
                  double* trans_label, double* rot_label, int n, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!poses_a || !poses_b || !trans_label || !rot_label || n < 0) return fail(c, SE3TN_ERR_INVALID, "se3tn_so3_log: null/invalid argument");
    DeviceGuard guard(c->device);
    CU_TRY(c, launch_so3_log(poses_a, poses_b, tn, rn, trans_label, rot_label, n, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

static int track_batch_launches(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                                const double* K, const double* poses_in, const double* object_width,
                                const uint8_t* rgbA, const uint16_t* depthA,
                                const int32_t* weight_ids_host, const int32_t* weight_ids_dev, int n,
                                double tn, double rn, int precision,
                                float* out_trans, float* out_rot, double* poses_out, bool multi, cudaStream_t s);

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
    if (n < 0 || n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_track_batch: n exceeds max_batch");
    // every id a track uses needs weights AND channel statistics (se3tn_set_stats is per weight id): checked here, where the
    // ids are visible on the host, so that the preprocess kernel never normalises with another set's (or no) statistics
    bool multi = false;
    for (int i = 0; i < n; ++i) {
        const int wid = weight_ids_host ? weight_ids_host[i] : 0;
        auto it = c->weights.find(wid);
        if (wid < 0 || it == c->weights.end() || !it->second.dev) return fail(c, SE3TN_ERR_STATE, "weight set " + std::to_string(wid) + " not loaded");
        if (!it->second.has_stats) return fail(c, SE3TN_ERR_STATE, "weight set " + std::to_string(wid) + " has no mean/std (se3tn_set_stats)");
        if (wid != (weight_ids_host ? weight_ids_host[0] : 0)) multi = true;
        if (!weight_ids_host) break;                // all tracks use set 0
    }
    if (n == 0) return SE3TN_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // ---- one CUDA graph per distinct step: every argument that ends up inside a kernel parameter is part of the key ----
    const bool graphable = c->use_graphs && !c->profiling && precision != SE3TN_PREC_FP32;
    std::vector<unsigned long long> key;
    c->last_was_graph = false;
    if (graphable) {
        auto bits = [](double d) { unsigned long long u; memcpy(&u, &d, 8); return u; };
        const void* ptrs[] = {frame_rgb, frame_depth, poses_in, object_width, rgbA, depthA, weight_ids_dev, out_trans, out_rot, poses_out};
        for (const void* p : ptrs) key.push_back(reinterpret_cast<unsigned long long>(p));
        key.push_back(static_cast<unsigned long long>(H)); key.push_back(static_cast<unsigned long long>(W));
        key.push_back(static_cast<unsigned long long>(n)); key.push_back(static_cast<unsigned long long>(precision));
        key.push_back(multi ? 1ull : 0ull); key.push_back(static_cast<unsigned long long>(weight_ids_host ? weight_ids_host[0] : 0));
        for (int i = 0; i < 4; ++i) key.push_back(bits(K[i]));
        key.push_back(bits(tn)); key.push_back(bits(rn));
        for (auto& g : c->graphs)
            if (g.key == key) {
                CU_TRY(c, cudaGraphLaunch(g.exec, s));
                g.last_use = ++c->graph_clock; c->launches = g.launches; c->last_was_graph = true;
                return SE3TN_OK;
            }
        // a new step shape: host-side table refreshes (synchronous copies) must not happen inside the capture
        int rc0 = sync_stats(c, s); if (rc0) return rc0;
        if (multi) { rc0 = sync_tables(c, s); if (rc0) return rc0; }
        if (c->sched_dirty) { CU_TRY(c, cudaMemsetAsync(c->sched, 0, trunk_sched_words(c->max_batch) * sizeof(unsigned), s)); c->sched_dirty = false; }
        if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); c->cap_stream = nullptr; c->use_graphs = 0; key.clear(); }
        if (!key.empty() && cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); c->use_graphs = 0; key.clear(); }
    }
    const bool capturing = graphable && !key.empty();
    auto end_capture = [&](int rc_launch) -> int {
        // turn what was recorded into an executable graph and run it; any failure falls back to plain stream launches for good
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->cap_stream, &graph);
        if (rc_launch != SE3TN_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError(); c->use_graphs = 0; c->sched_dirty = true;
            return rc_launch != SE3TN_OK ? rc_launch : 1;      // 1: capture failed, caller relaunches directly
        }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); c->use_graphs = 0; return 1; }
        if (c->graphs.size() >= 64) {                          // evict the least recently used step
            size_t lru = 0;
            for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].last_use < c->graphs[lru].last_use) lru = i;
            cudaGraphExecDestroy(c->graphs[lru].exec); c->graphs.erase(c->graphs.begin() + lru);
        }
        c->graphs.push_back({key, exec, c->launches, ++c->graph_clock});
        CU_TRY(c, cudaGraphLaunch(exec, s));
        c->last_was_graph = true;
        return SE3TN_OK;
    };
    if (capturing) {
        const int rc = track_batch_launches(c, frame_rgb, frame_depth, H, W, K, poses_in, object_width, rgbA, depthA, weight_ids_host, weight_ids_dev, n,
                                            tn, rn, precision, out_trans, out_rot, poses_out, multi, c->cap_stream);
        const int grc = end_capture(rc);
        if (grc == SE3TN_OK) return SE3TN_OK;
        if (grc != 1) return grc;                              // a real launch error
        // capture was not possible on this stream / driver: plain launches from here on
    }
    return track_batch_launches(c, frame_rgb, frame_depth, H, W, K, poses_in, object_width, rgbA, depthA, weight_ids_host, weight_ids_dev, n,
                                tn, rn, precision, out_trans, out_rot, poses_out, multi, s);
}

// the launches of one step, on stream s (being captured or not)
static int track_batch_launches(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                                const double* K, const double* poses_in, const double* object_width,
                                const uint8_t* rgbA, const uint16_t* depthA,
                                const int32_t* weight_ids_host, const int32_t* weight_ids_dev, int n,
                                double tn, double rn, int precision,
                                float* out_trans, float* out_rot, double* poses_out, bool multi, cudaStream_t s) {
    void* stream = s;
    c->launches = 0;
    int rc = se3tn_preprocess(c, frame_rgb, frame_depth, H, W, K, poses_in, object_width, rgbA, depthA, weight_ids_dev, n,
                              precision, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    PoseArgs pose; pose.in = poses_in; pose.out = poses_out; pose.tn = static_cast<float>(tn); pose.rn = static_cast<float>(rn);
    bool pose_done = false;
    if (precision != SE3TN_PREC_FP32) {
        // every track picks its own weight set inside the same launches; K6 runs inside the head kernel
        rc = run_network(c, weight_ids_host ? weight_ids_host[0] : 0, 0, n, precision, out_trans, out_rot, nullptr, s,
                         multi ? weight_ids_dev : nullptr, &pose, &pose_done);
        if (rc) return rc;
    } else {
        int first = 0;
        while (first < n) {
            const int wid = weight_ids_host ? weight_ids_host[first] : 0;
            int last = first + 1;
            while (last < n && (weight_ids_host ? weight_ids_host[last] : 0) == wid) ++last;
            rc = run_network(c, wid, first, last - first, precision, out_trans + first * 3, out_rot + first * 3, nullptr, s);
            if (rc) return rc;
            first = last;
        }
    }
    if (pose_done) return SE3TN_OK;
    return se3tn_pose_update(c, poses_in, out_trans, out_rot, tn, rn, poses_out, n, stream);
}

int se3tn_add_adi(se3tn_ctx* c, const double* model_pts, int m, const double* pred, const double* gt, int n,
                  double* out_add, double* out_adi, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!model_pts || !pred || !gt || m <= 0 || n < 0 || (!out_add && !out_adi)) return fail(c, SE3TN_ERR_INVALID, "se3tn_add_adi: null/invalid argument");
    DeviceGuard guard(c->device);
    CU_TRY(c, launch_add_adi(model_pts, m, pred, gt, n, out_add, out_adi, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_vocap(se3tn_ctx* c, const double* errs, int n, double* out_ap, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out_ap || n < 0 || (n > 0 && !errs)) return fail(c, SE3TN_ERR_INVALID, "se3tn_vocap: null/invalid argument");
    DeviceGuard guard(c->device);
    CU_TRY(c, vocap(errs, n, out_ap, static_cast<cudaStream_t>(stream)));
    return SE3TN_OK;
}

int se3tn_upload_frame_window(se3tn_ctx* c, const uint8_t* rgb_host, const uint16_t* depth_host, int H, int W,
                              int y0, int y1, int x0, int x1, uint8_t* rgb_dev, uint16_t* depth_dev, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (H <= 0 || W <= 0 || y0 < 0 || x0 < 0 || y1 > H || x1 > W || y0 > y1 || x0 > x1 || (rgb_host && !rgb_dev) || (depth_host && !depth_dev))
        return fail(c, SE3TN_ERR_INVALID, "se3tn_upload_frame_window: bad arguments");
    if (y0 == y1 || x0 == x1) return SE3TN_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t off = static_cast<size_t>(y0) * W + x0;
    if (rgb_host) CU_TRY(c, cudaMemcpy2DAsync(rgb_dev + off * 3, static_cast<size_t>(W) * 3, rgb_host + off * 3, static_cast<size_t>(W) * 3,
                                              static_cast<size_t>(x1 - x0) * 3, y1 - y0, cudaMemcpyHostToDevice, s));
    if (depth_host) CU_TRY(c, cudaMemcpy2DAsync(depth_dev + off, static_cast<size_t>(W) * 2, depth_host + off, static_cast<size_t>(W) * 2,
                                                static_cast<size_t>(x1 - x0) * 2, y1 - y0, cudaMemcpyHostToDevice, s));
    return SE3TN_OK;
}

namespace {
// crop window of one track in frame pixels: compute_bbox + crop_bbox's window (reference Utils.py:302-316, 324-327), the same
// arithmetic as bbox.cuh on the device
inline void host_crop_window(const double* pose, const double* K, double width, int& top, int& left, int& ch, int& cw) {
    const double ox = pose[3] * 1000.0, oy = pose[7] * 1000.0, oz = pose[11] * 1000.0, half = width / 2;
    const double u0 = std::nearbyint((ox - half) * K[0] / oz + K[2]), u1 = std::nearbyint((ox + half) * K[0] / oz + K[2]);
    const double v0 = std::nearbyint((oy - half) * K[1] / oz + K[3]), v1 = std::nearbyint((oy + half) * K[1] / oz + K[3]);
    const double umin = std::fmin(u0, u1), umax = std::fmax(u0, u1), vmin = std::fmin(v0, v1), vmax = std::fmax(v0, v1);
    const double lim = 1.0e9;
    if (!(umin == umin && umax == umax && vmin == vmin && vmax == vmax)) { top = left = ch = cw = 0; return; }
    left = static_cast<int>(std::fmax(-lim, std::fmin(lim, umin))); top = static_cast<int>(std::fmax(-lim, std::fmin(lim, vmin)));
    cw = static_cast<int>(std::fmax(-lim, std::fmin(lim, umax))) - left; ch = static_cast<int>(std::fmax(-lim, std::fmin(lim, vmax))) - top;
}
inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
}  // namespace

int se3tn_track_host(se3tn_ctx* c, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W, const double* K,
                     const double* poses, const double* object_width, const uint8_t* rgbA, const uint16_t* depthA,
                     const int32_t* weight_ids, int n, double tn, double rn, int precision,
                     double* poses_out, float* out_trans, float* out_rot, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!frame_rgb || !frame_depth || !K || !poses || !object_width || !rgbA || !depthA || !poses_out || H <= 0 || W <= 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_track_host: null argument or empty frame");
    if (n < 0 || n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_track_host: n exceeds max_batch");
    if (n == 0) return SE3TN_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto& io = c->hio;
    const size_t px = static_cast<size_t>(H) * W, img = static_cast<size_t>(kImg) * kImg;
    // ---- (re)size the context-owned buffers: stable addresses from then on, so the captured step is replayed ----
    if (io.H != H || io.W != W || io.n_cap < n) {
        CU_TRY(c, cudaStreamSynchronize(s));
        const int cap = std::max(n, io.n_cap);
        const size_t per = 128 + 8 + img * 3 + img * 2 + 4 + 128 + 12 + 12;
        const size_t dev_bytes = align256(px * 3) + align256(px * 2) + align256(per * cap) + 8 * 256;
        cudaFree(io.dev); io.dev = nullptr; if (io.pin) { cudaFreeHost(io.pin); io.pin = nullptr; }
        io.H = io.W = io.n_cap = 0;
        CU_TRY(c, cudaMalloc(&io.dev, dev_bytes));
        CU_TRY(c, cudaMemsetAsync(io.dev, 0, dev_bytes, s));    // on the caller's stream, ahead of the copies below; frame pixels outside the uploaded windows are never read, keep them defined
        CU_TRY(c, cudaHostAlloc(&io.pin, px * 5 + per * cap + 4096, cudaHostAllocDefault));
        io.dev_bytes = dev_bytes; io.pin_bytes = px * 5 + per * cap + 4096; io.H = H; io.W = W; io.n_cap = cap;
        drop_graphs(c);                                          // steps captured against the old addresses
    }
    uint8_t* d = io.dev;
    uint8_t* d_rgb = d; d += align256(px * 3);
    uint16_t* d_depth = reinterpret_cast<uint16_t*>(d); d += align256(px * 2);
    // the per-track arrays are packed by THIS call's n (a step's graph is keyed by n anyway), inputs first, outputs behind them:
    // one host -> device copy carries all inputs, one device -> host copy all outputs
    const size_t nn = static_cast<size_t>(n);
    const size_t o_ow = align256(nn * 128), o_rgbA = o_ow + align256(nn * 8), o_depthA = o_rgbA + align256(nn * img * 3),
                 o_wid = o_depthA + align256(nn * img * 2), in_bytes = o_wid + align256(nn * 4);
    const size_t o_tr = align256(nn * 128), o_ro = o_tr + align256(nn * 12), out_bytes = o_ro + align256(nn * 12);
    uint8_t* d_in = d;
    double* d_poses = reinterpret_cast<double*>(d_in);
    double* d_ow = reinterpret_cast<double*>(d_in + o_ow);
    uint8_t* d_rgbA = d_in + o_rgbA;
    uint16_t* d_depthA = reinterpret_cast<uint16_t*>(d_in + o_depthA);
    int32_t* d_wid = reinterpret_cast<int32_t*>(d_in + o_wid);
    uint8_t* d_res = d_in + in_bytes;
    double* d_out = reinterpret_cast<double*>(d_res);
    float* d_tr = reinterpret_cast<float*>(d_res + o_tr);
    float* d_ro = reinterpret_cast<float*>(d_res + o_ro);
    // ---- the part of the frame the tracks' crop windows touch (K0 reads nothing else) ----
    int y0 = H, y1 = 0, x0 = W, x1 = 0;
    for (int i = 0; i < n; ++i) {
        int top, left, ch, cw;
        host_crop_window(poses + 16 * i, K, object_width[i], top, left, ch, cw);
        if (ch <= 0 || cw <= 0) continue;
        y0 = std::min(y0, std::max(top - 1, 0)); y1 = std::max(y1, std::min(top + ch + 1, H));       // one pixel of margin
        x0 = std::min(x0, std::max(left - 1, 0)); x1 = std::max(x1, std::min(left + cw + 1, W));
    }
    if (y1 <= y0 || x1 <= x0) { y0 = y1 = x0 = x1 = 0; }         // every window misses the frame: nothing of it is read
    if (static_cast<size_t>(y1 - y0) * (x1 - x0) * 2 >= px) { y0 = 0; y1 = H; x0 = 0; x1 = W; }
    // ---- stage through pinned memory, one asynchronous copy per array ----
    uint8_t* hp = io.pin;
    const int wh = y1 - y0, ww = x1 - x0;
    if (wh > 0 && ww > 0) {
        uint8_t* st_rgb = hp; hp += static_cast<size_t>(wh) * ww * 3;
        uint8_t* st_dep = hp; hp += align256(static_cast<size_t>(wh) * ww * 2);
        for (int y = 0; y < wh; ++y) {
            memcpy(st_rgb + static_cast<size_t>(y) * ww * 3, frame_rgb + (static_cast<size_t>(y0 + y) * W + x0) * 3, static_cast<size_t>(ww) * 3);
            memcpy(st_dep + static_cast<size_t>(y) * ww * 2, frame_depth + static_cast<size_t>(y0 + y) * W + x0, static_cast<size_t>(ww) * 2);
        }
        const size_t off = static_cast<size_t>(y0) * W + x0;
        CU_TRY(c, cudaMemcpy2DAsync(d_rgb + off * 3, static_cast<size_t>(W) * 3, st_rgb, static_cast<size_t>(ww) * 3, static_cast<size_t>(ww) * 3, wh, cudaMemcpyHostToDevice, s));
        CU_TRY(c, cudaMemcpy2DAsync(d_depth + off, static_cast<size_t>(W) * 2, st_dep, static_cast<size_t>(ww) * 2, static_cast<size_t>(ww) * 2, wh, cudaMemcpyHostToDevice, s));
    }
    hp = io.pin + align256(static_cast<size_t>(hp - io.pin));
    memcpy(hp, poses, nn * 128);
    memcpy(hp + o_ow, object_width, nn * 8);
    memcpy(hp + o_rgbA, rgbA, nn * img * 3);
    memcpy(hp + o_depthA, depthA, nn * img * 2);
    if (weight_ids) memcpy(hp + o_wid, weight_ids, nn * 4);
    CU_TRY(c, cudaMemcpyAsync(d_in, hp, weight_ids ? in_bytes : o_wid, cudaMemcpyHostToDevice, s));
    hp += in_bytes;
    const int rc = se3tn_track_batch(c, d_rgb, d_depth, H, W, K, d_poses, d_ow, d_rgbA, d_depthA, weight_ids, weight_ids ? d_wid : nullptr, n,
                                     tn, rn, precision, d_tr, d_ro, d_out, stream);
    if (rc != SE3TN_OK) return rc;
    uint8_t* ho = hp;                                            // outputs come back through the same pinned block
    CU_TRY(c, cudaMemcpyAsync(ho, d_res, (out_trans || out_rot) ? out_bytes : nn * 128, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    memcpy(poses_out, ho, nn * 128);
    if (out_trans) memcpy(out_trans, ho + o_tr, nn * 12);
    if (out_rot) memcpy(out_rot, ho + o_ro, nn * 12);
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
    DeviceGuard guard(c->device);
    const int kNcclFloat64 = 8;
    const int rc = fn(local_poses, all_poses, static_cast<size_t>(n_local) * 16, kNcclFloat64, nccl_comm, static_cast<cudaStream_t>(stream));
    if (rc != 0) return fail(c, SE3TN_ERR_CUDA, "se3tn_allgather_poses: ncclAllGather returned " + std::to_string(rc));
    return SE3TN_OK;
}

int se3tn_fill_depth_ex(se3tn_ctx* c, const uint16_t* depth_mm, int H, int W, double max_depth, int extrapolate, int blur_type,
                        uint16_t* out_mm, float* out_m, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!depth_mm || H <= 0 || W <= 0 || (!out_mm && !out_m) || (blur_type != SE3TN_BLUR_BILATERAL && blur_type != SE3TN_BLUR_GAUSSIAN))
        return fail(c, SE3TN_ERR_INVALID, "se3tn_fill_depth: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DeviceGuard guard(c->device);
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
    CU_TRY(c, launch_fill_depth(depth_mm, H, W, static_cast<float>(max_depth), extrapolate != 0, blur_type == SE3TN_BLUR_GAUSSIAN, c->fill, out_mm, out_m, s));
    c->launches += 8;
    return SE3TN_OK;
}

int se3tn_fill_depth(se3tn_ctx* c, const uint16_t* depth_mm, int H, int W, double max_depth,
                     uint16_t* out_mm, float* out_m, void* stream) {
    return se3tn_fill_depth_ex(c, depth_mm, H, W, max_depth, 0, SE3TN_BLUR_BILATERAL, out_mm, out_m, stream);
}

int se3tn_set_mesh(se3tn_ctx* c, int mesh_id, const float* pos, const float* nrm, const uint8_t* col,
                   const int32_t* faces, int nv, int nf) {
    if (!c) return SE3TN_ERR_INVALID;
    if (mesh_id < 0 || mesh_id > 4095 || !pos || !nrm || !col || !faces || nv <= 0 || nf <= 0)
        return fail(c, SE3TN_ERR_INVALID, "se3tn_set_mesh: bad arguments");
    for (int i = 0; i < 3 * nf; ++i)
        if (faces[i] < 0 || faces[i] >= nv) return fail(c, SE3TN_ERR_INVALID, "se3tn_set_mesh: face index out of range");
    DeviceGuard guard(c->device);
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
    return se3tn_render_ex(c, K, poses, object_width, mesh_ids, n, SE3TN_RENDER_VISPY, 0, 0, rgbA, depthA, stream);
}

int se3tn_render_ex(se3tn_ctx* c, const double* K, const double* poses, const double* object_width,
                    const int32_t* mesh_ids, int n, int mode, int H, int W, uint8_t* rgbA, uint16_t* depthA, void* stream) {
    if (!c) return SE3TN_ERR_INVALID;
    if (n < 0 || !K || (n > 0 && (!poses || !object_width || !rgbA || !depthA))) return fail(c, SE3TN_ERR_INVALID, "se3tn_render: bad arguments");
    if (mode != SE3TN_RENDER_VISPY && mode != SE3TN_RENDER_PYRENDER) return fail(c, SE3TN_ERR_INVALID, "se3tn_render_ex: unknown mode");
    // the camera image of the pyrender-style mode: sample positions are kept in 1/256 pixel as int32
    if (mode == SE3TN_RENDER_PYRENDER && (H <= 0 || W <= 0 || H > 65536 || W > 65536)) return fail(c, SE3TN_ERR_INVALID, "se3tn_render_ex: camera image size out of range");
    if (n == 0) return SE3TN_OK;
    if (n > c->max_batch) return fail(c, SE3TN_ERR_INVALID, "se3tn_render: n exceeds the context's max_batch");
    if (c->meshes.empty()) return fail(c, SE3TN_ERR_STATE, "se3tn_render: no mesh loaded (se3tn_set_mesh)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DeviceGuard guard(c->device);
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
    a.mode = mode == SE3TN_RENDER_PYRENDER ? 1 : 0; a.vw = W; a.vh = H;
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
int se3tn_last_step_was_graph(se3tn_ctx* c) { return (c && c->last_was_graph) ? 1 : 0; }

int se3tn_get_trace(se3tn_ctx* c, unsigned long long* out) {
    if (!c) return SE3TN_ERR_INVALID;
    if (!out) return fail(c, SE3TN_ERR_INVALID, "se3tn_get_trace: null argument");
    if (!c->trace) return fail(c, SE3TN_ERR_STATE, "se3tn_get_trace: the context was created without SE3TN_TRACE=1");
    DeviceGuard guard(c->device);
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemcpy(out, c->trace, SE3TN_TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return SE3TN_OK;
}

int se3tn_set_profiling(se3tn_ctx* c, int enable) {
    if (!c) return SE3TN_ERR_INVALID;
    DeviceGuard guard(c->device);
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

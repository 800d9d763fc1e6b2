// Shared host/device descriptors for the convolution launches of the se(3)-TrackNet conv stack.
//
// Every conv on the path (reference se3_tracknet.py:57-78; 7x7 s2 stem, 3x3 s1, 3x3 s2) is
// described the same way: an NHWC activation tensor, a list of filter "taps", each tap a
// displacement in input pixels plus `k_per_tap` contiguous input channels, and a K-major weight
// matrix W[g*Cout + co][tap*k_per_tap + c] with the eval-mode BatchNorm folded in.
//
//  * 3x3:  9 taps (dy,dx) in {-1,0,1}^2, k_per_tap = Cin (channels of one pixel)
//  * stem: 7 taps (one per filter ROW r), k_per_tap = 32 = 8 pixels x 4 channels of the
//          zero-padded NHWC4 input starting at x = 2*ox (7 real filter columns + 1 zero column)
//
// Two tcgen05 kernels (conv_umma2.cu) cover the 14 launches' worth of layers:
//  * conv_resident_kernel: Cout = 64 layers (the two stems, the six 64-channel 3x3 convs); the whole weight
//    matrix lives in shared memory; one launch per layer, static tile ranges.
//  * conv_trunk_kernel: the Cout >= 256 layers (convAB1, convAB2.*, {trans,rot}_conv1, {trans,rot}_conv2.*) as ONE
//    launch: a persistent CTA per SM pulls (layer, image, tile) work units from a global counter and per-image
//    completion counters carry the layer-to-layer dependencies, so no SM idles at a layer boundary.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace se3tn {

constexpr int kMaxTaps = 9;
constexpr int kBlockM = 128;          // UMMA M (TMEM lanes)
constexpr int kChunkBytes = 128;      // one SWIZZLE_128B row of K: 32 tf32 words / 32 x [bf16 hi, bf16 lo] / 64 bf16

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SELU = 2 };
enum { KIND_S1 = 0, KIND_S2 = 1, KIND_STEM = 2 };            // conv kinds (compile-time unit tables in conv_umma2.cu)
enum { PREC_TF32 = 0, PREC_BF16X3 = 1, PREC_BF16 = 2 };      // arithmetic / storage of the tcgen05 kernels
// Activation storage per precision: PREC_TF32 fp32 words (tf32-rounded); PREC_BF16X3 the same 4 bytes per channel as
// [32 x bf16 hi | 32 x bf16 lo] per 32-channel chunk; PREC_BF16 2 bytes per channel, 64 channels per 128-byte chunk
// (the stem INPUT stays 16 bytes per pixel [4 x hi | 4 x lo] in both bf16 modes).
__host__ __device__ constexpr int prec_bytes_per_channel(int prec) { return prec == PREC_BF16 ? 2 : 4; }

struct Tap {
    int16_t dy, dx;      // input-pixel displacement of this tap relative to (oy*stride, ox*stride)
};

// geometry of one conv for the FFMA cross-check kernel (conv_direct.cu) -- fp32 NHWC, 4 bytes per channel
struct ConvGeom {
    int Hin, Win, in_cstride, in_coff;
    int Ho, Wo, stride;
    int cin;             // K floats per tap per group
    int cout;            // per group
    int groups;
    int num_taps;
    int n_img;
    Tap taps[kMaxTaps];
    int out_cstride, out_coff;
    int res_cstride, res_coff;
    int act;
};

struct ConvPtrs {
    const float* in;
    const float* w;      // [groups*cout][num_taps*cin]
    const float* bias;   // [groups*cout]
    const float* res;    // nullable, NHWC
    float* out;
};

constexpr int kLayersPerSet = 20;     // stride of the per-set device tables: rows 0..13 the 14 conv layers, rows 14..19 the trunk layers'
                                      // maps again with 128-row boxes (small-batch trunk tiles; same biases)

// ---- tcgen05 kernels ----------------------------------------------------------------------------------------
// One layer as the device sees it.  Channel counts are in CHANNELS; byte strides follow from the precision.
struct LayerDesc {
    CUtensorMap amap[4];   // activation views: S1 one map; S2 four parity views (py*2+px); stem two (even / odd input rows)
    CUtensorMap bmap;      // weights of weight set `single_wid` (multi-set launches take theirs from gbmaps)
    const float* bias;     // [groups*cout] fp32 (single-set)
    uint8_t* out;          // NHWC output buffer (image 0)
    const uint8_t* res;    // residual input (nullable), same storage format as out
    float* pool_part;      // non-null: fused AdaptiveAvgPool2d(1): column sums [image][4 row quadrants][out_c] instead of the activation
    int kind;              // KIND_*
    int chunks;            // 128-byte K chunks per pixel per group (cin * bytes / 128)
    int cin_words;         // 32-bit words of K per tap per group (weight-matrix K offset of a tap = tap * cin_words)
    int in_cbase_words;    // word offset of group 0's channels inside a pixel of the input buffer
    int in_gstride_words;  // word offset between groups
    int cout;              // per group
    int groups;
    int n_tiles;           // cout / BN
    int tiles_x, tiles_y;  // 11x11 output tiles per image
    int Ho, Wo;
    int act;
    int out_c, out_coff;   // channels per pixel of the output buffer, channel offset of this layer's channel 0
    int res_c;             // channels per pixel of the residual buffer
    int li;                // row of the per-set tables (layer index; trunk layers with 128-row weight boxes: 14 + layer - 8)
    // trunk scheduling
    int unit_base;         // first global work-unit index of this layer (units are K-split pieces when TrunkParams::ksplit > 1)
    int base_unit0;        // index of this layer's first UNSPLIT unit among all unsplit units of the launch (split-K scratch / counters)
    int units_per_image;   // tiles_x * tiles_y * n_tiles * groups (unsplit)
    int dep_layer;         // index (within the launch) of the layer whose per-image completion this layer waits for; -1: none
    unsigned dep_target;   // value done[dep_layer][image] reaches when that image is complete (one signal per epilogue warp that finishes part of a unit: 8 x units per image; 16 x in 4-piece latency mode)
};

constexpr int kTrunkMaxLayers = 6;

struct TrunkParams {
    LayerDesc layer[kTrunkMaxLayers];
    int n_layers;
    int total_units;
    int img_first, n_img;          // absolute image range [img_first, img_first + n_img)
    int max_batch;                 // row length of the done[] table
    unsigned* sched;               // [0] next work unit; then done[layer][image] counters (zeroed before the launch)
    const int* img_wid;            // nullable: weight-set id per absolute image index
    const CUtensorMap* gbmaps;     // per-set weight maps, entry [wid * kLayersPerSet + li]
    const float* const* gbias;     // per-set bias pointers, same indexing
    unsigned long long* trace;     // nullable (SE3TN_TRACE)
    // latency mode (a handful of tracks): every unit's K loop is cut into `ksplit` pieces run by different CTAs (units are dealt
    // round robin so that they are); each piece dumps its fp32 accumulator to `partial`, then finishes ITS share of the unit's
    // 32-column blocks: it waits for the other pieces' dumps, sums all pieces in a fixed order and runs the normal epilogue.  1 = off.
    int ksplit;
    float* partial;                // [unsplit unit][piece][8 warp slices][32-column block][float4 0..7][row]
    unsigned* slice_cnt;           // [unsplit unit][8 warp slices] number of pieces that have dumped the slice (zeroed before the launch)
};

struct ResidentParams {
    LayerDesc L;
    int img_first, n_img;
    int m_tiles;                   // n_img * tiles_x * tiles_y
    int step_x, step_y, off_x, off_y;   // tile origin in A-map coordinates = tile index * step + off (stem: pooled 5x5 blocks)
    const int* img_wid;
    const CUtensorMap* gbmaps;
    const float* const* gbias;
    unsigned long long* trace;
};

cudaError_t launch_conv_resident(const ResidentParams& p, int kind, int prec, int num_sms, bool pdl, cudaStream_t stream);
cudaError_t launch_conv_trunk(const TrunkParams& p, int prec, int block_n /*256 | 128*/, int num_sms, bool pdl, cudaStream_t stream);
// weights-stationary stem (conv_stem_t.cu): bf16 modes, single weight set; wstack = the stem's stacked weight matrix [128][224 words]
cudaError_t launch_conv_stem_ws(const ResidentParams& p, const void* wstack, int prec, int debug_flags, int num_sms, bool pdl, cudaStream_t stream);
// latency mode: at most kSplitMaxImages images, ksplit = kSplitK pieces, 128-channel units (at most 8 per image and layer)
constexpr int kSplitMaxImages = 4, kSplitK = 4, kSplitMaxUnits = kSplitMaxImages * 8 * kTrunkMaxLayers;
// 32-bit words of scheduler state a trunk launch needs: next-unit counter + done[layers][max_batch] + split-K slice counters
inline size_t trunk_sched_words(int max_batch) { return 1 + static_cast<size_t>(kTrunkMaxLayers) * max_batch + static_cast<size_t>(kSplitMaxUnits) * 8; }
inline size_t trunk_partial_floats() { return static_cast<size_t>(kSplitMaxUnits) * kSplitK * 128 * 128; }

cudaError_t launch_conv_direct(const ConvGeom& g, const ConvPtrs& p, cudaStream_t stream);

}  // namespace se3tn

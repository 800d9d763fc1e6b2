// Shared host/device descriptors for one convolution launch of the se(3)-TrackNet conv stack.
//
// Every conv on the path (reference se3_tracknet.py:57-78; 7x7 s2 stem, 3x3 s1, 3x3 s2) is
// described the same way: an NHWC activation tensor, a list of filter "taps", each tap a
// displacement in input pixels plus `k_per_tap` contiguous input floats, and a K-major weight
// matrix W[g*Cout + co][tap*k_per_tap + c] with the eval-mode BatchNorm folded in.
//
//  * 3x3:  9 taps (dy,dx) in {-1,0,1}^2, k_per_tap = Cin (channels of one pixel)
//  * stem: 7 taps (one per filter ROW r), k_per_tap = 32 = 8 pixels x 4 channels of the
//          zero-padded NHWC4 input starting at x = 2*ox (7 real filter columns + 1 zero column)
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace se3tn {

constexpr int kMaxTaps = 9;
constexpr int kBlockM = 128;          // UMMA M (TMEM lanes)
constexpr int kChunkBytes = 128;      // one SWIZZLE_128B row: 32 tf32 / 64 bf16 along K

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SELU = 2 };

struct Tap {
    int16_t dy, dx;      // input-pixel displacement of this tap relative to (oy*stride, ox*stride)
    int8_t  map;         // tcgen05 path: which A tensor map (s2: 4 parity views; stem: one per filter row)
    int8_t  c1, c2;      // tcgen05 path: TMA coordinate deltas for dims 1 (x) and 2 (y)
    int8_t  pad_;
};

struct ConvGeom {
    // input (NHWC, `in_cstride` floats per pixel, channel window starting at in_coff (+ g*cin))
    int Hin, Win, in_cstride, in_coff;
    // output
    int Ho, Wo, stride;
    int cin;             // K floats per tap per group
    int cout;            // per group
    int groups;
    int num_taps;
    int n_img;
    Tap taps[kMaxTaps];
    // epilogue
    int out_cstride, out_coff;
    int res_cstride, res_coff;
    int act;
    int round_tf32;      // fp32-word storage: round stored activations to tf32 (rna) so the next UMMA sees exact operands
};

constexpr int kLayersPerSet = 14;     // conv launches per weight set (stride of the per-set device tables)

struct ConvPtrs {
    const float* in;
    const float* w;      // [groups*cout][num_taps*cin]
    const float* bias;   // [groups*cout]
    const float* res;    // nullable, NHWC
    float* out;
    // multi-weight-set launches (v2 kernel): weight-set id per absolute image index, and per-set tables that are
    // already offset to this layer (entry of set w at [w * kLayersPerSet])
    const int* img_wid;            // nullable: single-set launch (maps.b / bias above)
    const CUtensorMap* gbmaps;     // device array of weight tensor maps
    const float* const* gbias;     // device array of bias pointers
    // stream-K (ring-weight kernels): one 128 x 256 fp32 partial-accumulator slot and one flag per CTA; null = off
    float* sk_part;
    int* sk_flags;
    // fused AdaptiveAvgPool2d(1) (last conv of the heads, one image per M tile): instead of storing the activation, every
    // epilogue warp writes the column sums over its 32 rows to pool_part[image][row quadrant][channel]; null = normal store
    float* pool_part;
    // device-side timeline (SE3TN_TRACE=1; null = off): 8 globaltimer stamps per CTA, see conv_umma2.cu trace_stamp()
    unsigned long long* trace;
};

// ---- tcgen05 path only ------------------------------------------------------------------
struct UmmaTiling {
    int bw, bh, bn;          // pixel box of one M tile: bw*bh*bn <= 128 rows
    int tiles_x, tiles_y;    // boxes per image
    int m_tiles;             // ceil(n_img/bn) * tiles_y * tiles_x
    int n_tiles;             // cout / BLOCK_N
    int chunks_per_tap;      // cin / 32
    int img_first;           // absolute index of the first image this launch covers
};

struct alignas(64) UmmaMaps {
    CUtensorMap a[7];
    CUtensorMap b;
};

// ---- tcgen05 path, second generation (conv_umma2.cu) ---------------------------------------
enum { KIND_S1 = 0, KIND_S2 = 1, KIND_STEM = 2 };   // conv kinds of the v2 kernel (compile-time unit tables)
struct UnitTap { int8_t row_shift; int8_t w_tap; };   // rows to advance the A descriptor; weight tap index
struct Unit {
    int8_t map;            // A tensor map index
    int8_t c1, c2;         // TMA coordinate deltas (x, y) relative to the tile origin
    int8_t ntaps;
    uint16_t rows;         // rows the TMA box writes (bw * extended height * bn): expect_tx = rows*128
    UnitTap taps[4];
};

struct Umma2Plan {
    int units_per_chunk;
    Unit units[6];
    int chunks;                          // cin / 32
    int step_x, step_y, off_x, off_y;    // tile origin in A-map coordinates = tile index * step + off
    int bw, bh, bn;                      // output pixel box per tile (normal epilogue)
    int tiles_x, tiles_y, m_tiles, n_tiles;
    int img_first;
    int base_off_mode;                   // 0: descriptor base_offset = 0; 1: (addr >> 7) & 7
    int pair;                            // resident-weight kernels: run as CTA pairs (cta_group::2); maps.b must then box BN/2 rows
    int pdl;                             // launch with programmatic stream serialization (overlap prologue with the previous conv's tail)
    int sk_seq;                          // stream-K: value a partial's flag takes in THIS launch (unique per launch, never 0)
    int debug;                           // timing experiments only (results are garbage): bit0 skip B fills, bit1 skip A fills
};

cudaError_t launch_conv_umma2(const UmmaMaps& maps, const ConvGeom& g, const Umma2Plan& t, const ConvPtrs& p,
                              int block_n, bool resident, int kind /*KIND_*/, int m_per_cta, int prec /*0 tf32, 1 bf16x3, 2 bf16*/,
                              int num_sms, cudaStream_t stream);

cudaError_t launch_conv_umma(const UmmaMaps& maps, const ConvGeom& g, const UmmaTiling& t,
                             const ConvPtrs& p, int block_n, int num_sms, cudaStream_t stream);
cudaError_t launch_conv_direct(const ConvGeom& g, const ConvPtrs& p, cudaStream_t stream);

}  // namespace se3tn

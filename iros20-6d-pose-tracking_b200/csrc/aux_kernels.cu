// HBM-bound fp32/fp64 kernels either side of the conv stack: frame crop + depth clip + channel
// normalisation (K0), layout packing, 3x3/s2 max-pool, avg-pool + FC + tanh head (K4), and the
// R^3 x so(3) pose update / label (K6 / K5).  Each replaces numpy/cv2 CPU code in the reference;
// the file:line each follows is cited at the kernel.
#include "aux_kernels.h"
#include "bbox.cuh"
#include "ptx.cuh"
#include <cfloat>
#include <cuda_bf16.h>

namespace se3tn {

// =============================================================================================
// K0: crop window + nearest resize + depth offset/clip + (x-mean)/std  -> padded NHWC4
// =============================================================================================
// bbox:   reference Utils.py:302-316 (compute_bbox): 4 corners of an object_width-mm square at the
//         object depth, projected in float64, np.round (half-to-even == rint) to int32.
// crop:   reference Utils.py:320-359 (crop_bbox): window copy with zero padding outside the frame,
//         then cv2.resize(INTER_NEAREST): src = min(floor(dst * (1/(dsize/ssize))), ssize-1).
// depth:  reference data_augmentation.py:134-144: invalid = d<=100 || d>=2000 (raw mm), then
//         float32(double(d) -/+ z*1000), invalid -> 2000.  (numpy>=2 semantics, see oracle header.)
// norm:   reference data_augmentation.py:154-164: (x - mean[c]) / std[c]; float32 arithmetic when
//         mean/std are float32 arrays (what train.py:121-125 saves), float64 otherwise.
// pack:   reference data_augmentation.py:179-189 builds CHW float32; here the result goes straight
//         into the stem conv's zero-padded NHWC4 layout (and optionally to NCHW for the drop-in API).

// conv-input storage modes: 0 raw fp32, 1 fp32 words rounded to tf32, 2 per pixel [4 x bf16 hi | 4 x bf16 lo]
__device__ __forceinline__ float4 pack_stem_pixel(float4 v, int mode) {
    if (mode == 1) return make_float4(ptx::to_tf32(v.x), ptx::to_tf32(v.y), ptx::to_tf32(v.z), ptx::to_tf32(v.w));
    if (mode == 2) {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
        float4 o;
        o.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h01)); o.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h23));
        o.z = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l01)); o.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l23));
        return o;
    }
    return v;
}

__device__ __forceinline__ float norm_f32(float x, float m, float s) { return __fdiv_rn(__fsub_rn(x, m), s); }
__device__ __forceinline__ float norm_f64(float x, double m, double s) { return static_cast<float>(__ddiv_rn(__dsub_rn(static_cast<double>(x), m), s)); }

__device__ __forceinline__ float depth_offset(unsigned d, double z1000, bool gl) {
    if (d <= 100u || d >= 2000u) return 2000.f;
    return static_cast<float>(gl ? __dadd_rn(static_cast<double>(d), z1000) : __dsub_rn(static_cast<double>(d), z1000));
}

// One CTA = one quarter (44 rows) of one track's 176x176 window.  Everything that depends only on the pixel VALUE is tabulated once
// per CTA with the same IEEE operations the per-pixel code would use (bit-identical results):
//   * (v - mean) / std of an 8-bit colour value: 6 channels x 256 entries;
//   * the whole depth chain  u16 mm -> clip -> float32(double(d) -+ z*1000) -> (x - mean) / std : a function of d alone for a given
//     track, non-constant only for 100 < d < 2000 -> 2 x 1899 entries (A and B use different statistics);
//   * cv2's nearest-neighbour source index floor(dst * (1 / (176 / size))) for the 176 columns and this CTA's 44 rows.
// The per-pixel work is then byte loads, table look-ups, the bf16 / tf32 packing and two 16-byte stores: ~3x fewer instructions
// than dividing per pixel (ncu, round 2: the kernel was issue-bound at 274 instructions per pixel, DRAM at 7 %).
constexpr int kPreRowsMax = 88;                   // most rows one CTA handles (rows per CTA is a launch parameter: 88, 22 or 11)
constexpr int kDepthLo = 101, kDepthN = 1899;     // valid raw depths 101..1999

// THREADS = 256 (strips of 11 / 22 rows) or 1024 (half an image per CTA: the per-CTA tables are built twice per track instead of eight times)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 4 : 1)
preprocess_kernel(PreprocessArgs a, int rows_per_cta)
{
    ptx::grid_dep_launch();
    const int n = blockIdx.y;
    const int row0 = blockIdx.x * rows_per_cta;
    const double* pose = a.poses + n * 16;
    __shared__ float s_lut[6][256];
    __shared__ float s_dlut[2][kDepthN];
    __shared__ float s_dinv[2];                    // normalised value of an invalid depth (2000)
    __shared__ short s_sx[kImg], s_sy[kPreRowsMax];
    __shared__ int s_win[4];
    const int wi = a.weight_ids ? min(max(a.weight_ids[n], 0), a.stats_rows - 1) : 0;   // ids without statistics are rejected on the host where it can see them; never index past the table
    if (threadIdx.x < 256) {
        const float v = static_cast<float>(threadIdx.x);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int mc = c < 3 ? c : c + 1;                   // A: mean[0..2], B: mean[4..6]
            s_lut[c][threadIdx.x] = a.stats_f64 ? norm_f64(v, a.mean64[wi * 8 + mc], a.std64[wi * 8 + mc])
                                                : norm_f32(v, a.mean32[wi * 8 + mc], a.std32[wi * 8 + mc]);
        }
    }
    ptx::grid_dep_wait();                                       // poses come from the previous step's pose update
    const double z = pose[11];
    const bool gl = z < 0;
    const double z1000 = __dmul_rn(z, 1000.0);
    for (int i = threadIdx.x; i < 2 * kDepthN + 2; i += blockDim.x) {
        const int which = i >= kDepthN + 1;                     // 0: A statistics (channel 3), 1: B statistics (channel 7)
        const int j = which ? i - (kDepthN + 1) : i;            // j == kDepthN: the invalid-depth value
        const float zf = (j == kDepthN) ? 2000.f : depth_offset(static_cast<unsigned>(kDepthLo + j), z1000, gl);
        const int mc = which ? 7 : 3;
        const float r = a.stats_f64 ? norm_f64(zf, a.mean64[wi * 8 + mc], a.std64[wi * 8 + mc]) : norm_f32(zf, a.mean32[wi * 8 + mc], a.std32[wi * 8 + mc]);
        if (j == kDepthN) s_dinv[which] = r; else s_dlut[which][j] = r;
    }
    if (!a.b_precropped) {
        if (threadIdx.x == 0) {
            int top, left, ch, cw;
            bbox_window(pose, a.fx, a.fy, a.cx, a.cy, a.object_width[n], 1000.0, 1000.0, 1000.0, top, left, ch, cw);
            s_win[0] = top; s_win[1] = left; s_win[2] = ch; s_win[3] = cw;
        }
        __syncthreads();
        const int ch = s_win[2], cw = s_win[3];
        // cv2 resizeNN index: floor(dst * ifx), ifx = 1/(dsize/ssize) in double, clamped to ssize-1
        const double ifx = (cw > 0) ? 1.0 / (static_cast<double>(kImg) / cw) : 0.0;
        const double ify = (ch > 0) ? 1.0 / (static_cast<double>(kImg) / ch) : 0.0;
        if (threadIdx.x < kImg) { int sx = static_cast<int>(floor(threadIdx.x * ifx)); if (sx > cw - 1) sx = cw - 1; s_sx[threadIdx.x] = static_cast<short>(sx); }
        constexpr int kSyFirst = THREADS == 256 ? 192 : 256;     // threads that fill the row table (the column table takes 0..175)
        if (threadIdx.x >= kSyFirst && threadIdx.x < kSyFirst + rows_per_cta) {
            const int y = row0 + threadIdx.x - kSyFirst;
            int sy = static_cast<int>(floor(y * ify)); if (sy > ch - 1) sy = ch - 1; s_sy[threadIdx.x - kSyFirst] = static_cast<short>(sy);
        }
    }
    __syncthreads();
    const int top = s_win[0], left = s_win[1], ch = s_win[2], cw = s_win[3];
    const size_t img0 = static_cast<size_t>(n) * kImg * kImg;
    auto depth_norm = [&](unsigned d, int which) -> float {
        const unsigned j = d - kDepthLo;
        return j < static_cast<unsigned>(kDepthN) ? s_dlut[which][j] : s_dinv[which];
    };
    // consecutive threads take consecutive pixels: every load / store instruction of a warp touches 32 consecutive pixels
    // (the 16-byte stem stores are 512 contiguous bytes per instruction); four pixels per thread and iteration so that four
    // sets of byte loads are in flight at once (the loop is latency-bound otherwise)
    const int npix = rows_per_cta * kImg;
    for (int lp0 = threadIdx.x; lp0 < npix; lp0 += 4 * blockDim.x) {
        unsigned rA[4], gA[4], bA[4], dA[4], rB[4], gB[4], bB[4], dB[4];
        int pixs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int lp = lp0 + u * blockDim.x;
            rA[u] = gA[u] = bA[u] = dA[u] = rB[u] = gB[u] = bB[u] = dB[u] = 0; pixs[u] = -1;
            if (lp >= npix) continue;
            const int ly = lp / kImg, x = lp - ly * kImg;
            const int pix = (row0 + ly) * kImg + x;
            pixs[u] = pix;
            const size_t ao = img0 + pix;
            // ---- B: observed frame crop ------------------------------------------------------------------
            if (a.b_precropped) {
                // frame_rgb / frame_depth already hold n 176x176 crops (TrackDataset.processData's inputs)
                const uint8_t* pr = a.frame_rgb + ao * 3;
                rB[u] = pr[0]; gB[u] = pr[1]; bB[u] = pr[2];
                dB[u] = a.frame_depth[ao];
            } else if (ch > 0 && cw > 0) {
                const int fy_ = top + s_sy[ly], fx_ = left + s_sx[x];
                if (fy_ >= 0 && fy_ < a.H && fx_ >= 0 && fx_ < a.W) {
                    const size_t fo = static_cast<size_t>(fy_) * a.W + fx_;
                    const uint8_t* pr = a.frame_rgb + fo * 3;
                    rB[u] = pr[0]; gB[u] = pr[1]; bB[u] = pr[2];
                    dB[u] = a.frame_depth[fo];
                }
            }
            // ---- A: rendered previous view --------------------------------------------------------------
            const uint8_t* pa = a.rgbA + ao * 3;
            rA[u] = pa[0]; gA[u] = pa[1]; bA[u] = pa[2]; dA[u] = a.depthA[ao];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pix = pixs[u];
            if (pix < 0) continue;
            const int y = pix / kImg, x = pix - y * kImg;
            const size_t ao = img0 + pix;
            if (a.crop_rgb) { uint8_t* o = a.crop_rgb + ao * 3; o[0] = static_cast<uint8_t>(rB[u]); o[1] = static_cast<uint8_t>(gB[u]); o[2] = static_cast<uint8_t>(bB[u]); }
            if (a.crop_depth) a.crop_depth[ao] = static_cast<uint16_t>(dB[u]);
            const float4 vA = make_float4(s_lut[0][rA[u]], s_lut[1][gA[u]], s_lut[2][bA[u]], depth_norm(dA[u], 0));
            const float4 vB = make_float4(s_lut[3][rB[u]], s_lut[4][gB[u]], s_lut[5][bB[u]], depth_norm(dB[u], 1));
            if (a.nchwA) {
                float* oa = a.nchwA + static_cast<size_t>(n) * 4 * kImg * kImg + pix;
                float* ob = a.nchwB + static_cast<size_t>(n) * 4 * kImg * kImg + pix;
                oa[0] = vA.x; oa[kImg * kImg] = vA.y; oa[2 * kImg * kImg] = vA.z; oa[3 * kImg * kImg] = vA.w;
                ob[0] = vB.x; ob[kImg * kImg] = vB.y; ob[2 * kImg * kImg] = vB.z; ob[3 * kImg * kImg] = vB.w;
            }
            if (a.stemA) {
                const size_t so = (static_cast<size_t>(n) * kStemH + (y + 3)) * kStemW + (x + 3);
                reinterpret_cast<float4*>(a.stemA)[so] = pack_stem_pixel(vA, a.round_tf32);
                reinterpret_cast<float4*>(a.stemB)[so] = pack_stem_pixel(vB, a.round_tf32);
            }
        }
    }
}


static cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, void** args, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, func, args);
}

cudaError_t launch_preprocess(const PreprocessArgs& a, int n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    // rows per CTA: big strips amortise the per-CTA tables, small ones fill the machine when there are few tracks
    int rows = n >= 64 ? 88 : (n >= 32 ? 22 : 11);
    dim3 grid(kImg / rows, n);
    PreprocessArgs aa = a;
    void* args[] = {&aa, &rows};
    if (rows == 88) return launch_pdl(reinterpret_cast<const void*>(preprocess_kernel<1024>), grid, dim3(1024), args, s);
    return launch_pdl(reinterpret_cast<const void*>(preprocess_kernel<256>), grid, dim3(256), args, s);
}

// =============================================================================================
// Stand-alone compute_bbox / crop_bbox (reference Utils.py:302-316, 320-359) for the drop-in
// Utils API: the same arithmetic as inside preprocess_kernel, with the bbox made visible.
// =============================================================================================
__global__ void bbox_kernel(const double* __restrict__ poses, double fx, double fy, double cx, double cy,
                            const double* __restrict__ widths, double sx, double sy, double sz,
                            int* __restrict__ out /* (n,4,2) rows (x-,y-),(x-,y+),(x+,y-),(x+,y+), cols (v,u) */, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* pose = poses + i * 16;
    const double ox = __dmul_rn(pose[3], sx), oy = __dmul_rn(pose[7], sy), oz = __dmul_rn(pose[11], sz);
    const double half = widths[i] / 2;
    const double lim = 2.0e9;
    auto proj = [&](double v, double f, double c) {
        const double p = rint(__dadd_rn(__ddiv_rn(__dmul_rn(v, f), oz), c));
        return static_cast<int>(fmax(-lim, fmin(lim, p == p ? p : 0.0)));
    };
    const int u0 = proj(ox - half, fx, cx), u1 = proj(ox + half, fx, cx);
    const int v0 = proj(oy - half, fy, cy), v1 = proj(oy + half, fy, cy);
    int* o = out + i * 8;
    o[0] = v0; o[1] = u0; o[2] = v1; o[3] = u0; o[4] = v0; o[5] = u1; o[6] = v1; o[7] = u1;
}

cudaError_t launch_bbox(const double* poses, const double* K4, const double* widths, const double* scale3,
                        int* out, int n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    bbox_kernel<<<(n + 127) / 128, 128, 0, s>>>(poses, K4[0], K4[1], K4[2], K4[3], widths, scale3[0], scale3[1], scale3[2], out, n);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
crop_kernel(const uint8_t* __restrict__ frame_rgb, const uint16_t* __restrict__ frame_depth, int H, int W,
            const int* __restrict__ bbox, int out_h, int out_w, uint8_t* __restrict__ crop_rgb, uint16_t* __restrict__ crop_depth)
{
    const int n = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out_h * out_w) return;
    const int y = pix / out_w, x = pix - y * out_w;
    const int* bb = bbox + n * 8;
    const int top = min(min(bb[0], bb[2]), min(bb[4], bb[6])), bottom = max(max(bb[0], bb[2]), max(bb[4], bb[6]));
    const int left = min(min(bb[1], bb[3]), min(bb[5], bb[7])), right = max(max(bb[1], bb[3]), max(bb[5], bb[7]));
    const int ch = bottom - top, cw = right - left;
    unsigned r = 0, g = 0, b = 0, d = 0;
    if (ch > 0 && cw > 0) {
        const double ifx = 1.0 / (static_cast<double>(out_w) / cw);
        const double ify = 1.0 / (static_cast<double>(out_h) / ch);
        int sx = static_cast<int>(floor(x * ifx)); if (sx > cw - 1) sx = cw - 1;
        int sy = static_cast<int>(floor(y * ify)); if (sy > ch - 1) sy = ch - 1;
        const int fy_ = top + sy, fx_ = left + sx;
        if (fy_ >= 0 && fy_ < H && fx_ >= 0 && fx_ < W) {
            const size_t fo = static_cast<size_t>(fy_) * W + fx_;
            r = frame_rgb[fo * 3]; g = frame_rgb[fo * 3 + 1]; b = frame_rgb[fo * 3 + 2];
            d = frame_depth[fo];
        }
    }
    const size_t o = static_cast<size_t>(n) * out_h * out_w + pix;
    crop_rgb[o * 3] = r; crop_rgb[o * 3 + 1] = g; crop_rgb[o * 3 + 2] = b;
    crop_depth[o] = static_cast<uint16_t>(d);
}

cudaError_t launch_crop(const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W, const int* bbox, int n,
                        int out_h, int out_w, uint8_t* crop_rgb, uint16_t* crop_depth, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dim3 grid((out_h * out_w + 255) / 256, n);
    crop_kernel<<<grid, 256, 0, s>>>(frame_rgb, frame_depth, H, W, bbox, out_h, out_w, crop_rgb, crop_depth);
    return cudaGetLastError();
}

// =============================================================================================
// NCHW float32 (N,4,176,176) -> zero-padded NHWC4 stem input (for Se3TrackNet.forward(A, B))
// =============================================================================================
__global__ void __launch_bounds__(256)
nchw_to_stem_kernel(const float* __restrict__ src, float* __restrict__ dst, int round_tf32)
{
    const int n = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= kImg * kImg) return;
    const int y = pix / kImg, x = pix - y * kImg;
    const float* s = src + static_cast<size_t>(n) * 4 * kImg * kImg + pix;
    float4 v = make_float4(s[0], s[kImg * kImg], s[2 * kImg * kImg], s[3 * kImg * kImg]);
    v = pack_stem_pixel(v, round_tf32);
    reinterpret_cast<float4*>(dst)[(static_cast<size_t>(n) * kStemH + (y + 3)) * kStemW + (x + 3)] = v;
}

cudaError_t launch_nchw_to_stem(const float* src, float* dst, int n, int round_tf32, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dim3 grid((kImg * kImg + 255) / 256, n);
    nchw_to_stem_kernel<<<grid, 256, 0, s>>>(src, dst, round_tf32);
    return cudaGetLastError();
}

// =============================================================================================
// MaxPool2d(3, 2, 1) on NHWC (reference se3_tracknet.py:58,62,85,89).  Padding behaves as -inf
// (the SELU output it follows can be negative).
// =============================================================================================
__global__ void __launch_bounds__(256)
maxpool_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n_img, int Hin, int Win, int C4)
{
    const int Ho = Hin / 2, Wo = Win / 2;
    const long long total = static_cast<long long>(n_img) * Ho * Wo * C4;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C4);
    long long r = i / C4;
    const int ox = static_cast<int>(r % Wo); r /= Wo;
    const int oy = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * oy + dy;
        if (y < 0 || y >= Hin) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int x = 2 * ox + dx;
            if (x < 0 || x >= Win) continue;
            const float4 v = __ldg(&in[((static_cast<size_t>(n) * Hin + y) * Win + x) * C4 + c]);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    out[i] = m;
}

cudaError_t launch_maxpool(const float* in, float* out, int n_img, int Hin, int Win, int C, cudaStream_t s) {
    const long long total = static_cast<long long>(n_img) * (Hin / 2) * (Win / 2) * (C / 4);
    if (total <= 0) return cudaSuccess;
    maxpool_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n_img, Hin, Win, C / 4);
    return cudaGetLastError();
}

// =============================================================================================
// K4 head: AdaptiveAvgPool2d(1) + Linear(512,3) + Tanh for both heads
// (reference se3_tracknet.py:100-102, 107-109).  x: NHWC (N, 11*11, 1024): channels [0,512) are
// the translation head, [512,1024) the rotation head.  One CTA (256 threads x 4 channels) per image.
// =============================================================================================
__global__ void __launch_bounds__(512)
head_kernel(const float4* __restrict__ x, const float* __restrict__ fcw /*[6][512]*/, const float* __restrict__ fcb /*[6]*/,
            float* __restrict__ out_trans, float* __restrict__ out_rot, int npix, int split_bf16,
            const int* __restrict__ img_wid, const float* const* __restrict__ fc_table)
{
    // grid (n, 2): blockIdx.y = head (0 trans: channels 0..511, 1 rot: 512..1023).  512 threads =
    // 4 pixel groups x 128 threads, each thread 4 channels.
    ptx::grid_dep_launch();
    __shared__ float4 part4[4][128];
    __shared__ float red[4][3];
    const int n = blockIdx.x, head = blockIdx.y, t = threadIdx.x;
    const int cq = t & 127, pg = t >> 7;
    const int c = head * 512 + cq * 4;                 // first of this thread's 4 channels
    ptx::grid_dep_wait();
    if (img_wid) { fcw = fc_table[img_wid[n]]; fcb = fcw + 6 * 512; }    // per-object head weights
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!split_bf16) {
        const float4* xp = x + static_cast<size_t>(n) * npix * 256 + (c >> 2);
        for (int p = pg; p < npix; p += 4) {
            const float4 v = __ldg(xp + static_cast<size_t>(p) * 256);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    } else {
        // channels c..c+3 live in chunk c/32 as bf16 hi at byte (c%32)*2 and lo 64 bytes further
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(x) + static_cast<size_t>(n) * npix * 4096 + (c >> 5) * 128 + (c & 31) * 2;
        for (int p = pg; p < npix; p += 4) {
            const uint2 h = __ldg(reinterpret_cast<const uint2*>(xb + static_cast<size_t>(p) * 4096));
            const uint2 l = __ldg(reinterpret_cast<const uint2*>(xb + static_cast<size_t>(p) * 4096 + 64));
            const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x)), h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
            const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x)), l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
            s.x += h0.x + l0.x; s.y += h0.y + l0.y; s.z += h1.x + l1.x; s.w += h1.y + l1.y;
        }
    }
    part4[pg][cq] = s;
    __syncthreads();
    if (t < 128) {
        const float4 a0 = part4[0][t], a1 = part4[1][t], a2 = part4[2][t], a3 = part4[3][t];
        const float inv = 1.0f / static_cast<float>(npix);
        const float mx = ((a0.x + a1.x) + (a2.x + a3.x)) * inv, my = ((a0.y + a1.y) + (a2.y + a3.y)) * inv;
        const float mz = ((a0.z + a1.z) + (a2.z + a3.z)) * inv, mw = ((a0.w + a1.w) + (a2.w + a3.w)) * inv;
        float part[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(fcw + (head * 3 + o) * 512 + t * 4));
            part[o] = mx * w.x + my * w.y + mz * w.z + mw * w.w;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int o = 0; o < 3; ++o) part[o] += __shfl_xor_sync(0xffffffffu, part[o], off);
        if ((t & 31) == 0) { red[t >> 5][0] = part[0]; red[t >> 5][1] = part[1]; red[t >> 5][2] = part[2]; }
    }
    __syncthreads();
    if (t < 3) {
        const float v = red[0][t] + red[1][t] + red[2][t] + red[3][t] + fcb[head * 3 + t];
        (head == 0 ? out_trans : out_rot)[n * 3 + t] = tanhf(v);
    }
}

// =============================================================================================
// K6 pose update (reference datasets.py:159-175): t' = t + float32(trans*tn);
// R' = float32(Rodrigues(float32(rot*rn))) . R, all remaining arithmetic in float64 (F9/F10).
// Rodrigues follows OpenCV's cvRodrigues2 vector->matrix branch:
//   theta = |r|; theta < DBL_EPSILON -> I; else R = cos*I + (1-cos)*rr^T + sin*[r]x, r <- r/theta.
// =============================================================================================
__device__ __forceinline__ void rodrigues_exp_f32in(float rx32, float ry32, float rz32, double R[9], bool round_f32)
{
    double rx = rx32, ry = ry32, rz = rz32;
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    rx *= it; ry *= it; rz *= it;
    // same association as cv::Matx: c*I + c1*(r r^T) + s*[r]x, outer products formed first
    const double xx = rx * rx, xy = rx * ry, xz = rx * rz, yy = ry * ry, yz = ry * rz, zz = rz * rz;
    R[0] = c + c1 * xx;      R[1] = c1 * xy - s * rz; R[2] = c1 * xz + s * ry;
    R[3] = c1 * xy + s * rz; R[4] = c + c1 * yy;      R[5] = c1 * yz - s * rx;
    R[6] = c1 * xz - s * ry; R[7] = c1 * yz + s * rx; R[8] = c + c1 * zz;
    if (round_f32)
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = static_cast<double>(static_cast<float>(R[i]));
}

// one track: A (row-major 4x4) and the network's 3+3 output -> B
__device__ __forceinline__ void pose_update_one(const double* A, const float tr[3], const float ro[3], float tn, float rn,
                                                double* B /* may alias A */)
{
    // float32 * python-float stays float32 in numpy
    const float t0 = __fmul_rn(tr[0], tn), t1 = __fmul_rn(tr[1], tn), t2 = __fmul_rn(tr[2], tn);
    const float r0 = __fmul_rn(ro[0], rn), r1 = __fmul_rn(ro[1], rn), r2 = __fmul_rn(ro[2], rn);
    double R[9];
    rodrigues_exp_f32in(r0, r1, r2, R, true);
    double out[16];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[r * 4 + c] = __dadd_rn(__dadd_rn(__dmul_rn(R[r * 3 + 0], A[0 * 4 + c]), __dmul_rn(R[r * 3 + 1], A[1 * 4 + c])),
                                       __dmul_rn(R[r * 3 + 2], A[2 * 4 + c]));
    out[3] = static_cast<double>(t0) + A[3];
    out[7] = static_cast<double>(t1) + A[7];
    out[11] = static_cast<double>(t2) + A[11];
    out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;        // B_in_cam starts as np.eye(4)
#pragma unroll
    for (int k = 0; k < 16; ++k) B[k] = out[k];
}

// Head on the fused average pool (conv_umma2.cu writes pool_part[image][4 row quadrants][1024] column sums): mean -> Linear -> tanh for
// BOTH heads of one image per CTA (threads 0-127 translation, 128-255 rotation), and -- when poses_in is given -- the pose update of that
// track by thread 0 (K4 + K6 in one launch: the update is a 650-instruction fp64 chain per track, pure latency as its own kernel).
// `zero_words` (nullable): scheduler / dependency counters of the step that just finished, cleared for the next one by block 0.
__global__ void __launch_bounds__(256)
head_pooled_kernel(const float4* __restrict__ part, const float* __restrict__ fcw, const float* __restrict__ fcb,
                   float* __restrict__ out_trans, float* __restrict__ out_rot, int npix,
                   const int* __restrict__ img_wid, const float* const* __restrict__ fc_table,
                   const double* poses_in, double* poses_out /* may alias */, float tn, float rn,
                   unsigned* __restrict__ zero_words, int n_zero)
{
    ptx::grid_dep_launch();
    __shared__ float red[8][3];
    __shared__ float six[6];
    const int n = blockIdx.x, head = threadIdx.x >> 7, t = threadIdx.x & 127;
    ptx::grid_dep_wait();
    if (zero_words && blockIdx.x == 0) for (int i = threadIdx.x; i < n_zero; i += blockDim.x) zero_words[i] = 0u;
    if (img_wid) { fcw = fc_table[img_wid[n]]; fcb = fcw + 6 * 512; }
    const float4* pp = part + static_cast<size_t>(n) * 4 * 256 + head * 128 + t;
    const float4 a0 = pp[0], a1 = pp[256], a2 = pp[512], a3 = pp[768];
    const float inv = 1.0f / static_cast<float>(npix);
    const float mx = ((a0.x + a1.x) + (a2.x + a3.x)) * inv, my = ((a0.y + a1.y) + (a2.y + a3.y)) * inv;
    const float mz = ((a0.z + a1.z) + (a2.z + a3.z)) * inv, mw = ((a0.w + a1.w) + (a2.w + a3.w)) * inv;
    float acc[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(fcw + (head * 3 + o) * 512 + t * 4));
        acc[o] = mx * w.x + my * w.y + mz * w.z + mw * w.w;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int o = 0; o < 3; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
    if ((t & 31) == 0) { red[threadIdx.x >> 5][0] = acc[0]; red[threadIdx.x >> 5][1] = acc[1]; red[threadIdx.x >> 5][2] = acc[2]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int h = threadIdx.x / 3, o = threadIdx.x - 3 * h;
        const float v = tanhf(red[4 * h + 0][o] + red[4 * h + 1][o] + red[4 * h + 2][o] + red[4 * h + 3][o] + fcb[h * 3 + o]);
        (h == 0 ? out_trans : out_rot)[n * 3 + o] = v;
        six[threadIdx.x] = v;
    }
    if (!poses_in) return;
    __syncthreads();
    if (threadIdx.x == 0) pose_update_one(poses_in + n * 16, six, six + 3, tn, rn, poses_out + n * 16);
}
cudaError_t launch_head_pooled(const float* part, const float* fcw, const float* fcb, float* out_trans, float* out_rot,
                               int n_img, int npix, const int* img_wid, const float* const* fc_table,
                               const double* poses_in, double* poses_out, float tn, float rn, unsigned* zero_words, int n_zero, cudaStream_t s) {
    if (n_img <= 0) return cudaSuccess;
    const float4* p4 = reinterpret_cast<const float4*>(part);
    void* args[] = {&p4, &fcw, &fcb, &out_trans, &out_rot, &npix, &img_wid, &fc_table, &poses_in, &poses_out, &tn, &rn, &zero_words, &n_zero};
    return launch_pdl(reinterpret_cast<const void*>(head_pooled_kernel), dim3(n_img), dim3(256), args, s);
}

cudaError_t launch_head(const float* x, const float* fcw, const float* fcb, float* out_trans, float* out_rot,
                        int n_img, int npix, int split_bf16, const int* img_wid, const float* const* fc_table, cudaStream_t s) {
    if (n_img <= 0) return cudaSuccess;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    void* args[] = {&x4, &fcw, &fcb, &out_trans, &out_rot, &npix, &split_bf16, &img_wid, &fc_table};
    return launch_pdl(reinterpret_cast<const void*>(head_kernel), dim3(n_img, 2), dim3(512), args, s);
}

// =============================================================================================
// NHWC -> NCHW (the 'feature' entry of the reference's output dict, se3_tracknet.py:96)
// =============================================================================================
// storage: 0 fp32 words, 1 [32 x bf16 hi | 32 x bf16 lo] per 32-channel chunk, 2 plain bf16 (conv_common.h)
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int HW, int C, int storage)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + tx;
        float val = 0.f;
        if (p < HW && c < C) {
            const size_t pix = static_cast<size_t>(n) * HW + p;
            if (storage == 0) val = reinterpret_cast<const float*>(in)[pix * C + c];
            else if (storage == 1) {
                const uint8_t* cb = in + (pix * C + (c & ~31)) * 4 + (c & 31) * 2;
                val = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(cb)) + __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(cb + 64));
            } else val = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[pix * C + c]);
        }
        tile[j][tx] = val;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + tx;
        if (p < HW && c < C) out[(static_cast<size_t>(n) * C + c) * HW + p] = tile[tx][j];
    }
}

cudaError_t launch_nhwc_to_nchw(const void* in, float* out, int n_img, int HW, int C, int storage, cudaStream_t s) {
    if (n_img <= 0) return cudaSuccess;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, n_img);
    nhwc_to_nchw_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(in), out, HW, C, storage);
    return cudaGetLastError();
}

// =============================================================================================
// Weight preparation for the bf16 modes (conv_umma2.cu PREC_BF16X3 / PREC_BF16).
// Trunk layers, PREC_BF16X3: every 32-word K chunk of a weight row becomes [32 x bf16 hi | 32 x bf16 lo].
// =============================================================================================
__global__ void split_weights_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t words)
{
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < words; i += stride) {
        const float v = src[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const size_t chunk = i >> 5, j = i & 31;
        *reinterpret_cast<__nv_bfloat16*>(dst + chunk * 128 + j * 2) = h;
        *reinterpret_cast<__nv_bfloat16*>(dst + chunk * 128 + 64 + j * 2) = l;
    }
}

// plain bf16 copy of a K-major weight matrix (PREC_BF16: 64 channels per 128-byte K chunk)
__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n)
{
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}
cudaError_t launch_to_bf16(const float* src, void* dst, size_t n, cudaStream_t s) {
    if (!n) return cudaSuccess;
    to_bf16_kernel<<<1024, 256, 0, s>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    return cudaGetLastError();
}

// STACK layouts for the resident-weight kernels (conv_umma2.cu): 128 rows, rows 0-63 carry the hi parts, rows 64-127 the lo parts.
// 64-channel 3x3 layers: src [64][9*64] (K-major, tap*64 + c) -> dst [128][9*32 words]; a tap's 128 bytes = 64 bf16 = both chunks.
__global__ void split_stack_weights_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 576) return;
    const int co = i / 576, k = i - co * 576;
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(dst + (static_cast<size_t>(co) * 288) * 4 + k * 2) = h;
    *reinterpret_cast<__nv_bfloat16*>(dst + (static_cast<size_t>(64 + co) * 288) * 4 + k * 2) = l;
}
// stem: src [64][7*32] -> dst [128][7*32 words]; pixel slot p of filter row r: rows 0-63 [h0..h3 h0..h3], rows 64-127 [l0..l3 0 0 0 0]
__global__ void split_stem_stack_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (co, r, p): 4 channels
    if (i >= 64 * 7 * 8) return;
    const int p = i & 7, r = (i >> 3) % 7, co = i / 56;
    const float* w = src + co * 224 + r * 32 + p * 4;
    __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(dst + (static_cast<size_t>(co) * 224 + r * 32 + p * 4) * 4);
    __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(dst + (static_cast<size_t>(64 + co) * 224 + r * 32 + p * 4) * 4);
    for (int c = 0; c < 4; ++c) {
        const __nv_bfloat16 h = __float2bfloat16_rn(w[c]);
        d0[c] = h; d0[4 + c] = h;
        d1[c] = __float2bfloat16_rn(w[c] - __bfloat162float(h)); d1[4 + c] = __float2bfloat16_rn(0.f);
    }
}
// Resident 64-channel layers (conv_umma2.cu, 16x256b epilogue): accumulator column 8j + 2m + e of every 32-column block
// must carry output channel 8m + 2j + e, so the weight ROWS are stored in that order.
__global__ void permute_rows64_kernel(const float* __restrict__ src, float* __restrict__ dst, int ktot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * ktot) return;
    const int col = i / ktot, k = i - col * ktot;
    const int ch = (col & 32) | (((col >> 1) & 3) << 3) | (((col >> 3) & 3) << 1) | (col & 1);
    dst[i] = src[ch * ktot + k];
}
cudaError_t launch_permute_rows64(const float* src, float* dst, int ktot, cudaStream_t s) {
    permute_rows64_kernel<<<(64 * ktot + 255) / 256, 256, 0, s>>>(src, dst, ktot);
    return cudaGetLastError();
}
cudaError_t launch_split_stack_weights(const float* src, void* dst, bool stem, cudaStream_t s) {
    if (stem) split_stem_stack_kernel<<<(64 * 7 * 8 + 127) / 128, 128, 0, s>>>(src, static_cast<uint8_t*>(dst));
    else split_stack_weights_kernel<<<(64 * 576 + 255) / 256, 256, 0, s>>>(src, static_cast<uint8_t*>(dst));
    return cudaGetLastError();
}

cudaError_t launch_split_weights(const float* src, void* dst, size_t words, cudaStream_t s) {
    if (!words) return cudaSuccess;
    split_weights_kernel<<<1024, 256, 0, s>>>(src, static_cast<uint8_t*>(dst), words);
    return cudaGetLastError();
}
// stand-alone K6 (se3tn_pose_update; the batched path runs it inside head_pooled_kernel)
__global__ void pose_update_kernel(const double* poses_in, const float* __restrict__ trans,
                                   const float* __restrict__ rot, float tn, float rn,
                                   double* poses_out /* may alias poses_in */, int n)
{
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float tr[3] = {trans[i * 3 + 0], trans[i * 3 + 1], trans[i * 3 + 2]};
    const float ro[3] = {rot[i * 3 + 0], rot[i * 3 + 1], rot[i * 3 + 2]};
    pose_update_one(poses_in + i * 16, tr, ro, tn, rn, poses_out + i * 16);
}

cudaError_t launch_pose_update(const double* poses_in, const float* trans, const float* rot, float tn, float rn,
                               double* poses_out, int n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    void* args[] = {&poses_in, &trans, &rot, &tn, &rn, &poses_out, &n};
    return launch_pdl(reinterpret_cast<const void*>(pose_update_kernel), dim3((n + 31) / 32), dim3(32), args, s);
}

// =============================================================================================
// K5 so(3) log label (reference datasets.py:141-150): trans = (tB - tA)/tn;
// rot = Rodrigues^-1(normalize_cols(R_B R_A^T))/rn  (Utils.py:363-367 + cvRodrigues2 matrix->vector
// branch: R <- U V^T from the SVD, then the antisymmetric-part formula with its small-angle cases).
// The SVD's U V^T is the orthogonal polar factor of R; it is computed here with the Newton
// iteration X <- (X + X^-T)/2, which converges quadratically to the same matrix.
// =============================================================================================
__device__ __forceinline__ void inv_transpose3(const double* m, double* o) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
    o[3] = (m[2] * m[7] - m[1] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[1] * m[6] - m[0] * m[7]) * id;
    o[6] = (m[1] * m[5] - m[2] * m[4]) * id; o[7] = (m[2] * m[3] - m[0] * m[5]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

__global__ void so3_log_kernel(const double* __restrict__ poses_a, const double* __restrict__ poses_b,
                               double tn, double rn, double* __restrict__ trans_label, double* __restrict__ rot_label, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* A = poses_a + i * 16; const double* B = poses_b + i * 16;
    trans_label[i * 3 + 0] = (B[3] - A[3]) / tn;
    trans_label[i * 3 + 1] = (B[7] - A[7]) / tn;
    trans_label[i * 3 + 2] = (B[11] - A[11]) / tn;
    double R[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)     // R_B . R_A^T
            R[r * 3 + c] = B[r * 4 + 0] * A[c * 4 + 0] + B[r * 4 + 1] * A[c * 4 + 1] + B[r * 4 + 2] * A[c * 4 + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {       // column normalise
        const double nr = sqrt(R[c] * R[c] + R[3 + c] * R[3 + c] + R[6 + c] * R[6 + c]);
        R[c] /= nr; R[3 + c] /= nr; R[6 + c] /= nr;
    }
    for (int it = 0; it < 12; ++it) {   // polar factor
        double T[9]; inv_transpose3(R, T);
        double diff = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double v = 0.5 * (R[k] + T[k]); diff = fmax(diff, fabs(v - R[k])); R[k] = v; }
        if (diff < 1e-16) break;
    }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : (c < -1. ? -1. : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { rx = ry = rz = 0; }
        else {
            double t;
            t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.));
            t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[5] > 0) != (ry * rz > 0))) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        const double vth = theta / (2 * s);
        rx *= vth; ry *= vth; rz *= vth;
    }
    rot_label[i * 3 + 0] = rx / rn; rot_label[i * 3 + 1] = ry / rn; rot_label[i * 3 + 2] = rz / rn;
}

cudaError_t launch_so3_log(const double* poses_a, const double* poses_b, double tn, double rn,
                           double* trans_label, double* rot_label, int n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    so3_log_kernel<<<(n + 127) / 128, 128, 0, s>>>(poses_a, poses_b, tn, rn, trans_label, rot_label, n);
    return cudaGetLastError();
}

}  // namespace se3tn

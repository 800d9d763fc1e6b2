// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, TMEM accumulators), im2col-free:
// the A operand (activations) is pulled straight out of the NHWC tensor by TMA tile loads, one
// shifted pixel box per filter tap, with TMA's out-of-bounds zero fill standing in for the conv
// padding.  Replaces every cuDNN conv call of the reference's Se3TrackNet.forward
// (se3_tracknet.py:84-106 via network_modules.py:63,82) with BatchNorm folded into W/bias and
// the activation / residual add (network_modules.py:65,105-120) fused into the epilogue.
//
// GEMM view per launch:  D[M = pixels, N = Cout] = sum_{tap, c} A_tap[pixel, c] * W[Cout, tap*cin + c]
//   M tile  = one pixel box (bw x bh x bn <= 128 rows) -> UMMA M = 128 (TMEM lanes)
//   N tile  = BN in {64, 128, 256} output channels     -> UMMA N
//   K step  = 128 bytes (32 tf32) of one tap           -> 4 x (UMMA K = 8)
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      TMA producer (one elected lane)          smem ring: full/empty mbarriers
//   warp 1      MMA issuer   (one elected lane)          tcgen05.mma + tcgen05.commit
//   warp 2      TMEM allocator (2 x BN columns: double-buffered accumulator)
//   warps 4..7  epilogue: tcgen05.ld -> +bias (+residual) -> act -> tf32 round -> NHWC store
#include "conv_common.h"
#include "ptx.cuh"

namespace se3tn {

namespace {

constexpr int kThreads = 256;
constexpr int kABytes = kBlockM * kChunkBytes;   // 16 KB per stage

template <int BN> struct Cfg {
    static constexpr int kBBytes = BN * kChunkBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
    static constexpr int kTmemCols = 2 * BN;     // 128 / 256 / 512: powers of two
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float selu_f(float x) {
    constexpr float kAlpha = 1.6732632423543772f, kScale = 1.0507009873554805f;
    return x > 0.f ? kScale * x : (kScale * kAlpha) * expm1f(x);
}

struct TileCoord { int x0, y0, n0, n_tile, grp; };

__device__ __forceinline__ TileCoord decode_tile(int tile, const UmmaTiling& t) {
    TileCoord c;
    int m = tile % t.m_tiles;
    int rest = tile / t.m_tiles;
    c.n_tile = rest % t.n_tiles;
    c.grp = rest / t.n_tiles;
    int tx = m % t.tiles_x;
    int r2 = m / t.tiles_x;
    int ty = r2 % t.tiles_y;
    int ib = r2 / t.tiles_y;
    c.x0 = tx * t.bw; c.y0 = ty * t.bh; c.n0 = t.img_first + ib * t.bn;
    return c;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ UmmaMaps maps, const ConvGeom g, const UmmaTiling t, const ConvPtrs p)
{
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // [stages][16 KB]
    uint8_t* sB = smem + C::kStages * kABytes;            // [stages][BN*128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = bars;                            // [stages]
    uint64_t* empty_bar = bars + C::kStages;              // [stages]
    uint64_t* tmem_full = bars + 2 * C::kStages;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = t.m_tiles * t.n_tiles * g.groups;
    const int num_k = g.num_taps * t.chunks_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 4); }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&maps.b);
        ptx::prefetch_tmap(&maps.a[0]);
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            const uint32_t a_bytes = static_cast<uint32_t>(t.bw * t.bh * t.bn) * kChunkBytes;
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(tile, t);
                const int cbase = g.in_coff + tc.grp * g.cin;
                const int wrow = tc.grp * g.cout + tc.n_tile * BN;
                for (int tap = 0; tap < g.num_taps; ++tap) {
                    const Tap tp = g.taps[tap];
                    const CUtensorMap* amap = &maps.a[tp.map];
                    for (int ch = 0; ch < t.chunks_per_tap; ++ch) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], a_bytes + C::kBBytes);
                        void* dA = sA + stage * kABytes;
                        void* dB = sB + stage * C::kBBytes;
                        ptx::tma_load_4d(dA, amap, &full_bar[stage], cbase + ch * 32, tc.x0 + tp.c1, tc.y0 + tp.c2, tc.n0);
                        ptx::tma_load_2d(dB, &maps.b, &full_bar[stage], tap * g.cin + ch * 32, wrow);
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc(2 /*tf32*/, kBlockM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < num_k; ++k) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + stage * kABytes);
                    const uint32_t b_addr = ptx::smem_u32(sB + stage * C::kBBytes);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_addr + kk * 32), ptx::umma_desc_sw128(b_addr + kk * 32),
                                       idesc, (k | kk) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);      // frees this smem stage when the MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);            // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ============================== epilogue ==================================
        const int q = warp - 4;                               // TMEM lane quadrant == warp % 4
        const int row = q * 32 + lane;
        const int box = t.bw * t.bh;
        const int pn = row / box;
        const int rem = row - pn * box;
        const int py = rem / t.bw;
        const int px = rem - py * t.bw;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const TileCoord tc = decode_tile(tile, t);
            const int n = tc.n0 + pn, y = tc.y0 + py, x = tc.x0 + px;
            const bool valid = (pn < t.bn) && (n < g.n_img) && (y < g.Ho) && (x < g.Wo);
            const size_t pix = (static_cast<size_t>(n) * g.Ho + y) * g.Wo + x;
            const int ch0 = tc.grp * g.cout + tc.n_tile * BN;
            float* outp = p.out + pix * g.out_cstride + g.out_coff + ch0;
            const float* resp = p.res ? p.res + pix * g.res_cstride + g.res_coff + ch0 : nullptr;
            const float* biasp = p.bias + ch0;

            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                ptx::tmem_ld16(taddr + c0, r);
                ptx::tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(biasp + c0 + j));
                    v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
                    v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
                }
                if (valid) {
                    if (resp) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(resp + c0 + j));
                            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float o = v[j];
                        if (g.act == ACT_RELU) o = fmaxf(o, 0.f);
                        else if (g.act == ACT_SELU) o = selu_f(o);
                        if (g.round_tf32) o = ptx::to_tf32(o);
                        v[j] = o;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(outp + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

template <int BN>
cudaError_t launch_bn(const UmmaMaps& maps, const ConvGeom& g, const UmmaTiling& t, const ConvPtrs& p,
                      int num_sms, cudaStream_t stream) {
    using C = Cfg<BN>;
    static bool attr_set_dev[64] = {};                       // per-device function attribute
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    bool& attr_set = attr_set_dev[dev];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int total = t.m_tiles * t.n_tiles * g.groups;
    const int grid = total < num_sms ? total : num_sms;
    conv_umma_kernel<BN><<<grid, kThreads, C::kSmemBytes, stream>>>(maps, g, t, p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_umma(const UmmaMaps& maps, const ConvGeom& g, const UmmaTiling& t,
                             const ConvPtrs& p, int block_n, int num_sms, cudaStream_t stream) {
    switch (block_n) {
        case 64:  return launch_bn<64>(maps, g, t, p, num_sms, stream);
        case 128: return launch_bn<128>(maps, g, t, p, num_sms, stream);
        case 256: return launch_bn<256>(maps, g, t, p, num_sms, stream);
        default:  return cudaErrorInvalidValue;
    }
}

}  // namespace se3tn

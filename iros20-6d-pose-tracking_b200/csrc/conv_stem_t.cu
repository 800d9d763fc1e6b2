// conv_stem_ws_kernel: the 7x7 stride-2 stem (reference se3_tracknet.py:57-58,61-62: ConvBNReLU(4,64,7,2) + MaxPool2d(3,2,1))
// in the WEIGHTS-STATIONARY formulation, for the bf16 hi/lo input format and one weight set per launch.
//
// Why: conv_resident_kernel<KIND_STEM> is bound by shared-memory operand bandwidth -- every 128 x 128 x 16 MMA reads 4 KB of
// pixels (A) and 4 KB of stacked weights (B) from shared memory, 64 cycles at 128 B/cycle = the whole MMA slot, and its pooling
// epilogue adds 90 KB of shared-memory traffic per tile on top (measured 92 cycles per MMA slot).  Here the
// roles are swapped:
//   D^T[m = stacked output channel, n = pixel] = sum_k Wst[m][k] * Act[n][k]
//   * A operand = the stacked weight matrix Wst = [w_hi|w_hi ; w_lo|0] (128 rows x 7 taps x 64 bf16), written ONCE per CTA into
//     TENSOR MEMORY (tcgen05.st, 224 columns) and read from there by every MMA (tcgen05.mma with [a_tmem]);
//   * B operand = the activation unit tile in shared memory exactly as before (K-major, SWIZZLE_128B, the 7 filter rows as
//     row shifts of two TMA units), now the N = 128 side: 4 KB of shared-memory reads per MMA instead of 8.
//   * the accumulator holds one stacked weight ROW per TMEM lane and the 11 x 11 conv positions along the columns, so
//     MaxPool2d is register arithmetic inside a thread (no staging tile).  The rows are ordered so that the hi-stack and the
//     lo-stack row of a channel sit 16 lanes apart in the SAME warp (lane l of quadrant q: channel 16q + l % 16, stack l / 16):
//     the two partial sums meet through one __shfl_xor per value, and the pair splits the pooling work -- the low lane takes
//     pooled columns 0..2, the high lane the mirrored columns 4..2 -- with identical code (no divergence, no shared memory,
//     no block barrier in the epilogue).
// Shared-memory traffic per tile: 28 x 4 KB = 112 KB (was 224 + 90 = 314 KB).  TMEM: 2 x 128 accumulator columns + 224
// weight columns = 480 of 512.
#include "conv_common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace se3tn {
namespace {

constexpr int kThreadsS = 512;                 // warps: 0 A-TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..15 epilogue (three per TMEM lane quadrant)
constexpr int kAUnitS = 21 * 1024;             // (33 + 128) rows * 128 B
constexpr int kStagesS = 8;
constexpr int kSmemS = kStagesS * kAUnitS + 1024 + 512;
static_assert(kSmemS <= 232448, "shared memory budget");
constexpr uint32_t kDescHiS = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr int kWCol = 256;                     // first TMEM column of the weights
constexpr float kNegInf = -3.0e38f;

__device__ __forceinline__ uint64_t mk_desc_s(uint32_t lo) { return (static_cast<uint64_t>(kDescHiS) << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo_s(const void* p) { return ((ptx::smem_u32(p) & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ float selu_s(float x) {
    constexpr float kAlpha = 1.6732632423543772f, kScale = 1.0507009873554805f;
    return x > 0.f ? kScale * x : (kScale * kAlpha) * (__expf(x) - 1.f);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory, B from a shared-memory descriptor
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// The epilogue of one tile for one warp.  TMEM lane = (channel ch, stack): lanes l and l ^ 16 of a warp hold the hi-stack
// (a_hi*w_hi + a_lo*w_hi) and lo-stack (a_hi*w_lo) partial sums of the same channel.  The three warps of a lane quadrant
// split the conv rows:
//   PART 0: rows 0..4  (pooled rows 0, 1), accumulator columns [0, 64)
//   PART 1: rows 4..8  (pooled rows 2, 3), accumulator columns [32, 112)
//   PART 2: rows 8..10 (pooled row 4),     accumulator columns [80, 128)
// and within a pair the low lane (grp 0) works on conv columns 0..6 (pooled columns 0, 1, 2), the high lane (grp 1) on the
// mirrored columns 10..4 (pooled columns 4, 3, 2): register (row, j) means column j for grp 0 and 10 - j for grp 1, so both
// run the same instructions.  Per value: select what the partner needs, one shuffle, one add; then -inf for positions outside
// the 88 x 88 conv output (MaxPool2d's padding), 3x3 / stride 2 max in registers, + bias, SELU (monotone: max first), store.
template <int PART, int PREC>
__device__ __forceinline__ void pool_tile(const LayerDesc& L, const ResidentParams& p, uint32_t taddr, int ch, int grp, float bias,
                                          int n0, int ty, int tx, int lane, uint64_t* tmem_empty_bar)
{
    constexpr int kCol0 = PART == 0 ? 0 : (PART == 1 ? 32 : 80), kCols = PART == 0 ? 64 : (PART == 1 ? 80 : 48);
    constexpr int kRow0 = PART * 4, kRows = PART == 2 ? 3 : 5;
    constexpr int kPr0 = PART * 2, kPr = PART == 2 ? 1 : 2;
    float own[kCols];
    {
        uint32_t t16[kCols / 16][16];
#pragma unroll
        for (int c0 = 0; c0 < kCols; c0 += 16) ptx::tmem_ld16(taddr + kCol0 + c0, t16[c0 / 16]);      // all loads in flight, one wait
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c0 = 0; c0 < kCols; c0 += 16)
#pragma unroll
            for (int j = 0; j < 16; ++j) own[c0 + j] = __uint_as_float(t16[c0 / 16][j]);
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);                   // accumulator is in registers
    // conv coordinates of tile position (lr, lc): (10*ty - 1 + lr, 10*tx - 1 + lc)
    const int cy0 = ty * p.step_y + p.off_y, cx0 = tx * p.step_x + p.off_x;
    float sum[kRows][7];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const bool rv = static_cast<unsigned>(cy0 + kRow0 + r) < 88u;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float a = own[(kRow0 + r) * 11 + j - kCol0], b = own[(kRow0 + r) * 11 + (10 - j) - kCol0];
            const float mine = grp ? b : a;
            const float recv = __shfl_xor_sync(0xffffffffu, grp ? a : b, 16);      // the partner stack's sum at MY position
            const bool ok = rv && static_cast<unsigned>(cx0 + (grp ? 10 - j : j)) < 88u;
            sum[r][j] = ok ? mine + recv : kNegInf;
        }
    }
    const int c = L.out_coff + ch;
    uint8_t* const obase = L.out + (PREC == PREC_BF16X3 ? static_cast<size_t>(c & ~31) * 4 + (c & 31) * 2 : static_cast<size_t>(c) * 2);
    const size_t pix_bytes = static_cast<size_t>(L.out_c) * (PREC == PREC_BF16X3 ? 4 : 2);
#pragma unroll
    for (int pr = 0; pr < kPr; ++pr) {
        const int oy = ty * 5 + kPr0 + pr;
#pragma unroll
        for (int jx = 0; jx < 3; ++jx) {
            const int ox = tx * 5 + (grp ? 4 - jx : jx);
            if (oy >= L.Ho || ox >= L.Wo || (grp && jx == 2)) continue;         // pooled column 2 is the low lane's
            float m = kNegInf;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, sum[2 * (kPr0 + pr) + dy - kRow0][2 * jx + dx]);
            const float o = selu_s(m + bias);
            uint8_t* po = obase + ((static_cast<size_t>(n0) * L.Ho + oy) * L.Wo + ox) * pix_bytes;
            if (PREC == PREC_BF16X3) {
                const __nv_bfloat16 h = __float2bfloat16_rn(o);
                const __nv_bfloat16 l = __float2bfloat16_rn(o - __bfloat162float(h));
                *reinterpret_cast<__nv_bfloat16*>(po) = h;
                *reinterpret_cast<__nv_bfloat16*>(po + 64) = l;
            } else {
                *reinterpret_cast<__nv_bfloat16*>(po) = __float2bfloat16_rn(o);
            }
        }
    }
}

// PREC_BF16X3: output [32 hi | 32 lo] chunks; PREC_BF16: plain bf16
template <int PREC>
__global__ void __launch_bounds__(kThreadsS, 1)
conv_stem_ws_kernel(const __grid_constant__ ResidentParams p, const uint32_t* __restrict__ wstack /*[128][224 words]*/, int debug_flags)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const LayerDesc& L = p.L;
    uint8_t* sA = smem;                                                 // [kStagesS][unit]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kStagesS * kAUnitS);
    uint64_t* a_full = bars;                       // [kStagesS]
    uint64_t* a_empty = a_full + kStagesS;         // [kStagesS]
    uint64_t* tmem_full = a_empty + kStagesS;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* w_ready = tmem_empty + 2;            // [1]: the four weight-writer warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_ready + 1);

    ptx::grid_dep_launch();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_img = L.tiles_x * L.tiles_y;
    const int w_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.m_tiles / gridDim.x);
    const int w_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.m_tiles / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesS; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 12); }
        ptx::mbar_init(&w_ready[0], 4);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== A producer ================================
        if (lane == 0) {
            ptx::grid_dep_wait();                   // the packed input comes from the preprocess kernel
            int stage = 0; uint32_t phase = 0;
            for (int tile = w_begin; tile < w_end; ++tile) {
                const int n0 = p.img_first + tile / tiles_img, r = tile % tiles_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                const int ox = tx * p.step_x + p.off_x, oy = ty * p.step_y + p.off_y;
#pragma unroll
                for (int u = 0; u < 2; ++u) {       // even input rows (14 rows tall), odd input rows (13)
                    ptx::mbar_wait(&a_empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&a_full[stage], static_cast<uint32_t>(u == 0 ? 11 * 14 : 11 * 13) * kChunkBytes);
                    ptx::tma_load_4d(sA + stage * kAUnitS, &L.amap[u], &a_full[stage], 0, ox, oy, n0);
                    if (++stage == kStagesS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        constexpr uint32_t idesc = ptx::umma_idesc(1u /*bf16*/, kBlockM, 128);
        ptx::mbar_wait(&w_ready[0], 0);             // the weights are in tensor memory
        ptx::tc_fence_after();
        int astage = 0; uint32_t aphase = 0;
        int it = 0;
        for (int tile = w_begin; tile < w_end; ++tile, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128;
            uint32_t fresh = 0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                ptx::mbar_wait(&a_full[astage], aphase);
                ptx::tc_fence_after();
                const uint32_t b_unit_lo = desc_lo_s(sA + astage * kAUnitS);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (u == 1 && k == 3) break;    // odd rows: filter rows 1, 3, 5
                    const int tap = 2 * k + u;      // filter row
                    const uint32_t b_lo = b_unit_lo + k * 11 * (kChunkBytes >> 4);          // +11 pixel rows per vertical tap
                    const uint32_t a_col = tmem_base + kWCol + tap * 32;                     // 64 bf16 of K per filter row = 32 columns
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_f16_ts(d_tmem, a_col + kk * 8, mk_desc_s(b_lo + 2 * kk), idesc, fresh | (kk ? 1u : 0u));
                    }
                    __syncwarp();
                    fresh = 1u;
                }
                if (ptx::elect_one()) ptx::umma_commit(&a_empty[astage]);
                __syncwarp();
                if (++astage == kStagesS) { astage = 0; aphase ^= 1; }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ============================== epilogue (12 warps) ==========================
        const int ew = warp - 4;
        const int q = ew & 3;                       // TMEM lane quadrant: channels 16q .. 16q + 15, hi stack in lanes 0-15, lo stack in lanes 16-31
        const int part = ew >> 2;                   // which conv rows of the tile (pool_tile)
        const int grp = lane >> 4, ch = q * 16 + (lane & 15);
        if (ew < 4) {
            // ---- weights -> tensor memory: this thread's lane holds row (stack grp, channel ch) of [w_hi|w_hi ; w_lo|0]: 7 filter rows x 32 columns
            const uint32_t* wrow = wstack + static_cast<size_t>(grp * 64 + ch) * 224;
#pragma unroll 1
            for (int tap = 0; tap < 7; ++tap) {
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + tap * 32 + j));
                    r[j] = v.x; r[j + 1] = v.y; r[j + 2] = v.z; r[j + 3] = v.w;
                }
                tmem_st32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kWCol + tap * 32, r);
            }
            tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&w_ready[0]);
        }
        ptx::grid_dep_wait();                       // the output buffers are still read by the previous step's kernels
        const float bias = __ldg(L.bias + ch);
        int it = 0;
        for (int tile = w_begin; tile < w_end; ++tile, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n0 = p.img_first + tile / tiles_img, r = tile % tiles_img;
            const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128;
            if (debug_flags & 1) {                    // timing experiment: free the accumulator, no pooling (results are garbage)
                ptx::tc_fence_before(); __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
                continue;
            }
            if (part == 0)      pool_tile<0, PREC>(L, p, taddr, ch, grp, bias, n0, ty, tx, lane, &tmem_empty[acc]);
            else if (part == 1) pool_tile<1, PREC>(L, p, taddr, ch, grp, bias, n0, ty, tx, lane, &tmem_empty[acc]);
            else                pool_tile<2, PREC>(L, p, taddr, ch, grp, bias, n0, ty, tx, lane, &tmem_empty[acc]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

template <int PREC>
cudaError_t launch_stem_ws_t(const ResidentParams& p, const uint32_t* wstack, int debug_flags, int num_sms, bool pdl, cudaStream_t stream) {
    if (p.L.kind != KIND_STEM || p.img_wid || p.m_tiles <= 0 || !wstack) return cudaErrorInvalidValue;
    static size_t attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (static_cast<size_t>(kSmemS) > attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_stem_ws_kernel<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemS);
        if (e != cudaSuccess) return e;
        attr[dev] = kSmemS;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(std::min(p.m_tiles, num_sms)); cfg.blockDim = dim3(kThreadsS); cfg.dynamicSmemBytes = kSmemS; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, conv_stem_ws_kernel<PREC>, p, wstack, debug_flags);
}

}  // namespace

cudaError_t launch_conv_stem_ws(const ResidentParams& p, const void* wstack, int prec, int debug_flags, int num_sms, bool pdl, cudaStream_t stream) {
    if (prec == PREC_BF16X3) return launch_stem_ws_t<PREC_BF16X3>(p, static_cast<const uint32_t*>(wstack), debug_flags, num_sms, pdl, stream);
    if (prec == PREC_BF16) return launch_stem_ws_t<PREC_BF16>(p, static_cast<const uint32_t*>(wstack), debug_flags, num_sms, pdl, stream);
    return cudaErrorInvalidValue;
}

}  // namespace se3tn

// tcgen05 implicit-GEMM convolutions of the se(3)-TrackNet conv stack (reference se3_tracknet.py:57-78,
// network_modules.py:59-66,86-120), im2col-free: D[pixels, Cout] = sum_tap sum_c A_tap[pixel, c] * W[Cout, tap*Cin + c].
//
// Common to both kernels
//   * M tile = an 11x11 pixel box of one image (121 of 128 UMMA rows; 44, 22 and 11 are multiples of 11).
//   * A operand by TMA "units": one cp.async.bulk.tensor.4d per (128-byte channel chunk, filter COLUMN) whose box is
//     two rows taller than the tile.  It lands as 143 rows x 128 B in the SWIZZLE_128B K-major layout UMMA wants and
//     the three vertical taps of that column are the SAME tile read through descriptors whose start address is advanced
//     by whole pixel rows (11 * 128 B; base_offset stays 0, the swizzle is a function of absolute address bits) -- no
//     data moves.  Out-of-image coordinates are zero-filled by TMA = the conv padding.  Stride-2 convs: six units per
//     chunk over four parity views of the input.  Stem (7x7 s2, Cin = 4): two units (even / odd input rows) of an
//     overlapping 8-pixel-window view; the 7 filter rows are row shifts 0/11/22/33.
//   * warp roles (384 threads): 0 activation TMA producer (+ work scheduler in the trunk kernel), 1 MMA issuer (the whole
//     warp walks the loop warp-uniformly so descriptors stay in uniform registers; one elected lane issues), 2 TMEM
//     allocator, 3 weight TMA producer, 4..11 epilogue (two warps per TMEM lane quadrant).  Two accumulator sets in
//     TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * PREC selects arithmetic and storage (conv_common.h): TF32 (4 x kind::tf32 per chunk-tap), BF16X3 (x = hi + lo as
//     two bf16, 3 products per MAC: fp32-faithful), BF16 (2-byte activations, 64 channels per chunk, 1 product).
//   * programmatic dependent launch: every CTA signals launch_dependents at entry; only the warps that touch
//     activations execute griddepcontrol.wait, so barrier init / TMEM alloc / weight TMA run under the previous
//     kernel's tail.
//   * per-object weights (reference README.md:132: one checkpoint per object class): with img_wid every work unit takes
//     its weight tensor map and bias from per-set device tables, so all tracks of a frame share the launches.
//
// conv_resident_kernel<KIND, PREC>  (Cout = 64: stems, 64-channel 3x3 convs)
//   * the whole K-major weight matrix (<= 147 KB) is TMA-loaded into shared memory once per CTA (again only when the
//     weight-set id changes between consecutive tiles of the CTA's contiguous range), one mbarrier per filter-column
//     unit so the first MMAs start when the first third has landed.
//   * N = 64 MMAs are bound by shared-memory operand bandwidth (A 4 KB + B 2 KB per 32-cycle MMA slot > 128 B/cycle),
//     so in the bf16 hi/lo modes the weight halves are STACKED along N: rows [w_hi ; w_lo] -> one N = 128 MMA forms
//     a_hi*w_hi and a_hi*w_lo, one N = 64 MMA adds a_lo*w_hi, the epilogue sums the two column halves (4 instead of 6
//     MMAs per chunk-tap; the stem's [w_hi|w_hi ; w_lo|0] rows give all three products in one N = 128 MMA per K step).
//   * 64-channel layers read the accumulator with tcgen05.ld.16x256b and keep their weight ROWS permuted so that each
//     lane owns 8 consecutive channels of four pixels: 16-byte pieces, 8 lines per store instruction, no staging.
//   * stem: fused MaxPool2d(3,2,1): the M tile is the 11x11 block of conv outputs that feeds a 5x5 block of pooled
//     outputs; max -> +bias -> SELU (monotone, so they commute) on 1/4.84 of the values; the 88x88x64 conv output never
//     reaches HBM.
//
// conv_trunk_kernel<PREC>  (Cout >= 256: convAB1, convAB2.{conv1,conv2}, {trans,rot}_conv1, {trans,rot}_conv2.{conv1,conv2})
//   * ONE launch for all six layers.  A work unit = (layer, image, 11x11 tile, 256 output channels); at batch 64 every
//     layer has 256 units for 148 SMs, which as separate launches left 27 % of the SMs idle for the second half of each
//     layer.  Here a persistent CTA per SM pulls the next unit from a global counter (layer-major, image-major order)
//     when its producer has issued the last loads of the current one, and a unit of layer l first waits until
//     done[l-1][image] says that image's previous-layer output is complete (release/acquire at gpu scope; the waits
//     point backwards in the pull order and all CTAs are co-resident, so the schedule cannot deadlock).
//   * weights stream through a 4-stage ring of {32 words, 256 rows} tiles fed by their own producer warp.
//   * epilogue: each 32-row x 32-column accumulator block is transposed through a per-warp shared-memory tile so every
//     global load / store instruction covers whole lines; the last layer reduces its 121 rows to per-quadrant column
//     sums instead (AdaptiveAvgPool2d(1) fused; fixed order -> deterministic).
#include "conv_common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace se3tn {
namespace {

constexpr int kThreads2 = 384;                 // warps: 0 A-TMA, 1 MMA, 2 TMEM alloc, 3 B-TMA, 4..11 epilogue
constexpr int kPoolPitch = 68;                 // floats per staged conv position (64 + 4: bank spread)
constexpr int kPoolStageBytes = 121 * kPoolPitch * 4;
constexpr int kPoolStageAlloc = (kPoolStageBytes + 1023) & ~1023;
constexpr int kAUnit3 = 19 * 1024;             // 3x3: (22 + 128) rows * 128 B = 19,200
constexpr int kAUnitStem = 21 * 1024;          // stem: (33 + 128) rows * 128 B = 20,608
// descriptor high word: SBO = 1024 B (>>4) | version 1 (bit 46) | SWIZZLE_128B (bits 61..63)
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);

// Compile-time unit / tap structure per conv kind, so the MMA issue loop is straight-line code with immediate row
// shifts / weight-tile indices, and the TMA producer's box coordinates are immediates too.
template <int KIND> struct KTab;
template <> struct KTab<KIND_S1> {            // 3x3 stride 1: unit = filter column s, taps = filter rows r
    static constexpr int NU = 3;
    __host__ __device__ static constexpr int ntaps(int) { return 3; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return k * 3 + u; }
    __host__ __device__ static constexpr int amap(int) { return 0; }
    __host__ __device__ static constexpr int c1(int u) { return u - 1; }
    __host__ __device__ static constexpr int c2(int) { return -1; }
    __host__ __device__ static constexpr int rows(int) { return 11 * 13; }
};
template <> struct KTab<KIND_S2> {            // 3x3 stride 2: per column s an even-row unit (r=1) and an odd-row unit (r=0,2)
    static constexpr int NU = 6;
    __host__ __device__ static constexpr int ntaps(int u) { return (u & 1) ? 2 : 1; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return (u & 1) ? (k == 0 ? (u >> 1) : 6 + (u >> 1)) : 3 + (u >> 1); }
    // iy = 2*oy + dy: dy = -1 -> odd row oy-1; dy = 0 -> even row oy; dy = +1 -> odd row oy; columns likewise
    __host__ __device__ static constexpr int amap(int u) { return (u & 1) * 2 + ((u >> 1) == 1 ? 0 : 1); }
    __host__ __device__ static constexpr int c1(int u) { return (u >> 1) == 0 ? -1 : 0; }
    __host__ __device__ static constexpr int c2(int u) { return (u & 1) ? -1 : 0; }
    __host__ __device__ static constexpr int rows(int u) { return (u & 1) ? 11 * 12 : 11 * 11; }
};
template <> struct KTab<KIND_STEM> {          // 7x7 stride 2 stem: even input rows (r=0,2,4,6), odd input rows (r=1,3,5)
    static constexpr int NU = 2;
    __host__ __device__ static constexpr int ntaps(int u) { return u == 0 ? 4 : 3; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return 2 * k + u; }
    __host__ __device__ static constexpr int amap(int u) { return u; }
    __host__ __device__ static constexpr int c1(int) { return 0; }
    __host__ __device__ static constexpr int c2(int) { return 0; }
    __host__ __device__ static constexpr int rows(int u) { return u == 0 ? 11 * 14 : 11 * 13; }
};
constexpr int kRowShift = 11;                  // rows a descriptor advances per vertical tap (tile width)

__device__ __forceinline__ float selu_fast(float x) {
    constexpr float kAlpha = 1.6732632423543772f, kScale = 1.0507009873554805f;
    return x > 0.f ? kScale * x : (kScale * kAlpha) * (__expf(x) - 1.f);
}
__device__ __forceinline__ float act_apply(float x, int act) {
    return act == ACT_RELU ? fmaxf(x, 0.f) : (act == ACT_SELU ? selu_fast(x) : x);
}

// timeline stamps (debug): slot 0 kernel entry, 1 setup done, 2 MMA warp has its first weights, 3 MMA warp has its first A unit,
// 4 MMA warp issued its last commit, 5 epilogue got its first accumulator, 6 epilogue finished its last tile, 7 CTA exit (low 8 bits: SM id)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_stamp(unsigned long long* tr, int slot) { if (tr) tr[blockIdx.x * 8 + slot] = gtimer(); }
__device__ __forceinline__ void trace_exit(unsigned long long* tr) {
    if (!tr) return;
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[blockIdx.x * 8 + 7] = (gtimer() & ~0xffull) | (smid & 0xff);
}

// fp32 -> (bf16 hi, bf16 lo) with x ~= hi + lo; packs two values per 32-bit word (element 0 in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(const void* smem_ptr) { return ((ptx::smem_u32(smem_ptr) & 0x3FFFFu) >> 4) | (1u << 16); }

// Byte address of 4 consecutive channels (c % 4 == 0) of pixel `pix` in an NHWC buffer of C channels per pixel:
//   TF32   : fp32 words                       -> 16 bytes at (pix*C + c) * 4
//   BF16X3 : chunk [32 x hi | 32 x lo]        -> 8 bytes (hi) at (pix*C + (c & ~31)) * 4 + (c & 31) * 2, lo 64 bytes further
//   BF16   : 2 bytes per channel              -> 8 bytes at (pix*C + c) * 2
template <int PREC>
__device__ __forceinline__ size_t chan_byte(size_t pix, int C, int c) {
    if (PREC == PREC_TF32) return (pix * C + c) * 4;
    if (PREC == PREC_BF16X3) return (pix * C + (c & ~31)) * 4 + (c & 31) * 2;
    return (pix * C + c) * 2;
}

// ================================================================================================================
// conv_resident_kernel
// ================================================================================================================
template <int KIND, int PREC> struct RCfg {
    static constexpr bool POOL = (KIND == KIND_STEM);
    static constexpr int BN = 64;
    // STACK: hi / lo weight rows stacked along N (header comment).  The stem input is [4 hi | 4 lo] per pixel in both bf16 modes.
    static constexpr int kStack = (PREC == PREC_BF16X3 || (PREC == PREC_BF16 && POOL)) ? 2 : 1;
    static constexpr int kBTile = BN * kStack * kChunkBytes;
    static constexpr int kAUnit = POOL ? kAUnitStem : kAUnit3;
    static constexpr int kAStages = POOL ? (PREC == PREC_TF32 ? 4 : 3) : (PREC == PREC_BF16 ? 6 : 4);
    static constexpr int kPoolBufs = POOL ? (PREC == PREC_TF32 ? 2 : 1) : 0;
    static constexpr int kAccCols = BN * kStack;
    static constexpr int kTmemCols = 2 * kAccCols;                  // two accumulator sets
    static constexpr int kMaxWTiles = POOL ? 7 : ((PREC == PREC_TF32) ? 18 : 9);
    static constexpr int kSmem = kAStages * kAUnit + kMaxWTiles * kBTile + kPoolBufs * kPoolStageAlloc + 1024 + 512;
    static_assert(kSmem <= 232448, "shared memory budget");
};

template <int KIND, int PREC>
__global__ void __launch_bounds__(kThreads2, 1)
conv_resident_kernel(const __grid_constant__ ResidentParams p)
{
    using C = RCfg<KIND, PREC>;
    using KT = KTab<KIND>;
    constexpr bool POOL = C::POOL;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const LayerDesc& L = p.L;
    const int chunks = L.chunks;
    // K tiles of the resident weight matrix: STACK keeps all chunks of a tap in one 128-row tile, otherwise one tile per (tap, chunk)
    const int tiles_per_tap = (C::kStack == 2) ? 1 : chunks;
    uint8_t* sA = smem;                                                 // [kAStages][unit]
    uint8_t* sB = sA + C::kAStages * C::kAUnit;                         // [num_taps * tiles_per_tap][BN*kStack rows x 128 B]
    uint8_t* sP = sB + C::kMaxWTiles * C::kBTile;                       // pool staging (stem only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::kPoolBufs * kPoolStageAlloc);
    uint64_t* a_full = bars;                       // [kAStages]
    uint64_t* a_empty = a_full + C::kAStages;      // [kAStages]
    uint64_t* b_full = a_empty + C::kAStages;      // [NU]: weights of filter-column unit u have landed
    uint64_t* b_empty = b_full + KT::NU;           // [1]: MMAs that read the current weights have retired (multi-set reload)
    uint64_t* tmem_full = b_empty + 1;             // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    ptx::grid_dep_launch();
    if (threadIdx.x == 0) trace_stamp(p.trace, 0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_img = L.tiles_x * L.tiles_y;
    // each CTA owns a contiguous range of tiles (consecutive tiles of the same image / weight set)
    const int w_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.m_tiles / gridDim.x);
    const int w_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.m_tiles / gridDim.x);
    auto img_of = [&](int tile) -> int { return p.img_first + tile / tiles_img; };
    auto wid_of = [&](int tile) -> int { return p.img_wid ? p.img_wid[img_of(tile)] : -1; };   // -1: single-set launch

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kAStages; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int u = 0; u < KT::NU; ++u) ptx::mbar_init(&b_full[u], 1);
        ptx::mbar_init(&b_empty[0], 1);
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 8); }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) { ptx::tmem_alloc(tmem_slot, C::kTmemCols); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) trace_stamp(p.trace, 1);

    if (warp == 0) {
        // ============================== A producer ================================
        if (lane == 0) {
            ptx::grid_dep_wait();                   // activations come from the previous kernel
            int stage = 0; uint32_t phase = 0;
            for (int tile = w_begin; tile < w_end; ++tile) {
                const int n0 = img_of(tile), r = tile % tiles_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                const int ox = tx * p.step_x + p.off_x, oy = ty * p.step_y + p.off_y;
                for (int ch = 0; ch < chunks; ++ch) {
#pragma unroll
                    for (int u = 0; u < KT::NU; ++u) {
                        ptx::mbar_wait(&a_empty[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&a_full[stage], static_cast<uint32_t>(KT::rows(u)) * kChunkBytes);
                        ptx::tma_load_4d(sA + stage * C::kAUnit, &L.amap[KT::amap(u)], &a_full[stage],
                                         L.in_cbase_words + ch * 32, ox + KT::c1(u), oy + KT::c2(u), n0);
                        if (++stage == C::kAStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ============================== weight loader ================================
        if (lane == 0) {
            int cur = -2; uint32_t gen = 0;
            for (int tile = w_begin; tile < w_end; ++tile) {
                const int wid = wid_of(tile);
                if (wid == cur) continue;
                if (gen) ptx::mbar_wait(&b_empty[0], (gen - 1) & 1);       // MMAs that read the previous weights have retired
                const CUtensorMap* bm = wid < 0 ? &L.bmap : p.gbmaps + wid * kLayersPerSet + L.li;
#pragma unroll
                for (int u = 0; u < KT::NU; ++u) {                          // in the order the MMA warp needs them: unit by unit
                    ptx::mbar_arrive_expect_tx(&b_full[u], static_cast<uint32_t>(KT::ntaps(u) * tiles_per_tap) * C::kBTile);
#pragma unroll
                    for (int k = 0; k < KT::ntaps(u); ++k)
                        for (int c = 0; c < tiles_per_tap; ++c) {
                            const int wt = KT::wtap(u, k) * tiles_per_tap + c;   // tile wt covers K words [wt*32, wt*32 + 32)
                            ptx::tma_load_2d(sB + wt * C::kBTile, bm, &b_full[u], wt * 32, 0);
                        }
                }
                cur = wid; ++gen;
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        constexpr uint32_t idesc = ptx::umma_idesc(PREC == PREC_TF32 ? 2u /*tf32*/ : 1u /*bf16*/, kBlockM, BN);
        constexpr uint32_t idesc_stack = ptx::umma_idesc(1u, kBlockM, BN * C::kStack);      // N = 128: [w_hi ; w_lo] rows
        int astage = 0; uint32_t aphase = 0;
        int w_cur = -2; uint32_t w_gen = 0;        // weight set currently in shared memory
        int it = 0;
        for (int tile = w_begin; tile < w_end; ++tile, ++it) {
            const int wid = wid_of(tile);
            const bool new_w = (wid != w_cur);     // wait for each unit's weights at its first use below
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
            uint32_t fresh = 0;                     // 0 until the first MMA of this tile has been issued
            for (int ch = 0; ch < chunks; ++ch) {
#pragma unroll
                for (int u = 0; u < KT::NU; ++u) {
                    if (new_w && ch == 0) {
                        ptx::mbar_wait(&b_full[u], w_gen & 1); ptx::tc_fence_after();
                        if (it == 0 && u == 0 && lane == 0) trace_stamp(p.trace, 2);
                    }
                    ptx::mbar_wait(&a_full[astage], aphase);
                    ptx::tc_fence_after();
                    if (it == 0 && ch == 0 && u == 0 && lane == 0) trace_stamp(p.trace, 3);
                    const uint32_t a_unit_lo = desc_lo(sA + astage * C::kAUnit);
#pragma unroll
                    for (int k = 0; k < KT::ntaps(u); ++k) {
                        // low descriptor words: +2 per 32-byte K step, +8 per pixel row
                        const uint32_t a_lo = a_unit_lo + k * kRowShift * (kChunkBytes >> 4);
                        uint32_t b_lo;
                        if (C::kStack == 2) b_lo = desc_lo(sB + KT::wtap(u, k) * C::kBTile) + ch * 4;   // chunk ch sits 64 bytes (4 x 16 B) further along K
                        else                b_lo = desc_lo(sB + (KT::wtap(u, k) * tiles_per_tap + ch) * C::kBTile);
                        if (ptx::elect_one()) {
                            if (PREC == PREC_TF32) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    ptx::umma_tf32(d_tmem, mk_desc(a_lo + 2 * kk), mk_desc(b_lo + 2 * kk), idesc, fresh | (kk ? 1u : 0u));
                            } else if (POOL) {
                                // stem window = 8 pixels x [hi4|lo4] against rows [w_hi|w_hi ; w_lo|0]: all three products in one N = 128 MMA per K step
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    ptx::umma_f16(d_tmem, mk_desc(a_lo + 2 * kk), mk_desc(b_lo + 2 * kk), idesc_stack, fresh | (kk ? 1u : 0u));
                            } else if (C::kStack == 2) {
                                // chunk = [32 hi | 32 lo] (A); weight rows [w_hi ; w_lo]: a_hi x both (N = 128), then a_lo x w_hi (N = 64, columns 0-63)
#pragma unroll
                                for (int sl = 0; sl < 2; ++sl) {
                                    ptx::umma_f16(d_tmem, mk_desc(a_lo + 2 * sl), mk_desc(b_lo + 2 * sl), idesc_stack, fresh | (sl ? 1u : 0u));
                                    ptx::umma_f16(d_tmem, mk_desc(a_lo + 4 + 2 * sl), mk_desc(b_lo + 2 * sl), idesc, 1u);
                                }
                            } else {
                                // PREC_BF16: chunk = 64 bf16 channels: four K = 16 steps
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    ptx::umma_f16(d_tmem, mk_desc(a_lo + 2 * kk), mk_desc(b_lo + 2 * kk), idesc, fresh | (kk ? 1u : 0u));
                            }
                        }
                        __syncwarp();
                        fresh = 1u;
                    }
                    if (ptx::elect_one()) ptx::umma_commit(&a_empty[astage]);
                    __syncwarp();
                    if (++astage == C::kAStages) { astage = 0; aphase ^= 1; }
                }
            }
            if (new_w) { w_cur = wid; ++w_gen; }
            if (ptx::elect_one()) {
                ptx::umma_commit(&tmem_full[acc]);
                // the next tile uses other weights -> tell the loader when these MMAs have retired
                if (tile + 1 < w_end && wid_of(tile + 1) != w_cur) ptx::umma_commit(&b_empty[0]);
            }
            __syncwarp();
        }
        if (lane == 0) trace_stamp(p.trace, 4);
    } else if (warp >= 4) {
        // ============================== epilogue (8 warps) ==========================
        ptx::grid_dep_wait();                       // residual reads / output writes must follow the previous kernel
        const int ew = warp - 4;
        const int q = ew & 3;                       // TMEM lane quadrant (== warp % 4)
        const int half = ew >> 2;                   // which half of the 64 columns
        constexpr int kCols = BN / 2;
        const int row = q * 32 + lane;
        int it = 0;
        if (!POOL) {
            // ---------------- 64-channel layers: 16x256b TMEM loads ----------------
            // Lane (R = lane/4, m = lane%4) receives, for each of its four rows q*32 + 8i + R, the accumulator columns
            // 8j + 2m + e (j < 4, e < 2) of this warp's 32-column block.  The weight rows of these layers are stored
            // permuted (column 8j + 2m + e carries output channel 8m + 2j + e, se3tn.cu), so the lane owns the 8
            // CONSECUTIVE channels 8m .. 8m+7 of each pixel: four lanes complete 64 (128) contiguous bytes and a
            // load / store instruction touches 8 lines instead of the 32 of a row-per-lane epilogue.
            const int mm = lane & 3;
            const int ch0 = half * kCols + mm * 8;                          // this lane's first channel
            for (int tile = w_begin; tile < w_end; ++tile, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                const int n0 = img_of(tile), r = tile % tiles_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                // this lane's four pixels and their residual pieces are known before the accumulator is: issue the residual
                // loads now so their latency hides behind the wait for the MMAs
                size_t rpix[4]; bool rvalid[4]; uint4 rres[4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                           // row q*32 + 8i + lane/4
                    const int rw = q * 32 + 8 * i + (lane >> 2);
                    const int py = rw / 11, px = rw - py * 11;
                    const int y = ty * 11 + py, x = tx * 11 + px;
                    rvalid[i] = (rw < 121) && (y < L.Ho) && (x < L.Wo);
                    rpix[i] = (static_cast<size_t>(n0) * L.Ho + y) * L.Wo + x;
                    rres[i][0] = make_uint4(0u, 0u, 0u, 0u); rres[i][1] = make_uint4(0u, 0u, 0u, 0u);
                    if (L.res && rvalid[i]) {
                        const uint8_t* rp = L.res + chan_byte<PREC>(rpix[i], L.res_c, ch0);
                        if (PREC == PREC_TF32)        { rres[i][0] = __ldg(reinterpret_cast<const uint4*>(rp)); rres[i][1] = __ldg(reinterpret_cast<const uint4*>(rp + 16)); }   // 8 fp32 words
                        else if (PREC == PREC_BF16X3) { rres[i][0] = __ldg(reinterpret_cast<const uint4*>(rp)); rres[i][1] = __ldg(reinterpret_cast<const uint4*>(rp + 64)); }   // hi piece, lo piece
                        else                          { rres[i][0] = __ldg(reinterpret_cast<const uint4*>(rp)); }                                                             // 8 bf16
                    }
                }
                const float* bias_base = (p.img_wid ? p.gbias[p.img_wid[n0] * kLayersPerSet + L.li] : L.bias) + ch0;
                const float4 bA = __ldg(reinterpret_cast<const float4*>(bias_base)), bB = __ldg(reinterpret_cast<const float4*>(bias_base + 4));
                const float bias8[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, 5);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols + half * kCols;
                uint32_t ra[16], rb[16];
                ptx::tmem_ld_16x256b_x4(taddr, ra);
                ptx::tmem_ld_16x256b_x4(taddr + (16u << 16), rb);
                if (C::kStack == 2) {
                    uint32_t rc[16], rd[16];
                    ptx::tmem_ld_16x256b_x4(taddr + BN, rc);
                    ptx::tmem_ld_16x256b_x4(taddr + (16u << 16) + BN, rd);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        ra[i] = __float_as_uint(__uint_as_float(ra[i]) + __uint_as_float(rc[i]));
                        rb[i] = __float_as_uint(__uint_as_float(rb[i]) + __uint_as_float(rd[i]));
                    }
                } else {
                    ptx::tmem_ld_wait();
                }
                // accumulator is in registers: hand it back to the MMA warp before the stores
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {                           // row q*32 + 8i + R; register 4j + 2(i&1) + e of ra (i < 2) / rb
                    if (!rvalid[i]) continue;
                    float v[8];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            v[2 * jj + e] = __uint_as_float((i < 2 ? ra : rb)[4 * jj + 2 * (i & 1) + e]) + bias8[2 * jj + e];
                    if (L.res) {
                        if (PREC == PREC_TF32) {
                            const uint4 r0 = rres[i][0], r1 = rres[i][1];
                            v[0] += __uint_as_float(r0.x); v[1] += __uint_as_float(r0.y); v[2] += __uint_as_float(r0.z); v[3] += __uint_as_float(r0.w);
                            v[4] += __uint_as_float(r1.x); v[5] += __uint_as_float(r1.y); v[6] += __uint_as_float(r1.z); v[7] += __uint_as_float(r1.w);
                        } else {
                            const uint4 h4 = rres[i][0], l4 = rres[i][1];       // l4 = 0 in PREC_BF16
                            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 hf = unpack2(hw[e]), lf = unpack2(lw[e]);
                                v[2 * e] += hf.x + lf.x; v[2 * e + 1] += hf.y + lf.y;
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = act_apply(v[e], L.act);
                    uint8_t* outp = L.out + chan_byte<PREC>(rpix[i], L.out_c, L.out_coff + ch0);
                    if (PREC == PREC_TF32) {
                        *reinterpret_cast<float4*>(outp) = make_float4(ptx::to_tf32(v[0]), ptx::to_tf32(v[1]), ptx::to_tf32(v[2]), ptx::to_tf32(v[3]));
                        *reinterpret_cast<float4*>(outp + 16) = make_float4(ptx::to_tf32(v[4]), ptx::to_tf32(v[5]), ptx::to_tf32(v[6]), ptx::to_tf32(v[7]));
                    } else if (PREC == PREC_BF16X3) {
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
                        *reinterpret_cast<uint4*>(outp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(outp + 64) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    } else {
                        *reinterpret_cast<uint4*>(outp) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                }
            }
        } else {
            // ---- stem: conv tile 11x11 -> 5x5 max-pooled outputs (MaxPool2d(3,2,1), -inf padding) ----
            const int cy_l = row / 11, cx_l = row - cy_l * 11;          // conv position inside the tile
            const int et = threadIdx.x - 128;                           // 0..255 among epilogue threads
            for (int tile = w_begin; tile < w_end; ++tile, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                const int n0 = img_of(tile), r = tile % tiles_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                float* stage = reinterpret_cast<float*>(sP + (C::kPoolBufs == 2 ? (it & 1) : 0) * kPoolStageAlloc);
                const int cy = ty * p.step_y + p.off_y + cy_l, cx = tx * p.step_x + p.off_x + cx_l;   // conv output coordinates
                const bool cvalid = (row < 121) && cy >= 0 && cy < 88 && cx >= 0 && cx < 88;

                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, 5);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols + half * kCols;
                {
                    float v[32];
                    uint32_t r0[16], r1[16];
                    ptx::tmem_ld16(taddr, r0);
                    ptx::tmem_ld16(taddr + 16, r1);
                    if (C::kStack == 2) {
                        uint32_t r2[16], r3[16];
                        ptx::tmem_ld16(taddr + BN, r2);
                        ptx::tmem_ld16(taddr + BN + 16, r3);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) { v[jj] = __uint_as_float(r0[jj]) + __uint_as_float(r2[jj]); v[16 + jj] = __uint_as_float(r1[jj]) + __uint_as_float(r3[jj]); }
                    } else {
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) { v[jj] = __uint_as_float(r0[jj]); v[16 + jj] = __uint_as_float(r1[jj]); }
                    }
                    if (row < 121) {
                        float* srow = stage + row * kPoolPitch + half * kCols;
                        const float ninf = -3.0e38f;
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 4)
                            *reinterpret_cast<float4*>(srow + jj) = cvalid ? make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3])
                                                                           : make_float4(ninf, ninf, ninf, ninf);
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);           // accumulator drained
                asm volatile("bar.sync 1, 256;" ::: "memory");             // staging tile complete (epilogue warps only)

                const float* bias = p.img_wid ? p.gbias[p.img_wid[n0] * kLayersPerSet + L.li] : L.bias;
                // 25 pooled pixels x 16 float4 channel groups = 400 vectors over 256 threads
                for (int v = et; v < 400; v += 256) {
                    const int pp = v >> 4, c4 = (v & 15) * 4;
                    const int ppy = pp / 5, ppx = pp - ppy * 5;
                    const int oy = ty * 5 + ppy, ox = tx * 5 + ppx;     // pooled output coordinates
                    if (oy >= L.Ho || ox >= L.Wo) continue;
                    float4 m = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float4 s4 = *reinterpret_cast<const float4*>(stage + ((2 * ppy + dy) * 11 + 2 * ppx + dx) * kPoolPitch + c4);
                            m.x = fmaxf(m.x, s4.x); m.y = fmaxf(m.y, s4.y); m.z = fmaxf(m.z, s4.z); m.w = fmaxf(m.w, s4.w);
                        }
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c4));
                    const float v0 = selu_fast(m.x + b4.x), v1 = selu_fast(m.y + b4.y), v2 = selu_fast(m.z + b4.z), v3 = selu_fast(m.w + b4.w);
                    uint8_t* po = L.out + chan_byte<PREC>((static_cast<size_t>(n0) * L.Ho + oy) * L.Wo + ox, L.out_c, L.out_coff + c4);
                    if (PREC == PREC_TF32) {
                        *reinterpret_cast<float4*>(po) = make_float4(ptx::to_tf32(v0), ptx::to_tf32(v1), ptx::to_tf32(v2), ptx::to_tf32(v3));
                    } else if (PREC == PREC_BF16X3) {
                        uint32_t h0, l0, h1, l1;
                        split2(v0, v1, h0, l0); split2(v2, v3, h1, l1);
                        *reinterpret_cast<uint2*>(po) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(po + 64) = make_uint2(l0, l1);
                    } else {
                        *reinterpret_cast<uint2*>(po) = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
                    }
                }
                if (C::kPoolBufs == 1) asm volatile("bar.sync 1, 256;" ::: "memory");   // single staging buffer: readers done before the next tile writes
            }
        }
        if (threadIdx.x == 128) trace_stamp(p.trace, 6);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) trace_exit(p.trace);
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ================================================================================================================
// conv_trunk_kernel
// ================================================================================================================
template <int PREC, int BN_> struct TCfg {
    static constexpr int BN = BN_;                                      // 256; 128 for small batches (twice the work units per layer)
    static constexpr int kAStages = 3;
    static constexpr int kBStages = BN == 256 ? 4 : 6;
    static constexpr int kBTile = BN * kChunkBytes;                     // 32 KB / 16 KB
    static constexpr int kEpiPitch = 36;                                // words per staged row (32 + 4: conflict-free 16 B accesses)
    static constexpr int kEpiWarpBytes = 32 * kEpiPitch * 4 + 128;      // 32 rows + 32-entry pixel-index table
    static constexpr int kEpiBytes = 8 * kEpiWarpBytes;
    static constexpr int kSched = 4;                                    // work-unit ring between the scheduler (A producer) and the other roles
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kSmem = kAStages * kAUnit3 + kBStages * kBTile + ((kEpiBytes + 1023) & ~1023) + 1024 + 512;
    static_assert(kSmem <= 232448, "shared memory budget");
};

struct UnitCoord { int l, img, tx, ty, n_tile, grp, c0, c1, piece, gidx; };
// per-unit timeline (SE3TN_TRACE, small launches only): 5 stamps per work unit behind the trunk's per-CTA stamps:
// 0 dependency satisfied (producer), 1 first A unit landed (MMA warp), 2 last MMA committed, 3 accumulator seen by epilogue warp 0,
// 4 epilogue warp 0 finished (stores + completion signal)
__device__ __forceinline__ void unit_stamp(const TrunkParams& p, int u, int k) {
    if (p.trace && u < 2048) p.trace[256 * 8 + u * 5 + k] = gtimer();
}   // [c0, c1): K chunks of this piece; gidx: unsplit unit index in the launch

__device__ __forceinline__ int unit_layer(const TrunkParams& p, int u) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kTrunkMaxLayers; ++i) if (i < p.n_layers && u >= p.layer[i].unit_base) l = i;
    return l;
}

__device__ __forceinline__ UnitCoord decode_unit(const TrunkParams& p, int u) {
    UnitCoord c;
    c.l = unit_layer(p, u);
    const LayerDesc& L = p.layer[c.l];
    int local = u - L.unit_base;
    c.piece = 0; c.c0 = 0; c.c1 = L.chunks;
    if (p.ksplit > 1) {                            // consecutive indices = the pieces of one unit (pulled by different CTAs at about the same time)
        c.piece = local % p.ksplit; local /= p.ksplit;
        c.c0 = c.piece * L.chunks / p.ksplit; c.c1 = (c.piece + 1) * L.chunks / p.ksplit;
    }
    c.gidx = L.base_unit0 + local;
    const int im = local / L.units_per_image;
    int r = local - im * L.units_per_image;
    c.img = p.img_first + im;
    c.n_tile = r % L.n_tiles; r /= L.n_tiles;
    c.grp = r % L.groups; r /= L.groups;
    c.ty = r / L.tiles_x; c.tx = r - c.ty * L.tiles_x;
    return c;
}

// activation loads of one work unit (A producer thread)
template <int KIND>
__device__ __forceinline__ void trunk_load_unit(const LayerDesc& L, const UnitCoord& c, uint8_t* sA, uint64_t* a_full, uint64_t* a_empty,
                                                int& stage, uint32_t& phase, int n_stages)
{
    using KT = KTab<KIND>;
    const int cbase = L.in_cbase_words + c.grp * L.in_gstride_words;
    const int ox = c.tx * 11, oy = c.ty * 11;
    for (int ch = c.c0; ch < c.c1; ++ch) {
#pragma unroll
        for (int u = 0; u < KT::NU; ++u) {
            ptx::mbar_wait(&a_empty[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&a_full[stage], static_cast<uint32_t>(KT::rows(u)) * kChunkBytes);
            ptx::tma_load_4d(sA + stage * kAUnit3, &L.amap[KT::amap(u)], &a_full[stage], cbase + ch * 32, ox + KT::c1(u), oy + KT::c2(u), c.img);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    }
}

// weight-tile loads of one work unit (B producer thread)
template <int KIND, int PREC, int BN>
__device__ __forceinline__ void trunk_load_weights(const LayerDesc& L, const CUtensorMap* bm, const UnitCoord& c, uint8_t* sB,
                                                   uint64_t* b_full, uint64_t* b_empty, int& stage, uint32_t& phase)
{
    using KT = KTab<KIND>;
    using C = TCfg<PREC, BN>;
    const int wrow = c.grp * L.cout + c.n_tile * C::BN;
    for (int ch = c.c0; ch < c.c1; ++ch) {
#pragma unroll
        for (int u = 0; u < KT::NU; ++u) {
#pragma unroll
            for (int k = 0; k < KT::ntaps(u); ++k) {
                ptx::mbar_wait(&b_empty[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx(&b_full[stage], C::kBTile);
                ptx::tma_load_2d(sB + stage * C::kBTile, bm, &b_full[stage], KT::wtap(u, k) * L.cin_words + ch * 32, wrow);
                if (++stage == C::kBStages) { stage = 0; phase ^= 1; }
            }
        }
    }
}

// MMAs of one work unit (whole MMA warp, warp-uniform)
template <int KIND, int PREC, int BN>
__device__ __forceinline__ void trunk_mma_unit(int chunks, uint32_t d_tmem, uint8_t* sA, uint8_t* sB, uint64_t* a_full, uint64_t* a_empty,
                                               uint64_t* b_full, uint64_t* b_empty, int& astage, uint32_t& aphase, int& bstage, uint32_t& bphase,
                                               unsigned long long* trace, bool first_unit, unsigned long long* ustamp)
{
    using KT = KTab<KIND>;
    using C = TCfg<PREC, BN>;
    constexpr uint32_t idesc = ptx::umma_idesc(PREC == PREC_TF32 ? 2u : 1u, kBlockM, C::BN);
    uint32_t fresh = 0;
    for (int ch = 0; ch < chunks; ++ch) {
#pragma unroll
        for (int u = 0; u < KT::NU; ++u) {
            ptx::mbar_wait(&a_full[astage], aphase);
            ptx::tc_fence_after();
            if (first_unit && ch == 0 && u == 0 && (threadIdx.x & 31) == 0) trace_stamp(trace, 3);
            if (ustamp && ch == 0 && u == 0 && (threadIdx.x & 31) == 0) *ustamp = gtimer();
            const uint32_t a_unit_lo = desc_lo(sA + astage * kAUnit3);
#pragma unroll
            for (int k = 0; k < KT::ntaps(u); ++k) {
                ptx::mbar_wait(&b_full[bstage], bphase);
                ptx::tc_fence_after();
                if (first_unit && ch == 0 && u == 0 && k == 0 && (threadIdx.x & 31) == 0) trace_stamp(trace, 2);
                const uint32_t b_lo = desc_lo(sB + bstage * C::kBTile);
                const uint32_t a_lo = a_unit_lo + k * kRowShift * (kChunkBytes >> 4);
                if (ptx::elect_one()) {
                    if (PREC == PREC_TF32) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            ptx::umma_tf32(d_tmem, mk_desc(a_lo + 2 * kk), mk_desc(b_lo + 2 * kk), idesc, fresh | (kk ? 1u : 0u));
                    } else if (PREC == PREC_BF16X3) {
                        // chunk = [32 hi | 32 lo] bf16 (A) x [32 w_hi | 32 w_lo] (B); offsets in 16-byte units
                        constexpr int AO[6] = {0, 2, 4, 6, 0, 2};      // hi, hi, lo, lo, hi, hi
                        constexpr int BO[6] = {0, 2, 0, 2, 4, 6};      // w_hi x4,        w_lo x2
#pragma unroll
                        for (int i = 0; i < 6; ++i)
                            ptx::umma_f16(d_tmem, mk_desc(a_lo + AO[i]), mk_desc(b_lo + BO[i]), idesc, fresh | (i ? 1u : 0u));
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            ptx::umma_f16(d_tmem, mk_desc(a_lo + 2 * kk), mk_desc(b_lo + 2 * kk), idesc, fresh | (kk ? 1u : 0u));
                    }
                    ptx::umma_commit(&b_empty[bstage]);
                }
                __syncwarp();
                fresh = 1u;
                if (++bstage == C::kBStages) { bstage = 0; bphase ^= 1; }
            }
            if (ptx::elect_one()) ptx::umma_commit(&a_empty[astage]);
            __syncwarp();
            if (++astage == C::kAStages) { astage = 0; aphase ^= 1; }
        }
    }
}

template <int PREC, int BN>
__global__ void __launch_bounds__(kThreads2, 1)
conv_trunk_kernel(const __grid_constant__ TrunkParams p)
{
    using C = TCfg<PREC, BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                                 // [kAStages][unit]
    uint8_t* sB = sA + C::kAStages * kAUnit3;                           // [kBStages][256 rows x 128 B]
    uint8_t* sT = sB + C::kBStages * C::kBTile;                         // epilogue transpose tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + ((C::kEpiBytes + 1023) & ~1023));
    uint64_t* a_full = bars;                       // [kAStages]
    uint64_t* a_empty = a_full + C::kAStages;
    uint64_t* b_full = a_empty + C::kAStages;      // [kBStages]
    uint64_t* b_empty = b_full + C::kBStages;
    uint64_t* tmem_full = b_empty + C::kBStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* sched_full = tmem_empty + 2;         // [kSched]
    uint64_t* sched_empty = sched_full + C::kSched;
    int* sched_slot = reinterpret_cast<int*>(sched_empty + C::kSched);   // [kSched] work-unit indices
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_slot + C::kSched);

    ptx::grid_dep_launch();
    if (threadIdx.x == 0) trace_stamp(p.trace, 0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned* done = p.sched + 1;                  // done[layer * max_batch + image]

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kAStages; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < C::kBStages; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 8); }
        for (int s = 0; s < C::kSched; ++s) { ptx::mbar_init(&sched_full[s], 1); ptx::mbar_init(&sched_empty[s], 10); }   // B producer + MMA warp + 8 epilogue warps
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) { ptx::tmem_alloc(tmem_slot, C::kTmemCols); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) trace_stamp(p.trace, 1);

    // every consumer role walks the same ring of work units
    int sslot = 0; uint32_t sphase = 0;
    auto next_unit = [&](bool whole_warp) -> int {  // a consumer role's next work unit: called by a whole warp, or by lane 0 alone
        ptx::mbar_wait(&sched_full[sslot], sphase);
        int u = 0;
        if (lane == 0) u = sched_slot[sslot];       // read by the one thread that hands the slot back (a direct mbarrier edge to the writer)
        if (whole_warp) u = __shfl_sync(0xffffffffu, u, 0);
        if (lane == 0) ptx::mbar_arrive(&sched_empty[sslot]);
        if (++sslot == C::kSched) { sslot = 0; sphase ^= 1; }
        return u;
    };

    if (warp == 0) {
        // ============================== scheduler + A producer ================================
        if (lane == 0) {
            for (int l = 0; l < p.n_layers; ++l)    // every layer brings its own tensor maps: fetch the descriptors now, not at each layer's first load
                for (int m = 0; m < (p.layer[l].kind == KIND_S2 ? 4 : 1); ++m) ptx::prefetch_tmap(&p.layer[l].amap[m]);
            ptx::grid_dep_wait();                   // the first layer's input comes from the previous kernel
            int stage = 0; uint32_t phase = 0;
            int ps = 0; uint32_t pph = 0;
            // latency mode: the pieces of a unit (consecutive indices) must sit on DIFFERENT CTAs, because a piece's epilogue waits for
            // the others' partial sums -- units are dealt round robin (index i goes to CTA i mod grid; every wait then points at a
            // smaller index or at a piece whose CTA only has smaller indices left to finish: no cycle).  Throughput mode: first come, first served.
            const bool dealt = BN == 128 && p.ksplit > 1;
            int u = dealt ? static_cast<int>(blockIdx.x) : static_cast<int>(atomicAdd(p.sched, 1u));
            for (;;) {
                ptx::mbar_wait(&sched_empty[ps], pph ^ 1);
                sched_slot[ps] = u;
                ptx::mbar_arrive(&sched_full[ps]);  // release: the slot write is visible to the waiters
                if (++ps == C::kSched) { ps = 0; pph ^= 1; }
                if (u >= p.total_units) break;
                const UnitCoord c = decode_unit(p, u);
                const LayerDesc& L = p.layer[c.l];
                if (L.dep_layer >= 0) {
                    // this image's previous-layer output is complete once all its units' epilogue warps have signalled
                    const int* flag = reinterpret_cast<const int*>(done + L.dep_layer * p.max_batch + c.img);
                    if (static_cast<unsigned>(ptx::ld_acquire_gpu(flag)) < L.dep_target) {
                        const long long t0 = clock64();
                        while (static_cast<unsigned>(ptx::ld_acquire_gpu(flag)) < L.dep_target) {
                            __nanosleep(64);
                            if (clock64() - t0 > (1ll << 34)) __trap();     // ~10 s: a scheduling bug becomes an error, not a hung GPU (long enough for compute-sanitizer's slowdown)
                        }
                    }
                    ptx::fence_proxy_async_all();   // the TMA (async proxy) reads below must observe what the acquire made visible
                }
                unit_stamp(p, u, 0);
                if (L.kind == KIND_S1) trunk_load_unit<KIND_S1>(L, c, sA, a_full, a_empty, stage, phase, C::kAStages);
                else                   trunk_load_unit<KIND_S2>(L, c, sA, a_full, a_empty, stage, phase, C::kAStages);
                u = dealt ? u + static_cast<int>(gridDim.x) : static_cast<int>(atomicAdd(p.sched, 1u));   // pull the next unit only now: look-ahead = the A pipeline depth
            }
        }
    } else if (warp == 3) {
        // ============================== B producer ================================
        if (lane == 0) {
            if (!p.img_wid) for (int l = 0; l < p.n_layers; ++l) ptx::prefetch_tmap(&p.layer[l].bmap);
            int stage = 0; uint32_t phase = 0;
            for (;;) {
                const int u = next_unit(false);
                if (u >= p.total_units) break;
                const UnitCoord c = decode_unit(p, u);
                const LayerDesc& L = p.layer[c.l];
                const CUtensorMap* bm = p.img_wid ? p.gbmaps + p.img_wid[c.img] * kLayersPerSet + L.li : &L.bmap;
                if (L.kind == KIND_S1) trunk_load_weights<KIND_S1, PREC, BN>(L, bm, c, sB, b_full, b_empty, stage, phase);
                else                   trunk_load_weights<KIND_S2, PREC, BN>(L, bm, c, sB, b_full, b_empty, stage, phase);
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        int astage = 0; uint32_t aphase = 0;
        int bstage = 0; uint32_t bphase = 0;
        int it = 0;
        for (;; ++it) {
            const int u = next_unit(true);
            if (u >= p.total_units) break;
            const int l = unit_layer(p, u);
            const int kind = p.layer[l].kind, chunks = p.layer[l].chunks / p.ksplit;      // K chunks of this piece (the host makes chunks divisible)
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            unsigned long long* ust = (p.trace && u < 2048) ? p.trace + 256 * 8 + u * 5 + 1 : nullptr;
            if (kind == KIND_S1) trunk_mma_unit<KIND_S1, PREC, BN>(chunks, d_tmem, sA, sB, a_full, a_empty, b_full, b_empty, astage, aphase, bstage, bphase, p.trace, it == 0, ust);
            else                 trunk_mma_unit<KIND_S2, PREC, BN>(chunks, d_tmem, sA, sB, a_full, a_empty, b_full, b_empty, astage, aphase, bstage, bphase, p.trace, it == 0, ust);
            if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);
            __syncwarp();
            if (lane == 0) unit_stamp(p, u, 2);
        }
        if (lane == 0) trace_stamp(p.trace, 4);
    } else if (warp >= 4) {
        // ============================== epilogue (8 warps) ==========================
        ptx::grid_dep_wait();
        const int ew = warp - 4;
        const int q = ew & 3;                       // TMEM lane quadrant (== warp % 4)
        const int half = ew >> 2;                   // which half of the 256 columns
        constexpr int kCols = BN / 2;
        const int row = q * 32 + lane;
        const int py = row / 11, px = row - py * 11;
        float (*stg)[C::kEpiPitch] = reinterpret_cast<float (*)[C::kEpiPitch]>(sT + ew * C::kEpiWarpBytes);
        int* rowtab = reinterpret_cast<int*>(sT + ew * C::kEpiWarpBytes + 32 * C::kEpiPitch * 4);
        const int grp = lane & 7, sub = lane >> 3;  // lane = (16-byte column group, pixel within a group of 4)
        int it = 0;
        for (;; ++it) {
            const int u = next_unit(true);
            if (u >= p.total_units) break;
            const UnitCoord c = decode_unit(p, u);
            const LayerDesc& L = p.layer[c.l];
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            {
                const int y = c.ty * 11 + py, x = c.tx * 11 + px;
                const bool valid = (row < 121) && (y < L.Ho) && (x < L.Wo);
                __syncwarp();
                rowtab[lane] = valid ? (c.img * L.Ho + y) * L.Wo + x : -1;   // pixel index of TMEM row q*32 + lane
                __syncwarp();
            }
            int pix[8];                                          // pixels this lane post-processes: rows 4k + sub (same for every block)
#pragma unroll
            for (int k = 0; k < 8; ++k) pix[k] = rowtab[4 * k + sub];
            const int ch0 = c.grp * L.cout + c.n_tile * BN + half * kCols;   // first output channel of this warp's 128 columns
            const float* bias_base = p.img_wid ? p.gbias[p.img_wid[c.img] * kLayersPerSet + L.li] : L.bias;
            if (L.res && L.dep_layer >= 0) {
                // the residual was written earlier in THIS launch by other CTAs (an ancestor layer of this unit): order this
                // warp's loads after the completion counter the producer thread already observed
                if (lane == 0) (void)ptx::ld_acquire_gpu(reinterpret_cast<const int*>(done + L.dep_layer * p.max_batch + c.img));
                __syncwarp();
            }
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, 5);
            if (threadIdx.x == 128) unit_stamp(p, u, 3);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * kCols;
            const bool split = BN == 128 && p.ksplit > 1;      // latency mode exists for the 128-channel units only (checked at launch)
            // this warp's slice (32 rows x kCols columns) of piece `pc` of this unit in the split-K scratch: row-major, 32-column blocks
            auto slice_of = [&](int pc) -> float* {
                return p.partial + ((static_cast<size_t>(c.gidx) * p.ksplit + pc) * 8 + ew) * (32 * kCols);
            };
            unsigned own = 0xFu;                                 // which of this warp's 32-column blocks it post-processes (bit per block)
            if (split) {
                // latency mode: this CTA only summed K chunks [c0, c1).  Dump the fp32 slice and hand the accumulator back.  The unit's
                // four 32-column blocks are then finished by its pieces in parallel: block b by piece b * ksplit / 4, which waits until
                // every piece has dumped and adds them in piece order 0..ksplit-1 (the result does not depend on arrival order).
                float* dst = slice_of(c.piece);
#pragma unroll 1
                for (int c0 = 0; c0 < kCols; c0 += 32) {
                    uint32_t r0[16], r1[16];
                    ptx::tmem_ld16(taddr + c0, r0);
                    ptx::tmem_ld16(taddr + c0 + 16, r1);
                    ptx::tmem_ld_wait();
                    // float4 number jj of row `lane` lives at [block][jj][lane]: every store / load instruction of the warp is 512 contiguous bytes
                    float4* d4 = reinterpret_cast<float4*>(dst + (c0 / 32) * 1024) + lane;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        d4[jj * 32] = make_float4(__uint_as_float(r0[4 * jj]), __uint_as_float(r0[4 * jj + 1]), __uint_as_float(r0[4 * jj + 2]), __uint_as_float(r0[4 * jj + 3]));
                        d4[(4 + jj) * 32] = make_float4(__uint_as_float(r1[4 * jj]), __uint_as_float(r1[4 * jj + 1]), __uint_as_float(r1[4 * jj + 2]), __uint_as_float(r1[4 * jj + 3]));
                    }
                }
                ptx::tc_fence_before();
                __threadfence();
                __syncwarp();
                if (lane == 0) { ptx::mbar_arrive(&tmem_empty[acc]); atomicAdd(p.slice_cnt + c.gidx * 8 + ew, 1u); }
                own = 0u;
#pragma unroll
                for (int bi = 0; bi < kCols / 32; ++bi)
                    if ((half * (kCols / 32) + bi) * p.ksplit / (BN / 32) == c.piece) own |= 1u << bi;
                if (!own) { if (threadIdx.x == 128) unit_stamp(p, u, 4); continue; }   // other pieces finish this warp's blocks
                if (lane == 0) {
                    const int* cnt = reinterpret_cast<const int*>(p.slice_cnt + c.gidx * 8 + ew);
                    const long long t0 = clock64();
                    while (ptx::ld_acquire_gpu(cnt) < p.ksplit) {
                        __nanosleep(32);
                        if (clock64() - t0 > (1ll << 34)) __trap();
                    }
                }
                __syncwarp();
                __threadfence();                                                     // order the reads below after the other pieces' dumps
            }
#pragma unroll 1
            for (int c0 = 0; c0 < kCols; c0 += 32) {
                if (!((own >> (c0 / 32)) & 1u)) continue;
                const int chan = ch0 + c0;                        // first channel of this 32-channel block
                // bias and residual pieces first: their L2 latency overlaps the TMEM load / the split-K reads and the staging round trip below
                float4 b4;
                if (BN == 128) b4 = __ldg(reinterpret_cast<const float4*>(bias_base + chan + grp * 4));
                float4 r4[8];                                    // TF32: 4 fp32 words
                uint2 rh[8], rl[8];                              // bf16: 4 channels hi (and lo, BF16X3)
                if (L.res) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        r4[k] = make_float4(0.f, 0.f, 0.f, 0.f); rh[k] = make_uint2(0u, 0u); rl[k] = make_uint2(0u, 0u);
                        if (pix[k] >= 0) {
                            const uint8_t* rp = L.res + chan_byte<PREC>(static_cast<size_t>(pix[k]), L.res_c, chan + grp * 4);
                            if (PREC == PREC_TF32) r4[k] = __ldcg(reinterpret_cast<const float4*>(rp));
                            else { rh[k] = __ldcg(reinterpret_cast<const uint2*>(rp)); if (PREC == PREC_BF16X3) rl[k] = __ldcg(reinterpret_cast<const uint2*>(rp + 64)); }
                        }
                    }
                }
                if (!split) {
                    uint32_t r0[16], r1[16];
                    ptx::tmem_ld16(taddr + c0, r0);
                    ptx::tmem_ld16(taddr + c0 + 16, r1);
                    ptx::tmem_ld_wait();
                    __syncwarp();                                // previous block's readers are done with stg
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) {
                        *reinterpret_cast<uint4*>(&stg[lane][jj]) = make_uint4(r0[jj], r0[jj + 1], r0[jj + 2], r0[jj + 3]);
                        *reinterpret_cast<uint4*>(&stg[lane][16 + jj]) = make_uint4(r1[jj], r1[jj + 1], r1[jj + 2], r1[jj + 3]);
                    }
                } else {
                    // sum of the pieces, row `lane` of this block, in piece order.  All pieces' loads of half a row are issued before the
                    // first add (16 x 16 bytes in flight per lane): two L2 round trips per block instead of one per piece
                    __syncwarp();
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float4 v[kSplitK][4];
#pragma unroll
                        for (int pc = 0; pc < kSplitK; ++pc) {
                            if (pc < p.ksplit) {
                                const float4* s4 = reinterpret_cast<const float4*>(slice_of(pc) + (c0 / 32) * 1024) + lane + hh * 128;
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) v[pc][jj] = __ldcg(s4 + jj * 32);
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int pc = 0; pc < kSplitK; ++pc)
                                if (pc < p.ksplit) { a.x += v[pc][jj].x; a.y += v[pc][jj].y; a.z += v[pc][jj].z; a.w += v[pc][jj].w; }
                            *reinterpret_cast<float4*>(&stg[lane][4 * (hh * 4 + jj)]) = a;
                        }
                    }
                }
                if (!split && c0 + 32 == kCols) {                 // last TMEM read of this unit: hand the accumulator back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
                } else {
                    __syncwarp();
                }
                if (BN != 128) b4 = __ldg(reinterpret_cast<const float4*>(bias_base + chan + grp * 4));   // 256-channel units: registers are tight, and the latency hides behind the next unit's MMAs
                float4 a4[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a4[k] = *reinterpret_cast<const float4*>(&stg[4 * k + sub][grp * 4]);
                    a4[k].x += b4.x; a4[k].y += b4.y; a4[k].z += b4.z; a4[k].w += b4.w;
                }
                if (L.res) {
                    if (PREC == PREC_TF32) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { a4[k].x += r4[k].x; a4[k].y += r4[k].y; a4[k].z += r4[k].z; a4[k].w += r4[k].w; }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float2 h0 = unpack2(rh[k].x), h1 = unpack2(rh[k].y), l0 = unpack2(rl[k].x), l1 = unpack2(rl[k].y);
                            a4[k].x += h0.x + l0.x; a4[k].y += h0.y + l0.y; a4[k].z += h1.x + l1.x; a4[k].w += h1.y + l1.y;
                        }
                    }
                }
                float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);   // fused average pool: this lane's rows, 4 channels
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 o = a4[k];
                    o.x = act_apply(o.x, L.act); o.y = act_apply(o.y, L.act); o.z = act_apply(o.z, L.act); o.w = act_apply(o.w, L.act);
                    if (pix[k] < 0) continue;
                    if (L.pool_part) { psum.x += o.x; psum.y += o.y; psum.z += o.z; psum.w += o.w; continue; }
                    uint8_t* po = L.out + chan_byte<PREC>(static_cast<size_t>(pix[k]), L.out_c, L.out_coff + chan + grp * 4);
                    if (PREC == PREC_TF32) {
                        *reinterpret_cast<float4*>(po) = make_float4(ptx::to_tf32(o.x), ptx::to_tf32(o.y), ptx::to_tf32(o.z), ptx::to_tf32(o.w));
                    } else if (PREC == PREC_BF16X3) {
                        uint32_t h0, l0, h1, l1;
                        split2(o.x, o.y, h0, l0); split2(o.z, o.w, h1, l1);
                        *reinterpret_cast<uint2*>(po) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(po + 64) = make_uint2(l0, l1);
                    } else {
                        *reinterpret_cast<uint2*>(po) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                    }
                }
                if (L.pool_part) {                                // rows 4k + sub summed above; fold the four `sub` groups (fixed order: deterministic)
#pragma unroll
                    for (int off = 8; off <= 16; off <<= 1) {
                        psum.x += __shfl_xor_sync(0xffffffffu, psum.x, off); psum.y += __shfl_xor_sync(0xffffffffu, psum.y, off);
                        psum.z += __shfl_xor_sync(0xffffffffu, psum.z, off); psum.w += __shfl_xor_sync(0xffffffffu, psum.w, off);
                    }
                    if (sub == 0)
                        *reinterpret_cast<float4*>(L.pool_part + (static_cast<size_t>(c.img) * 4 + q) * L.out_c + L.out_coff + chan + grp * 4) = psum;
                }
            }
            // this warp's part of the unit is in memory: publish it to the units of the next layer that wait for this image
            if (c.l + 1 < p.n_layers) {
                ptx::fence_proxy_async_all();       // consumers read these bytes through TMA (async proxy)
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(done + c.l * p.max_batch + c.img, 1u);
            }
            if (threadIdx.x == 128) unit_stamp(p, u, 4);
        }
        if (threadIdx.x == 128) trace_stamp(p.trace, 6);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) trace_exit(p.trace);
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename K>
cudaError_t set_smem(K kernel, size_t smem, size_t* cache) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > cache[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        cache[dev] = smem;
    }
    return cudaSuccess;
}

template <int KIND, int PREC>
cudaError_t launch_resident_t(const ResidentParams& p, int num_sms, bool pdl, cudaStream_t stream) {
    using C = RCfg<KIND, PREC>;
    const int tiles_per_tap = (C::kStack == 2) ? 1 : p.L.chunks;
    if ((KIND == KIND_STEM ? 7 : 9) * tiles_per_tap > C::kMaxWTiles) return cudaErrorInvalidValue;
    if (C::kStack == 2 && KIND == KIND_S1 && p.L.chunks != 2) return cudaErrorInvalidValue;    // a stacked tile row is exactly two 32-channel chunks
    if (p.L.cout != 64 || p.L.groups != 1 || p.L.n_tiles != 1 || p.m_tiles <= 0) return cudaErrorInvalidValue;
    static size_t attr[64] = {};                   // the dynamic shared-memory limit is a per-device function attribute
    cudaError_t e = set_smem(conv_resident_kernel<KIND, PREC>, C::kSmem, attr);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(std::min(p.m_tiles, num_sms)); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = C::kSmem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, conv_resident_kernel<KIND, PREC>, p);
}

template <int PREC, int BN>
cudaError_t launch_trunk_t(const TrunkParams& p, int num_sms, bool pdl, cudaStream_t stream) {
    using C = TCfg<PREC, BN>;
    if (p.n_layers < 1 || p.n_layers > kTrunkMaxLayers || p.total_units <= 0 || !p.sched) return cudaErrorInvalidValue;
    if (p.ksplit < 1 || (p.ksplit > 1 && (!p.partial || !p.slice_cnt || p.ksplit > kSplitK || C::BN != 128))) return cudaErrorInvalidValue;
    for (int l = 0; l < p.n_layers; ++l) {
        const LayerDesc& L = p.layer[l];
        if ((L.kind != KIND_S1 && L.kind != KIND_S2) || L.cout % C::BN || L.n_tiles != L.cout / C::BN) return cudaErrorInvalidValue;
        if (L.pool_part && (L.tiles_x != 1 || L.tiles_y != 1)) return cudaErrorInvalidValue;   // fused avg-pool: tile = whole image
        if (L.dep_layer >= l) return cudaErrorInvalidValue;                                    // dependencies point backwards in the pull order
        if (L.chunks % p.ksplit) return cudaErrorInvalidValue;
    }
    static size_t attr[64] = {};
    cudaError_t e = set_smem(conv_trunk_kernel<PREC, BN>, C::kSmem, attr);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    // all CTAs must be co-resident (one per SM): the dependency waits rely on it
    cfg.gridDim = dim3(std::min(p.total_units, num_sms)); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = C::kSmem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, conv_trunk_kernel<PREC, BN>, p);
}

}  // namespace

cudaError_t launch_conv_resident(const ResidentParams& p, int kind, int prec, int num_sms, bool pdl, cudaStream_t stream) {
    if (kind == KIND_STEM) {
        switch (prec) {
            case PREC_TF32:   return launch_resident_t<KIND_STEM, PREC_TF32>(p, num_sms, pdl, stream);
            case PREC_BF16X3: return launch_resident_t<KIND_STEM, PREC_BF16X3>(p, num_sms, pdl, stream);
            case PREC_BF16:   return launch_resident_t<KIND_STEM, PREC_BF16>(p, num_sms, pdl, stream);
        }
    } else if (kind == KIND_S1) {
        switch (prec) {
            case PREC_TF32:   return launch_resident_t<KIND_S1, PREC_TF32>(p, num_sms, pdl, stream);
            case PREC_BF16X3: return launch_resident_t<KIND_S1, PREC_BF16X3>(p, num_sms, pdl, stream);
            case PREC_BF16:   return launch_resident_t<KIND_S1, PREC_BF16>(p, num_sms, pdl, stream);
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_conv_trunk(const TrunkParams& p, int prec, int block_n, int num_sms, bool pdl, cudaStream_t stream) {
    if (block_n == 256) {
        switch (prec) {
            case PREC_TF32:   return launch_trunk_t<PREC_TF32, 256>(p, num_sms, pdl, stream);
            case PREC_BF16X3: return launch_trunk_t<PREC_BF16X3, 256>(p, num_sms, pdl, stream);
            case PREC_BF16:   return launch_trunk_t<PREC_BF16, 256>(p, num_sms, pdl, stream);
        }
    } else if (block_n == 128) {
        switch (prec) {
            case PREC_TF32:   return launch_trunk_t<PREC_TF32, 128>(p, num_sms, pdl, stream);
            case PREC_BF16X3: return launch_trunk_t<PREC_BF16X3, 128>(p, num_sms, pdl, stream);
            case PREC_BF16:   return launch_trunk_t<PREC_BF16, 128>(p, num_sms, pdl, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace se3tn

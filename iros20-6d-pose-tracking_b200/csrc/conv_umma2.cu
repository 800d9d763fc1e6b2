// conv_umma2: tcgen05 implicit-GEMM convolution, second generation.
//
// v1 (conv_umma.cu) pulls one TMA box per (filter tap, 32-channel chunk): every activation patch
// crosses L2 -> shared memory 9 times and every weight tile once per M tile, and the measured
// forward is bound by that fill traffic (~11 GB per 64-pair forward), not by the tensor pipe.
// v2 removes most of it:
//   * "units": one TMA box per (chunk, filter COLUMN) that is `bh + 2` rows tall.  The three
//     vertical taps of that column are the same shared-memory tile read through UMMA descriptors
//     whose start address is advanced by whole pixel rows (bw*128 bytes) -- no data moves.
//     A fill traffic: 9x -> 3x (+2/11 halo).  The stem does the same over its 7 filter rows with
//     two units (even / odd input rows), stride-2 convs with six (parity views).
//   * weights either RESIDENT in shared memory for the whole kernel (64-channel layers and the
//     stem: the complete K-major matrix is <= 147 KB) or streamed through their own ring.
//   * the stem's epilogue fuses MaxPool2d(3,2,1): an M tile is an 11x11 block of conv outputs that
//     yields a 5x5 block of pooled outputs; max -> +bias -> SELU (monotone, so they commute) run
//     on 1/4.84 of the values and the 88x88x64 intermediate never reaches HBM.
//   * 8 epilogue warps (two per TMEM lane quadrant).
//   * Ring-weight kernels (BN = 256) transpose each 32-row x 32-column accumulator block through a
//     per-warp shared-memory tile so that every global load/store instruction of the epilogue covers 4
//     whole 128-byte lines (8 lanes per pixel) instead of 32 different lines (one 16-byte piece per
//     lane): the row-per-lane epilogue was LSU-bound at ~25-30 us per dual-M work unit.
//   * Programmatic dependent launch: every CTA signals launch_dependents at entry, so the next conv's
//     CTAs start on SMs as they drain and run their prologue (barrier init, TMEM alloc, weight TMA --
//     weights are never written during a forward) under the tail of this one; only the activation
//     producer and the epilogue warps execute griddepcontrol.wait before touching activations.
//   * Per-object weights in ONE launch (reference README.md:132: one checkpoint per object class): with
//     p.img_wid set, each work unit looks up its image's weight-set id and takes its weight tensor map and
//     bias from device tables; RESIDENT kernels reload their shared-memory weights when the id changes
//     between consecutive tiles (CTAs own CONTIGUOUS tile ranges, so with tracks grouped by id that is rare).
//   * PAIR (resident-weight kernels: stem, 64-channel layers): clusters of two CTAs run ONE tcgen05.mma.cta_group::2
//     of M = 256 per step -- each CTA supplies its own pixel tile (A) and HALF of the weight rows (B) from its own
//     shared memory, the leader CTA issues, commits are multicast to both.  Measured (profiles/
//     r01_umma_rate3_cta_pair_probe.txt): an N = 64 pair MMA costs 43 cycles (74 % of the per-SM tensor peak)
//     against 52-65 cycles (50-60 %) for a single-CTA one -- but in the network the cross-CTA barrier round trips
//     eat that, so PAIR is opt-in (SE3TN_PAIR=1).
//   * N = 64 tiles (stem, 64-channel layers) are capped by the hardware: one 128xNx32B tcgen05.mma costs
//     52-65 cycles at N = 64, 68-74 at N = 128, 128 at N = 256 (scripts/umma_rate2.cu) -- a per-instruction floor.
//   * STACK (resident-weight kernels, bf16 modes): because of that floor, the hi/lo weight halves are stacked along N:
//     the weight tile of a tap has 128 rows, [w_hi rows 0-63 ; w_lo rows 64-127], so ONE N = 128 MMA forms a_hi*w_hi
//     (accumulator columns 0-63) and a_hi*w_lo (columns 64-127), a second N = 64 MMA adds a_lo*w_hi into columns
//     0-63, and the epilogue sums the two column halves: 4 instead of 6 MMAs per chunk-tap (stem: 4 instead of 8).
//   * MT = 2 ("dual-M", BN = 256 layers): one CTA carries TWO M tiles (two accumulators, all 512
//     TMEM columns) through the K loop, so every weight tile fetched from L2 feeds 8 MMAs instead
//     of 4 -- the weight stream, which is >80% of the fill traffic of the deep layers, halves.
//     At batch 64 each deep layer is exactly 128 work units: one wave on 148 SMs.
//   * Stream-K (ring-weight kernels): at batch 64 every BN = 256 layer has 256 work units for 148 CTAs -- two waves, the second
//     73 % full.  When there are more units than CTAs, the (unit, 32-channel chunk) steps are dealt out evenly instead: a
//     CTA's contiguous range starts with the TAIL chunks of one unit (accumulate, dump the fp32 partial to its slot in
//     p.sk_part, raise its flag), runs whole units, and ends with the HEAD chunks of another, whose epilogue first adds
//     the partial its neighbour dumped long before.  Makespan 1.75 instead of 2 units.
//   * PREC selects the arithmetic without touching the byte layout.  Every 128-byte K chunk of an
//     activation pixel / weight row is either 32 fp32 words holding TF32 values (PREC_TF32) or
//     [32 x bf16 hi | 32 x bf16 lo] of the same 32 channels (x = hi + lo to 16 mantissa bits):
//       PREC_TF32    4 x kind::tf32 MMAs per chunk
//       PREC_BF16X3  6 x kind::f16 (bf16) MMAs per chunk: hi*w_hi, lo*w_hi, hi*w_lo via descriptor
//                    offsets 0/64 bytes into the same tiles -- fp32-faithful (error ~2^-16) at 1.5x
//                    the tensor time of TF32 and identical fill traffic
//       PREC_BF16    2 MMAs (hi*w_hi): the plain bf16 tensor-core path (BASELINE configs[2])
//     The stem's 8-pixel window interleaves hi/lo per pixel, so it uses two weight tiles per filter
//     row ([w_hi|w_hi] and [w_lo|0]) -- 8 MMAs.
// Same tensors, packed weights, tile boxes and epilogue semantics as v1 (see conv_common.h).
#include "conv_common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace se3tn {
namespace {

constexpr int kThreads2 = 384;                 // warps: 0 A-TMA, 1 MMA, 2 TMEM alloc, 3 B-TMA, 4..11 epilogue
// A unit buffer: (max row shift + 128) rows * 128 B, rounded to 1 KB: stem 33+128 rows -> 21 KB, 3x3 22+128 -> 19 KB
constexpr int kPoolPitch = 68;                 // floats per staged conv position (64 + 4: bank spread)
constexpr int kPoolStageBytes = 121 * kPoolPitch * 4;

enum { PREC_TF32 = 0, PREC_BF16X3 = 1, PREC_BF16 = 2 };

// Compile-time unit/tap structure per conv kind, so the MMA issue loop is straight-line code with
// immediate row shifts / weight-tile indices (a runtime plan table cost ~200 cycles of dependent
// constant-bank loads per MMA group).  Must agree with the host plan (checked in launch2).
template <int KIND> struct KTab;
template <> struct KTab<KIND_S1> {            // 3x3 stride 1: unit = filter column s, taps = filter rows r
    static constexpr int NU = 3;
    __host__ __device__ static constexpr int ntaps(int) { return 3; }
    __host__ __device__ static constexpr int shift(int, int k) { return k * 11; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return k * 3 + u; }
};
template <> struct KTab<KIND_S2> {            // 3x3 stride 2: per column s an even-row unit (r=1) and an odd-row unit (r=0,2)
    static constexpr int NU = 6;
    __host__ __device__ static constexpr int ntaps(int u) { return (u & 1) ? 2 : 1; }
    __host__ __device__ static constexpr int shift(int, int k) { return k * 11; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return (u & 1) ? (k == 0 ? (u >> 1) : 6 + (u >> 1)) : 3 + (u >> 1); }
};
template <> struct KTab<KIND_STEM> {          // 7x7 stride 2 stem: even input rows (r=0,2,4,6), odd input rows (r=1,3,5)
    static constexpr int NU = 2;
    __host__ __device__ static constexpr int ntaps(int u) { return u == 0 ? 4 : 3; }
    __host__ __device__ static constexpr int shift(int, int k) { return k * 11; }
    __host__ __device__ static constexpr int wtap(int u, int k) { return 2 * k + u; }
};

template <int BN, bool RESIDENT, int KIND, int MT, int PREC = 0, bool PAIR = false> struct Cfg2 {
    static_assert(!PAIR || (RESIDENT && MT == 1), "CTA pairs are implemented for the resident-weight kernels");
    static constexpr bool POOL = (KIND == KIND_STEM);
    // STACK: hi / lo weight rows stacked along N (see the header comment); the stem's bf16 path always needs all three products
    static constexpr int kStack = (RESIDENT && !PAIR && PREC != PREC_TF32 && (KIND == KIND_STEM || PREC == PREC_BF16X3)) ? 2 : 1;
    static constexpr int kBTile = (PAIR ? BN / 2 : BN * kStack) * kChunkBytes;    // pair: each CTA holds half of the weight rows
    static constexpr int kAUnit = POOL ? 21 * 1024 : 19 * 1024;     // 3x3: (22 + 128) rows * 128 B = 19,200
    static constexpr int kAStage = MT * kAUnit;
    static constexpr int kAStages = RESIDENT ? (PAIR ? (POOL ? 4 : 6) : ((POOL && PREC != PREC_TF32) ? 3 : 4)) : (MT == 2 ? 2 : 3);
    static constexpr bool kEpiT = !RESIDENT;                            // transposed (coalesced) epilogue through per-warp smem tiles
    static constexpr int kEpiPitch = 36;                                // words per staged row (32 + 4: conflict-free 16 B accesses)
    static constexpr int kEpiWarpBytes = 32 * kEpiPitch * 4 + 128;      // 32 rows + 32-entry pixel-index table
    static constexpr int kEpiBytes = kEpiT ? 8 * kEpiWarpBytes : 0;
    static constexpr int kWPerTap = (POOL && PREC != PREC_TF32 && kStack == 1) ? 2 : 1;  // weight tiles per (tap, chunk)
    static constexpr int kPoolBufs = POOL ? (PREC == PREC_TF32 ? 2 : 1) : 0;
    static constexpr int kBStages = RESIDENT ? 0 : (BN == 256 ? (MT == 2 ? 3 : 4) : 6);
    // Partial accumulators per tile (independent MMA chains).  Measured (profiles/r01_umma_rate_microbench.txt):
    // a 128xNx(32 B) tcgen05.mma costs ~90 cycles for N <= 128 whether or not consecutive MMAs share an
    // accumulator, so splitting buys nothing -- kept at 1 (the code path stays for experiments).
    static constexpr int kSplit = 1;
    static constexpr int kAccCols = MT * BN * kSplit * kStack;          // TMEM columns of one accumulator set
    static constexpr int kNAcc = (2 * kAccCols <= 512) ? 2 : 1;         // accumulator sets (double-buffered when they fit)
    static constexpr int kTmemCols = kNAcc * kAccCols;                  // 256 / 512
};

__device__ __forceinline__ float selu_fast(float x) {
    constexpr float kAlpha = 1.6732632423543772f, kScale = 1.0507009873554805f;
    return x > 0.f ? kScale * x : (kScale * kAlpha) * (__expf(x) - 1.f);
}

// timeline stamps (debug): slot 0 kernel entry, 1 setup done, 2 MMA warp has its first weights, 3 MMA warp has its first A unit,
// 4 MMA warp issued its last commit, 5 epilogue got its first accumulator, 6 epilogue finished its last tile, 7 CTA exit (low 8 bits: SM id)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_stamp(unsigned long long* tr, int slot) { if (tr) tr[blockIdx.x * 8 + slot] = gtimer(); }

struct TileCoord2 { int ox, oy, n0, tx, ty; };
struct WorkUnit { int mp, n_tile, grp; };

__device__ __forceinline__ WorkUnit decode_work(int w, int m_units, const Umma2Plan& t) {
    WorkUnit u;
    u.mp = w % m_units;
    const int rest = w / m_units;
    u.n_tile = rest % t.n_tiles;
    u.grp = rest / t.n_tiles;
    return u;
}

__device__ __forceinline__ TileCoord2 decode2(int m, const Umma2Plan& t) {
    TileCoord2 c;
    c.tx = m % t.tiles_x;
    const int r2 = m / t.tiles_x;
    c.ty = r2 % t.tiles_y;
    c.n0 = t.img_first + (r2 / t.tiles_y) * t.bn;
    c.ox = c.tx * t.step_x + t.off_x;           // tile origin in A-map coordinates
    c.oy = c.ty * t.step_y + t.off_y;
    return c;
}

// fp32 -> (bf16 hi, bf16 lo) with x ~= hi + lo; packs two values per 32-bit word (element 0 in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

template <int BN, bool RESIDENT, int KIND, int MT, int PREC, bool PAIR>
__global__ void __launch_bounds__(kThreads2, 1)
conv_umma2_kernel(const __grid_constant__ UmmaMaps maps, const ConvGeom g, const Umma2Plan t, const ConvPtrs p)
{
    using C = Cfg2<BN, RESIDENT, KIND, MT, PREC, PAIR>;
    using KT = KTab<KIND>;
    constexpr bool POOL = C::POOL;
    static_assert(!(POOL && MT != 1), "pool epilogue is single-tile");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int w_tiles = C::kStack == 2 ? g.num_taps : g.num_taps * t.chunks * C::kWPerTap;   // K tiles of the weight matrix (STACK: one per tap, all chunks)
    uint8_t* sA = smem;                                                 // [kAStages][MT][unit]
    uint8_t* sB = sA + C::kAStages * C::kAStage;                       // resident: [w_tiles][BN*128]; ring: [kBStages][BN*128]
    uint8_t* sP = sB + (RESIDENT ? w_tiles : C::kBStages) * C::kBTile;  // pool staging (POOL only): 2 x 121 x 68 floats
    uint8_t* sT = sP + C::kPoolBufs * ((kPoolStageBytes + 1023) & ~1023);   // epilogue transpose tiles (kEpiT only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + ((C::kEpiBytes + 1023) & ~1023));
    uint64_t* a_full = bars;                       // [kAStages]
    uint64_t* a_empty = a_full + C::kAStages;      // [kAStages]
    uint64_t* b_full = a_empty + C::kAStages;      // [max(kBStages,1)]  (resident: b_full[0] = "weights landed")
    uint64_t* b_empty = b_full + (RESIDENT ? 1 : C::kBStages);
    uint64_t* tmem_full = b_empty + (RESIDENT ? 1 : C::kBStages);   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    ptx::grid_dep_launch();
    if (threadIdx.x == 0) trace_stamp(p.trace, 0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_units = (t.m_tiles + MT - 1) / MT;
    const int total_tiles = m_units * t.n_tiles * g.groups;       // work units
    // each CTA (pair) owns a contiguous range of work units (consecutive tiles of the same image / weight set);
    // in a pair, CTA rank r handles tile 2*work + r of every work unit and both walk the range in lockstep
    const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;
    const bool leader = !PAIR || crank == 0;
    const int n_split = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int my_split = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int total_work = PAIR ? (total_tiles + 1) / 2 : total_tiles;
    const int w_begin = static_cast<int>(static_cast<long long>(my_split) * total_work / n_split);
    const int w_end = static_cast<int>(static_cast<long long>(my_split + 1) * total_work / n_split);
    auto tile_of = [&](int work) -> int {          // this CTA's tile of a work unit (odd tail of a pair: clamp, result discarded)
        if (!PAIR) return work;
        const int tl = 2 * work + static_cast<int>(crank);
        return tl < total_tiles ? tl : total_tiles - 1;
    };
    auto tile_valid = [&](int work) -> bool { return !PAIR || 2 * work + static_cast<int>(crank) < total_tiles; };
    // Work items: (unit, chunk range).  Without stream-K every item is a whole unit of [w_begin, w_end).
    const int chunks = t.chunks;
    const bool streamk = !RESIDENT && MT == 1 && p.sk_part != nullptr && total_tiles > static_cast<int>(gridDim.x);
    int s_begin = w_begin * chunks, s_end = w_end * chunks;
    if (streamk) {
        const long long S = static_cast<long long>(total_tiles) * chunks;
        s_begin = static_cast<int>(blockIdx.x * S / gridDim.x);
        s_end = static_cast<int>((blockIdx.x + 1) * S / gridDim.x);
    }
    struct Item { int tile, c0, c1, next; };
    auto item_at = [&](int sidx) -> Item {
        Item im; im.tile = sidx / chunks;
        const int base = im.tile * chunks;
        im.c0 = sidx - base; im.c1 = min(chunks, s_end - base); im.next = base + im.c1;
        return im;
    };
    auto wid_of = [&](int work) -> int {            // weight-set id of a work unit (-1: single-set launch)
        if (!p.img_wid) return -1;
        const WorkUnit wu = decode_work(tile_of(work), m_units, t);
        int m = wu.mp * MT; if (m >= t.m_tiles) m = t.m_tiles - 1;
        return p.img_wid[decode2(m, t).n0];
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kAStages; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < (RESIDENT ? 1 : C::kBStages); ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < C::kNAcc; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], PAIR ? 16 : 8); }   // pair: both CTAs' epilogue warps arrive on the leader's
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) {
        if (PAIR) { ptx::tmem_alloc_2sm(tmem_slot, C::kTmemCols); ptx::tmem_relinquish_2sm(); }
        else      { ptx::tmem_alloc(tmem_slot, C::kTmemCols); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();     // barrier inits must be visible to the peer before any remote signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) trace_stamp(p.trace, 1);

    if (warp == 0) {
        // ============================== A producer ================================
        if (lane == 0) {
            ptx::grid_dep_wait();                   // activations come from the previous kernel
            int stage = 0; uint32_t phase = 0;
            for (int sidx = s_begin; sidx < s_end;) {
                const Item im = item_at(sidx); sidx = im.next;
                const int tile = im.tile;
                const WorkUnit wu = decode_work(tile_of(tile), m_units, t);
                TileCoord2 tc[MT];
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    int m = wu.mp * MT + j; if (m >= t.m_tiles) m = t.m_tiles - 1;   // odd tail: reload a valid tile, result discarded
                    tc[j] = decode2(m, t);
                }
                const int cbase = g.in_coff + wu.grp * g.cin;
                for (int ch = im.c0; ch < im.c1; ++ch) {
                    for (int u = 0; u < t.units_per_chunk; ++u) {
                        const Unit un = t.units[u];
                        ptx::mbar_wait(&a_empty[stage], phase ^ 1);
                        if (!PAIR && (t.debug & 2)) { ptx::mbar_arrive(&a_full[stage]); if (++stage == C::kAStages) { stage = 0; phase ^= 1; } continue; }   // timing experiment: no A fill
                        if (PAIR) {
                            // the leader's barrier collects the bytes of BOTH CTAs' boxes
                            if (leader) ptx::mbar_arrive_expect_tx(&a_full[stage], static_cast<uint32_t>(un.rows) * kChunkBytes * 2);
                            ptx::tma_load_4d_2sm(sA + stage * C::kAStage, &maps.a[un.map], &a_full[stage],
                                                 cbase + ch * 32, tc[0].ox + un.c1, tc[0].oy + un.c2, tc[0].n0);
                        } else {
                            ptx::mbar_arrive_expect_tx(&a_full[stage], static_cast<uint32_t>(un.rows) * kChunkBytes * MT);
#pragma unroll
                            for (int j = 0; j < MT; ++j)
                                ptx::tma_load_4d(sA + stage * C::kAStage + j * C::kAUnit, &maps.a[un.map], &a_full[stage],
                                                 cbase + ch * 32, tc[j].ox + un.c1, tc[j].oy + un.c2, tc[j].n0);
                        }
                        if (++stage == C::kAStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ============================== B producer ================================
        if (lane == 0) {
            if (RESIDENT) {
                // whole K-major weight matrix of this CTA's (only) N tile; reloaded only when the weight-set id changes
                int cur = -2; uint32_t gen = 0;
                for (int tile = w_begin; tile < w_end; ++tile) {
                    const int wid = wid_of(tile);
                    if (wid == cur) continue;
                    if (gen) ptx::mbar_wait(&b_empty[0], (gen - 1) & 1);       // MMAs that read the previous weights have retired
                    const CUtensorMap* bm = wid < 0 ? &maps.b : p.gbmaps + wid * kLayersPerSet;
                    if (PAIR) {                               // each CTA loads its half of the weight rows; the leader's barrier counts both
                        if (leader) ptx::mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(w_tiles) * C::kBTile * 2);
                        for (int wt = 0; wt < w_tiles; ++wt)
                            ptx::tma_load_2d_2sm(sB + wt * C::kBTile, bm, &b_full[0], wt * 32, static_cast<int>(crank) * (BN / 2));
                    } else {
                        ptx::mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(w_tiles) * C::kBTile);
                        for (int wt = 0; wt < w_tiles; ++wt)      // tile wt = (tap*chunks + ch)*kWPerTap + pass, 32 words of K each
                            ptx::tma_load_2d(sB + wt * C::kBTile, bm, &b_full[0], wt * 32, 0);
                    }
                    cur = wid; ++gen;
                }
            } else {
                int stage = 0; uint32_t phase = 0;
                for (int sidx = s_begin; sidx < s_end;) {
                    const Item im = item_at(sidx); sidx = im.next;
                    const int tile = im.tile;
                    const WorkUnit wu = decode_work(tile_of(tile), m_units, t);
                    const int wrow = wu.grp * g.cout + wu.n_tile * BN;
                    const int wid = wid_of(tile);
                    const CUtensorMap* bm = wid < 0 ? &maps.b : p.gbmaps + wid * kLayersPerSet;
                    for (int ch = im.c0; ch < im.c1; ++ch)
                        for (int u = 0; u < t.units_per_chunk; ++u) {
                            const Unit un = t.units[u];
                            for (int k = 0; k < un.ntaps; ++k) {
                                ptx::mbar_wait(&b_empty[stage], phase ^ 1);
                                if (t.debug & 1) { ptx::mbar_arrive(&b_full[stage]); if (++stage == C::kBStages) { stage = 0; phase ^= 1; } continue; }   // timing experiment: no B fill
                                ptx::mbar_arrive_expect_tx(&b_full[stage], C::kBTile);
                                ptx::tma_load_2d(sB + stage * C::kBTile, bm, &b_full[stage], un.taps[k].w_tap * g.cin + ch * 32, wrow);
                                if (++stage == C::kBStages) { stage = 0; phase ^= 1; }
                            }
                        }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ============================== MMA issuer (pair: leader CTA only) ========
        // The WHOLE warp walks the loop (warp-uniform control flow lets ptxas keep descriptors and
        // barrier addresses in uniform registers); one elected lane issues the tcgen05 instructions.
        constexpr uint32_t idesc = ptx::umma_idesc(PREC == PREC_TF32 ? 2u /*tf32*/ : 1u /*bf16*/, PAIR ? 2 * kBlockM : kBlockM, BN);
        constexpr uint32_t idesc_stack = ptx::umma_idesc(1u, kBlockM, BN * C::kStack);      // N = 128: [w_hi ; w_lo] rows
        auto mma_tf32 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc_) { if (PAIR) ptx::umma_tf32_2sm(d, a, b, id, acc_); else ptx::umma_tf32(d, a, b, id, acc_); };
        auto mma_f16 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc_) { if (PAIR) ptx::umma_f16_2sm(d, a, b, id, acc_); else ptx::umma_f16(d, a, b, id, acc_); };
        auto commit = [](uint64_t* bar) { if (PAIR) ptx::umma_commit_2sm(bar, 3); else ptx::umma_commit(bar); };
        // descriptor high word: SBO = 1024 B (>>4) | version 1 (bit 46) | SWIZZLE_128B (bits 61..63)
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        int astage = 0; uint32_t aphase = 0;
        int bstage = 0; uint32_t bphase = 0;
        int w_cur = -2; uint32_t w_gen = 0;        // RESIDENT: weight-set currently in shared memory
        int it = 0;
        for (int sidx = s_begin; sidx < s_end; ++it) {
            const Item im = item_at(sidx); sidx = im.next;
            const int tile = im.tile;
            if (RESIDENT) {
                const int wid = wid_of(tile);
                if (wid != w_cur) { ptx::mbar_wait(&b_full[0], w_gen & 1); ptx::tc_fence_after(); w_cur = wid; ++w_gen; if (it == 0 && lane == 0) trace_stamp(p.trace, 2); }
            }
            const int acc = it % C::kNAcc;
            const uint32_t acc_phase = (it / C::kNAcc) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
            uint32_t cnt = 0;                       // MMAs issued into this item so far (per M tile)
            for (int ch = im.c0; ch < im.c1; ++ch) {
#pragma unroll
                for (int u = 0; u < KT::NU; ++u) {
                    ptx::mbar_wait(&a_full[astage], aphase);
                    ptx::tc_fence_after();
                    if (it == 0 && ch == im.c0 && u == 0 && lane == 0) trace_stamp(p.trace, 3);
                    // low descriptor words: (addr >> 4) | LBO(=1) << 16; +2 per 32-byte K step, +8 per pixel row.
                    // base_offset stays 0: the 128B swizzle is a function of absolute smem address bits
                    // (profiles/r01_umma_desc_rowshift_probe.txt)
                    const uint32_t a_unit_lo = ((ptx::smem_u32(sA + astage * C::kAStage) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
                    for (int k = 0; k < KT::ntaps(u); ++k) {
                        uint32_t b_lo;
                        if (RESIDENT && C::kStack == 2) {     // one 128-row tile per tap; chunk ch sits 64 bytes (4 x 16 B) further along K
                            b_lo = (((ptx::smem_u32(sB + KT::wtap(u, k) * C::kBTile) & 0x3FFFFu) >> 4) | (1u << 16)) + ch * 4;
                        } else if (RESIDENT) {
                            b_lo = ((ptx::smem_u32(sB + (KT::wtap(u, k) * t.chunks + ch) * C::kWPerTap * C::kBTile) & 0x3FFFFu) >> 4) | (1u << 16);
                        } else {
                            ptx::mbar_wait(&b_full[bstage], bphase);
                            ptx::tc_fence_after();
                            if (it == 0 && ch == im.c0 && u == 0 && k == 0 && lane == 0) trace_stamp(p.trace, 2);
                            b_lo = ((ptx::smem_u32(sB + bstage * C::kBTile) & 0x3FFFFu) >> 4) | (1u << 16);
                        }
                        const uint32_t a_lo = a_unit_lo + KT::shift(u, k) * (kChunkBytes >> 4);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int j = 0; j < MT; ++j) {
                                const uint32_t aj = a_lo + j * (C::kAUnit >> 4);
                                const uint32_t dj = d_tmem + j * (BN * C::kSplit);
                                auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
                                // MMA number i of this tile goes to partial accumulator i % kSplit; its first visit zero-initialises
                                auto dst = [&](uint32_t i) { return dj + ((cnt + i) & (C::kSplit - 1)) * BN; };
                                auto accf = [&](uint32_t i) { return (cnt + i) >= static_cast<uint32_t>(C::kSplit) ? 1u : 0u; };
                                if (PREC == PREC_TF32) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        mma_tf32(dst(kk), desc(aj + 2 * kk), desc(b_lo + 2 * kk), idesc, accf(kk));
                                } else if (C::kStack == 2 && POOL) {
                                    // stem window = 8 pixels x [hi4|lo4] against rows [w_hi|w_hi ; w_lo|0]: all three products in one N = 128 MMA per K step
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        mma_f16(dj, desc(aj + 2 * kk), desc(b_lo + 2 * kk), idesc_stack, (cnt + kk) ? 1u : 0u);
                                } else if (C::kStack == 2) {
                                    // chunk = [32 hi | 32 lo] (A); weight rows [w_hi ; w_lo]: a_hi x both (N = 128), then a_lo x w_hi (N = 64, columns 0-63)
#pragma unroll
                                    for (int sl = 0; sl < 2; ++sl) {
                                        mma_f16(dj, desc(aj + 2 * sl), desc(b_lo + 2 * sl), idesc_stack, (cnt + sl) ? 1u : 0u);
                                        mma_f16(dj, desc(aj + 4 + 2 * sl), desc(b_lo + 2 * sl), idesc, 1u);
                                    }
                                } else if (POOL) {
                                    // stem window = 8 pixels x [hi4|lo4]: pass 0 against [w_hi|w_hi], pass 1 against [w_lo|0]
#pragma unroll
                                    for (int ps = 0; ps < 2; ++ps)
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            mma_f16(dst(ps * 4 + kk), desc(aj + 2 * kk), desc(b_lo + ps * (C::kBTile >> 4) + 2 * kk), idesc, accf(ps * 4 + kk));
                                } else if (PREC == PREC_BF16X3) {
                                    // chunk = [32 hi | 32 lo] bf16 (A) x [32 w_hi | 32 w_lo] (B); offsets in 16-byte units
                                    constexpr int AO[6] = {0, 2, 4, 6, 0, 2};      // hi, hi, lo, lo, hi, hi
                                    constexpr int BO[6] = {0, 2, 0, 2, 4, 6};      // w_hi x4,        w_lo x2
#pragma unroll
                                    for (int i = 0; i < 6; ++i)
                                        mma_f16(dst(i), desc(aj + AO[i]), desc(b_lo + BO[i]), idesc, accf(i));
                                } else {
                                    mma_f16(dst(0), desc(aj), desc(b_lo), idesc, accf(0));
                                    mma_f16(dst(1), desc(aj + 2), desc(b_lo + 2), idesc, accf(1));
                                }
                            }
                            if (!RESIDENT) commit(&b_empty[bstage]);
                        }
                        __syncwarp();
                        cnt += (PREC == PREC_TF32 || C::kStack == 2) ? 4u : (POOL ? 8u : (PREC == PREC_BF16X3 ? 6u : 2u));
                        if (!RESIDENT) { if (++bstage == C::kBStages) { bstage = 0; bphase ^= 1; } }
                    }
                    if (ptx::elect_one()) commit(&a_empty[astage]);
                    __syncwarp();
                    if (++astage == C::kAStages) { astage = 0; aphase ^= 1; }
                }
            }
            if (ptx::elect_one()) {
                commit(&tmem_full[acc]);
                // RESIDENT: the next tile uses other weights -> tell the loader when these MMAs have retired
                if (RESIDENT && tile + 1 < w_end && wid_of(tile + 1) != w_cur) commit(&b_empty[0]);
            }
            __syncwarp();
        }
        if (lane == 0) trace_stamp(p.trace, 4);
    } else if (warp >= 4) {
        // ============================== epilogue (8 warps) ==========================
        ptx::grid_dep_wait();                       // residual reads / output writes must follow the previous kernel
        const int ew = warp - 4;
        const int q = ew & 3;                       // TMEM lane quadrant (== warp % 4)
        const int half = ew >> 2;                   // which half of the BN columns
        constexpr int kCols = BN / 2;
        const int row = q * 32 + lane;
        int it = 0;
        if (!POOL) {
            const int box = t.bw * t.bh;
            const int pn = row / box;
            const int rem = row - pn * box;
            const int py = rem / t.bw;
            const int px = rem - py * t.bw;
            for (int sidx = s_begin; sidx < s_end; ++it) {
                const Item im = item_at(sidx); sidx = im.next;
                const int tile = im.tile;
                const int acc = it % C::kNAcc;
                const uint32_t acc_phase = (it / C::kNAcc) & 1;
                const WorkUnit wu = decode_work(tile_of(tile), m_units, t);
                // stream-K roles of a cut unit: its TAIL chunks are the first item of a CTA's range (dump the partial into this
                // CTA's slot), its HEAD chunks the last item of the previous CTA's range (add that partial, then the normal epilogue)
                const bool sk_dump = im.c0 > 0, sk_fix = im.c1 < chunks;
                // resident layers: this lane's four pixels and their residual pieces are known before the accumulator is --
                // issue the residual loads now so their latency hides behind the wait for the MMAs
                size_t rpix[4]; bool rvalid[4]; uint4 rres[4][2];
                if constexpr (!C::kEpiT) {
                    const TileCoord2 tc = decode2(wu.mp, t);
                    const int chr = wu.grp * g.cout + wu.n_tile * BN + half * kCols;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                           // row q*32 + 8i + lane/4
                        const int rw = q * 32 + 8 * i + (lane >> 2);
                        const int pn_ = rw / box, rem_ = rw - pn_ * box, py_ = rem_ / t.bw, px_ = rem_ - py_ * t.bw;
                        const int n = tc.n0 + pn_, y = tc.ty * t.bh + py_, x = tc.tx * t.bw + px_;
                        rvalid[i] = (pn_ < t.bn) && (n < g.n_img) && (y < g.Ho) && (x < g.Wo) && tile_valid(tile) && !(t.debug & 8);   // bit3: TMEM drain only
                        rpix[i] = (static_cast<size_t>(n) * g.Ho + y) * g.Wo + x;
                        rres[i][0] = make_uint4(0u, 0u, 0u, 0u); rres[i][1] = make_uint4(0u, 0u, 0u, 0u);
                        if (p.res && rvalid[i]) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix[i] * g.res_cstride + g.res_coff + chr);
                            if (PREC == PREC_TF32) { rres[i][0] = __ldg(rp + 2 * (lane & 3)); rres[i][1] = __ldg(rp + 2 * (lane & 3) + 1); }   // 8 fp32 words
                            else                   { rres[i][0] = __ldg(rp + (lane & 3));     rres[i][1] = __ldg(rp + 4 + (lane & 3)); }       // hi piece, lo piece
                        }
                    }
                }
                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, 5);
                if (t.debug & 4) {                  // timing experiment: free the accumulator at once, no epilogue work
                    ptx::tc_fence_before(); __syncwarp();
                    if (lane == 0) { if (leader) ptx::mbar_arrive(&tmem_empty[acc]); else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); }
                    continue;
                }
                if constexpr (C::kEpiT) {
                    // ---------------- transposed epilogue: lane = (pixel-in-group-of-4 sub, 16-byte column group grp) ----------------
                    float (*stg)[C::kEpiPitch] = reinterpret_cast<float (*)[C::kEpiPitch]>(sT + ew * C::kEpiWarpBytes);
                    int* rowtab = reinterpret_cast<int*>(sT + ew * C::kEpiWarpBytes + 32 * C::kEpiPitch * 4);
                    const int grp = lane & 7, sub = lane >> 3;
#pragma unroll 1
                    for (int j = 0; j < MT; ++j) {
                        const int m = wu.mp * MT + j;
                        if (m >= t.m_tiles) break;                           // odd tail (warp-uniform)
                        const TileCoord2 tc = decode2(m, t);
                        {
                            const int n = tc.n0 + pn, y = tc.ty * t.bh + py, x = tc.tx * t.bw + px;
                            const bool valid = (pn < t.bn) && (n < g.n_img) && (y < g.Ho) && (x < g.Wo);
                            __syncwarp();
                            rowtab[lane] = valid ? (n * g.Ho + y) * g.Wo + x : -1;   // pixel index of TMEM row q*32 + lane
                            __syncwarp();
                        }
                        const int ch0 = wu.grp * g.cout + wu.n_tile * BN + half * kCols;
                        const float* bias_base = p.img_wid ? p.gbias[p.img_wid[tc.n0] * kLayersPerSet] : p.bias;
                        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols + j * (BN * C::kSplit) + half * kCols;
                        // partial slot layout: [32-column block][16-byte piece][row] -> every load/store instruction covers 512 contiguous bytes
                        if (sk_dump) {
                            uint4* slot = reinterpret_cast<uint4*>(p.sk_part + static_cast<size_t>(blockIdx.x) * (kBlockM * BN));
#pragma unroll 1
                            for (int c0 = 0; c0 < kCols; c0 += 32) {
                                uint32_t r0[16], r1[16];
                                ptx::tmem_ld16(taddr + c0, r0);
                                ptx::tmem_ld16(taddr + c0 + 16, r1);
                                ptx::tmem_ld_wait();
                                uint4* sb = slot + static_cast<size_t>((half * kCols + c0) >> 5) * 8 * kBlockM + row;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    sb[k * kBlockM] = make_uint4(r0[4 * k], r0[4 * k + 1], r0[4 * k + 2], r0[4 * k + 3]);
                                    sb[(4 + k) * kBlockM] = make_uint4(r1[4 * k], r1[4 * k + 1], r1[4 * k + 2], r1[4 * k + 3]);
                                }
                            }
                            __threadfence();
                            continue;                                        // flag raised below, once all 8 warps have stored
                        }
                        const uint4* fix = nullptr;
                        if (sk_fix) {
                            if (lane == 0) while (ptx::ld_acquire_gpu(p.sk_flags + blockIdx.x + 1) != t.sk_seq) { }
                            __syncwarp();
                            fix = reinterpret_cast<const uint4*>(p.sk_part + static_cast<size_t>(blockIdx.x + 1) * (kBlockM * BN));
                        }
                        int pix[8];                                          // pixels this lane post-processes: rows 4k + sub (same for every block)
#pragma unroll
                        for (int k = 0; k < 8; ++k) pix[k] = rowtab[4 * k + sub];
#pragma unroll 1
                        for (int c0 = 0; c0 < kCols; c0 += 32) {
                            const int chan = ch0 + c0;                        // first channel (word index) of this 32-channel chunk
                            // residual pieces first: their L2 latency overlaps the TMEM load and the staging round trip below
                            float4 r4[8];                                    // tf32 storage: 4 fp32 words
                            uint2 rh[8], rl[8];                              // bf16 storage: 4 channels, hi at byte grp*8 of the chunk, lo 64 bytes further
                            if (p.res) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    if (PREC == PREC_TF32) {
                                        r4[k] = pix[k] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.res + static_cast<size_t>(pix[k]) * g.res_cstride + g.res_coff + chan + grp * 4))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                                    } else {
                                        rh[k] = make_uint2(0u, 0u); rl[k] = make_uint2(0u, 0u);
                                        if (pix[k] >= 0) {
                                            const uint8_t* cb = reinterpret_cast<const uint8_t*>(p.res + static_cast<size_t>(pix[k]) * g.res_cstride + g.res_coff + chan) + grp * 8;
                                            rh[k] = __ldg(reinterpret_cast<const uint2*>(cb));
                                            rl[k] = __ldg(reinterpret_cast<const uint2*>(cb + 64));
                                        }
                                    }
                                }
                            }
                            {
                                uint32_t r0[16], r1[16];
                                ptx::tmem_ld16(taddr + c0, r0);
                                ptx::tmem_ld16(taddr + c0 + 16, r1);
                                if (fix) {
                                    const uint4* fb = fix + static_cast<size_t>((half * kCols + c0) >> 5) * 8 * kBlockM + row;
                                    uint4 f[8];
#pragma unroll
                                    for (int k = 0; k < 8; ++k) f[k] = __ldcg(fb + k * kBlockM);
                                    ptx::tmem_ld_wait();
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        r0[4 * k]     = __float_as_uint(__uint_as_float(r0[4 * k])     + __uint_as_float(f[k].x));
                                        r0[4 * k + 1] = __float_as_uint(__uint_as_float(r0[4 * k + 1]) + __uint_as_float(f[k].y));
                                        r0[4 * k + 2] = __float_as_uint(__uint_as_float(r0[4 * k + 2]) + __uint_as_float(f[k].z));
                                        r0[4 * k + 3] = __float_as_uint(__uint_as_float(r0[4 * k + 3]) + __uint_as_float(f[k].w));
                                        r1[4 * k]     = __float_as_uint(__uint_as_float(r1[4 * k])     + __uint_as_float(f[4 + k].x));
                                        r1[4 * k + 1] = __float_as_uint(__uint_as_float(r1[4 * k + 1]) + __uint_as_float(f[4 + k].y));
                                        r1[4 * k + 2] = __float_as_uint(__uint_as_float(r1[4 * k + 2]) + __uint_as_float(f[4 + k].z));
                                        r1[4 * k + 3] = __float_as_uint(__uint_as_float(r1[4 * k + 3]) + __uint_as_float(f[4 + k].w));
                                    }
                                }
                                ptx::tmem_ld_wait();
                                __syncwarp();                                // previous block's readers are done with stg
#pragma unroll
                                for (int jj = 0; jj < 16; jj += 4) {
                                    *reinterpret_cast<uint4*>(&stg[lane][jj]) = make_uint4(r0[jj], r0[jj + 1], r0[jj + 2], r0[jj + 3]);
                                    *reinterpret_cast<uint4*>(&stg[lane][16 + jj]) = make_uint4(r1[jj], r1[jj + 1], r1[jj + 2], r1[jj + 3]);
                                }
                            }
                            __syncwarp();
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_base + chan + grp * 4));
                            float4 a4[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                a4[k] = *reinterpret_cast<const float4*>(&stg[4 * k + sub][grp * 4]);
                                a4[k].x += b4.x; a4[k].y += b4.y; a4[k].z += b4.z; a4[k].w += b4.w;
                            }
                            if (p.res) {
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
                                if (g.act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                                else if (g.act == ACT_SELU) { o.x = selu_fast(o.x); o.y = selu_fast(o.y); o.z = selu_fast(o.z); o.w = selu_fast(o.w); }
                                if (pix[k] < 0) continue;
                                if (p.pool_part) { psum.x += o.x; psum.y += o.y; psum.z += o.z; psum.w += o.w; continue; }
                                float* po = p.out + static_cast<size_t>(pix[k]) * g.out_cstride + g.out_coff + chan;
                                if (PREC == PREC_TF32) {
                                    if (g.round_tf32) o = make_float4(ptx::to_tf32(o.x), ptx::to_tf32(o.y), ptx::to_tf32(o.z), ptx::to_tf32(o.w));
                                    *reinterpret_cast<float4*>(po + grp * 4) = o;
                                } else {
                                    uint32_t h0, l0, h1, l1;
                                    split2(o.x, o.y, h0, l0); split2(o.z, o.w, h1, l1);
                                    uint8_t* cb = reinterpret_cast<uint8_t*>(po) + grp * 8;
                                    *reinterpret_cast<uint2*>(cb) = make_uint2(h0, h1);
                                    *reinterpret_cast<uint2*>(cb + 64) = make_uint2(l0, l1);
                                }
                            }
                            if (p.pool_part) {                                // rows 4k + sub summed above; fold the four `sub` groups (fixed order: deterministic)
#pragma unroll
                                for (int off = 8; off <= 16; off <<= 1) {
                                    psum.x += __shfl_xor_sync(0xffffffffu, psum.x, off); psum.y += __shfl_xor_sync(0xffffffffu, psum.y, off);
                                    psum.z += __shfl_xor_sync(0xffffffffu, psum.z, off); psum.w += __shfl_xor_sync(0xffffffffu, psum.w, off);
                                }
                                if (sub == 0)
                                    *reinterpret_cast<float4*>(p.pool_part + (static_cast<size_t>(tc.n0) * 4 + q) * g.out_cstride + g.out_coff + chan + grp * 4) = psum;
                            }
                        }
                    }
                } else {
                    // ---------------- resident 64-channel layers: 16x256b TMEM loads ----------------
                    // Lane (R = lane/4, m = lane%4) receives, for each of its four rows q*32 + 8i + R, the accumulator columns
                    // 8j + 2m + e (j < 4, e < 2) of this warp's 32-column block.  The weight rows of these layers are stored
                    // permuted (column 8j + 2m + e carries output channel 8m + 2j + e, se3tn.cu), so the lane owns the 8
                    // CONSECUTIVE channels 8m .. 8m+7 of each pixel: one 16-byte hi and one 16-byte lo piece (or 32 bytes of fp32)
                    // per row, four lanes complete 64 (128) contiguous bytes, and a load/store instruction touches 8 lines
                    // instead of the 32 of a row-per-lane epilogue.
                    static_assert(MT == 1 && C::kSplit == 1, "resident epilogue: one tile, one accumulator");
                    const int mm = lane & 3;
                    const TileCoord2 tc = decode2(wu.mp, t);
                    const int ch0 = wu.grp * g.cout + wu.n_tile * BN + half * kCols;
                    const float* bias_base = (p.img_wid ? p.gbias[p.img_wid[tc.n0] * kLayersPerSet] : p.bias) + ch0 + mm * 8;
                    const float4 bA = __ldg(reinterpret_cast<const float4*>(bias_base)), bB = __ldg(reinterpret_cast<const float4*>(bias_base + 4));
                    const float bias8[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
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
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                           // row q*32 + 8i + Rr; register 4j + 2(i&1) + e of ra (i < 2) / rb
                        if (!rvalid[i]) continue;
                        const size_t pix = rpix[i];
                        float v[8];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                v[2 * jj + e] = __uint_as_float((i < 2 ? ra : rb)[4 * jj + 2 * (i & 1) + e]) + bias8[2 * jj + e];
                        if (p.res) {
                            if (PREC == PREC_TF32) {
                                const uint4 r0 = rres[i][0], r1 = rres[i][1];
                                v[0] += __uint_as_float(r0.x); v[1] += __uint_as_float(r0.y); v[2] += __uint_as_float(r0.z); v[3] += __uint_as_float(r0.w);
                                v[4] += __uint_as_float(r1.x); v[5] += __uint_as_float(r1.y); v[6] += __uint_as_float(r1.z); v[7] += __uint_as_float(r1.w);
                            } else {                                         // chunk = [32 bf16 hi | 32 bf16 lo]
                                const uint4 h4 = rres[i][0], l4 = rres[i][1];
                                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 hf = unpack2(hw[e]), lf = unpack2(lw[e]);
                                    v[2 * e] += hf.x + lf.x; v[2 * e + 1] += hf.y + lf.y;
                                }
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (g.act == ACT_RELU) v[e] = fmaxf(v[e], 0.f);
                            else if (g.act == ACT_SELU) v[e] = selu_fast(v[e]);
                        }
                        float* outp = p.out + pix * g.out_cstride + g.out_coff + ch0;
                        if (PREC == PREC_TF32) {
                            float4 o0 = make_float4(v[0], v[1], v[2], v[3]), o1 = make_float4(v[4], v[5], v[6], v[7]);
                            if (g.round_tf32) {
                                o0 = make_float4(ptx::to_tf32(o0.x), ptx::to_tf32(o0.y), ptx::to_tf32(o0.z), ptx::to_tf32(o0.w));
                                o1 = make_float4(ptx::to_tf32(o1.x), ptx::to_tf32(o1.y), ptx::to_tf32(o1.z), ptx::to_tf32(o1.w));
                            }
                            *reinterpret_cast<float4*>(outp + mm * 8) = o0;
                            *reinterpret_cast<float4*>(outp + mm * 8 + 4) = o1;
                        } else {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
                            reinterpret_cast<uint4*>(outp)[mm] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            reinterpret_cast<uint4*>(outp)[4 + mm] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (leader) ptx::mbar_arrive(&tmem_empty[acc]); else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); }
                if (C::kEpiT && sk_dump) {                                   // all 8 epilogue warps have stored (and fenced) their part of the partial
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 128) ptx::st_release_gpu(p.sk_flags + blockIdx.x, t.sk_seq);
                }
            }
        } else {
            // ---- stem: conv tile 11x11 -> 5x5 max-pooled outputs (MaxPool2d(3,2,1), -inf padding) ----
            const int cy_l = row / 11, cx_l = row - cy_l * 11;          // conv position inside the tile
            const int et = threadIdx.x - 128;                           // 0..255 among epilogue threads
            for (int tile = w_begin; tile < w_end; ++tile, ++it) {
                const int acc = it % C::kNAcc;
                const uint32_t acc_phase = (it / C::kNAcc) & 1;
                const WorkUnit wu = decode_work(tile_of(tile), m_units, t);
                const TileCoord2 tc = decode2(wu.mp, t);
                float* stage = reinterpret_cast<float*>(sP + (C::kPoolBufs == 2 ? (it & 1) : 0) * ((kPoolStageBytes + 1023) & ~1023));
                const int cy = tc.oy + cy_l, cx = tc.ox + cx_l;             // conv output coordinates
                const bool cvalid = (row < 121) && cy >= 0 && cy < 88 && cx >= 0 && cx < 88;

                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                if (it == 0 && threadIdx.x == 128) trace_stamp(p.trace, 5);
                if (t.debug & 4) {                  // timing experiment: free the accumulator at once, no epilogue work
                    ptx::tc_fence_before(); __syncwarp();
                    if (lane == 0) { if (leader) ptx::mbar_arrive(&tmem_empty[acc]); else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); }
                    continue;
                }
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols + half * kCols;
                {
                    float v[32];
                    {
                        uint32_t r0[16], r1[16];
                        ptx::tmem_ld16(taddr, r0);
                        ptx::tmem_ld16(taddr + 16, r1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) { v[jj] = __uint_as_float(r0[jj]); v[16 + jj] = __uint_as_float(r1[jj]); }
                    }
#pragma unroll
                    for (int sp = 1; sp < C::kSplit * C::kStack; ++sp) {
                        uint32_t r0[16], r1[16];
                        ptx::tmem_ld16(taddr + sp * BN, r0);
                        ptx::tmem_ld16(taddr + sp * BN + 16, r1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) { v[jj] += __uint_as_float(r0[jj]); v[16 + jj] += __uint_as_float(r1[jj]); }
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
                if (lane == 0) { if (leader) ptx::mbar_arrive(&tmem_empty[acc]); else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); }   // accumulator drained
                asm volatile("bar.sync 1, 256;" ::: "memory");             // staging tile complete (epilogue warps only)

                // 25 pooled pixels x 16 float4 channel groups = 400 vectors over 256 threads
                for (int v = et; v < 400; v += 256) {
                    const int pp = v >> 4, c4 = (v & 15) * 4;
                    const int ppy = pp / 5, ppx = pp - ppy * 5;
                    const int oy = tc.ty * 5 + ppy, ox = tc.tx * 5 + ppx;     // pooled output coordinates
                    if (oy >= g.Ho || ox >= g.Wo || (t.debug & 8)) continue;
                    float4 m = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float4 s4 = *reinterpret_cast<const float4*>(stage + ((2 * ppy + dy) * 11 + 2 * ppx + dx) * kPoolPitch + c4);
                            m.x = fmaxf(m.x, s4.x); m.y = fmaxf(m.y, s4.y); m.z = fmaxf(m.z, s4.z); m.w = fmaxf(m.w, s4.w);
                        }
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>((p.img_wid ? p.gbias[p.img_wid[tc.n0] * kLayersPerSet] : p.bias) + c4));
                    const float v0 = selu_fast(m.x + b4.x), v1 = selu_fast(m.y + b4.y), v2 = selu_fast(m.z + b4.z), v3 = selu_fast(m.w + b4.w);
                    const int n = tc.n0;
                    if (n >= g.n_img || !tile_valid(tile)) continue;
                    float* po = p.out + ((static_cast<size_t>(n) * g.Ho + oy) * g.Wo + ox) * g.out_cstride + g.out_coff;
                    if (PREC == PREC_TF32) {
                        float4 o = make_float4(v0, v1, v2, v3);
                        if (g.round_tf32) o = make_float4(ptx::to_tf32(o.x), ptx::to_tf32(o.y), ptx::to_tf32(o.z), ptx::to_tf32(o.w));
                        *reinterpret_cast<float4*>(po + c4) = o;
                    } else {
                        // channel c4..c4+3 of chunk c4/32: hi at byte (c4%32)*2, lo 64 bytes further
                        uint32_t h0, l0, h1, l1;
                        split2(v0, v1, h0, l0); split2(v2, v3, h1, l1);
                        uint8_t* cb = reinterpret_cast<uint8_t*>(po + (c4 & ~31)) + (c4 & 31) * 2;
                        *reinterpret_cast<uint2*>(cb) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(cb + 64) = make_uint2(l0, l1);
                    }
                }
                if (C::kPoolBufs == 1) asm volatile("bar.sync 1, 256;" ::: "memory");   // single staging buffer: readers done before the next tile writes
            }
        }
    }

    if (threadIdx.x == 128) trace_stamp(p.trace, 6);
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    if (threadIdx.x == 0 && p.trace) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); p.trace[blockIdx.x * 8 + 7] = (gtimer() & ~0xffull) | (smid & 0xff); }
    if (warp == 2) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_2sm(tmem_base, C::kTmemCols); else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

template <int KIND>
bool plan_matches(const Umma2Plan& t) {
    using KT = KTab<KIND>;
    if (t.units_per_chunk != KT::NU) return false;
    for (int u = 0; u < KT::NU; ++u) {
        if (t.units[u].ntaps != KT::ntaps(u)) return false;
        for (int k = 0; k < KT::ntaps(u); ++k)
            if (t.units[u].taps[k].row_shift != KT::shift(u, k) || t.units[u].taps[k].w_tap != KT::wtap(u, k)) return false;
    }
    return true;
}

template <int BN, bool RESIDENT, int KIND, int MT, int PREC, bool PAIR = false>
cudaError_t launch2(const UmmaMaps& maps, const ConvGeom& g, const Umma2Plan& t, const ConvPtrs& p, int num_sms, cudaStream_t stream) {
    using C = Cfg2<BN, RESIDENT, KIND, MT, PREC, PAIR>;
    if (!plan_matches<KIND>(t)) return cudaErrorInvalidValue;
    if (PAIR && (p.img_wid || t.n_tiles != 1 || g.groups != 1)) return cudaErrorInvalidValue;
    if (p.pool_part && (!C::kEpiT || MT != 1 || t.bn != 1 || t.tiles_x != 1 || t.tiles_y != 1)) return cudaErrorInvalidValue;   // fused avg-pool: tile = whole image
    if (C::kStack == 2 && KIND == KIND_S1 && t.chunks != 2) return cudaErrorInvalidValue;    // a stacked tile row is exactly two 32-channel chunks
    const int w_tiles = C::kStack == 2 ? g.num_taps : g.num_taps * t.chunks * C::kWPerTap;
    const size_t smem = static_cast<size_t>(C::kAStages) * C::kAStage + static_cast<size_t>(RESIDENT ? w_tiles : C::kBStages) * C::kBTile +
                        C::kPoolBufs * ((kPoolStageBytes + 1023) & ~1023) + ((C::kEpiBytes + 1023) & ~1023) + 1024 + 512;
    if (smem > 232448) return cudaErrorInvalidConfiguration;
    // the dynamic shared-memory limit is a per-device function attribute: cache what was set for each device
    static size_t attr_smem_dev[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    size_t& attr_smem = attr_smem_dev[dev];
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma2_kernel<BN, RESIDENT, KIND, MT, PREC, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    const int total = ((t.m_tiles + MT - 1) / MT) * t.n_tiles * g.groups;
    int grid = total < num_sms ? total : num_sms;
    if (PAIR) { const int pairs = std::min((total + 1) / 2, num_sms / 2); grid = 2 * pairs; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
    if (t.pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, conv_umma2_kernel<BN, RESIDENT, KIND, MT, PREC, PAIR>, maps, g, t, p);
}

template <int PREC>
cudaError_t dispatch2(const UmmaMaps& maps, const ConvGeom& g, const Umma2Plan& t, const ConvPtrs& p,
                      int block_n, bool resident, int kind, int m_per_cta, int num_sms, cudaStream_t stream) {
    if (kind == KIND_STEM) {
        if (block_n != 64 || !resident) return cudaErrorInvalidValue;
        return t.pair ? launch2<64, true, KIND_STEM, 1, PREC, true>(maps, g, t, p, num_sms, stream) : launch2<64, true, KIND_STEM, 1, PREC>(maps, g, t, p, num_sms, stream);
    }
    if (resident) {
        if (block_n != 64 || kind != KIND_S1) return cudaErrorInvalidValue;
        return t.pair ? launch2<64, true, KIND_S1, 1, PREC, true>(maps, g, t, p, num_sms, stream) : launch2<64, true, KIND_S1, 1, PREC>(maps, g, t, p, num_sms, stream);
    }
    if (block_n != 256) return cudaErrorInvalidValue;
    if (kind == KIND_S1)
        return m_per_cta == 2 ? launch2<256, false, KIND_S1, 2, PREC>(maps, g, t, p, num_sms, stream)
                              : launch2<256, false, KIND_S1, 1, PREC>(maps, g, t, p, num_sms, stream);
    return m_per_cta == 2 ? launch2<256, false, KIND_S2, 2, PREC>(maps, g, t, p, num_sms, stream)
                          : launch2<256, false, KIND_S2, 1, PREC>(maps, g, t, p, num_sms, stream);
}

}  // namespace

cudaError_t launch_conv_umma2(const UmmaMaps& maps, const ConvGeom& g, const Umma2Plan& t, const ConvPtrs& p,
                              int block_n, bool resident, int kind, int m_per_cta, int prec, int num_sms, cudaStream_t stream) {
    switch (prec) {
        case PREC_TF32:   return dispatch2<PREC_TF32>(maps, g, t, p, block_n, resident, kind, m_per_cta, num_sms, stream);
        case PREC_BF16X3: return dispatch2<PREC_BF16X3>(maps, g, t, p, block_n, resident, kind, m_per_cta, num_sms, stream);
        case PREC_BF16:   return dispatch2<PREC_BF16>(maps, g, t, p, block_n, resident, kind, m_per_cta, num_sms, stream);
        default:          return cudaErrorInvalidValue;
    }
}

}  // namespace se3tn

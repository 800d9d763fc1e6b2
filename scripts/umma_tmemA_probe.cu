// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (weights-stationary formulation for the stem):
// D[m][n] = sum_k A[m][k] * B[n][k], A (M = 128 x K = 64 bf16) written into TMEM with tcgen05.st (thread = row, 2 bf16 per
// 32-bit column), B (N = 128 x K = 64 bf16) in shared memory by TMA (K-major, SWIZZLE_128B).  Which packing / column
// stride per K = 16 step does the hardware expect?
#include "../iros20-6d-pose-tracking_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

constexpr int M = 128, N = 128, K = 64;

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
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// mode bit0: swap the two bf16 inside a column; col_step: TMEM columns A advances per K = 16 step
__global__ void probe(const __grid_constant__ CUtensorMap mapB, const __nv_bfloat16* __restrict__ A, int swap_pack, int col_step, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_barrier_init(); ptx::fence_proxy_async(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t a_tmem = tmem + 128;            // D: columns [0,128); A: columns [128, 160)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const int row = warp * 32 + lane;
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) {
            const unsigned short e0 = __bfloat16_as_ushort(A[row * K + 2 * j]), e1 = __bfloat16_as_ushort(A[row * K + 2 * j + 1]);
            r[j] = swap_pack ? (static_cast<uint32_t>(e0) << 16 | e1) : (static_cast<uint32_t>(e1) << 16 | e0);
        }
        tmem_st32(a_tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
        tmem_st_wait();
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(&bar[0], N * 128);
        ptx::tma_load_2d(sB, &mapB, &bar[0], 0, 0);
        ptx::mbar_wait(&bar[0], 0);
        ptx::tc_fence_after();
        const uint32_t b0 = ptx::smem_u32(sB);
        constexpr uint32_t idesc = ptx::umma_idesc(1, 128, N);
        for (int kk = 0; kk < 4; ++kk)
            umma_f16_ts(tmem, a_tmem + kk * col_step, ptx::umma_desc_sw128(b0 + kk * 32), idesc, kk != 0);
        ptx::umma_commit(&bar[1]);
    }
    ptx::mbar_wait(&bar[1], 0);
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> A(M * K), B(N * K);
    std::vector<float> Af(M * K), Bf(N * K);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { Af[r * K + k] = float(((r * 7 + k * 3) % 13) - 6); A[r * K + k] = __float2bfloat16(Af[r * K + k]); }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { Bf[n * K + k] = float(((n * 5 + k) % 7) - 3); B[n * K + k] = __float2bfloat16(Bf[n * K + k]); }
    __nv_bfloat16 *dA, *dB; float* dOut;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dOut, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    CUtensorMap mB;
    cuuint64_t dimsB[2] = {K / 2, N}, str[1] = {K * 2};          // as 32-bit words: 32 words = 128 B per row
    cuuint32_t boxB[2] = {K / 2, N}, es[2] = {1, 1};
    CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r2);
    const int smem = N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(M * N);
    const int steps[] = {8, 16, 4};
    for (int swap_pack = 0; swap_pack < 2; ++swap_pack)
        for (int cs : steps) {
            cudaMemset(dOut, 0, M * N * 4);
            probe<<<1, 128, smem>>>(mB, dA, swap_pack, cs, dOut);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swap %d col_step %d: CUDA error %s\n", swap_pack, cs, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int i = 0; i < M; ++i) for (int n = 0; n < N; ++n) {
                float ref = 0; for (int k = 0; k < K; ++k) ref += Af[i * K + k] * Bf[n * K + k];
                if (ref != out[i * N + n]) ++bad;
            }
            printf("A in TMEM: pack %s, %2d columns per K=16 step: %s (%d mismatches; out[0][0..3] = %g %g %g %g)\n",
                   swap_pack ? "(k even -> HIGH half)" : "(k even -> low half)", cs, bad ? "MISMATCH" : "exact", bad, out[0], out[1], out[2], out[3]);
        }
    return 0;
}

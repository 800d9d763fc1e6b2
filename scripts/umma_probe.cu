// Probe: does a K-major SWIZZLE_128B UMMA smem descriptor accept a start address shifted by a
// whole number of 128-byte rows (not 1024-aligned), and does it need the base_offset field?
// D[i][n] = sum_k A[i + shift][k] * B[n][k], A tile written by TMA (absolute-address swizzle).
#include "../iros20-6d-pose-tracking_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int R = 192, KC = 32, N = 64;

__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t base_off) {
    return ptx::umma_desc_sw128(addr) | (static_cast<uint64_t>(base_off & 7) << 49);
}

__global__ void probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      int shift, int bo_mode, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + R * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_barrier_init(); ptx::fence_proxy_async(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(&bar[0], R * 128 + N * 128);
        ptx::tma_load_2d(sA, &mapA, &bar[0], 0, 0);
        ptx::tma_load_2d(sB, &mapB, &bar[0], 0, 0);
        ptx::mbar_wait(&bar[0], 0);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(sA) + shift * 128, b0 = ptx::smem_u32(sB);
        const uint32_t bo = bo_mode == 0 ? 0 : (bo_mode == 1 ? ((a0 >> 7) & 7) : ((8 - ((a0 >> 7) & 7)) & 7));
        constexpr uint32_t idesc = ptx::umma_idesc(2, 128, N);
        for (int kk = 0; kk < 4; ++kk)
            ptx::umma_tf32(tmem, desc_bo(a0 + kk * 32, bo), desc_bo(b0 + kk * 32, 0), idesc, kk != 0);
        ptx::umma_commit(&bar[1]);
    }
    ptx::mbar_wait(&bar[1], 0);
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<float> A(R * KC), B(N * KC);
    for (int r = 0; r < R; ++r) for (int k = 0; k < KC; ++k) A[r * KC + k] = float(((r * 7 + k * 3) % 13) - 6);
    for (int n = 0; n < N; ++n) for (int k = 0; k < KC; ++k) B[n * KC + k] = float(((n * 5 + k) % 7) - 3);
    float *dA, *dB, *dOut;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dOut, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    CUtensorMap mA, mB;
    cuuint64_t dimsA[2] = {KC, R}, dimsB[2] = {KC, N}, str[1] = {KC * 4};
    cuuint32_t boxA[2] = {KC, R}, boxB[2] = {KC, N}, es[2] = {1, 1};
    CUresult r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d %d\n", (int)r1, (int)r2);
    const int smem = R * 128 + N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(128 * N);
    const int shifts[] = {0, 1, 3, 8, 11, 22, 33, 44};
    for (int bo_mode = 0; bo_mode < 3; ++bo_mode)
        for (int shift : shifts) {
            cudaMemset(dOut, 0, 128 * N * 4);
            probe<<<1, 128, smem>>>(mA, mB, shift, bo_mode, dOut);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("bo_mode %d shift %d: CUDA error %s\n", bo_mode, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, first_bad = -1;
            for (int i = 0; i < 128; ++i) for (int n = 0; n < N; ++n) {
                float ref = 0; for (int k = 0; k < KC; ++k) ref += A[(i + shift) * KC + k] * B[n * KC + k];
                if (ref != out[i * N + n]) { if (first_bad < 0) first_bad = i; ++bad; }
            }
            printf("bo_mode %d (%s) shift %2d rows: %s (%d mismatches, first bad row %d)\n", bo_mode,
                   bo_mode == 0 ? "base_offset=0" : bo_mode == 1 ? "base_offset=(addr>>7)&7" : "base_offset=8-((addr>>7)&7)",
                   shift, bad ? "MISMATCH" : "exact", bad, first_bad);
        }
    return 0;
}

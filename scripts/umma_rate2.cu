// Issue-cost micro-benchmark, second cut: fully unrolled groups of G tcgen05.mma with precomputed descriptor
// words (no per-MMA integer work), optionally separated by the things the real K loop does between groups:
//   mode 0: nothing              mode 1: tcgen05.commit to a spare mbarrier after every group
//   mode 2: + mbarrier try_wait on an already-completed barrier before every group (+ tcgen05 fence)
#include "../iros20-6d-pose-tracking_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int KIND, int G>
__global__ void rate(int groups, int mode, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + 192 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);   // [0] final, [1] spare commit target, [2] pre-completed
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    for (int i = threadIdx.x; i < (192 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1 << 20); ptx::mbar_init(&bar[2], 1);
        ptx::fence_barrier_init(); ptx::fence_proxy_async();
        ptx::mbar_arrive(&bar[2]);                 // phase 0 of bar[2] completes: waiting on parity 0 succeeds at once
    }
    if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = ptx::umma_idesc(KIND == 0 ? 2u : 1u, 128, N);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const uint32_t a_lo = ((ptx::smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo = ((ptx::smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
        __syncwarp();
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (mode >= 2) { ptx::mbar_wait(&bar[2], 0); ptx::tc_fence_after(); }
            if (ptx::elect_one()) {
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const uint64_t ad = (static_cast<uint64_t>(kHi) << 32) | (a_lo + 2 * (i & 3));
                    const uint64_t bd = (static_cast<uint64_t>(kHi) << 32) | (b_lo + 2 * (i & 3));
                    const uint32_t d = tmem + ((N <= 256 && (i & 1)) ? (512 - N >= N ? N : 0) : 0);
                    if (KIND == 0) ptx::umma_tf32(d, ad, bd, idesc, 1u); else ptx::umma_f16(d, ad, bd, idesc, 1u);
                }
                if (mode >= 1) ptx::umma_commit(&bar[1]);
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(&bar[0]);
        __syncwarp();
        ptx::mbar_wait(&bar[0], 0);
        t1 = clock64();
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int N, int KIND, int G>
void run(const char* name, long long* d_out) {
    const int smem = (192 + 256) * 128 + 2048;
    cudaFuncSetAttribute(rate<N, KIND, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int total = 6144;
    for (int mode = 0; mode < 3; ++mode) {
        rate<N, KIND, G><<<148, 128, smem>>>(total / G, mode, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        const double per = double(cyc) / total;
        printf("%-5s N=%3d group=%2d mode=%d (%s): %6.1f cycles/MMA (tensor floor %d) %s\n", name, N, G, mode,
               mode == 0 ? "bare" : mode == 1 ? "commit/group" : "try_wait+fence+commit/group", per, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    run<64, 1, 6>("bf16", d_out);  run<64, 1, 12>("bf16", d_out); run<64, 1, 24>("bf16", d_out);
    run<128, 1, 6>("bf16", d_out); run<128, 1, 12>("bf16", d_out);
    run<256, 1, 6>("bf16", d_out); run<256, 1, 12>("bf16", d_out); run<256, 1, 24>("bf16", d_out);
    run<64, 0, 4>("tf32", d_out);  run<64, 0, 12>("tf32", d_out);
    run<256, 0, 4>("tf32", d_out); run<256, 0, 8>("tf32", d_out);
    return 0;
}

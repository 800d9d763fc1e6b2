// cta_group::2 probe: cycles per tcgen05.mma.cta_group::2 (M=256 across a CTA pair) vs N, issued by the leader CTA.
// Operands are whatever is in shared memory (finite bf16 patterns); only timing matters.
#include "../iros20-6d-pose-tracking_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(mask) : "memory");
}

template <int N, int G>
__global__ void __cluster_dims__(2, 1, 1) rate2(int groups, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + 192 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < (192 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::fence_barrier_init(); ptx::fence_proxy_async(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before(); cluster.sync(); ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = ptx::umma_idesc(1u, 256, N);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        if (rank == 0) {
            const uint32_t a_lo = ((ptx::smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = ((ptx::smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
            __syncwarp();
            t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        const uint64_t ad = (static_cast<uint64_t>(kHi) << 32) | (a_lo + 2 * (i & 3));
                        const uint64_t bd = (static_cast<uint64_t>(kHi) << 32) | (b_lo + 2 * (i & 3));
                        umma2_f16(tmem + ((i & 1) ? (N <= 256 ? N : 0) : 0), ad, bd, idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if (ptx::elect_one()) commit2(&bar[0], 3);
            __syncwarp();
        }
        ptx::mbar_wait(&bar[0], 0);        // both CTAs: the multicast commit arrives on each CTA's barrier
        t1 = clock64();
    }
    ptx::tc_fence_before(); cluster.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int G> void run(long long* d_out) {
    const int smem = (192 + 256) * 128 + 2048;
    cudaFuncSetAttribute(rate2<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int total = 6144;
    rate2<N, G><<<148, 128, smem>>>(total / G, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const double per = double(cyc) / total;
    printf("bf16 cta_group::2 M=256 N=%3d group=%2d: %6.1f cycles/MMA -> %6.0f MAC/clk per SM (peak 4096) %s\n", N, G, per,
           256.0 * N * 16 / per / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    run<64, 6>(d_out); run<64, 12>(d_out); run<64, 24>(d_out);
    run<128, 6>(d_out); run<128, 12>(d_out);
    run<256, 6>(d_out); run<256, 12>(d_out);
    return 0;
}

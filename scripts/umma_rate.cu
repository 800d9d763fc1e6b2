// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M=128, operands in shared memory, SWIZZLE_128B
// K-major) as a function of N, kind (tf32 K=8 / bf16 K=16), accumulator rotation and A row shift.
// One CTA per SM on all SMs (so smem/tensor contention is per-SM only), one elected lane issues.
#include "../iros20-6d-pose-tracking_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int KIND /*0 tf32, 1 bf16*/>
__global__ void rate(int iters, int rot, int shift_rows, int whole_warp, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 192 rows x 128 B
    uint8_t* sB = smem + 192 * 128;     // N rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (192 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // finite in both kinds
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); ptx::fence_proxy_async(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = ptx::umma_idesc(KIND == 0 ? 2u : 1u, 128, N);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const uint32_t a_lo = (((ptx::smem_u32(sA) + shift_rows * 128) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo = ((ptx::smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
        const int nacc = 512 / N < rot ? 512 / N : rot;
        __syncwarp();
        t0 = clock64();
        if (whole_warp) {
            for (int i = 0; i < iters; i += 4) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t d = tmem + ((i + kk) % nacc) * N;
                        const uint64_t ad = (static_cast<uint64_t>(kHi) << 32) | (a_lo + 2 * kk), bd = (static_cast<uint64_t>(kHi) << 32) | (b_lo + 2 * kk);
                        if (KIND == 0) ptx::umma_tf32(d, ad, bd, idesc, 1u); else ptx::umma_f16(d, ad, bd, idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if (ptx::elect_one()) ptx::umma_commit(bar);
            __syncwarp();
        } else if (threadIdx.x == 0) {
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t d = tmem + ((i + kk) % nacc) * N;
                    const uint64_t ad = (static_cast<uint64_t>(kHi) << 32) | (a_lo + 2 * kk), bd = (static_cast<uint64_t>(kHi) << 32) | (b_lo + 2 * kk);
                    if (KIND == 0) ptx::umma_tf32(d, ad, bd, idesc, 1u); else ptx::umma_f16(d, ad, bd, idesc, 1u);
                }
            }
            ptx::umma_commit(bar);
        }
        ptx::mbar_wait(bar, 0);
        t1 = clock64();
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int N, int KIND>
void run(const char* name, long long* d_out) {
    const int smem = (192 + 256) * 128 + 2048;
    cudaFuncSetAttribute(rate<N, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    for (int whole = 1; whole >= 0; --whole)
        for (int rot : {1, 2, 4})
            for (int shift : {0, 11}) {
                rate<N, KIND><<<148, 128, smem>>>(iters, rot, shift, whole, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
                const double per = double(cyc) / iters;
                const double macs = 128.0 * N * (KIND == 0 ? 8 : 16);
                printf("%-5s N=%3d %s rot=%d shift=%2d: %7.1f cycles/MMA  -> %6.0f MAC/clk/SM (ideal %d)  %s\n", name, N, whole ? "warp-uniform" : "lane0-only  ", rot, shift,
                       per, macs / per, KIND == 0 ? 2048 : 4096, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    run<64, 1>("bf16", d_out); run<128, 1>("bf16", d_out); run<256, 1>("bf16", d_out);
    run<64, 0>("tf32", d_out); run<128, 0>("tf32", d_out); run<256, 0>("tf32", d_out);
    return 0;
}

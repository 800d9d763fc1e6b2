// Pose-error metrics on the GPU -- the first row of SURVEY.md 8(f) ("next"): what the reference's evaluation
// scripts compute on the CPU with open3d + scipy for every key-frame pose (eval_ycb.py:96-106).
//   ADD   (reference Utils.py:72-82):  mean_i || (R_p x_i + t_p) - (R_g x_i + t_g) ||
//   ADD-S (reference Utils.py:84-98):  mean_i min_j || (R_g x_i + t_g) - (R_p x_j + t_p) ||   (cKDTree, k=1)
//   VOCap (reference eval_ycb.py:45-64): area under the accuracy-threshold curve below 0.1 m, x10
// float64 throughout, no FMA contraction (-fmad=false) so distances match numpy/scipy to the last few ulps.
// The nearest-neighbour search is exhaustive (m^2 distance evaluations per pose, a few 1e6): exact by
// construction, so there is no kd-tree to mirror.
#include "metrics.h"
#include "ptx.cuh"
#include <cub/cub.cuh>

namespace se3tn {

namespace {
constexpr int kMetricThreads = 256;
constexpr int kPredTile = 512;

__device__ __forceinline__ void xform(const double* T, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = T[0] * x + T[1] * y + T[2] * z + T[3];
    oy = T[4] * x + T[5] * y + T[6] * z + T[7];
    oz = T[8] * x + T[9] * y + T[10] * z + T[11];
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    if (threadIdx.x == 0) for (int w = 0; w < kMetricThreads / 32; ++w) s += red[w];
    __syncthreads();
    return s;                       // valid in thread 0
}

__global__ void __launch_bounds__(kMetricThreads)
add_adi_kernel(const double* __restrict__ model, int m, const double* __restrict__ pred, const double* __restrict__ gt,
               double* __restrict__ out_add, double* __restrict__ out_adi)
{
    __shared__ double sp[kPredTile * 3];
    __shared__ double red[kMetricThreads / 32];
    __shared__ double Tp[12], Tg[12];
    const int pose = blockIdx.x;
    if (threadIdx.x < 12) { Tp[threadIdx.x] = pred[pose * 16 + threadIdx.x]; Tg[threadIdx.x] = gt[pose * 16 + threadIdx.x]; }
    __syncthreads();
    double sum_add = 0, sum_adi = 0;
    for (int base = 0; base < m; base += kMetricThreads) {
        const int i = base + threadIdx.x;
        const bool have = i < m;
        double gx = 0, gy = 0, gz = 0;
        if (have) {
            const double x = model[i * 3], y = model[i * 3 + 1], z = model[i * 3 + 2];
            double px, py, pz;
            xform(Tp, x, y, z, px, py, pz);
            xform(Tg, x, y, z, gx, gy, gz);
            const double dx = px - gx, dy = py - gy, dz = pz - gz;
            sum_add += sqrt(dx * dx + dy * dy + dz * dz);
        }
        if (out_adi) {
            double best = 1.0e300;
            for (int t0 = 0; t0 < m; t0 += kPredTile) {
                __syncthreads();
                for (int j = threadIdx.x; j < kPredTile && t0 + j < m; j += kMetricThreads) {
                    double px, py, pz;
                    xform(Tp, model[(t0 + j) * 3], model[(t0 + j) * 3 + 1], model[(t0 + j) * 3 + 2], px, py, pz);
                    sp[j * 3] = px; sp[j * 3 + 1] = py; sp[j * 3 + 2] = pz;
                }
                __syncthreads();
                const int cnt = min(kPredTile, m - t0);
                if (have)
                    for (int j = 0; j < cnt; ++j) {
                        const double dx = sp[j * 3] - gx, dy = sp[j * 3 + 1] - gy, dz = sp[j * 3 + 2] - gz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        best = d2 < best ? d2 : best;
                    }
            }
            if (have) sum_adi += sqrt(best);
        }
    }
    const double a = block_sum(sum_add, red);
    if (threadIdx.x == 0 && out_add) out_add[pose] = a / m;
    if (out_adi) {
        const double b = block_sum(sum_adi, red);
        if (threadIdx.x == 0) out_adi[pose] = b / m;
    }
}

// errs sorted ascending; ap = 10 * [ sum_j (r_j - r_{j-1}) * j/n  +  (0.1 - r_c) * c/n ],  r_0 = 0, c = #(r < 0.1)
__global__ void __launch_bounds__(1024)
vocap_kernel(const double* __restrict__ rec, int n, double* __restrict__ out)
{
    __shared__ double red[32];
    __shared__ int s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    int c_local = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c_local += rec[i] < 0.1 ? 1 : 0;
    atomicAdd(&s_c, c_local);
    __syncthreads();
    const int c = s_c;
    double s = 0;
    for (int j = threadIdx.x + 1; j <= c; j += blockDim.x) {
        const double prev = (j == 1) ? 0.0 : rec[j - 2];
        s += (rec[j - 1] - prev) * (static_cast<double>(j) / static_cast<double>(n));
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0;
        for (int w = 0; w < 32; ++w) tot += red[w];
        if (c > 0) tot += (0.1 - rec[c - 1]) * (static_cast<double>(c) / static_cast<double>(n));
        out[0] = c > 0 ? tot * 10.0 : 0.0;
    }
}
}  // namespace

cudaError_t launch_add_adi(const double* model, int m, const double* pred, const double* gt, int n,
                           double* out_add, double* out_adi, cudaStream_t s) {
    if (n <= 0 || m <= 0) return cudaSuccess;
    add_adi_kernel<<<n, kMetricThreads, 0, s>>>(model, m, pred, gt, out_add, out_adi);
    return cudaGetLastError();
}

cudaError_t vocap(const double* errs, int n, double* out_host, cudaStream_t s) {
    if (n <= 0) { *out_host = 0.0; return cudaSuccess; }
    double *sorted = nullptr, *d_out = nullptr; void* tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, errs, sorted, n, 0, 64, s);
    if (e != cudaSuccess) return e;
    if ((e = cudaMalloc(&sorted, sizeof(double) * n)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_out, sizeof(double))) != cudaSuccess) { cudaFree(sorted); return e; }
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1)) != cudaSuccess) { cudaFree(sorted); cudaFree(d_out); return e; }
    e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, errs, sorted, n, 0, 64, s);
    if (e == cudaSuccess) {
        vocap_kernel<<<1, 1024, 0, s>>>(sorted, n, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp); cudaFree(sorted); cudaFree(d_out);
    return e;
}

}  // namespace se3tn

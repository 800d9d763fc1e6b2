// ADD / ADD-S / VOCap on the GPU (see metrics.cu).
#pragma once
#include <cuda_runtime.h>
namespace se3tn {
cudaError_t launch_add_adi(const double* model, int m, const double* pred, const double* gt, int n,
                           double* out_add, double* out_adi, cudaStream_t s);
cudaError_t vocap(const double* errs, int n, double* out_host, cudaStream_t s);   // synchronises the stream
}

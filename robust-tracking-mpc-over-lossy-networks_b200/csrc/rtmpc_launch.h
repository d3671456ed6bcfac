// Host-side launchers of the QP kernels (one translation unit per kernel family so they compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "rtmpc_ipm.cuh"

namespace rtmpc {

struct QPLaunch {
    int B;
    const double *x_init, *ref;
    const int* sel;
    int sel_value;
    double *z, *U;
    int *status, *iters, *warm;
    unsigned long long* work;
    cudaStream_t stream;
};

// interior-point kernel: picks the instantiation for (n, mpad); returns false if none fits
bool ipm_configure(const QPDev& P, int max_smem, int* wpb, size_t* smem, cudaError_t* err);
cudaError_t ipm_launch(const QPDev& P, int wpb, size_t smem, int num_sms, const QPLaunch& a);

// dual active-set kernel
bool as_configure(const QPDev& P, int max_smem, int* wpb, size_t* smem, int* g_in_smem, cudaError_t* err);
cudaError_t as_launch(const QPDev& P, int wpb, size_t smem, int g_in_smem, int num_sms, const QPLaunch& a);

}  // namespace rtmpc

// Instantiations and launcher of the interior-point kernel (fallback path of the QP solve).
#include "rtmpc_launch.h"

namespace rtmpc {

typedef void (*ipm_fn)(QPDev, int, const double*, const double*, const int*, int, double*, double*, int*, int*, int*);
struct IpmChoice { int bs, r; ipm_fn fn; };
static const IpmChoice kIpm[] = {
    {2, 4, ipm_solve_kernel<2, 4>},   {3, 9, ipm_solve_kernel<3, 9>},   {3, 16, ipm_solve_kernel<3, 16>},
    {4, 24, ipm_solve_kernel<4, 24>}, {5, 32, ipm_solve_kernel<5, 32>},
};

static const IpmChoice* pick(int n, int mpad) {
    const int bs_need = (n + 6) / 7, r_need = mpad / 32;
    for (const auto& c : kIpm)
        if (c.bs >= bs_need && c.r >= r_need) return &c;
    return nullptr;
}

bool ipm_configure(const QPDev& P, int max_smem, int* wpb_out, size_t* smem_out, cudaError_t* err) {
    *err = cudaSuccess;
    const IpmChoice* kc = pick(P.n, P.mpad);
    if (!kc) return false;
    int wpb = 8;
    size_t smem = 0;
    for (; wpb >= 1; wpb >>= 1) {
        smem = ((size_t)ipm_block_doubles(P) + 2 + (size_t)wpb * (ipm_warp_doubles(P) + 2)) * sizeof(double);
        if (smem <= (size_t)max_smem) break;
    }
    if (wpb < 1) return false;
    // several QPs may share one instantiation: always opt in to the device maximum
    *err = cudaFuncSetAttribute((const void*)kc->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    *wpb_out = wpb;
    *smem_out = smem;
    return *err == cudaSuccess;
}

cudaError_t ipm_launch(const QPDev& P, int wpb, size_t smem, int num_sms, const QPLaunch& a) {
    const IpmChoice* kc = pick(P.n, P.mpad);
    int blocks = (a.B + wpb - 1) / wpb;
    if (blocks > num_sms) blocks = num_sms;   // one resident CTA per SM, warps loop over instances
    kc->fn<<<blocks, wpb * 32, smem, a.stream>>>(P, a.B, a.x_init, a.ref, a.sel, a.sel_value, a.z, a.U, a.status,
                                                 a.iters, a.warm);
    return cudaGetLastError();
}

}  // namespace rtmpc

// Instantiations and launcher of the dual active-set kernel (main path of the QP solve).
#include "rtmpc_as.cuh"
#include "rtmpc_launch.h"

namespace rtmpc {

typedef void (*as_fn)(QPDev, int, const double*, const double*, const int*, int, double*, double*, int*, int*, int*,
                      unsigned long long*);
struct AsChoice { int r2, maxw; as_fn fn; };
// R2 pairs of rows per lane (mpad = 64 R2); MAXW = resident warps per SM the register budget is sized for
static const AsChoice kAs[] = {
    {2, 24, as_solve_kernel<2, 24>},  {5, 20, as_solve_kernel<5, 20>},  {9, 16, as_solve_kernel<9, 16>},
    {12, 16, as_solve_kernel<12, 16>}, {16, 16, as_solve_kernel<16, 16>},
};

Tuning& tuning() {
    static Tuning t;
    return t;
}

static const AsChoice* pick(int mpad) {
    const int r_need = (mpad + 63) / 64;
    for (const auto& c : kAs)
        if (c.r2 >= r_need) return &c;
    return nullptr;
}

int as_padded_rows(int mpad) {
    const int r_need = (mpad + 63) / 64;
    for (const auto& c : kAs)
        if (c.r2 >= r_need) return 64 * c.r2;
    return -1;
}

bool as_configure(const QPDev& P, int max_smem, int* wpb_out, size_t* smem_out, cudaError_t* err) {
    *err = cudaSuccess;
    const AsChoice* kc = pick(P.mpad);
    if (!kc || kc->r2 * 64 != P.mpad) return false;
    const size_t per_warp = (size_t)as_warp_doubles(P) * sizeof(double);
    int wpb = (int)((size_t)max_smem / per_warp);
    if (wpb > kc->maxw) wpb = kc->maxw;
    if (wpb < 1) return false;
    *err = cudaFuncSetAttribute((const void*)kc->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    *wpb_out = wpb;
    *smem_out = per_warp * wpb;
    return *err == cudaSuccess;
}

cudaError_t as_launch(const QPDev& P, int wpb, size_t smem, int num_sms, const QPLaunch& a) {
    const AsChoice* kc = pick(P.mpad);
    if (!kc || P.mpad != 64 * kc->r2) return cudaErrorInvalidValue;      // (the kernels take mpad = 64 * R2 as a constant)
    // spread the instances over all SMs first, then fill the warps of each CTA
    if (tuning().as_warps > 0 && tuning().as_warps < wpb) wpb = tuning().as_warps;   // RTMPC_TUNE_AS_WARPS
    int warps = balanced_warps(a.B, num_sms, wpb);
    int blocks = (a.B + warps - 1) / warps;
    if (blocks > num_sms) blocks = num_sms;
    const size_t per_warp = (size_t)as_warp_doubles(P) * sizeof(double);
    QPDev Pl = P;
    if (!tuning().cert_factored) Pl.kap = nullptr;      // RTMPC_TUNE_CERT_FACTORED
    kc->fn<<<blocks, warps * 32, per_warp * warps, a.stream>>>(Pl, a.B, a.x_init, a.ref, a.sel, a.sel_value, a.z, a.U,
                                                             a.status, a.iters, a.warm, a.work);
    return cudaGetLastError();
}

}  // namespace rtmpc

// Instantiations and launcher of the dual active-set kernel (main path of the QP solve).
#include <cstdlib>

#include "rtmpc_as.cuh"
#include "rtmpc_launch.h"

namespace rtmpc {

typedef void (*as_fn)(QPDev, int, const double*, const double*, const int*, int, double*, double*, int*, int*, int*,
                      unsigned long long*, int);
struct AsChoice { int r, maxw; as_fn fn; };
// R row slots per lane (mpad <= 32 R); MAXW = resident warps the register budget is sized for
static const AsChoice kAs[] = {
    {4, 32, as_solve_kernel<4, 32>},   {9, 28, as_solve_kernel<9, 28>},   {16, 16, as_solve_kernel<16, 16>},
    {24, 12, as_solve_kernel<24, 12>}, {32, 8, as_solve_kernel<32, 8>},
};
static const AsChoice kAsExp[] = {{9, 16, as_solve_kernel<9, 16>}, {9, 20, as_solve_kernel<9, 20>}, {9, 24, as_solve_kernel<9, 24>}};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static const AsChoice* pick(int mpad) {
    const int r_need = mpad / 32;
    const int want = env_int("RTMPC_AS_MAXW", 0);     // experiment knob: register budget of the R = 9 kernel
    if (want && r_need <= 9)
        for (const auto& c : kAsExp)
            if (c.maxw == want) return &c;
    for (const auto& c : kAs)
        if (c.r >= r_need) return &c;
    return nullptr;
}

bool as_configure(const QPDev& P, int max_smem, int* wpb_out, size_t* smem_out, int* g_in_smem, cudaError_t* err) {
    *err = cudaSuccess;
    const AsChoice* kc = pick(P.mpad);
    if (!kc) return false;
    const size_t per_warp = (size_t)as_warp_doubles(P) * sizeof(double);
    const size_t g_bytes = (size_t)as_block_doubles(P, true) * sizeof(double);
    // G on chip if that still leaves room for at least half of the warps the register budget allows
    int in_smem = (g_bytes + per_warp * ((kc->maxw + 1) / 2) <= (size_t)max_smem) ? 1 : 0;
    in_smem = env_int("RTMPC_AS_GSMEM", in_smem);
    size_t avail = (size_t)max_smem - (in_smem ? g_bytes : 0);
    int wpb = (int)(avail / per_warp);
    if (wpb > kc->maxw) wpb = kc->maxw;
    if (env_int("RTMPC_AS_WPB", 0) > 0 && env_int("RTMPC_AS_WPB", 0) < wpb) wpb = env_int("RTMPC_AS_WPB", 0);
    if (wpb < 1) return false;
    *err = cudaFuncSetAttribute((const void*)kc->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    *wpb_out = wpb;
    *smem_out = (in_smem ? g_bytes : 0) + per_warp * wpb;
    *g_in_smem = in_smem;
    return *err == cudaSuccess;
}

cudaError_t as_launch(const QPDev& P, int wpb, size_t smem, int g_in_smem, int num_sms, const QPLaunch& a) {
    const AsChoice* kc = pick(P.mpad);
    // spread the instances over all SMs first, then fill the warps of each CTA
    int per_cta = (a.B + num_sms - 1) / num_sms;
    int warps = per_cta < wpb ? per_cta : wpb;
    if (warps < 1) warps = 1;
    int blocks = (a.B + warps - 1) / warps;
    if (blocks > num_sms) blocks = num_sms;
    const size_t per_warp = (size_t)as_warp_doubles(P) * sizeof(double);
    const size_t bytes = smem - per_warp * (wpb - warps);
    kc->fn<<<blocks, warps * 32, bytes, a.stream>>>(P, a.B, a.x_init, a.ref, a.sel, a.sel_value, a.z, a.U, a.status,
                                                    a.iters, a.warm, a.work, g_in_smem);
    return cudaGetLastError();
}

}  // namespace rtmpc

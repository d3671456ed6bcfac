// C ABI of librtmpc_b200.so (see include/rtmpc.h).  Host-side glue only: uploads the
// once-per-problem data, picks the kernel instantiation, launches on the caller's stream.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <algorithm>
#include <cmath>
#include <vector>

#define RTMPC_LOOP_KERNELS
#include "rtmpc_launch.h"
#include "rtmpc_loop.cuh"

using namespace rtmpc;

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(const char* what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
    return -1;
}
namespace rtmpc { int set_last_error(const char* what, cudaError_t e) { return fail(what, e); } }
// a handle's arrays live on the device that was current when it was created
static int wrong_device(int handle_device, const char* who) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != handle_device) {
        g_err = std::string(who) + ": the handle belongs to device " + std::to_string(handle_device) +
                " but device " + std::to_string(cur) + " is current (rtmpc_set_device)";
        return 1;
    }
    return 0;
}
#define CU(call)                                                    \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return fail(#call, _e);              \
    } while (0)

struct rtmpc_qp {
    QPDev dev;
    std::vector<void*> allocs;
    int device = 0, num_sms = 0;
    int ipm_wpb = 8, as_wpb = 0;
    size_t ipm_smem = 0, as_smem = 0;
    int method = RTMPC_METHOD_ACTIVE_SET;
    unsigned long long* d_work = nullptr;   // algorithmic flop counter of the active-set kernel
    // grow-only scratch (status of callers that pass none; staging of the host-buffer entry point)
    int cap = 0, warm_cap = 0;
    int* s_tmp_status = nullptr;
    double *s_x = nullptr, *s_ref = nullptr, *s_z = nullptr, *s_U = nullptr;
    int *s_sel = nullptr, *s_status = nullptr, *s_iters = nullptr, *s_warm = nullptr;
    int tmp_cap = 0;
    cudaStream_t stream = nullptr;
};

template <typename T>
static int upload(rtmpc_qp* q, const T* src, size_t count, const T** dst) {
    *dst = nullptr;
    if (!src || count == 0) return 0;
    void* p = nullptr;
    CU(cudaMalloc(&p, count * sizeof(T)));
    q->allocs.push_back(p);
    CU(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *dst = static_cast<const T*>(p);
    return 0;
}

// An unselected instance whose status slot happens to hold the transient hand-over mark (stale or uninitialised caller
// memory) must not be picked up by the interior-point launch that follows the active-set launch.
__global__ void clear_stale_handover_kernel(int B, const int* __restrict__ sel, int sel_value, int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && sel[b] != sel_value && status[b] <= RTMPC_FALLBACK_STATUS) status[b] = -1;
}

// start of a rollout: every instance at the loop's common time, nothing pending, counters / ticket flags cleared
__global__ void rollout_prepare_kernel(int B, int t, int* __restrict__ inst_t, int* __restrict__ pending,
                                       int* __restrict__ npend) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { inst_t[b] = t; pending[b] = 0; npend[2 + b] = 0; }
    if (b < 2) npend[b] = 0;
}

extern "C" {

int rtmpc_abi_version(void) { return RTMPC_ABI_VERSION; }
const char* rtmpc_last_error(void) { return g_err.c_str(); }
int64_t rtmpc_launch_count(void) { return (int64_t)g_launches.load(); }

int rtmpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { g_err = "cudaGetDeviceCount failed: no CUDA device"; return -1; }
    return n;
}
int rtmpc_set_device(int device) {
    CU(cudaSetDevice(device));
    return 0;
}

int rtmpc_set_tuning(int32_t knob, int32_t value) {
    Tuning& t = tuning();
    switch (knob) {
        case RTMPC_TUNE_ROLLOUT_QUANTUM: t.rollout_quantum = value < 0 ? 25 : value; return 0;
        case RTMPC_TUNE_ROLLOUT_WARPS: t.rollout_warps = value < 0 ? 0 : value; return 0;
        case RTMPC_TUNE_AS_WARPS: t.as_warps = value < 0 ? 0 : value; return 0;
        case RTMPC_TUNE_ROLLOUT_CARRY: t.rollout_carry = value < 0 ? 1 : (value ? 1 : 0); return 0;
        case RTMPC_TUNE_ROLLOUT_FIXED_DIMS: t.rollout_fixed_dims = value < 0 ? 1 : (value ? 1 : 0); return 0;
        case RTMPC_TUNE_CERT_FACTORED: t.cert_factored = value < 0 ? 1 : (value ? 1 : 0); return 0;
        default: return fail("rtmpc_set_tuning: unknown knob");
    }
}
int32_t rtmpc_get_tuning(int32_t knob) {
    const Tuning& t = tuning();
    switch (knob) {
        case RTMPC_TUNE_ROLLOUT_QUANTUM: return t.rollout_quantum;
        case RTMPC_TUNE_ROLLOUT_WARPS: return t.rollout_warps;
        case RTMPC_TUNE_AS_WARPS: return t.as_warps;
        case RTMPC_TUNE_ROLLOUT_CARRY: return t.rollout_carry;
        case RTMPC_TUNE_ROLLOUT_FIXED_DIMS: return t.rollout_fixed_dims;
        case RTMPC_TUNE_CERT_FACTORED: return t.cert_factored;
        default: return -1;
    }
}

int rtmpc_qp_create(const rtmpc_qp_desc* d, rtmpc_qp** out) {
    if (!d || !out) return fail("rtmpc_qp_create: null argument");
    *out = nullptr;
    if (d->n < 1 || d->n > 32 || d->npad < d->n || d->npad > 32 || (d->npad & 3))
        return fail("rtmpc_qp_create: need 1 <= n <= npad <= 32 and npad % 4 == 0");
    if (d->nx < 1 || d->nx > 8 || d->nu < 1 || d->nu > 4) return fail("rtmpc_qp_create: need nx <= 8, nu <= 4");
    if (d->mpad < 32 || (d->mpad & 31) || d->m > d->mpad || d->mpad > 1024)
        return fail("rtmpc_qp_create: mpad must be a multiple of 32 with m <= mpad <= 1024");
    if (d->n > 30) return fail("rtmpc_qp_create: the active-set kernel keeps one working-set slot per lane: n <= 30");
    if (!d->Hs || !d->Hinv || !d->G || !d->Y || !d->Fx || !d->Fr || !d->lo0 || !d->up0 || !d->Lx || !d->Ux ||
        !d->has_lo || !d->has_up || !d->Dscale || !d->Phi || !d->Psi)
        return fail("rtmpc_qp_create: null matrix in the description");
    // rows are padded to what the active-set kernel's instantiation works on (a multiple of 64)
    const int mpad = as_padded_rows(d->mpad > d->min_rows ? d->mpad : d->min_rows);
    if (mpad < 0) return fail("rtmpc_qp_create: too many rows for the compiled kernel set (mpad <= 1024)");
    if (d->nss > 0 && !d->Kss) return fail("rtmpc_qp_create: Kss required when nss > 0");
    rtmpc_qp* q = new rtmpc_qp();
    {
        cudaError_t e0 = cudaGetDevice(&q->device);
        if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&q->num_sms, cudaDevAttrMultiProcessorCount, q->device);
        if (e0 != cudaSuccess) { delete q; return fail("rtmpc_qp_create: no usable CUDA device", e0); }
    }
    QPDev& P = q->dev;
    std::memset(&P, 0, sizeof(P));
    P.nx = d->nx; P.nu = d->nu; P.N = d->N; P.n = d->n; P.npad = d->npad; P.m = d->m; P.mpad = mpad;
    P.np = d->np; P.nz = d->nz; P.nss = d->nss;
    P.gs = d->npad + 2;
    P.ss = d->npad + 1;
    P.va_len = ((mpad > d->nz ? mpad : d->nz) + 1) & ~1;
    P.s_floor = d->s_floor; P.sc_b = d->sc_b; P.max_iter = d->max_iter > 0 ? d->max_iter : 60;
    P.as_max_steps = 16 * d->npad + 128;
    int mtot = 0;
    for (int i = 0; i < d->mpad; ++i) mtot += (d->has_lo[i] ? 1 : 0) + (d->has_up[i] ? 1 : 0);
    P.mtot = mtot > 0 ? mtot : 1;
    const int npad = d->npad, nx = d->nx, m0 = d->mpad;
    const size_t nn = (size_t)npad * npad, mn = (size_t)mpad * npad;
    // row-indexed inputs, zero-padded to mpad rows (padding rows carry no bound)
    std::vector<double> G(mn, 0.0), Y(mn, 0.0), lo0(mpad, 0.0), up0(mpad, 0.0), Lx((size_t)mpad * nx, 0.0),
        Ux((size_t)mpad * nx, 0.0);
    std::vector<unsigned char> has_lo(mpad, 0), has_up(mpad, 0);
    std::vector<int> shift(mpad, -1);
    std::memcpy(G.data(), d->G, (size_t)m0 * npad * sizeof(double));
    std::memcpy(Y.data(), d->Y, (size_t)m0 * npad * sizeof(double));
    std::memcpy(lo0.data(), d->lo0, (size_t)m0 * sizeof(double));
    std::memcpy(up0.data(), d->up0, (size_t)m0 * sizeof(double));
    std::memcpy(Lx.data(), d->Lx, (size_t)m0 * nx * sizeof(double));
    std::memcpy(Ux.data(), d->Ux, (size_t)m0 * nx * sizeof(double));
    std::memcpy(has_lo.data(), d->has_lo, (size_t)m0);
    std::memcpy(has_up.data(), d->has_up, (size_t)m0);
    if (d->shift) std::memcpy(shift.data(), d->shift, (size_t)m0 * sizeof(int));
    // operators derived once on the host (shared by every instance):
    //   W = G Hinv G' (= G Y'), Zx = -Hinv Fx, Zr = -Hinv Fr, Tx = -Y Fx, Tr = -Y Fr; G, Tx, Tr, Ux, Lx transposed
    std::vector<double> W((size_t)mpad * mpad), Zx((size_t)npad * nx), Zr((size_t)npad * nx), ExT((size_t)mpad * nx),
        TrT((size_t)mpad * nx), UxT((size_t)mpad * nx), LxT((size_t)mpad * nx), GT(mn), upI(mpad), loI(mpad), wid(mpad);
    bool as_ok = true;
    for (int a = 0; a < mpad; ++a)
        for (int b = a; b < mpad; ++b) {
            double acc = 0.0;
            for (int k = 0; k < npad; ++k) acc += G[(size_t)a * npad + k] * Y[(size_t)b * npad + k];
            W[(size_t)a * mpad + b] = acc;
            W[(size_t)b * mpad + a] = acc;
        }
    for (int j = 0; j < npad; ++j)
        for (int k = 0; k < nx; ++k) {
            double ax = 0.0, ar = 0.0;
            for (int i = 0; i < npad; ++i) {
                ax -= d->Hinv[(size_t)j * npad + i] * d->Fx[(size_t)i * nx + k];
                ar -= d->Hinv[(size_t)j * npad + i] * d->Fr[(size_t)i * nx + k];
            }
            Zx[(size_t)j * nx + k] = ax;
            Zr[(size_t)j * nx + k] = ar;
        }
    // Row values at z_u:  G z_u - up = (G Zx - Ux) x + (G Zr) r - up0 with the Zx, Zr the kernels compute z_u from
    // (accumulated in long double, rounded once: the tables agree with G z to the last bit or two, which is what the
    // certification's bound below relies on; -Y Fx would differ from G Zx by the rounding of Y = G Hinv, ~1e-7 here).
    std::vector<double> kap(mpad, 0.0);
    double zy_max = 0.0;
    for (size_t i = 0; i < Zx.size(); ++i) zy_max = std::max(zy_max, std::max(std::fabs(Zx[i]), std::fabs(Zr[i])));
    for (size_t i = 0; i < mn; ++i) zy_max = std::max(zy_max, std::fabs(Y[i]));
    for (int r = 0; r < mpad; ++r) {
        for (int k = 0; k < nx; ++k) {
            long double ax = 0.0L, ar = 0.0L;
            for (int i = 0; i < npad; ++i) {
                ax += (long double)G[(size_t)r * npad + i] * (long double)Zx[(size_t)i * nx + k];
                ar += (long double)G[(size_t)r * npad + i] * (long double)Zr[(size_t)i * nx + k];
            }
            ExT[(size_t)k * mpad + r] = (double)(ax - (long double)Ux[(size_t)r * nx + k]);
            TrT[(size_t)k * mpad + r] = (double)ar;
            UxT[(size_t)k * mpad + r] = Ux[(size_t)r * nx + k];
            LxT[(size_t)k * mpad + r] = Lx[(size_t)r * nx + k];
        }
        for (int k = 0; k < npad; ++k) GT[(size_t)k * mpad + r] = G[(size_t)r * npad + k];
        upI[r] = has_up[r] ? up0[r] : RTMPC_INF;
        loI[r] = has_lo[r] ? lo0[r] : -RTMPC_INF;
        wid[r] = (has_up[r] && has_lo[r]) ? up0[r] - lo0[r] : 3.0 * RTMPC_INF;   // -e - wid < 0 even for e = -1e30
        if (r < d->m && !has_up[r]) as_ok = false;                  // the active-set kernel measures every row from its upper bound
        if (has_up[r] && has_lo[r])
            for (int k = 0; k < nx; ++k)
                if (Lx[(size_t)r * nx + k] != Ux[(size_t)r * nx + k]) as_ok = false;   // ... and needs a constant width
    }
    // kap[i] (S = 1 + |x|_1 + |ref|_1 + sum |multipliers|):  | (G z - up)_i - (Ex x + Tr r - up0 - W[:,A] (s lam))_i | <= kap[i] S
    // for the z and the factored row value the kernels compute - rounding of the two evaluations (at most 32 operations
    // each on terms bounded by the row's largest table entry) plus what the stored tables differ by from G Zx - Ux,
    // G Zr, G Y' (measured here in long double); doubled for margin.  as_certify accepts the factored values only where
    // they clear the tolerance by this bound and recomputes the rows from G' z otherwise.
    for (int r = 0; r < mpad; ++r) {
        long double g1 = 0.0L, dmax = 0.0L;
        double big = has_up[r] ? std::fabs(up0[r]) : 0.0;
        for (int k = 0; k < npad; ++k) g1 += std::fabs((long double)G[(size_t)r * npad + k]);
        for (int k = 0; k < nx; ++k) {
            long double ax = 0.0L, ar = 0.0L;
            for (int i = 0; i < npad; ++i) {
                ax += (long double)G[(size_t)r * npad + i] * (long double)Zx[(size_t)i * nx + k];
                ar += (long double)G[(size_t)r * npad + i] * (long double)Zr[(size_t)i * nx + k];
            }
            dmax = std::max(dmax, std::fabs(ax - (long double)Ux[(size_t)r * nx + k] - (long double)ExT[(size_t)k * mpad + r]));
            dmax = std::max(dmax, std::fabs(ar - (long double)TrT[(size_t)k * mpad + r]));
            big = std::max(big, std::max(std::fabs(ExT[(size_t)k * mpad + r]), std::fabs(TrT[(size_t)k * mpad + r])));
        }
        for (int a = 0; a < mpad; ++a) {
            long double acc = 0.0L;          // effect on row r of a unit step along Y_a, as W[a][r] is read by the kernels
            for (int k = 0; k < npad; ++k) acc += (long double)G[(size_t)r * npad + k] * (long double)Y[(size_t)a * npad + k];
            dmax = std::max(dmax, std::fabs(acc - (long double)W[(size_t)a * mpad + r]));
            big = std::max(big, std::fabs(W[(size_t)a * mpad + r]));
        }
        const long double u = 1.1102230246251566e-16L;      // 2^-53
        kap[r] = (double)(2.0L * (64.0L * u * ((long double)big + g1 * (long double)zy_max) + dmax)) * (1.0 + 1e-9);
    }
    if (!as_ok) {
        delete q;
        return fail("rtmpc_qp_create: every row needs an upper bound and two-sided rows the same x_init dependence on "
                    "both sides (Lx == Ux), as produced by rtmpc_b200.condense");
    }
    // packet payload map from the scaled decision (unscaling and the steady-state gain folded in)
    const int nrow = (d->N + 1) * d->nu, ou = nx * (d->N + 1), oxb = ou + d->N * d->nu, oub = oxb + nx;
    std::vector<double> UPhiT((size_t)npad * nrow, 0.0), UPsiT((size_t)nx * nrow, 0.0);
    for (int i = 0; i < nrow; ++i) {
        const bool last = i >= d->N * d->nu;
        if (last && d->nss == 0) continue;
        const int j = i - d->N * d->nu;
        for (int k = 0; k < npad; ++k) {
            double v = last ? d->Phi[(size_t)(oub + j) * npad + k] : d->Phi[(size_t)(ou + i) * npad + k];
            if (last) for (int c = 0; c < nx; ++c) v += d->Kss[(size_t)j * nx + c] * d->Phi[(size_t)(oxb + c) * npad + k];
            UPhiT[(size_t)k * nrow + i] = v * d->Dscale[k];
        }
        for (int k = 0; k < nx; ++k) {
            double v = last ? d->Psi[(size_t)(oub + j) * nx + k] : d->Psi[(size_t)(ou + i) * nx + k];
            if (last) for (int c = 0; c < nx; ++c) v += d->Kss[(size_t)j * nx + c] * d->Psi[(size_t)(oxb + c) * nx + k];
            UPsiT[(size_t)k * nrow + i] = v;
        }
    }
    // usually every payload row is a single (scaled) decision variable: u_k itself, u_bar + K x_bar = c * theta
    std::vector<int> Uidx(nrow, 0);
    std::vector<double> Ucoef(nrow, 0.0);
    bool single = true;
    for (int i = 0; i < nrow && single; ++i) {
        int nnz = 0;
        for (int k = 0; k < npad; ++k)
            if (UPhiT[(size_t)k * nrow + i] != 0.0) { ++nnz; Uidx[i] = k; Ucoef[i] = UPhiT[(size_t)k * nrow + i]; }
        for (int k = 0; k < nx; ++k)
            if (UPsiT[(size_t)k * nrow + i] != 0.0) nnz = 2;
        if (nnz > 1) single = false;
    }
    int rc = 0;
    if (single) {
        rc |= upload(q, Uidx.data(), Uidx.size(), &P.Uidx);
        rc |= upload(q, Ucoef.data(), Ucoef.size(), &P.Ucoef);
    }
    rc |= upload(q, d->Hs, nn, &P.Hs);
    rc |= upload(q, d->Hinv, nn, &P.Hinv);
    rc |= upload(q, G.data(), mn, &P.G);
    rc |= upload(q, Y.data(), mn, &P.Y);
    rc |= upload(q, d->Fx, (size_t)npad * nx, &P.Fx);
    rc |= upload(q, d->Fr, (size_t)npad * nx, &P.Fr);
    rc |= upload(q, lo0.data(), (size_t)mpad, &P.lo0);
    rc |= upload(q, up0.data(), (size_t)mpad, &P.up0);
    rc |= upload(q, Lx.data(), (size_t)mpad * nx, &P.Lx);
    rc |= upload(q, Ux.data(), (size_t)mpad * nx, &P.Ux);
    rc |= upload(q, has_lo.data(), (size_t)mpad, &P.has_lo);
    rc |= upload(q, has_up.data(), (size_t)mpad, &P.has_up);
    rc |= upload(q, d->parC, (size_t)d->np * nx, &P.parC);
    rc |= upload(q, d->parh, (size_t)d->np, &P.parh);
    rc |= upload(q, d->Dscale, (size_t)npad, &P.D);
    rc |= upload(q, d->Phi, (size_t)d->nz * npad, &P.Phi);
    rc |= upload(q, d->Psi, (size_t)d->nz * nx, &P.Psi);
    rc |= upload(q, d->Kss, (size_t)d->nu * nx, &P.Kss);
    if (d->shift) rc |= upload(q, shift.data(), (size_t)mpad, &P.shift);
    rc |= upload(q, W.data(), W.size(), &P.W);
    rc |= upload(q, GT.data(), GT.size(), &P.GT);
    rc |= upload(q, Zx.data(), Zx.size(), &P.Zx);
    rc |= upload(q, Zr.data(), Zr.size(), &P.Zr);
    rc |= upload(q, ExT.data(), ExT.size(), &P.ExT);
    rc |= upload(q, TrT.data(), TrT.size(), &P.TrT);
    rc |= upload(q, UxT.data(), UxT.size(), &P.UxT);
    rc |= upload(q, LxT.data(), LxT.size(), &P.LxT);
    rc |= upload(q, upI.data(), upI.size(), &P.upI);
    rc |= upload(q, loI.data(), loI.size(), &P.loI);
    rc |= upload(q, wid.data(), wid.size(), &P.wid);
    rc |= upload(q, kap.data(), kap.size(), &P.kap);
    rc |= upload(q, UPhiT.data(), UPhiT.size(), &P.UPhiT);
    rc |= upload(q, UPsiT.data(), UPsiT.size(), &P.UPsiT);
    if (rc) { rtmpc_qp_destroy(q); return -1; }
    if (P.nss > 0 && !P.Kss) { rtmpc_qp_destroy(q); return fail("rtmpc_qp_create: Kss required when nss > 0"); }

    int max_smem = 0;
    if (cudaError_t e1 = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, q->device)) {
        rtmpc_qp_destroy(q);
        return fail("rtmpc_qp_create: cudaDeviceGetAttribute", e1);
    }
    cudaError_t e = cudaSuccess;
    if (!ipm_configure(P, max_smem, &q->ipm_wpb, &q->ipm_smem, &e)) {
        rtmpc_qp_destroy(q);
        return fail("rtmpc_qp_create: problem does not fit the interior-point kernel (shared memory / mpad <= 1024)", e);
    }
    if (!as_configure(P, max_smem, &q->as_wpb, &q->as_smem, &e)) {
        rtmpc_qp_destroy(q);
        return fail("rtmpc_qp_create: problem does not fit the active-set kernel", e);
    }
    if (!rollout_configure(P, max_smem, &e)) {
        rtmpc_qp_destroy(q);
        return fail("rtmpc_qp_create: problem does not fit the rollout kernel", e);
    }
    *out = q;
    return 0;
}

void rtmpc_qp_destroy(rtmpc_qp* q) {
    if (!q) return;
    for (void* p : q->allocs) cudaFree(p);
    cudaFree(q->s_x); cudaFree(q->s_ref); cudaFree(q->s_z); cudaFree(q->s_U);
    cudaFree(q->s_sel); cudaFree(q->s_status); cudaFree(q->s_iters); cudaFree(q->s_warm);
    cudaFree(q->s_tmp_status);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
}

int rtmpc_qp_set_method(rtmpc_qp* q, int32_t method) {
    if (!q) return fail("rtmpc_qp_set_method: null handle");
    if (method != RTMPC_METHOD_ACTIVE_SET && method != RTMPC_METHOD_INTERIOR_POINT)
        return fail("rtmpc_qp_set_method: unknown method");
    q->method = method;
    return 0;
}

int rtmpc_qp_set_step_cap(rtmpc_qp* q, int32_t max_steps) {
    if (!q) return fail("rtmpc_qp_set_step_cap: null handle");
    q->dev.as_max_steps = max_steps > 0 ? max_steps : 16 * q->dev.npad + 128;
    return 0;
}

int rtmpc_qp_set_work_counter(rtmpc_qp* q, uint64_t* d_counter) {
    if (!q) return fail("rtmpc_qp_set_work_counter: null handle");
    q->d_work = reinterpret_cast<unsigned long long*>(d_counter);
    return 0;
}

int32_t rtmpc_qp_warm_stride(rtmpc_qp* q) { return q ? q->dev.npad + 1 : -1; }
const char* rtmpc_qp_rollout_kernel(rtmpc_qp* q) { return q ? rollout_kernel_name(q->dev) : ""; }
int32_t rtmpc_qp_rows(rtmpc_qp* q) { return q ? q->dev.mpad : -1; }

int rtmpc_qp_solve(rtmpc_qp* q, int32_t B, const double* d_x_init, const double* d_ref, const int32_t* d_sel,
                   int32_t sel_value, int32_t* d_warm, double* d_z, double* d_U_t, int32_t* d_status, int32_t* d_iters,
                   void* stream) {
    if (!q) return fail("rtmpc_qp_solve: null handle");
    if (wrong_device(q->device, "rtmpc_qp_solve")) return -1;
    if (B <= 0) return 0;
    if (!d_x_init) return fail("rtmpc_qp_solve: d_x_init is null");
    QPLaunch a;
    a.B = B; a.x_init = d_x_init; a.ref = d_ref; a.sel = d_sel; a.sel_value = sel_value; a.z = d_z; a.U = d_U_t;
    a.status = d_status; a.iters = d_iters; a.warm = d_warm; a.work = q->d_work; a.stream = (cudaStream_t)stream;
    if (q->method == RTMPC_METHOD_INTERIOR_POINT) {
        CU(ipm_launch(q->dev, q->ipm_wpb, q->ipm_smem, q->num_sms, a));
        g_launches.fetch_add(1);
        return 0;
    }
    if (!d_status) {
        // the hand-over to the interior-point kernel goes through the status array
        if (B > q->tmp_cap) {
            cudaFree(q->s_tmp_status);
            q->s_tmp_status = nullptr; q->tmp_cap = 0;
            CU(cudaMalloc(&q->s_tmp_status, (size_t)B * sizeof(int)));
            q->tmp_cap = B;
        }
        a.status = q->s_tmp_status;
    }
    if (d_sel) {
        clear_stale_handover_kernel<<<(B + 255) / 256, 256, 0, a.stream>>>(B, d_sel, sel_value, a.status);
        g_launches.fetch_add(1);
        CU(cudaGetLastError());
    }
    CU(as_launch(q->dev, q->as_wpb, q->as_smem, q->num_sms, a));
    // instances the active-set kernel handed over (status RTMPC_FALLBACK); exits at once when there are none
    QPLaunch f = a;
    f.sel = a.status; f.sel_value = RTMPC_FALLBACK_STATUS; f.work = nullptr;
    CU(ipm_launch(q->dev, q->ipm_wpb, q->ipm_smem, q->num_sms, f));
    g_launches.fetch_add(2);
    return 0;
}

static int ensure_staging(rtmpc_qp* q, int B) {
    if (!q->stream) CU(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    if (B <= q->cap) return 0;
    cudaFree(q->s_x); cudaFree(q->s_ref); cudaFree(q->s_z); cudaFree(q->s_U);
    cudaFree(q->s_sel); cudaFree(q->s_status); cudaFree(q->s_iters); cudaFree(q->s_warm);
    q->s_x = q->s_ref = q->s_z = q->s_U = nullptr;
    q->s_sel = q->s_status = q->s_iters = q->s_warm = nullptr;
    q->cap = 0;
    q->warm_cap = 0;
    const QPDev& P = q->dev;
    CU(cudaMalloc(&q->s_x, (size_t)B * P.nx * sizeof(double)));
    CU(cudaMalloc(&q->s_ref, (size_t)B * P.nx * sizeof(double)));
    CU(cudaMalloc(&q->s_z, (size_t)B * P.nz * sizeof(double)));
    CU(cudaMalloc(&q->s_U, (size_t)B * (P.N + 1) * P.nu * sizeof(double)));
    CU(cudaMalloc(&q->s_sel, (size_t)B * sizeof(int)));
    CU(cudaMalloc(&q->s_status, (size_t)B * sizeof(int)));
    CU(cudaMalloc(&q->s_iters, (size_t)B * sizeof(int)));
    CU(cudaMalloc(&q->s_warm, (size_t)B * (P.npad + 1) * sizeof(int)));
    q->cap = B;
    return 0;
}

int rtmpc_qp_solve_host(rtmpc_qp* q, int32_t B, const double* h_x_init, const double* h_ref, const int32_t* h_sel,
                        int32_t sel_value, int32_t warm, double* h_z, double* h_U_t, int32_t* h_status,
                        int32_t* h_iters) {
    if (!q) return fail("rtmpc_qp_solve_host: null handle");
    if (wrong_device(q->device, "rtmpc_qp_solve_host")) return -1;
    if (B <= 0) return 0;
    if (ensure_staging(q, B)) return -1;
    const QPDev& P = q->dev;
    cudaStream_t s = q->stream;
    if (warm && q->warm_cap != B) {
        // no state yet (or another batch size): every instance starts cold
        CU(cudaMemsetAsync(q->s_warm, 0xFF, (size_t)B * (P.npad + 1) * sizeof(int), s));
        q->warm_cap = B;
    }
    CU(cudaMemcpyAsync(q->s_x, h_x_init, (size_t)B * P.nx * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_ref) CU(cudaMemcpyAsync(q->s_ref, h_ref, (size_t)B * P.nx * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_sel) {
        CU(cudaMemcpyAsync(q->s_sel, h_sel, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync(q->s_status, 0xFF, (size_t)B * sizeof(int), s));   // unselected instances report -1
        CU(cudaMemsetAsync(q->s_iters, 0, (size_t)B * sizeof(int), s));
    }
    if (rtmpc_qp_solve(q, B, q->s_x, h_ref ? q->s_ref : nullptr, h_sel ? q->s_sel : nullptr, sel_value,
                       warm ? q->s_warm : nullptr, h_z ? q->s_z : nullptr, h_U_t ? q->s_U : nullptr, q->s_status,
                       q->s_iters, s))
        return -1;
    if (h_z) CU(cudaMemcpyAsync(h_z, q->s_z, (size_t)B * P.nz * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_U_t) CU(cudaMemcpyAsync(h_U_t, q->s_U, (size_t)B * (P.N + 1) * P.nu * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_status) CU(cudaMemcpyAsync(h_status, q->s_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (h_iters) CU(cudaMemcpyAsync(h_iters, q->s_iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return 0;
}

int rtmpc_qp_warm_reset(rtmpc_qp* q) {
    if (!q) return fail("rtmpc_qp_warm_reset: null handle");
    q->warm_cap = 0;
    return 0;
}

}  // extern "C"

// ---- closed loop ---------------------------------------------------------------------------

struct rtmpc_loop {
    LoopDev dev;
    std::vector<void*> allocs;
    int B = 0, t = 0, device = 0;
    // scratch of rtmpc_loop_rollout
    double *r_U = nullptr, *r_ref = nullptr;
    int *r_status = nullptr, *r_iters = nullptr, *r_inst_t = nullptr, *r_pending = nullptr, *r_npend = nullptr;
    int *r_warm = nullptr, *r_warm1 = nullptr;
    int r_warm_stride = 0, r_warm1_stride = 0;
    double* r_z = nullptr;
    int r_z_stride = 0;
    int* h_npend = nullptr;     // pinned
};

template <typename T>
static int lalloc(rtmpc_loop* l, size_t count, T** dst, const T* src = nullptr) {
    *dst = nullptr;
    if (count == 0) return 0;
    void* p = nullptr;
    CU(cudaMalloc(&p, count * sizeof(T)));
    l->allocs.push_back(p);
    if (src) CU(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    else CU(cudaMemset(p, 0, count * sizeof(T)));
    *dst = static_cast<T*>(p);
    return 0;
}

extern "C" {

int rtmpc_loop_create(const rtmpc_loop_desc* d, int32_t B, rtmpc_loop** out) {
    if (!d || !out || B <= 0) return fail("rtmpc_loop_create: bad argument");
    *out = nullptr;
    if (d->nx < 1 || d->nx > LOOP_MAX_NX || d->nu < 1 || d->nu > LOOP_MAX_NU) return fail("rtmpc_loop_create: need nx <= 8, nu <= 4");
    if (d->plant == RTMPC_PLANT_CARTPOLE && (d->nx != 4 || d->nu != 1)) return fail("rtmpc_loop_create: cartpole plant needs nx=4, nu=1");
    if (!d->A || !d->B || !d->K) return fail("rtmpc_loop_create: A, B, K required");
    rtmpc_loop* l = new rtmpc_loop();
    l->B = B;
    if (cudaGetDevice(&l->device) != cudaSuccess) { delete l; return fail("rtmpc_loop_create: no CUDA device"); }
    LoopDev& L = l->dev;
    std::memset(&L, 0, sizeof(L));
    L.nx = d->nx; L.nu = d->nu; L.N = d->N; L.actuator = d->actuator; L.plant = d->plant; L.nz_rows = d->nz_rows;
    std::memcpy(L.cart, d->cart_params, sizeof(L.cart));
    const int nx = d->nx, nu = d->nu;
    double *A, *Bm, *K, *Kp, *Hz = nullptr, *hz = nullptr, *wh;
    std::vector<double> zeros(nx, 0.0);
    int rc = 0;
    rc |= lalloc(l, (size_t)nx * nx, &A, d->A);
    rc |= lalloc(l, (size_t)nx * nu, &Bm, d->B);
    rc |= lalloc(l, (size_t)nu * nx, &K, d->K);
    rc |= lalloc(l, (size_t)nu * nx, &Kp, d->K_plant ? d->K_plant : d->K);
    if (d->nz_rows > 0) {
        // Facets in pairs with opposite normals (a'x <= h+, -a'x <= h-; the mRPI set of a symmetric disturbance box has
        // them, with h+ and h- equal up to the rounding of two support evaluations) need one dot product per pair:
        // max(a'd - h+, -a'd - h-).  The normals have to be exact negatives, the bounds may differ.
        const int nr = d->nz_rows;
        std::vector<int> mate(nr, -1);
        bool sym = (nr % 2) == 0;
        for (int i = 0; i < nr && sym; ++i) {
            if (mate[i] >= 0) continue;
            for (int j = i + 1; j < nr; ++j) {
                if (mate[j] >= 0) continue;
                bool opp = true;
                for (int k = 0; k < nx && opp; ++k) opp = d->Hz[(size_t)j * nx + k] == -d->Hz[(size_t)i * nx + k];
                if (opp) { mate[i] = j; mate[j] = i; break; }
            }
            if (mate[i] < 0) sym = false;
        }
        // rows (one per facet, or one per pair of opposite facets: normal, h+, h-)
        std::vector<double> Hh, hh;
        const int per = sym ? 2 : 1;                    // bounds per row
        if (sym) {
            for (int i = 0; i < nr; ++i)
                if (mate[i] > i) {
                    Hh.insert(Hh.end(), d->Hz + (size_t)i * nx, d->Hz + (size_t)(i + 1) * nx);
                    hh.push_back(d->hz[i]);
                    hh.push_back(d->hz[mate[i]]);
                }
            L.tube_sym = 1;
        } else {
            Hh.assign(d->Hz, d->Hz + (size_t)nr * nx);
            hh.assign(d->hz, d->hz + nr);
        }
        // Sorted by the distance of the facet from the origin, nearest first, and padded to whole blocks of 32 rows
        // (one row per lane of the rollout kernel's warp).  A row's value a'd - h is at most |a| (|d| - h / |a|): once
        // some row already seen is above that bound for every row still to come, the scan stops (tube_cut[b] = smallest
        // distance and smallest |a| over the rows from block b on).  The maximum is the same, bit for bit: skipped rows
        // cannot reach it.  For the cartpole tube (427 pairs) 3 of 14 blocks are read on average.
        const int rows = (int)hh.size() / per;
        std::vector<int> order(rows);
        std::vector<double> key(rows), nrm(rows);
        for (int r = 0; r < rows; ++r) {
            double s2 = 0.0;
            for (int k = 0; k < nx; ++k) s2 += Hh[(size_t)r * nx + k] * Hh[(size_t)r * nx + k];
            nrm[r] = std::sqrt(s2);
            const double hmin = sym ? std::min(hh[2 * r], hh[2 * r + 1]) : hh[r];
            key[r] = nrm[r] > 0.0 ? hmin / nrm[r] : 1e300;
            order[r] = r;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
        const int nb = (rows + 31) / 32, padded = nb * 32;
        std::vector<double> Hs((size_t)padded * nx, 0.0), hs((size_t)padded * per, 1e300), cut((size_t)2 * (nb + 1));
        for (int r = 0; r < rows; ++r) {
            std::copy(Hh.begin() + (size_t)order[r] * nx, Hh.begin() + (size_t)(order[r] + 1) * nx, Hs.begin() + (size_t)r * nx);
            for (int k = 0; k < per; ++k) hs[(size_t)r * per + k] = hh[(size_t)order[r] * per + k];
        }
        cut[2 * nb] = -1e300; cut[2 * nb + 1] = 1.0;          // (never read as a reason to stop)
        double kmin = 1e300, nmin = 1e300;
        for (int b = nb - 1; b >= 0; --b) {
            for (int r = 32 * b; r < std::min(rows, 32 * (b + 1)); ++r) {
                kmin = std::min(kmin, key[order[r]]);
                nmin = std::min(nmin, nrm[order[r]]);
            }
            cut[2 * b] = kmin; cut[2 * b + 1] = nmin;
        }
        L.nz_rows = padded;
        rc |= lalloc(l, Hs.size(), &Hz, Hs.data());
        rc |= lalloc(l, hs.size(), &hz, hs.data());
        double* cutd = nullptr;
        rc |= lalloc(l, cut.size(), &cutd, cut.data());
        L.tube_cut = cutd;
    }
    rc |= lalloc(l, (size_t)nx, &wh, d->w_half ? d->w_half : zeros.data());
    L.A = A; L.Bm = Bm; L.K = K; L.Kp = Kp; L.Hz = Hz; L.hz = hz; L.w_half = wh;
    rc |= lalloc(l, (size_t)B * nx, &L.x);
    rc |= lalloc(l, (size_t)B * nx, &L.x_nom);
    rc |= lalloc(l, (size_t)B * nx, &L.x_hat);
    rc |= lalloc(l, (size_t)B * (d->N + 1) * nu, &L.buf);
    rc |= lalloc(l, (size_t)B * nu, &L.u_last);
    rc |= lalloc(l, (size_t)B, &L.err_acc);
    rc |= lalloc(l, (size_t)B, &L.tube_max);
    rc |= lalloc(l, (size_t)B, &L.q_t);
    rc |= lalloc(l, (size_t)B, &L.s_t);
    rc |= lalloc(l, (size_t)B, &L.Theta);
    rc |= lalloc(l, (size_t)B, &L.alive);
    rc |= lalloc(l, (size_t)B, &L.last_loss);
    rc |= lalloc(l, (size_t)B, &L.gamma_last);
    rc |= lalloc(l, (size_t)B * (d->N + 1) * nu, &l->r_U);
    rc |= lalloc(l, (size_t)B * nx, &l->r_ref);
    rc |= lalloc(l, (size_t)B, &l->r_status);
    rc |= lalloc(l, (size_t)B, &l->r_iters);
    rc |= lalloc(l, (size_t)B, &l->r_inst_t);
    rc |= lalloc(l, (size_t)B, &l->r_pending);
    rc |= lalloc(l, (size_t)B + 2, &l->r_npend);      // [instances parked by a launch, next ticket, done[B]]
    if (rc) { rtmpc_loop_destroy(l); return -1; }
    if (cudaMallocHost(&l->h_npend, sizeof(int)) != cudaSuccess) { rtmpc_loop_destroy(l); return fail("cudaMallocHost"); }
    *out = l;
    std::vector<double> x0((size_t)B * nx, 0.0);
    return rtmpc_loop_reset(l, x0.data());
}

void rtmpc_loop_destroy(rtmpc_loop* l) {
    if (!l) return;
    for (void* p : l->allocs) cudaFree(p);
    cudaFree(l->r_warm); cudaFree(l->r_warm1); cudaFree(l->r_z);
    if (l->h_npend) cudaFreeHost(l->h_npend);
    delete l;
}

int rtmpc_loop_reset_device(rtmpc_loop* l, const double* d_x0, void* stream) {
    if (!l) return fail("rtmpc_loop_reset_device: null handle");
    if (wrong_device(l->device, "rtmpc_loop_reset_device")) return -1;
    LoopDev& L = l->dev;
    const size_t B = l->B, nx = L.nx;
    cudaStream_t s = (cudaStream_t)stream;
    if (d_x0) CU(cudaMemcpyAsync(L.x, d_x0, B * nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
    else CU(cudaMemsetAsync(L.x, 0, B * nx * sizeof(double), s));
    const int threads = 128;
    loop_reset_kernel<<<(int)((B + threads - 1) / threads), threads, 0, s>>>(L, (int)B, l->r_warm, l->r_warm_stride, l->r_warm1,
                                                                              l->r_warm1_stride);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    l->t = 0;
    return 0;
}

int rtmpc_loop_reset(rtmpc_loop* l, const double* h_x0) {
    if (!l || !h_x0) return fail("rtmpc_loop_reset: null argument");
    if (wrong_device(l->device, "rtmpc_loop_reset")) return -1;
    LoopDev& L = l->dev;
    const size_t B = l->B, nx = L.nx;
    // host buffer in: the copy is synchronous with respect to the host (pageable memory), the kernel runs on the
    // default stream and only that stream is waited for - callers may continue on any stream afterwards
    CU(cudaMemcpy(L.x, h_x0, B * nx * sizeof(double), cudaMemcpyHostToDevice));
    const int threads = 128;
    loop_reset_kernel<<<(int)((B + threads - 1) / threads), threads>>>(L, (int)B, l->r_warm, l->r_warm_stride, l->r_warm1,
                                                                        l->r_warm1_stride);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(nullptr));
    l->t = 0;
    return 0;
}

double* rtmpc_loop_x(rtmpc_loop* l) { return l ? l->dev.x : nullptr; }
double* rtmpc_loop_x_nom(rtmpc_loop* l) { return l ? l->dev.x_nom : nullptr; }
double* rtmpc_loop_x_hat(rtmpc_loop* l) { return l ? l->dev.x_hat : nullptr; }
int32_t* rtmpc_loop_q_t(rtmpc_loop* l) { return l ? l->dev.q_t : nullptr; }
int32_t* rtmpc_loop_s_t(rtmpc_loop* l) { return l ? l->dev.s_t : nullptr; }
int32_t* rtmpc_loop_Theta(rtmpc_loop* l) { return l ? l->dev.Theta : nullptr; }
int32_t* rtmpc_loop_alive(rtmpc_loop* l) { return l ? l->dev.alive : nullptr; }
double* rtmpc_loop_err_acc(rtmpc_loop* l) { return l ? l->dev.err_acc : nullptr; }
double* rtmpc_loop_tube_max(rtmpc_loop* l) { return l ? l->dev.tube_max : nullptr; }
double* rtmpc_loop_u(rtmpc_loop* l) { return l ? l->dev.u_last : nullptr; }
int32_t* rtmpc_loop_gamma(rtmpc_loop* l) { return l ? l->dev.gamma_last : nullptr; }
int32_t rtmpc_loop_time(rtmpc_loop* l) { return l ? l->t : -1; }

int rtmpc_loop_step(rtmpc_loop* l, const double* d_U_t, const int32_t* d_status, const double* d_x_nom0,
                    int64_t x_nom0_stride, const double* d_ref, const int32_t* d_theta, const int32_t* d_gamma,
                    const double* d_w, const double* d_p_loss, uint64_t seed, int64_t id_offset, double* d_traj_x,
                    int64_t traj_stride, void* stream) {
    if (!l || !d_U_t) return fail("rtmpc_loop_step: null argument");
    if (wrong_device(l->device, "rtmpc_loop_step")) return -1;
    if ((d_theta == nullptr) != (d_gamma == nullptr)) return fail("rtmpc_loop_step: theta and gamma must be given together");
    const int threads = 128;
    const int blocks = (l->B + threads - 1) / threads;
    loop_step_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        l->dev, l->B, l->t, d_U_t, d_status, d_x_nom0, (long long)x_nom0_stride, d_ref, d_theta, d_gamma, d_w,
        d_p_loss, (unsigned long long)seed, (long long)id_offset, d_traj_x, (long long)traj_stride);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    l->t += 1;
    return 0;
}

static int ensure_warm(int** buf, int* stride, int want, size_t B, cudaStream_t s) {
    if (*buf && *stride == want) return 0;
    cudaFree(*buf);
    *buf = nullptr;
    *stride = want;
    CU(cudaMalloc(buf, B * (size_t)want * sizeof(int)));
    CU(cudaMemsetAsync(*buf, 0xFF, B * (size_t)want * sizeof(int), s));
    return 0;
}

int rtmpc_loop_rollout(rtmpc_loop* l, rtmpc_qp* q, rtmpc_qp* q1, int32_t T, const double* d_ref, int64_t ref_stride_t,
                       int64_t ref_stride_b, const int32_t* d_theta, const int32_t* d_gamma, const double* d_w,
                       const double* d_p_loss, uint64_t seed, int64_t id_offset, double* d_traj_x, int64_t traj_stride,
                       uint64_t* d_stats, void* stream) {
    if (!l || !q) return fail("rtmpc_loop_rollout: null handle");
    if (wrong_device(l->device, "rtmpc_loop_rollout") || wrong_device(q->device, "rtmpc_loop_rollout") ||
        (q1 && wrong_device(q1->device, "rtmpc_loop_rollout"))) return -1;
    if (T <= 0) return 0;
    if ((d_theta == nullptr) != (d_gamma == nullptr)) return fail("rtmpc_loop_rollout: theta and gamma must be given together");
    const QPDev& P = q->dev;
    const bool ext = l->dev.actuator == RTMPC_ACT_EXTENDED;
    if (ext && !q1) return fail("rtmpc_loop_rollout: the extended variant needs the 'packet received' problem as well");
    if (!ext && q1) return fail("rtmpc_loop_rollout: a second problem is only meaningful for RTMPC_ACT_EXTENDED");
    const QPDev& P1 = q1 ? q1->dev : q->dev;
    if (P.nx != l->dev.nx || P.nu != l->dev.nu || P.N != l->dev.N || P1.nx != P.nx || P1.nu != P.nu || P1.N != P.N)
        return fail("rtmpc_loop_rollout: QP and loop sizes differ");
    if (P1.mpad != P.mpad) return fail("rtmpc_loop_rollout: both problems must be padded to the same row count (rtmpc_qp_desc.min_rows, rtmpc_qp_rows)");
    if (q->method != RTMPC_METHOD_ACTIVE_SET || (q1 && q1->method != RTMPC_METHOD_ACTIVE_SET))
        return fail("rtmpc_loop_rollout: needs RTMPC_METHOD_ACTIVE_SET");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = l->B;
    if (ensure_warm(&l->r_warm, &l->r_warm_stride, P.npad + 1, B, s)) return -1;
    if (q1 && ensure_warm(&l->r_warm1, &l->r_warm1_stride, P1.npad + 1, B, s)) return -1;
    if (ext) {
        const int nzmax = P.nz > P1.nz ? P.nz : P1.nz;       // x_nom_0 of a handed-over step comes back through z
        if (l->r_z_stride < nzmax) {
            cudaFree(l->r_z);
            l->r_z = nullptr;
            CU(cudaMalloc(&l->r_z, B * (size_t)nzmax * sizeof(double)));
            l->r_z_stride = nzmax;
        }
    }
    int max_smem = 0;
    CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, q->device));
    RolloutArgs a;
    a.B = l->B; a.t0 = l->t; a.T = l->t + T;
    a.ref = d_ref; a.ref_stride_t = ref_stride_t; a.ref_stride_b = ref_stride_b;
    a.theta = d_theta; a.gamma = d_gamma; a.w = d_w; a.p_loss = d_p_loss;
    a.seed = seed; a.id_offset = id_offset; a.traj = d_traj_x; a.traj_stride = traj_stride;
    a.warm = l->r_warm; a.warm1 = l->r_warm1; a.two = q1 ? 1 : 0;
    a.U = l->r_U; a.z = l->r_z; a.z_stride = l->r_z_stride; a.status = l->r_status; a.iters = l->r_iters;
    a.inst_t = l->r_inst_t; a.pending = l->r_pending; a.ref_pending = l->r_ref; a.n_pending = l->r_npend; a.next = l->r_npend + 1; a.done = l->r_npend + 2; a.quantum = 0;
    a.stats = reinterpret_cast<unsigned long long*>(d_stats);
    // every instance starts at the loop's common time (filled on the device: nothing is staged on the host)
    rollout_prepare_kernel<<<(int)((B + 255) / 256), 256, 0, s>>>((int)B, l->t, l->r_inst_t, l->r_pending, l->r_npend);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    for (int round = 0;; ++round) {
        if (round > 0) CU(cudaMemsetAsync(l->r_npend, 0, ((size_t)B + 2) * sizeof(int), s));
        CU(rollout_launch(P, P1, l->dev, q->as_wpb, q->num_sms, max_smem, a, s));
        g_launches.fetch_add(1);
        CU(cudaMemcpyAsync(l->h_npend, l->r_npend, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (*l->h_npend == 0) break;
        if (round > T * 4 + 16) return fail("rtmpc_loop_rollout: parked instances do not make progress");
        // parked instances: interior-point solve of their current step (status says for which problem), then resume
        for (int which = 0; which < (q1 ? 2 : 1); ++which) {
            rtmpc_qp* qq = which ? q1 : q;
            QPLaunch f;
            f.B = l->B; f.x_init = l->dev.x_hat; f.ref = d_ref ? l->r_ref : nullptr; f.sel = l->r_status;
            f.sel_value = RTMPC_FALLBACK_STATUS - which; f.z = ext ? l->r_z : nullptr; f.U = l->r_U;
            f.status = l->r_status; f.iters = l->r_iters; f.warm = which ? l->r_warm1 : l->r_warm; f.work = nullptr;
            f.stream = s;
            CU(ipm_launch(qq->dev, qq->ipm_wpb, qq->ipm_smem, qq->num_sms, f));
            g_launches.fetch_add(1);
        }
#ifdef RTMPC_AS_DEBUG
        {   // development builds: which hand-overs did the interior-point kernel not solve either?
            std::vector<int> st(B), pd(B), tt(B);
            const int nx = l->dev.nx;
            std::vector<double> xh((size_t)B * nx), rf((size_t)B * nx);
            CU(cudaStreamSynchronize(s));
            CU(cudaMemcpy(st.data(), l->r_status, B * sizeof(int), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(pd.data(), l->r_pending, B * sizeof(int), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(tt.data(), l->r_inst_t, B * sizeof(int), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(xh.data(), l->dev.x_hat, (size_t)B * nx * sizeof(double), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(rf.data(), l->r_ref, (size_t)B * nx * sizeof(double), cudaMemcpyDeviceToHost));
            for (long long b = 0; b < B; ++b)
                if (pd[b] && st[b] != RTMPC_OPTIMAL) {
                    std::fprintf(stderr, "HANDOVER inst %lld t %d status %d xhat", b, tt[b], st[b]);
                    for (int k = 0; k < nx; ++k) std::fprintf(stderr, " %.17g", xh[(size_t)b * nx + k]);
                    std::fprintf(stderr, " ref");
                    for (int k = 0; k < nx; ++k) std::fprintf(stderr, " %.17g", rf[(size_t)b * nx + k]);
                    std::fprintf(stderr, "\n");
                }
        }
#endif
    }
    l->t += T;
    return 0;
}

// ---- split actuator / estimator calls ---------------------------------------------------------

int rtmpc_actuator_process(int32_t B, int32_t nx, int32_t nu, int32_t N, int32_t kind, int32_t t,
                           const double* d_A, const double* d_B, const double* d_K, const double* d_K_plant,
                           const double* d_x_t, const double* d_U_t, const double* d_x_nom0, const int32_t* d_q_pkt,
                           const int32_t* d_theta, double* d_buf, double* d_x_nom, int32_t* d_s_t, int32_t* d_Theta,
                           int32_t* d_last_loss, double* d_u_out, double* d_pkt_x, double* d_pkt_xnom, void* stream) {
    if (B <= 0) return 0;
    if (nx < 1 || nx > LOOP_MAX_NX || nu < 1 || nu > LOOP_MAX_NU) return fail("rtmpc_actuator_process: need nx <= 8, nu <= 4");
    if (!d_K || !d_x_t || !d_U_t || !d_q_pkt || !d_theta || !d_buf || !d_s_t || !d_Theta || !d_last_loss || !d_u_out || !d_pkt_x)
        return fail("rtmpc_actuator_process: null argument");
    if (kind != RTMPC_ACT_SMART && (!d_A || !d_B || !d_K_plant || !d_x_nom)) return fail("rtmpc_actuator_process: consistent actuator needs A, B, K_plant, x_nom");
    ActArgs a;
    a.nx = nx; a.nu = nu; a.N = N; a.kind = kind; a.t = t; a.has_xnom0 = d_x_nom0 != nullptr;
    a.A = d_A; a.Bm = d_B; a.K = d_K; a.Kp = d_K_plant; a.x_t = d_x_t; a.U_t = d_U_t; a.x_nom0 = d_x_nom0;
    a.q_pkt = d_q_pkt; a.theta = d_theta; a.buf = d_buf; a.x_nom = d_x_nom; a.u_out = d_u_out; a.pkt_x = d_pkt_x;
    a.pkt_xnom = d_pkt_xnom; a.s_t = d_s_t; a.Theta = d_Theta; a.last_loss = d_last_loss;
    actuator_process_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, B);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    return 0;
}

int rtmpc_estimator_update(int32_t B, int32_t nx, int32_t nu, int32_t N, int32_t robust, int32_t t, int32_t n_hist,
                           const double* d_A, const double* d_B, const double* d_K, const double* d_K_plant,
                           const double* d_pkt_x, const double* d_pkt_xnom, const int32_t* d_pkt_s,
                           const int32_t* d_gamma, const double* d_hist, const double* d_x_nom0_mpc, double* d_x_hat,
                           int32_t* d_q_t, void* stream) {
    if (B <= 0) return 0;
    if (nx < 1 || nx > LOOP_MAX_NX || nu < 1 || nu > LOOP_MAX_NU) return fail("rtmpc_estimator_update: need nx <= 8, nu <= 4");
    if (!d_A || !d_B || !d_K || !d_gamma || !d_hist || !d_x_hat || !d_q_t || n_hist < 1) return fail("rtmpc_estimator_update: null argument / empty history");
    if (robust && (!d_K_plant || !d_x_nom0_mpc)) return fail("rtmpc_estimator_update: robust estimator needs K_plant and x_nom0");
    EstArgs e;
    e.nx = nx; e.nu = nu; e.N = N; e.robust = robust; e.t = t; e.n_hist = n_hist;
    e.A = d_A; e.Bm = d_B; e.K = d_K; e.Kp = d_K_plant; e.pkt_x = d_pkt_x; e.pkt_xnom = d_pkt_xnom;
    e.x_nom0_mpc = d_x_nom0_mpc; e.hist = d_hist; e.pkt_s = d_pkt_s; e.gamma = d_gamma; e.x_hat = d_x_hat; e.q_t = d_q_t;
    estimator_update_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e, B);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    return 0;
}

// ---- support sweep -------------------------------------------------------------------------

int rtmpc_support_sweep(const double* d_V, int32_t nv, int32_t dim, const double* d_dirs, int64_t M, double* d_out,
                        void* stream) {
    if (!d_V || !d_dirs || !d_out) return fail("rtmpc_support_sweep: null argument");
    if (dim < 1 || dim > 16) return fail("rtmpc_support_sweep: need 1 <= dim <= 16");
    if (M <= 0) return 0;
    const int kc = (dim + 3) / 4, nv8 = (nv + 7) & ~7;
    const size_t smem = (size_t)kc * nv8 * 4 * sizeof(double);
    if (nv < 1 || smem > 200 * 1024) return fail("rtmpc_support_sweep: vertex set does not fit in shared memory");
    typedef void (*sweep_fn)(const double*, int, int, const double*, long long, double*);
    static const sweep_fn fns[4] = {support_sweep_kernel<4, 1>, support_sweep_kernel<4, 2>, support_sweep_kernel<2, 3>,
                                    support_sweep_kernel<2, 4>};
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    static std::atomic<unsigned long long> configured{0};       // function attributes are per device
    if (dev >= 64 || !((configured.load() >> dev) & 1ull)) {
        for (int i = 0; i < 4; ++i)
            CU(cudaFuncSetAttribute((const void*)fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one block per SM stages the vertices once; a warp takes 8 * RT directions per trip
    const int threads = 512;
    const int per_block = (threads / 32) * 8 * (kc <= 2 ? 4 : 2);
    long long blocks = (M + per_block - 1) / per_block;
    if (blocks > sms) blocks = sms;
    fns[kc - 1]<<<(int)blocks, threads, smem, (cudaStream_t)stream>>>(d_V, nv, dim, d_dirs, (long long)M, d_out);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    return 0;
}

// device staging of the host-buffer entry points: grown on demand, kept for the life of the process (a cudaMalloc /
// cudaFree pair per call costs more than the sweep itself)
namespace {
struct HostStage {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
    cudaError_t need(size_t bytes) {
        int cur = 0;
        cudaError_t e0 = cudaGetDevice(&cur);
        if (e0 != cudaSuccess) return e0;
        if (bytes <= cap && cur == dev) return cudaSuccess;
        if (p) {
            // (a buffer that lives on another device is released there)
            if (dev != cur) { cudaSetDevice(dev); cudaFree(p); cudaSetDevice(cur); } else cudaFree(p);
        }
        dev = cur;
        p = nullptr; cap = 0;
        const cudaError_t e = cudaMalloc(&p, bytes + bytes / 4);
        if (e == cudaSuccess) cap = bytes + bytes / 4;
        return e;
    }
};
std::mutex g_stage_mu;
HostStage g_stage[5];
}  // namespace

int rtmpc_support_sweep_host(const double* h_V, int32_t nv, int32_t dim, const double* h_dirs, int64_t M,
                             double* h_out) {
    if (!h_V || !h_dirs || !h_out) return fail("rtmpc_support_sweep_host: null argument");
    if (M <= 0) return 0;
    std::lock_guard<std::mutex> lock(g_stage_mu);
    const size_t bV = (size_t)nv * dim * sizeof(double), bD = (size_t)M * dim * sizeof(double), bO = (size_t)M * sizeof(double);
    CU(g_stage[0].need(bV));
    CU(g_stage[1].need(bD));
    CU(g_stage[2].need(bO));
    double *dV = (double*)g_stage[0].p, *dD = (double*)g_stage[1].p, *dO = (double*)g_stage[2].p;
    CU(cudaMemcpy(dV, h_V, bV, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dD, h_dirs, bD, cudaMemcpyHostToDevice));
    if (rtmpc_support_sweep(dV, nv, dim, dD, M, dO, nullptr)) return -1;
    CU(cudaMemcpy(h_out, dO, bO, cudaMemcpyDeviceToHost));
    return 0;
}

int rtmpc_model_error_sweep(const double cart_params[8], int32_t B, int32_t T, const double* d_x0, const double* d_K,
                            const double* d_Acl, double* d_w, double* d_x_final, void* stream) {
    if (!cart_params || !d_x0 || !d_K || !d_Acl || !d_w) return fail("rtmpc_model_error_sweep: null argument");
    if (B < 0 || T < 0) return fail("rtmpc_model_error_sweep: negative size");
    if (B == 0 || T == 0) return 0;
    double* d_c = nullptr;
    CU(cudaMalloc(&d_c, 8 * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_c, cart_params, 8 * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess) {
        const int threads = 128;
        model_error_kernel<<<(B + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(d_c, B, T, d_x0, d_K, d_Acl, d_w, d_x_final);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);      // d_c is freed below
    cudaFree(d_c);
    if (e != cudaSuccess) return fail("rtmpc_model_error_sweep", e);
    return 0;
}

int rtmpc_model_error_sweep_host(const double cart_params[8], int32_t B, int32_t T, const double* h_x0,
                                 const double* h_K, const double* h_Acl, double* h_w, double* h_x_final) {
    if (!cart_params || !h_x0 || !h_K || !h_Acl || !h_w) return fail("rtmpc_model_error_sweep_host: null argument");
    if (B <= 0 || T <= 0) return 0;
    std::lock_guard<std::mutex> lock(g_stage_mu);
    const size_t nw = (size_t)B * T * 4 * sizeof(double), nxb = (size_t)B * 4 * sizeof(double);
    CU(g_stage[0].need(nxb));
    CU(g_stage[1].need(nw));
    CU(g_stage[2].need(nxb));
    CU(g_stage[3].need(4 * sizeof(double)));
    CU(g_stage[4].need(16 * sizeof(double)));
    double *dx = (double*)g_stage[0].p, *dw = (double*)g_stage[1].p, *df = (double*)g_stage[2].p;
    double *dk = (double*)g_stage[3].p, *da = (double*)g_stage[4].p;
    CU(cudaMemcpy(dx, h_x0, nxb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dk, h_K, 4 * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(da, h_Acl, 16 * sizeof(double), cudaMemcpyHostToDevice));
    if (rtmpc_model_error_sweep(cart_params, B, T, dx, dk, da, dw, df, nullptr)) return -1;
    CU(cudaMemcpy(h_w, dw, nw, cudaMemcpyDeviceToHost));
    if (h_x_final) CU(cudaMemcpy(h_x_final, df, nxb, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"

// Batched low-dimensional linear programs over ONE shared H-representation:
//
//     value_b = max  c_b' x   s.t.  H x <= h (+ relax_by on row relax_row[b]),  E_b x <= f_b (n_extra own rows),  |x_i| <= box
//
// This is the LP behind the reference's offline set pipeline once the sets have no tractable vertex representation
// (the 9-D terminal sets): `support` (utils_polytope.py:12-23) inside the maximal-output-admissible-set iteration
// (utils_polytope.py:247-268 -> polytope.intersect / == -> one LP per row), the Chebyshev-ball LPs of the subset test,
// and `polytope.reduce` (TubeRegulatorMPC.py:74; one LP per row against all rows with its own bound relaxed by 0.1).
// Thousands of LPs share the matrix H and differ in the objective, in one relaxed bound and at most in a few own rows -
// the same "shared operators, one warp per instance" shape as the QP kernels.
//
// Method: dual simplex in the primal space (an active-set method on vertices).  A basis is `dim` rows holding with
// equality; Binv = H_B^-1 gives the vertex x = Binv h_B and the multipliers lam = Binv' c.  The artificial box supplies
// the dual-feasible start (vertex box * sign(c), lam = |c|).  Per iteration: price every row at x (lanes over rows,
// coalesced reads of the transposed H), let the most violated row p enter, w' = H_p Binv, the ratio test
// min lam_j / w_j (w_j > 0) picks the leaving row, lam / x / Binv are updated in O(dim^2).  Ties (dual degeneracy) are
// handled by a Harris ratio test (largest pivot among near-minimal ratios) and Bland's rule after `dim` stalled steps;
// the basis is refactorised (Gauss-Jordan, partial pivoting) every 32 iterations and before the final optimality check,
// so the reported vertex is exact for its basis.  A singular basis / the iteration limit restarts the LP once with
// Bland's rule from the first step.  tools/lp_model.py is the numpy model of this control flow.
// One warp per LP, state in shared memory (2 dim^2 + 5 dim doubles), FP64.
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtmpc.h"
#include "rtmpc_common.cuh"

namespace rtmpc { int set_last_error(const char* what, cudaError_t e); }   // rtmpc_capi.cu: text behind rtmpc_last_error()

namespace rtmpc {

constexpr int LP_MAX_DIM = 16;
constexpr int LP_WARPS = 8;

struct LPArgs {
    const double* HT;      // [dim][mpad]  shared rows, transposed
    const double* h;       // [mpad]
    int m, mpad, dim;
    const double* obj;     // [B][dim]
    const int* relax_row;  // [B] or NULL: row whose bound is relaxed by relax_by for instance b (-1: none)
    double relax_by;
    const double* extra;   // [B][n_extra][dim + 1] own rows (coefficients, then bound) or NULL
    int n_extra;
    long long B;
    double box;
    double* val;           // [B]
    double* x;             // [B][dim] or NULL
    int* status;           // [B]
    int* iters;            // [B] or NULL
    int max_iter;
};

__host__ __device__ inline int lp_warp_doubles(int dim) { return 2 * dim * dim + 5 * dim + 2; }

struct LPRow {
    const LPArgs& a;
    long long inst;
    int relax;
    __device__ __forceinline__ double coef(int r, int k) const {
        if (r < a.m) return a.HT[(size_t)k * a.mpad + r];
        r -= a.m;
        if (r < a.n_extra) return a.extra[((size_t)inst * a.n_extra + r) * (a.dim + 1) + k];
        r -= a.n_extra;
        return ((r >> 1) == k) ? ((r & 1) ? -1.0 : 1.0) : 0.0;
    }
    __device__ __forceinline__ double rhs(int r) const {
        if (r < a.m) return a.h[r] + (r == relax ? a.relax_by : 0.0);
        r -= a.m;
        if (r < a.n_extra) return a.extra[((size_t)inst * a.n_extra + r) * (a.dim + 1) + a.dim];
        return a.box;
    }
};

// Binv <- (rows bidx)^-1, x <- Binv h_B, lam <- Binv' c.  Returns false when the basis is singular.
__device__ bool lp_refactor(const LPRow& R, int d, int lane, int bidx, double* Binv, double* T, double* xs, const double* cs,
                            double* hB, double& lam) {
    // T = H_B (lane j < d writes row j), Binv = I
    if (lane < d) {
        for (int k = 0; k < d; ++k) { T[lane * d + k] = R.coef(bidx, k); Binv[lane * d + k] = (k == lane) ? 1.0 : 0.0; }
        hB[lane] = R.rhs(bidx);
    }
    __syncwarp();
    bool ok = true;
    for (int k = 0; k < d; ++k) {
        // partial pivoting over rows k..d-1 of column k
        double v = (lane >= k && lane < d) ? fabs(T[lane * d + k]) : -1.0;
        int idx = lane;
        v = warp_argmax(v, idx);
        if (!(v > 1e-13)) { ok = false; break; }
        if (idx != k && lane < d) {       // swap rows idx and k (lane = column)
            double t0 = T[k * d + lane]; T[k * d + lane] = T[idx * d + lane]; T[idx * d + lane] = t0;
            double t1 = Binv[k * d + lane]; Binv[k * d + lane] = Binv[idx * d + lane]; Binv[idx * d + lane] = t1;
        }
        __syncwarp();
        const double ip = 1.0 / T[k * d + k];
        __syncwarp();
        if (lane < d) { T[k * d + lane] *= ip; Binv[k * d + lane] *= ip; }
        __syncwarp();
        if (lane < d && lane != k) {      // lane = row to eliminate
            const double f = T[lane * d + k];
            if (f != 0.0)
                for (int c = 0; c < d; ++c) {
                    T[lane * d + c] = fma(-f, T[k * d + c], T[lane * d + c]);
                    Binv[lane * d + c] = fma(-f, Binv[k * d + c], Binv[lane * d + c]);
                }
        }
        __syncwarp();
    }
    if (!ok) return false;
    // Binv now holds (P H_B)^-1 P = H_B^-1 with columns in the ORIGINAL basis order (row swaps were applied to both sides)
    if (lane < d) {
        double acc = 0.0, l = 0.0;
        for (int j = 0; j < d; ++j) acc = fma(Binv[lane * d + j], hB[j], acc);
        xs[lane] = acc;
        for (int k = 0; k < d; ++k) l = fma(cs[k], Binv[k * d + lane], l);
        lam = l;
    }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(LP_WARPS * 32) lp_dual_simplex_kernel(LPArgs a) {
    extern __shared__ __align__(16) double lp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = a.dim;
    double* base = lp_smem + (size_t)warp * lp_warp_doubles(d);
    double *Binv = base, *T = base + d * d, *xs = T + d * d, *cs = xs + d, *rowp = cs + d, *us = rowp + d, *hB = us + d;
    const int total = a.m + a.n_extra + 2 * d;
    for (long long inst = (long long)blockIdx.x * LP_WARPS + warp; inst < a.B; inst += (long long)gridDim.x * LP_WARPS) {
        LPRow R{a, inst, a.relax_row ? a.relax_row[inst] : -1};
        double lam = 0.0;
        int bidx = 0;
        int status = 1, it = 0;
        // attempt 1 (only after a singular basis / the iteration limit): Bland's rule from the first step
        for (int attempt = 0; attempt < 2 && status == 1; ++attempt) {
        if (lane < d) {
            const double c = a.obj[(size_t)inst * d + lane];
            cs[lane] = c;
            const int neg = c < 0.0;
            bidx = a.m + a.n_extra + 2 * lane + neg;         // box row  +x_i <= box  or  -x_i <= box
            lam = fabs(c);
            xs[lane] = neg ? -a.box : a.box;
            for (int k = 0; k < d; ++k) Binv[lane * d + k] = (k == lane) ? (neg ? -1.0 : 1.0) : 0.0;
        }
        __syncwarp();
        int stalled = 0;
        bool fresh = true;
        const int it_end = it + a.max_iter;
        for (; it < it_end; ++it) {
            if (!fresh && (it & 31) == 0) {
                if (!lp_refactor(R, d, lane, bidx, Binv, T, xs, cs, hB, lam)) { status = 1; break; }
                fresh = true;
            }
            // ---- pricing ------------------------------------------------------------------------------------------
            const bool bland = attempt > 0 || stalled > d;
            double best = -1e300;
            int bp = 0x7fffffff;
            for (int r = lane; r < total; r += 32) {
                double v = -R.rhs(r);
                if (r < a.m) {
                    for (int k = 0; k < d; ++k) v = fma(a.HT[(size_t)k * a.mpad + r], xs[k], v);
                } else {
                    for (int k = 0; k < d; ++k) v = fma(R.coef(r, k), xs[k], v);
                }
                const double tol = 1e-9 * (1.0 + fabs(R.rhs(r)));
                if (bland) {
                    if (v > tol && r < bp) { bp = r; best = v; }
                } else if (v - tol > best) { best = v - tol; bp = r; }
            }
            int p;
            double viol;
            if (bland) {
                p = __reduce_min_sync(RTMPC_FULL_MASK, bp);
                viol = (p == 0x7fffffff) ? -1.0 : 1.0;
            } else {
                int idx = bp;
                viol = warp_argmax(best, idx);                   // violation beyond the tolerance
                p = idx;
            }
            if (!(viol > 0.0)) {
                if (fresh) { status = 0; break; }
                if (!lp_refactor(R, d, lane, bidx, Binv, T, xs, cs, hB, lam)) { status = 1; break; }
                fresh = true;
                continue;                                        // price again at the exact vertex
            }
            fresh = false;
            // exact violation of the entering row at x
            if (lane < d) rowp[lane] = R.coef(p, lane);
            __syncwarp();
            double vp = -R.rhs(p);
            for (int k = 0; k < d; ++k) vp = fma(rowp[k], xs[k], vp);
            // ---- ratio test (Harris: among the rows within a small tolerance of the smallest ratio, the largest pivot) -----
            double w = 0.0;
            if (lane < d) for (int k = 0; k < d; ++k) w = fma(rowp[k], Binv[k * d + lane], w);
            const double wmax = warp_max(lane < d ? fabs(w) : 0.0);
            const bool cand = lane < d && w > 1e-9 * (1.0 + wmax);
            if (!__any_sync(RTMPC_FULL_MASK, cand)) { status = 2; break; }      // no leaving row: the rows contradict each other
            const double lamp = fmax(lam, 0.0);
            const double lmax = warp_max(lane < d ? lamp : 0.0);
            const double ratio = cand ? lamp / w : 1e300;
            const double bound = warp_min(cand ? (lamp + 1e-9 * (1.0 + lmax)) / w : 1e300);
            int jl = lane;
            (void)warp_argmax((cand && ratio <= bound) ? w : -1e300, jl);
            const double theta = __shfl_sync(RTMPC_FULL_MASK, ratio, jl);
            const double wj = __shfl_sync(RTMPC_FULL_MASK, w, jl);
            stalled = (theta <= 1e-14 * (1.0 + lmax)) ? stalled + 1 : 0;
            if (lane < d) {
                lam = (lane == jl) ? theta : fmax(fma(-theta, w, lam), 0.0);
                us[lane] = Binv[lane * d + jl];
            }
            __syncwarp();
            if (lane < d) {
                xs[lane] = fma(-us[lane], vp / wj, xs[lane]);
                const double g = (w - (lane == jl ? 1.0 : 0.0)) / wj;      // lane = column
                for (int k = 0; k < d; ++k) Binv[k * d + lane] = fma(-us[k], g, Binv[k * d + lane]);
                if (lane == jl) bidx = p;
            }
            __syncwarp();
        }
        }
        // ---- outputs ----------------------------------------------------------------------------------------------
        double v = 0.0;
        if (lane < d) v = cs[lane] * xs[lane];
        v = warp_sum(v);
        // a box row with a positive multiplier in the final basis: the LP is unbounded in the polytope itself
        const bool onbox = lane < d && bidx >= a.m + a.n_extra && lam > 1e-9 * (1.0 + fabs(cs[lane]));
        if (status == 0 && __any_sync(RTMPC_FULL_MASK, onbox)) status = 4;
        if (lane == 0) {
            a.val[inst] = (status == 0) ? v : (status == 4 ? 1e300 : (status == 2 ? -1e300 : v));
            a.status[inst] = status;
            if (a.iters) a.iters[inst] = it;
        }
        if (a.x && lane < d) a.x[(size_t)inst * d + lane] = xs[lane];
        __syncwarp();
    }
}

}  // namespace rtmpc

using namespace rtmpc;

static int lp_fail(const char* what, cudaError_t e = cudaSuccess) { return rtmpc::set_last_error(what, e); }
#define LCU(call)                                                   \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return lp_fail(#call, _e);           \
    } while (0)

extern "C" int rtmpc_lp_solve(const double* d_HT, const double* d_h, int32_t m, int32_t mpad, int32_t dim, const double* d_obj,
                              const int32_t* d_relax_row, double relax_by, const double* d_extra, int32_t n_extra, int64_t B,
                              double box, double* d_val, double* d_x, int32_t* d_status, int32_t* d_iters, void* stream) {
    if (!d_HT || !d_h || !d_obj || !d_val || !d_status) return lp_fail("rtmpc_lp_solve: null argument");
    if (dim < 1 || dim > LP_MAX_DIM) return lp_fail("rtmpc_lp_solve: need 1 <= dim <= 16");
    if (m < 0 || mpad < m || n_extra < 0 || (n_extra > 0 && !d_extra)) return lp_fail("rtmpc_lp_solve: bad sizes");
    if (!(box > 0.0)) return lp_fail("rtmpc_lp_solve: box must be positive");
    if (B <= 0) return 0;
    LPArgs a;
    a.HT = d_HT; a.h = d_h; a.m = m; a.mpad = mpad; a.dim = dim; a.obj = d_obj; a.relax_row = d_relax_row; a.relax_by = relax_by;
    a.extra = d_extra; a.n_extra = n_extra; a.B = B; a.box = box; a.val = d_val; a.x = d_x; a.status = d_status; a.iters = d_iters;
    a.max_iter = 50 * (dim + 2) + (m + n_extra) / 2;
    int dev = 0, sms = 0;
    LCU(cudaGetDevice(&dev));
    LCU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (B + LP_WARPS - 1) / LP_WARPS;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    const size_t smem = (size_t)LP_WARPS * lp_warp_doubles(dim) * sizeof(double);
    lp_dual_simplex_kernel<<<(int)blocks, LP_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
    LCU(cudaGetLastError());
    return 0;
}

namespace {
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
    cudaError_t need(size_t bytes) {
        int cur = 0;
        cudaError_t e0 = cudaGetDevice(&cur);
        if (e0 != cudaSuccess) return e0;
        if (bytes <= cap && cur == dev) return cudaSuccess;
        if (p) { if (dev != cur) { cudaSetDevice(dev); cudaFree(p); cudaSetDevice(cur); } else cudaFree(p); }
        dev = cur; p = nullptr; cap = 0;
        const cudaError_t e = cudaMalloc(&p, bytes + bytes / 4 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
        return e;
    }
};
std::mutex g_lp_mu;
Buf g_lp[9];
}  // namespace

// Host buffers in and out (numpy arrays): H row-major [m*dim].  Copies, solves, copies back, synchronises.
extern "C" int rtmpc_lp_solve_host(const double* h_H, const double* h_h, int32_t m, int32_t dim, const double* h_obj,
                                   const int32_t* h_relax_row, double relax_by, const double* h_extra, int32_t n_extra,
                                   int64_t B, double box, double* h_val, double* h_x, int32_t* h_status, int32_t* h_iters) {
    if (!h_H || !h_h || !h_obj || !h_val || !h_status) return lp_fail("rtmpc_lp_solve_host: null argument");
    if (dim < 1 || dim > LP_MAX_DIM || m < 0) return lp_fail("rtmpc_lp_solve_host: need 1 <= dim <= 16, m >= 0");
    if (B <= 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return lp_fail("rtmpc_lp_solve_host: no CUDA device (there is no CPU fallback)");
    std::lock_guard<std::mutex> lock(g_lp_mu);
    const int mpad = (m + 31) & ~31;
    std::vector<double> HT((size_t)dim * (mpad > 0 ? mpad : 32), 0.0), hp(mpad > 0 ? mpad : 32, 0.0);
    const int mp = mpad > 0 ? mpad : 32;
    for (int r = 0; r < m; ++r) {
        for (int k = 0; k < dim; ++k) HT[(size_t)k * mp + r] = h_H[(size_t)r * dim + k];
        hp[r] = h_h[r];
    }
    const size_t bHT = HT.size() * 8, bh = hp.size() * 8, bo = (size_t)B * dim * 8, bv = (size_t)B * 8, bi = (size_t)B * 4;
    const size_t be = (size_t)B * n_extra * (dim + 1) * 8;
    LCU(g_lp[0].need(bHT)); LCU(g_lp[1].need(bh)); LCU(g_lp[2].need(bo)); LCU(g_lp[3].need(bv)); LCU(g_lp[4].need(bo));
    LCU(g_lp[5].need(bi)); LCU(g_lp[6].need(bi)); LCU(g_lp[7].need(bi)); LCU(g_lp[8].need(be ? be : 8));
    LCU(cudaMemcpy(g_lp[0].p, HT.data(), bHT, cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(g_lp[1].p, hp.data(), bh, cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(g_lp[2].p, h_obj, bo, cudaMemcpyHostToDevice));
    if (h_relax_row) LCU(cudaMemcpy(g_lp[7].p, h_relax_row, bi, cudaMemcpyHostToDevice));
    if (be) LCU(cudaMemcpy(g_lp[8].p, h_extra, be, cudaMemcpyHostToDevice));
    if (rtmpc_lp_solve((const double*)g_lp[0].p, (const double*)g_lp[1].p, m, mp, dim, (const double*)g_lp[2].p,
                       h_relax_row ? (const int32_t*)g_lp[7].p : nullptr, relax_by, be ? (const double*)g_lp[8].p : nullptr,
                       n_extra, B, box, (double*)g_lp[3].p, (double*)g_lp[4].p, (int32_t*)g_lp[5].p, (int32_t*)g_lp[6].p, nullptr))
        return -1;
    LCU(cudaMemcpy(h_val, g_lp[3].p, bv, cudaMemcpyDeviceToHost));
    if (h_x) LCU(cudaMemcpy(h_x, g_lp[4].p, bo, cudaMemcpyDeviceToHost));
    LCU(cudaMemcpy(h_status, g_lp[5].p, bi, cudaMemcpyDeviceToHost));
    if (h_iters) LCU(cudaMemcpy(h_iters, g_lp[6].p, bi, cudaMemcpyDeviceToHost));
    return 0;
}

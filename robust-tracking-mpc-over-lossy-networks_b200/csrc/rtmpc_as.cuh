// Batched dual active-set QP kernel (Goldfarb-Idnani on operators shared by the whole batch), FP64, sm_100a.
//
// One warp owns one instance of the condensed, equilibrated problem of rtmpc_ipm.cuh
//     min 1/2 z'Hs z + q'z   s.t.  lo <= G z <= up,   q = Fx x_init + Fr ref,  lo/up affine in x_init.
// Everything that depends only on (A, B, Q, R, sets) is prepared once per problem and shared by all
// instances:  Hinv,  Y = G Hinv,  W = G Hinv G'  and the parameter maps
//     z_u = Zx x_init + Zr ref  (unconstrained minimiser),   G z_u = Tx x_init + Tr ref.
// With those, a dual active-set step never touches H or G: for the working set A (signed rows n_a)
//     multipliers  lam_A = S^-1 (N_A' z_u - b_A),  S = N_A' Hinv N_A = signed sub-matrix of W,
//     adding the violated row p moves every row value by  -step * (s_p W[:,p] - W[:,A] (s_A r)),  r = S^-1 W[A,p].
// Per lane: the violations  vu = G z - up,  vl = lo - G z  of its R rows (row = slot*32 + lane) live
// in registers; per warp: the Cholesky factor of S (appended to on every add, rebuilt on a drop),
// multipliers and the active list live in shared memory.
//
// Per instance (mirrored by tools/as_model.py):
//   0. parameter rows; z_u; violations at z_u; none -> z_u is optimal.
//   1. warm start: last control step's certified active set moved one stage earlier (QPDev.shift),
//      equality solve, drop rows with negative multipliers until the set is dual feasible.
//   2. Goldfarb-Idnani: add the most violated row with primal/dual ratio test (partial steps drop
//      the blocking row) until no row is violated by more than 1e-11 (scaled).  Strictly increasing
//      dual objective: no cycling.  Linear dependence without a blocking row = primal infeasible.
//   3. certification: fresh multiplier solve with two steps of iterative refinement against the
//      true rows of G, z = z_u - Y_A' (s lam), every row and multiplier re-checked (KKT certificate,
//      same tolerances as the interior-point kernel's endgame).  A violated row found here sends
//      the instance back to 2 with exact row values.
// Anything unexpected (step cap, lost rank, negative multiplier at certification) returns
// RTMPC_FALLBACK; the C ABI then runs the interior-point kernel on those instances only.
#pragma once
#include "rtmpc_ipm.cuh"

namespace rtmpc {

constexpr int RTMPC_FALLBACK = RTMPC_FALLBACK_STATUS;

// per-warp shared memory, in doubles: S, zu, z, coef, lam, 16 parameters, act lists (2*npad ints)
__host__ __device__ inline int as_warp_doubles(const QPDev& P) {
    return P.npad * P.ss + 4 * P.npad + 16 + P.npad;
}
__host__ __device__ inline int as_block_doubles(const QPDev& P, bool g_in_smem) {
    return g_in_smem ? P.mpad * P.gs : 0;
}

struct ASWarp {
    double *S, *zu, *z, *coef, *lam, *xr;
    int *act_row, *act_sgn;
};

__device__ __forceinline__ ASWarp as_carve(double* base, const QPDev& P) {
    ASWarp w;
    w.S = base; base += P.npad * P.ss;
    w.zu = base; base += P.npad;
    w.z = base; base += P.npad;
    w.coef = base; base += P.npad;
    w.lam = base; base += P.npad;
    w.xr = base; base += 16;
    w.act_row = reinterpret_cast<int*>(base);
    w.act_sgn = w.act_row + P.npad;
    return w;
}

// forward substitution L y = r (inverse diagonal stored); lane k holds r_k / y_k
__device__ __forceinline__ double tri_fwd(const double* __restrict__ S, int ss, int na, int lane, double r) {
    for (int k = 0; k < na; ++k) {
        const double yk = __shfl_sync(RTMPC_FULL_MASK, r, k) * S[k * ss + k];
        if (lane == k) r = yk;
        else if (lane > k && lane < na) r = fma(-S[lane * ss + k], yk, r);
    }
    return (lane < na) ? r : 0.0;
}
// backward substitution L' x = y
__device__ __forceinline__ double tri_bwd(const double* __restrict__ S, int ss, int na, int lane, double r) {
    for (int k = na - 1; k >= 0; --k) {
        const double xk = __shfl_sync(RTMPC_FULL_MASK, r, k) * S[k * ss + k];
        if (lane == k) r = xk;
        else if (lane < k) r = fma(-S[k * ss + lane], xk, r);
    }
    return (lane < na) ? r : 0.0;
}

// S = signed sub-matrix of W on the active list (lower triangle), then its Cholesky factor with
// dependent rows neutralised.  Returns the mask of independent rows.
__device__ __forceinline__ unsigned as_factor(const QPDev& P, ASWarp& w, int na, int lane) {
    const int ss = P.ss;
    double* S = w.S;
    const bool mine = lane < na;
    if (mine) {
        const int ra = w.act_row[lane];
        const double sa = (double)w.act_sgn[lane];
        const double* Wa = P.W + (size_t)ra * P.mpad;
        for (int b = 0; b <= lane; ++b) S[lane * ss + b] = Wa[w.act_row[b]] * sa * (double)w.act_sgn[b];
    }
    __syncwarp();
    const double sdiag = mine ? S[lane * ss + lane] : 0.0;
    const double dmax = warp_max(sdiag);
    unsigned keep = 0;
    for (int k = 0; k < na; ++k) {
        double s = 0.0;
        if (lane >= k && mine) {
            s = S[lane * ss + k];
            double s2 = 0.0;
            int j = 0;
            for (; j + 1 < k; j += 2) {
                s = fma(-S[lane * ss + j], S[k * ss + j], s);
                s2 = fma(-S[lane * ss + j + 1], S[k * ss + j + 1], s2);
            }
            if (j < k) s = fma(-S[lane * ss + j], S[k * ss + j], s);
            s += s2;
        }
        const double pk = __shfl_sync(RTMPC_FULL_MASK, s, k);
        const double skk = __shfl_sync(RTMPC_FULL_MASK, sdiag, k);
        const bool good = (pk > 1e-11 * skk) && (pk > 1e-14 * dmax);
        if (good) {
            keep |= (1u << k);
            const double idk = __drcp_rn(sqrt(pk));
            if (lane >= k && mine) S[lane * ss + k] = (lane == k) ? idk : s * idk;
        } else {
            if (lane >= k && mine) S[lane * ss + k] = (lane == k) ? 1.0 : 0.0;
            if (lane == k) for (int j = 0; j < k; ++j) S[k * ss + j] = 0.0;
        }
        __syncwarp();
    }
    return keep;
}

// bound of the signed row (row, sgn) for this instance:  sgn > 0: up,  sgn < 0: -lo
__device__ __forceinline__ double as_bound(const QPDev& P, const double* __restrict__ xr, int row, int sgn) {
    const int nx = P.nx;
    if (sgn > 0) {
        double up = P.up0[row];
        for (int k = 0; k < nx; ++k) up = fma(P.Ux[row * nx + k], xr[k], up);
        return up;
    }
    double lo = P.lo0[row];
    for (int k = 0; k < nx; ++k) lo = fma(P.Lx[row * nx + k], xr[k], lo);
    return -lo;
}

// remove the entries flagged in `bad` from the active list (and lam), keeping the order
__device__ __forceinline__ int as_compact(ASWarp& w, int na, int lane, bool bad, double lamv) {
    const bool mine = lane < na;
    const unsigned good = __ballot_sync(RTMPC_FULL_MASK, mine && !bad);
    const int pos = __popc(good & ((1u << lane) - 1u));
    const int rr = mine ? w.act_row[lane] : 0, sg = mine ? w.act_sgn[lane] : 0;
    __syncwarp();
    if (mine && !bad) { w.act_row[pos] = rr; w.act_sgn[pos] = sg; w.lam[pos] = lamv; }
    __syncwarp();
    return __popc(good);
}

struct ASCounters { int steps, rounds; unsigned long long flops; };

// violations of every row move by  -sum_a coef_a W[row_a][row]  (coef in w.coef[0..na))
template <int R>
__device__ __forceinline__ void as_apply_rows(const QPDev& P, const ASWarp& w, int na, int lane, int nslots,
                                              double (&vu)[R], double (&vl)[R]) {
    for (int a = 0; a < na; ++a) {
        const double c = w.coef[a];
        const double* __restrict__ Wa = P.W + (size_t)w.act_row[a] * P.mpad + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nslots) {
                const double g = Wa[r * 32];
                vu[r] = fma(-c, g, vu[r]);
                vl[r] = fma(c, g, vl[r]);
            }
        }
    }
}

// Goldfarb-Idnani iteration.  Returns 0 when no row is violated by more than tolp.
template <int R>
__device__ __forceinline__ int as_gi(const QPDev& P, ASWarp& w, int& na, int lane, int nslots, double (&vu)[R],
                                     double (&vl)[R], unsigned& actu, unsigned& actl, double tolp, int max_steps,
                                     ASCounters& cnt) {
    const int ss = P.ss, n = P.n, mpad = P.mpad;
    double* S = w.S;
    while (true) {
        double best = -RTMPC_INF;
        int code = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nslots) {
                const int row = r * 32 + lane;
                if (!((actu >> r) & 1u) && vu[r] > best) { best = vu[r]; code = 2 * row; }
                if (!((actl >> r) & 1u) && vl[r] > best) { best = vl[r]; code = 2 * row + 1; }
            }
        }
        best = warp_argmax(best, code);
        if (best <= tolp) return 0;
        const int p = code >> 1;
        const double sp = (code & 1) ? -1.0 : 1.0;
        const double* __restrict__ Wp = P.W + (size_t)p * mpad;
        const double wpp = Wp[p];
        double cp = best, lam_p = 0.0;
        while (true) {
            if (cnt.steps >= max_steps) return RTMPC_FALLBACK;
            cnt.steps += 1;
            const bool mine = lane < na;
            const int ra = mine ? w.act_row[lane] : 0;
            const double sa = mine ? (double)w.act_sgn[lane] : 0.0;
            const double v = mine ? sa * sp * Wp[ra] : 0.0;
            const double l = tri_fwd(S, ss, na, lane, v);
            const double kappa = wpp - warp_sum(l * l);
            const double rr = tri_bwd(S, ss, na, lane, l);
            const bool dependent = !(kappa > 1e-11 * wpp);
            const double lam_a = mine ? w.lam[lane] : 0.0;
            const double rmax = warp_max(fabs(rr));
            double ratio = (mine && rr > 1e-13 * (1.0 + rmax)) ? lam_a / rr : RTMPC_INF;
            int j1 = lane;
            const double t1 = warp_argmin(ratio, j1);
            const bool has_j = t1 < 0.5 * RTMPC_INF;
            double step;
            bool full;
            if (dependent) {
                // n_p is a combination of the active rows: without a blocking multiplier the
                // constraints contradict each other (Farkas: y = (-r, 1) >= 0, N y = 0, b'y = -c_p < 0)
                if (!has_j) return (cp > 1e-6 * P.sc_b) ? RTMPC_INFEASIBLE : RTMPC_FALLBACK;
                step = t1;
                full = false;
            } else {
                const double t2 = cp / kappa;
                full = !(has_j && t1 < t2);
                step = full ? t2 : t1;
            }
            if (mine) w.lam[lane] = (!full && lane == j1) ? 0.0 : fma(-step, rr, lam_a);
            lam_p += step;
            cnt.flops += 4ull * na * na + 2ull * na;
            if (!dependent) {
                if (lane < P.npad) w.coef[lane] = mine ? -step * sa * rr : 0.0;
                __syncwarp();
                const double c = step * sp;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < nslots) {
                        const double g = Wp[r * 32 + lane];
                        vu[r] = fma(-c, g, vu[r]);
                        vl[r] = fma(c, g, vl[r]);
                    }
                }
                as_apply_rows<R>(P, w, na, lane, nslots, vu, vl);
                cp = fma(-step, kappa, cp);
                cnt.flops += 2ull * P.m * (na + 1);
                __syncwarp();
            }
            if (full) {
                if (na >= n) return RTMPC_FALLBACK;
                // the factor grows by one row: [l', sqrt(kappa)]
                if (mine) S[na * ss + lane] = l;
                if (lane == 0) {
                    S[na * ss + na] = __drcp_rn(sqrt(kappa));
                    w.act_row[na] = p;
                    w.act_sgn[na] = (int)sp;
                    w.lam[na] = lam_p;
                }
                if (lane == (p & 31)) {
                    if (sp > 0) actu |= 1u << (p >> 5); else actl |= 1u << (p >> 5);
                }
                na += 1;
                __syncwarp();
                break;
            }
            // partial step: the blocking row leaves the working set
            {
                const int rowj = w.act_row[j1], sgj = w.act_sgn[j1];
                if (lane == (rowj & 31)) {
                    if (sgj > 0) actu &= ~(1u << (rowj >> 5)); else actl &= ~(1u << (rowj >> 5));
                }
                const double lamv = mine ? w.lam[lane] : 0.0;
                na = as_compact(w, na, lane, lane == j1, lamv);
                const unsigned keep = as_factor(P, w, na, lane);
                cnt.flops += (unsigned long long)na * na * na / 3;
                if (keep != ((na >= 32) ? 0xffffffffu : ((1u << na) - 1u))) return RTMPC_FALLBACK;
            }
        }
    }
}

// Certification on the working set: refined multipliers, z, exact row values.  Returns 0 when the
// KKT conditions hold, 1 when a row is still violated (vu/vl hold exact values: go back to as_gi),
// 2 on a negative multiplier.
template <int R>
__device__ __forceinline__ int as_certify(const QPDev& P, const double* __restrict__ Gs, ASWarp& w, int na, int lane,
                                          int nslots, double (&vu)[R], double (&vl)[R], unsigned actu, unsigned actl,
                                          double tolp, ASCounters& cnt) {
    const int n = P.n, npad = P.npad, gs = P.gs, ss = P.ss, nx = P.nx;
    const bool mine = lane < na;
    const int ra = mine ? w.act_row[lane] : 0;
    const int sgi = mine ? w.act_sgn[lane] : 1;
    const double sa = mine ? (double)sgi : 0.0;
    const double ba = mine ? as_bound(P, w.xr, ra, sgi) : 0.0;
    const double* __restrict__ Ga = Gs + (size_t)ra * gs;
    double resid = 0.0;
    if (mine) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = fma(Ga[k], w.zu[k], acc);
        resid = sa * acc - ba;
    }
    double lam = 0.0;
    double zj = (lane < n) ? w.zu[lane] : 0.0;
    for (int pass = 0; pass < 3; ++pass) {
        const double dl = tri_bwd(w.S, ss, na, lane, tri_fwd(w.S, ss, na, lane, resid));
        lam += dl;
        if (lane < npad) w.coef[lane] = dl * sa;
        __syncwarp();
        if (lane < n) {
            for (int a = 0; a < na; ++a) zj = fma(-w.coef[a], P.Y[(size_t)w.act_row[a] * npad + lane], zj);
        }
        if (lane < npad) w.z[lane] = zj;
        __syncwarp();
        if (pass < 2) {
            resid = 0.0;
            if (mine) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc = fma(Ga[k], w.z[k], acc);
                resid = sa * acc - ba;
            }
        }
    }
    cnt.flops += 3ull * (4ull * na * na + 4ull * na * n) + 2ull * P.m * n;
    cnt.rounds += 1;
    // exact row values at z
    double t[R];
    gemv_rows<R>(Gs, gs, npad, nslots, w.z, lane, t);
    double worst = -RTMPC_INF;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nslots) {
            const int row = r * 32 + lane;
            double lo = P.lo0[row], up = P.up0[row];
            for (int k = 0; k < nx; ++k) {
                lo = fma(P.Lx[row * nx + k], w.xr[k], lo);
                up = fma(P.Ux[row * nx + k], w.xr[k], up);
            }
            vu[r] = P.has_up[row] ? t[r] - up : -RTMPC_INF;
            vl[r] = P.has_lo[row] ? lo - t[r] : -RTMPC_INF;
            if (!((actu >> r) & 1u)) worst = fmax(worst, vu[r]);
            if (!((actl >> r) & 1u)) worst = fmax(worst, vl[r]);
        }
    }
    worst = warp_max(worst);
    const double lmin = warp_min(mine ? lam : RTMPC_INF);
    const double lmaxabs = warp_max(mine ? fabs(lam) : 0.0);
    if (mine) w.lam[lane] = fmax(lam, 0.0);
    __syncwarp();
    if (lmin < -1e-9 * (1.0 + lmaxabs)) return 2;
    return (worst > tolp) ? 1 : 0;
}

// One instance, one warp.  x_init / ref: this instance's parameters (ref may be NULL); warm_inst: this
// instance's warm-start record (npad + 1 ints) or NULL; z_out_inst / U_out_inst: this instance's
// outputs or NULL.  Returns the status; on RTMPC_FALLBACK nothing has been written.
template <int R>
__device__ __forceinline__ int as_solve_instance(const QPDev& P, const double* __restrict__ Gs, ASWarp& w, int lane,
                                                 const double* x_init, const double* ref, int* warm_inst,
                                                 double* z_out_inst, double* U_out_inst, ASCounters& cnt) {
    const int n = P.n, npad = P.npad, mpad = P.mpad, nx = P.nx, ss = P.ss;
    const int nslots = mpad >> 5;
    const double tolp = 1e-11 * P.sc_b;
    if (lane < nx) {
        w.xr[lane] = x_init[lane];
        w.xr[8 + lane] = ref ? ref[lane] : 0.0;
    }
    __syncwarp();
    bool par_bad = false;
    for (int i = lane; i < P.np; i += 32) {
        double acc = -P.parh[i];
        for (int k = 0; k < nx; ++k) acc = fma(P.parC[i * nx + k], w.xr[k], acc);
        if (acc > 1e-9 * (1.0 + fabs(P.parh[i]))) par_bad = true;
    }
    par_bad = __any_sync(RTMPC_FULL_MASK, par_bad);
    double zuj = 0.0;
    if (lane < n) {
        for (int k = 0; k < nx; ++k) {
            zuj = fma(P.Zx[lane * nx + k], w.xr[k], zuj);
            zuj = fma(P.Zr[lane * nx + k], w.xr[8 + k], zuj);
        }
    }
    if (lane < npad) { w.zu[lane] = zuj; w.z[lane] = zuj; }
    // row values and violations at z_u
    double vu[R], vl[R];
    double vmax = -RTMPC_INF;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        vu[r] = -RTMPC_INF; vl[r] = -RTMPC_INF;
        if (r < nslots) {
            const int row = r * 32 + lane;
            double t = 0.0, lo = P.lo0[row], up = P.up0[row];
            for (int k = 0; k < nx; ++k) {
                t = fma(P.Tx[row * nx + k], w.xr[k], t);
                t = fma(P.Tr[row * nx + k], w.xr[8 + k], t);
                lo = fma(P.Lx[row * nx + k], w.xr[k], lo);
                up = fma(P.Ux[row * nx + k], w.xr[k], up);
            }
            if (P.has_up[row]) vu[r] = t - up;
            if (P.has_lo[row]) vl[r] = lo - t;
            vmax = fmax(vmax, fmax(vu[r], vl[r]));
        }
    }
    vmax = warp_max(vmax);
    cnt.flops += 2ull * n * 2 * nx + 2ull * P.m * 4 * nx;
    __syncwarp();

    int status = RTMPC_OPTIMAL, na = 0;
    unsigned actu = 0, actl = 0;
    if (par_bad) status = RTMPC_INFEASIBLE;
    else if (vmax < 0.0) status = RTMPC_OPTIMAL;      // the unconstrained minimiser is feasible
    else {
        // ---- 1. warm start ------------------------------------------------------------------
        if (warm_inst && P.shift) {
            const int wn = warm_inst[0];
            if (wn > 0) {
                int srow = -1, ssg = 1;
                if (lane < wn && lane < npad) {
                    const int code = warm_inst[1 + lane];
                    const int row = code >> 1;
                    ssg = (code & 1) ? -1 : 1;
                    srow = (row >= 0 && row < mpad) ? P.shift[row] : -1;
                    if (srow >= 0 && !(ssg > 0 ? P.has_up[srow] : P.has_lo[srow])) srow = -1;
                }
                const unsigned okm = __ballot_sync(RTMPC_FULL_MASK, srow >= 0);
                const int pos = __popc(okm & ((1u << lane) - 1u));
                if (srow >= 0) { w.act_row[pos] = srow; w.act_sgn[pos] = ssg; }
                na = __popc(okm);
                __syncwarp();
                while (na > 0) {
                    const unsigned keep = as_factor(P, w, na, lane);
                    const bool mine = lane < na;
                    const bool kept = mine && ((keep >> lane) & 1u);
                    double rhs = 0.0;
                    if (kept) {
                        const int ra = w.act_row[lane], sg = w.act_sgn[lane];
                        double t = 0.0;
                        for (int k = 0; k < nx; ++k) {
                            t = fma(P.Tx[ra * nx + k], w.xr[k], t);
                            t = fma(P.Tr[ra * nx + k], w.xr[8 + k], t);
                        }
                        rhs = (double)sg * t - as_bound(P, w.xr, ra, sg);
                    }
                    double lamv = tri_bwd(w.S, ss, na, lane, tri_fwd(w.S, ss, na, lane, rhs));
                    if (!kept) lamv = 0.0;
                    const double lmaxabs = warp_max(fabs(lamv));
                    const bool bad = mine && (!kept || lamv < -1e-9 * (1.0 + lmaxabs));
                    cnt.flops += (unsigned long long)na * na * na / 3 + 4ull * na * na;
                    cnt.steps += 1;
                    if (!__any_sync(RTMPC_FULL_MASK, bad)) {
                        if (mine) w.lam[lane] = fmax(lamv, 0.0);
                        __syncwarp();
                        break;
                    }
                    na = as_compact(w, na, lane, bad, lamv);
                }
                if (na > 0) {
                    if (lane < npad) w.coef[lane] = (lane < na) ? (double)w.act_sgn[lane] * w.lam[lane] : 0.0;
                    __syncwarp();
                    as_apply_rows<R>(P, w, na, lane, nslots, vu, vl);
                    cnt.flops += 2ull * P.m * na;
                    for (int a = 0; a < na; ++a) {
                        const int row = w.act_row[a];
                        if (lane == (row & 31)) {
                            if (w.act_sgn[a] > 0) actu |= 1u << (row >> 5); else actl |= 1u << (row >> 5);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        // ---- 2./3. Goldfarb-Idnani, then certification -----------------------------------------
        const int max_steps = 8 * npad + 32;
        for (int refresh = 0;; ++refresh) {
            status = as_gi<R>(P, w, na, lane, nslots, vu, vl, actu, actl, tolp, max_steps, cnt);
            if (status != 0) break;
            const int c = as_certify<R>(P, Gs, w, na, lane, nslots, vu, vl, actu, actl, tolp, cnt);
            if (c == 0) { status = RTMPC_OPTIMAL; break; }
            if (c == 2 || refresh >= 3) { status = RTMPC_FALLBACK; break; }
        }
    }

    // ---- outputs ------------------------------------------------------------------------------
    if (status != RTMPC_FALLBACK) {
        const bool has_sol = (status == RTMPC_OPTIMAL);
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        if (lane < npad) w.coef[lane] = (lane < n) ? w.z[lane] * P.D[lane] : 0.0;   // unscaled decision
        __syncwarp();
        const int nu = P.nu, N = P.N;
        const int ou = nx * (N + 1);
        const int first = z_out_inst ? 0 : ou;              // only the input rows are needed for the packet
        for (int i = first + lane; i < P.nz; i += 32) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(P.Phi[(size_t)i * npad + k], w.coef[k], acc);
            for (int k = 0; k < nx; ++k) acc = fma(P.Psi[(size_t)i * nx + k], w.xr[k], acc);
            if (!has_sol) acc = nanv;
            if (z_out_inst) z_out_inst[i] = acc;
            if (U_out_inst && i >= ou && i < ou + N * nu) U_out_inst[i - ou] = acc;
        }
        if (U_out_inst && P.nss > 0) {
            // last column of the packet: u_bar + K x_bar
            const int oxb = ou + N * nu;
            double val = 0.0;
            if (lane < nx + nu) {
                const int i = oxb + lane;
                for (int k = 0; k < n; ++k) val = fma(P.Phi[(size_t)i * npad + k], w.coef[k], val);
                for (int k = 0; k < nx; ++k) val = fma(P.Psi[(size_t)i * nx + k], w.xr[k], val);
            }
            for (int j = 0; j < nu; ++j) {
                double acc = __shfl_sync(RTMPC_FULL_MASK, val, nx + j);
                for (int k = 0; k < nx; ++k) acc = fma(P.Kss[j * nx + k], __shfl_sync(RTMPC_FULL_MASK, val, k), acc);
                if (lane == 0) U_out_inst[N * nu + j] = has_sol ? acc : nanv;
            }
        }
        cnt.flops += 2ull * (P.nz - first) * (n + nx);
    }
    if (warm_inst) {
        const int nw = (status == RTMPC_OPTIMAL) ? na : -1;
        if (lane == 0) warm_inst[0] = nw;
        if (lane < nw) warm_inst[1 + lane] = 2 * w.act_row[lane] + (w.act_sgn[lane] < 0 ? 1 : 0);
    }
    __syncwarp();
    return status;
}

// G staged on chip (or the padded global copy), and the per-warp scratch
__device__ __forceinline__ const double* as_stage(const QPDev& P, double* smem, int g_in_smem, double** wbase) {
    if (!g_in_smem) { *wbase = smem; return P.Gpad; }
    const int npad = P.npad, mpad = P.mpad;
    for (int idx = threadIdx.x; idx < mpad * npad; idx += blockDim.x) {
        int i = idx / npad, j = idx - i * npad;
        smem[i * P.gs + j] = P.G[idx];
    }
    for (int idx = threadIdx.x; idx < mpad * 2; idx += blockDim.x) smem[(size_t)(idx >> 1) * P.gs + npad + (idx & 1)] = 0.0;
    __syncthreads();
    *wbase = smem + (size_t)mpad * P.gs;
    return smem;
}

__device__ __forceinline__ int as_pack_iters(const ASCounters& cnt) {
    return ((cnt.steps > 4095 ? 4095 : cnt.steps) << 12) | ((cnt.rounds > 255 ? 255 : cnt.rounds) << 24);
}

template <int R, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
as_solve_kernel(QPDev P, int B, const double* __restrict__ x_init, const double* __restrict__ ref,
                const int* __restrict__ sel, int sel_value, double* __restrict__ z_out, double* __restrict__ U_out,
                int* __restrict__ status_out, int* __restrict__ iters_out, int* __restrict__ warm,
                unsigned long long* __restrict__ work, int g_in_smem) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nx = P.nx;
    double* wbase;
    const double* Gs = as_stage(P, smem, g_in_smem, &wbase);
    ASWarp w = as_carve(wbase + (size_t)warp * as_warp_doubles(P), P);
    const size_t usz = (size_t)(P.N + 1) * P.nu;
    for (int inst = blockIdx.x * wpb + warp; inst < B; inst += gridDim.x * wpb) {
        if (sel && sel[inst] != sel_value) continue;
        ASCounters cnt;
        cnt.steps = 0; cnt.rounds = 0; cnt.flops = 0;
        const int status = as_solve_instance<R>(P, Gs, w, lane, x_init + (size_t)inst * nx,
                                                ref ? ref + (size_t)inst * nx : nullptr,
                                                warm ? warm + (size_t)inst * (P.npad + 1) : nullptr,
                                                z_out ? z_out + (size_t)inst * P.nz : nullptr,
                                                U_out ? U_out + inst * usz : nullptr, cnt);
        if (lane == 0) {
            if (status_out) status_out[inst] = status;
            if (iters_out) iters_out[inst] = as_pack_iters(cnt);
            if (work) atomicAdd(work, cnt.flops);
        }
        __syncwarp();
    }
}

}  // namespace rtmpc

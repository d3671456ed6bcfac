// Batched primal-dual interior-point QP kernel with an active-set endgame (FP64, sm_100a).
//
// One warp owns one instance of
//     min 1/2 z'Hs z + q'z   s.t.  lo <= G z <= up,      q = Fx x_init + Fr ref,  lo/up affine in x_init
// (the condensed, equilibrated form of the reference's cvxpy problems, see condense.py).  Hs and G
// are shared by the whole batch and staged once per CTA in shared memory; the per-instance
// slack / multiplier state lives in registers (row i = slot*32 + lane, R slots per lane).
//
// Algorithm per instance (mirrored step for step by tools/ipm_model.py):
//   0. parameter rows, q, bounds; start at the unconstrained minimiser z_u = -Hinv q; if it is
//      feasible it is optimal -> done (0 iterations).
//   1. Mehrotra predictor-corrector on the reduced system  (Hs + G' diag(d) G) dz = rhs,
//      d = lam_u/s_u + lam_l/s_l (one G row per two-sided constraint), until a moderate tolerance.
//   2. Active-set endgame: solve the equality-constrained QP on {lam > s} through the Schur
//      complement  (G_A Hinv G_A') lam = G_A z_u - b_A  with a dependency-dropping Cholesky and
//      iterative refinement, then add the most violated row / drop the most negative multiplier
//      until the KKT conditions hold to 1e-9.  This is what makes the answer match the exact
//      minimiser: the MPC Hessian is so flat along late-horizon inputs (cond ~1e7) that an
//      interior-point iterate with a 1e-12 relative gap can still be 1e-3 away in those directions.
//   3. If the endgame does not certify, continue the interior-point iteration to its numerical
//      floor and try once more; report RTMPC_OPTIMAL_INACCURATE if that fails too.
#pragma once
#include "rtmpc_common.cuh"
#include "../../include/rtmpc.h"

namespace rtmpc {

struct QPDev {
    int nx, nu, N, n, npad, m, mpad, np, nz, nss;
    int gs;   // shared-memory row stride of G (npad + 2: 16-byte aligned rows, odd number of 16B units)
    int ss;   // row stride of the per-warp n x n work matrix (npad + 1)
    int va_len;  // max(mpad, nz)
    int mtot;    // number of finite bounds
    const double *Hs, *Hinv, *G, *Y, *Fx, *Fr, *lo0, *up0, *Lx, *Ux;
    const unsigned char *has_lo, *has_up;
    const double *parC, *parh, *D, *Phi, *Psi, *Kss;
    // operators of the active-set kernel (rtmpc_as.cuh), derived once per problem in rtmpc_qp_create
    const double* W;         // [mpad*mpad]  G Hinv G'
    const double* GT;        // [npad*mpad]  G transposed (coalesced row-value recomputation)
    const double *Zx, *Zr;   // [npad*nx]    z_u = Zx x_init + Zr ref    (= -Hinv Fx, -Hinv Fr)
    const double *ExT, *TrT; // [nx*mpad]    G z_u - up = Ex x_init + Tr ref - up0  (Ex = Tx - Ux), transposed
    const double *UxT, *LxT; // [nx*mpad]    Ux, Lx transposed
    const double *upI, *loI; // [mpad]       up0 / lo0 with +-1e30 where the row has no such bound
    const double* wid;       // [mpad]       up - lo (independent of x_init), 3e30 where the row has no lower bound
    const double* kap;       // [mpad]       bound per unit of (1 + |x|_1 + |ref|_1 + sum |multipliers|) on the difference between a
                             //              row value from G' z and through the factored tables (rtmpc_qp_create)
    const int* Uidx;               // [(N+1)nu] or NULL: payload row i = Ucoef[i] * z[Uidx[i]] (every row of UPhi has one entry,
    const double* Ucoef;           //                    UPsi = 0: inputs are decision variables; else the dense maps below)
    const double *UPhiT, *UPsiT;   // [npad*(N+1)nu], [nx*(N+1)nu]  packet payload from the scaled decision:
                                   // U_t = UPhi z + UPsi x_init, last column u_bar + K x_bar folded in
    const int* shift;        // [mpad] warm-start map: same constraint one stage earlier, -1 = none
    double s_floor, sc_b;
    int max_iter;
    int as_max_steps;     // active-set steps (rows added + dropped) before an instance is handed to the interior-point kernel
};

// per-warp shared memory, in doubles
__host__ __device__ inline int ipm_warp_doubles(const QPDev& P) {
    return 9 * P.npad + 16 + P.npad * P.ss + 2 * P.va_len + 2 * P.mpad + 4 * P.npad /*act lists as ints*/;
}
__host__ __device__ inline int ipm_block_doubles(const QPDev& P) { return P.mpad * P.gs + P.npad * P.npad; }

template <int R>
struct RowRegs {
    double su[R], sl[R], lu[R], ll[R], t[R], ta[R], tz[R];
};

// out[r] = sum_j G[row(r)][j] * vec[j]
template <int R>
__device__ __forceinline__ void gemv_rows(const double* __restrict__ Gs, int gs, int npad, int nslots,
                                          const double* __restrict__ vec, int lane, double (&out)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = 0.0;
    const double* __restrict__ grow = Gs + lane * gs;
    const int rstride = 32 * gs;
    for (int j = 0; j < npad; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(vec + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nslots) {
                const double2 g = *reinterpret_cast<const double2*>(grow + r * rstride + j);
                out[r] = fma(g.x, v.x, out[r]);
                out[r] = fma(g.y, v.y, out[r]);
            }
        }
    }
}

// lane j:  o1 = sum_i G[i][j] a[i],  o2 = sum_i G[i][j] b[i]   (b may be nullptr)
__device__ __forceinline__ void gemv_cols2(const double* __restrict__ Gs, int gs, int npad, int m,
                                           const double* __restrict__ a, const double* __restrict__ b,
                                           int lane, double& o1, double& o2) {
    double s0 = 0, s1 = 0, u0 = 0, u1 = 0;
    const bool on = lane < npad;
    const double* g = Gs + (on ? lane : 0);
    int i = 0;
    if (b) {
        for (; i + 1 < m; i += 2) {
            double g0 = g[i * gs], g1 = g[(i + 1) * gs];
            s0 = fma(g0, a[i], s0);
            u0 = fma(g0, b[i], u0);
            s1 = fma(g1, a[i + 1], s1);
            u1 = fma(g1, b[i + 1], u1);
        }
        if (i < m) { double g0 = g[i * gs]; s0 = fma(g0, a[i], s0); u0 = fma(g0, b[i], u0); }
    } else {
        double s2 = 0, s3 = 0;
        for (; i + 3 < m; i += 4) {
            s0 = fma(g[i * gs], a[i], s0);
            s1 = fma(g[(i + 1) * gs], a[i + 1], s1);
            s2 = fma(g[(i + 2) * gs], a[i + 2], s2);
            s3 = fma(g[(i + 3) * gs], a[i + 3], s3);
        }
        for (; i < m; ++i) s0 = fma(g[i * gs], a[i], s0);
        s0 += s2; s1 += s3;
    }
    o1 = on ? (s0 + s1) : 0.0;
    o2 = on ? (u0 + u1) : 0.0;
}

// S = Hs + sum_i d_i g_i g_i'   -- lane owns one BS x BS block (a >= b) of the symmetric matrix
template <int BS>
__device__ __forceinline__ void form_schur(const double* __restrict__ Gs, int gs, int m, int npad,
                                           const double* __restrict__ Hs_s, const double* __restrict__ dvec,
                                           double* __restrict__ S, int ss, int n, int lane, int blk_a, int blk_b,
                                           bool blk_on) {
    double acc[BS][BS];
#pragma unroll
    for (int x = 0; x < BS; ++x)
#pragma unroll
        for (int y = 0; y < BS; ++y) acc[x][y] = 0.0;
    if (blk_on) {
        const double* ga = Gs + blk_a * BS;
        const double* gb = Gs + blk_b * BS;
        for (int i = 0; i < m; ++i) {
            const double di = dvec[i];
            double av[BS], bv[BS];
#pragma unroll
            for (int x = 0; x < BS; ++x) { av[x] = di * ga[i * gs + x]; bv[x] = gb[i * gs + x]; }
#pragma unroll
            for (int x = 0; x < BS; ++x)
#pragma unroll
                for (int y = 0; y < BS; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < BS; ++x)
#pragma unroll
            for (int y = 0; y < BS; ++y) {
                int row = blk_a * BS + x, col = blk_b * BS + y;
                if (row < n && col < n) {
                    double v = acc[x][y] + Hs_s[row * npad + col];
                    S[row * ss + col] = v;
                    S[col * ss + row] = v;
                }
            }
    }
}

// In-place lower Cholesky of the n x n matrix S (lanes = rows, left-looking).  Returns false on a
// non-positive pivot (after a tiny relative regularisation).
__device__ __forceinline__ bool chol_warp(double* __restrict__ S, int ss, int n, int lane) {
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        double s = 0.0;
        if (lane >= k && lane < n) {
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
        double pk = __shfl_sync(RTMPC_FULL_MASK, s, k);
        if (!(pk > 0.0)) { ok = false; pk = 1e-300; }
        // the diagonal stores 1/L_kk: the substitutions multiply instead of divide
        const double idk = __drcp_rn(sqrt(pk));
        if (lane >= k && lane < n) S[lane * ss + k] = (lane == k) ? idk : s * idk;
        __syncwarp();
    }
    return ok;
}

// x = (L L')^{-1} rhs ; lane j holds rhs_j on entry and x_j on exit.
__device__ __forceinline__ double chol_solve_warp(const double* __restrict__ S, int ss, int n, int lane, double r) {
    for (int k = 0; k < n; ++k) {                      // forward: L y = r
        double yk = __shfl_sync(RTMPC_FULL_MASK, r, k) * S[k * ss + k];
        if (lane == k) r = yk;
        else if (lane > k && lane < n) r = fma(-S[lane * ss + k], yk, r);
    }
    for (int k = n - 1; k >= 0; --k) {                 // backward: L' x = y
        double xk = __shfl_sync(RTMPC_FULL_MASK, r, k) * S[k * ss + k];
        if (lane == k) r = xk;
        else if (lane < k) r = fma(-S[k * ss + lane], xk, r);
    }
    return (lane < n) ? r : 0.0;
}

struct WarpSmem {
    double *zeta, *zu, *q, *hz, *rhs, *dz, *best, *xr, *S, *va, *vb, *vlo, *vup;
    int *act_row, *act_sgn, *best_row, *best_sgn;
};

__device__ __forceinline__ WarpSmem carve(double* base, const QPDev& P) {
    WarpSmem w;
    w.zeta = base; base += P.npad;
    w.zu = base; base += P.npad;
    w.q = base; base += P.npad;
    w.hz = base; base += P.npad;
    w.rhs = base; base += P.npad;
    w.dz = base; base += P.npad;
    w.best = base; base += P.npad;
    w.xr = base; base += 16;
    base += P.npad;  // spare (keeps 16-byte alignment of what follows irrespective of npad parity)
    w.S = base; base += P.npad * P.ss;
    if (((P.npad * P.ss) & 1) != 0) base += 1;
    w.va = base; base += P.va_len;
    w.vb = base; base += P.va_len;
    w.vlo = base; base += P.mpad;
    w.vup = base; base += P.mpad;
    w.act_row = reinterpret_cast<int*>(base);
    w.act_sgn = w.act_row + 2 * P.npad;
    w.best_row = w.act_sgn + 2 * P.npad;
    w.best_sgn = w.best_row + 2 * P.npad;
    return w;
}

// ----------------------------------------------------------------------------------------------
// Active-set endgame.  On entry act_row/act_sgn[0..na) hold the estimate; on success w.zeta holds
// the certified minimiser (scaled coordinates).  Returns true on success.
// ----------------------------------------------------------------------------------------------
template <int R>
__device__ bool polish_warp(const QPDev& P, const double* __restrict__ Gs, WarpSmem& w, int na, int lane,
                            unsigned mask_u, unsigned mask_l, int nslots, int max_rounds, int* rounds_out,
                            int* na_out) {
    const int n = P.n, ss = P.ss, gs = P.gs, npad = P.npad;
    double* S = w.S;
    const double tol_p = 1e-11 * P.sc_b;   // a row left inactive may be violated by at most this (scaled units)
    int rounds = 0;
    bool success = false;
    for (; rounds < max_rounds; ++rounds) {
        if (na > npad) break;      // more rows than unknowns is fine (a degenerate vertex): the factorisation drops the dependent ones
        const bool mine = lane < na;
        const int ra = mine ? w.act_row[lane] : 0;
        const double sa = mine ? (double)w.act_sgn[lane] : 0.0;
        const double ba = mine ? (sa > 0 ? w.vup[ra] : -w.vlo[ra]) : 0.0;
        // S_A row a, and r_a = sa * G_a z_u - b_a
        double r0 = 0.0;
        if (mine) {
            const double* Wa = P.W + (size_t)ra * P.mpad;
            for (int b = 0; b < na; ++b) S[lane * ss + b] = Wa[w.act_row[b]] * sa * (double)w.act_sgn[b];
            const double* Ga = Gs + (size_t)ra * gs;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(Ga[k], w.zu[k], acc);
            r0 = sa * acc - ba;
        }
        __syncwarp();
        // dependency-dropping Cholesky (lanes = rows)
        double dmax = warp_max(mine ? S[lane * ss + lane] : 0.0);
        const double sdiag = mine ? S[lane * ss + lane] : 1.0;
        unsigned keep = 0;
        for (int k = 0; k < na; ++k) {
            double s = 0.0;
            if (lane >= k && mine) {
                s = S[lane * ss + k];
                for (int j = 0; j < k; ++j) s = fma(-S[lane * ss + j], S[k * ss + j], s);
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
        const bool kept = mine && ((keep >> lane) & 1u);
        // lam = S^{-1} r, z = z_u - sum_a lam_a sa Y_a ; two refinement passes on the true residual
        double lam = 0.0;
        double zj = (lane < n) ? w.zu[lane] : 0.0;
        double resid = kept ? r0 : 0.0;
        for (int pass = 0; pass < 3; ++pass) {
            double dl = chol_solve_warp(S, ss, na, lane, resid);
            if (!kept) dl = 0.0;
            lam += dl;
            if (lane < npad) w.rhs[lane] = dl * sa;      // signed step, indexed by active position
            __syncwarp();
            if (lane < n) {
                for (int a = 0; a < na; ++a) zj = fma(-w.rhs[a], P.Y[(size_t)w.act_row[a] * npad + lane], zj);
            }
            if (lane < npad) w.dz[lane] = zj;
            __syncwarp();
            if (pass < 2) {
                resid = 0.0;
                if (kept) {
                    const double* Ga = Gs + (size_t)ra * gs;
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc = fma(Ga[k], w.dz[k], acc);
                    resid = sa * acc - ba;
                }
            }
            __syncwarp();
        }
        // feasibility of every row at z = w.dz
        double t[R];
        gemv_rows<R>(Gs, gs, npad, nslots, w.dz, lane, t);
        double vworst = -RTMPC_INF;
        int iworst = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nslots) {
                const int row = r * 32 + lane;
                if ((mask_u >> r) & 1u) { double v = t[r] - w.vup[row]; if (v > vworst) { vworst = v; iworst = 2 * row; } }
                if ((mask_l >> r) & 1u) { double v = w.vlo[row] - t[r]; if (v > vworst) { vworst = v; iworst = 2 * row + 1; } }
            }
        }
        vworst = warp_argmax(vworst, iworst);
        int imin = lane;
        double lmin = warp_argmin(kept ? lam : RTMPC_INF, imin);
        const double lmaxabs = warp_max(kept ? fabs(lam) : 0.0);
        if (lmin < -1e-9 * (1.0 + lmaxabs)) {
            // drop the most negative multiplier (compact the list)
            int rr = 0, sg = 0;
            if (lane >= imin && lane + 1 < na) { rr = w.act_row[lane + 1]; sg = w.act_sgn[lane + 1]; }
            __syncwarp();
            if (lane >= imin && lane + 1 < na) { w.act_row[lane] = rr; w.act_sgn[lane] = sg; }
            na -= 1;
            __syncwarp();
            continue;
        }
        if (vworst > tol_p) {
            // compact away dependent rows, then append the most violated one
            const int wrow = iworst >> 1, wsgn = (iworst & 1) ? -1 : 1;
            const unsigned dup = __ballot_sync(RTMPC_FULL_MASK, kept && ra == wrow && (int)sa == wsgn);
            if (dup) break;  // most violated row is already active: numerical trouble
            const unsigned keptmask = __ballot_sync(RTMPC_FULL_MASK, kept);
            const int pos = __popc(keptmask & ((1u << lane) - 1u));
            const int myrow = ra, mysg = (int)sa;
            __syncwarp();
            if (kept) { w.act_row[pos] = myrow; w.act_sgn[pos] = mysg; }
            na = __popc(keptmask);
            if (lane == 0) { w.act_row[na] = wrow; w.act_sgn[na] = wsgn; }
            na += 1;
            __syncwarp();
            continue;
        }
        {
            // certified: leave the kept rows compacted at the front of the list (next step's warm start)
            const unsigned keptmask = __ballot_sync(RTMPC_FULL_MASK, kept);
            const int pos = __popc(keptmask & ((1u << lane) - 1u));
            const int myrow = ra, mysg = (int)sa;
            __syncwarp();
            if (kept) { w.act_row[pos] = myrow; w.act_sgn[pos] = mysg; }
            na = __popc(keptmask);
            __syncwarp();
        }
        success = true;
        rounds += 1;
        break;
    }
    if (success && lane < npad) w.zeta[lane] = w.dz[lane];
    __syncwarp();
    *rounds_out += rounds;
    *na_out = na;
    return success;
}

// build the active-set estimate {lam > s} from the interior-point state
template <int R>
__device__ __forceinline__ int build_active(const RowRegs<R>& st, int* __restrict__ out_row,
                                            int* __restrict__ out_sgn, int lane, unsigned mask_u,
                                            unsigned mask_l, int nslots, int cap) {
    int na = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nslots) {
            const int row = r * 32 + lane;
            const bool au = ((mask_u >> r) & 1u) && st.lu[r] > st.su[r];
            const bool al = ((mask_l >> r) & 1u) && st.ll[r] > st.sl[r];
            unsigned bu = __ballot_sync(RTMPC_FULL_MASK, au);
            int pu = na + __popc(bu & ((1u << lane) - 1u));
            if (au && pu < cap) { out_row[pu] = row; out_sgn[pu] = 1; }
            na += __popc(bu);
            unsigned bl = __ballot_sync(RTMPC_FULL_MASK, al);
            int pl = na + __popc(bl & ((1u << lane) - 1u));
            if (al && pl < cap) { out_row[pl] = row; out_sgn[pl] = -1; }
            na += __popc(bl);
        }
    }
    __syncwarp();
    return na;
}

template <int BS, int R>
__global__ void __launch_bounds__(256, 1)
ipm_solve_kernel(QPDev P, int B, const double* __restrict__ x_init, const double* __restrict__ ref,
                 const int* __restrict__ sel, int sel_value, double* __restrict__ z_out,
                 double* __restrict__ U_out, int* __restrict__ status_out, int* __restrict__ iters_out,
                 int* __restrict__ warm) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int n = P.n, npad = P.npad, m = P.m, mpad = P.mpad, gs = P.gs, ss = P.ss, nx = P.nx;
    const int nslots = mpad >> 5;

    double* Gs = smem;
    double* Hs_s = Gs + (size_t)mpad * gs;
    double* wbase = Hs_s + npad * npad;
    if ((((size_t)mpad * gs + npad * npad) & 1) != 0) wbase += 1;
    WarpSmem w = carve(wbase + (size_t)warp * (ipm_warp_doubles(P) + 2), P);

    if (sel) {
        // fallback launches usually select nothing: leave before staging anything
        bool any = false;
        for (int idx = threadIdx.x;; idx += blockDim.x) {
            const int inst = ((idx / wpb) * gridDim.x + blockIdx.x) * wpb + (idx % wpb);
            if (inst >= B) break;
            if (sel[inst] == sel_value) any = true;
        }
        if (!__syncthreads_or(any)) return;
    }
    // stage the shared matrices once per CTA
    for (int idx = threadIdx.x; idx < mpad * npad; idx += blockDim.x) {
        int i = idx / npad, j = idx - i * npad;
        Gs[i * gs + j] = P.G[idx];
    }
    for (int idx = threadIdx.x; idx < mpad * 2; idx += blockDim.x) Gs[(size_t)(idx >> 1) * gs + npad + (idx & 1)] = 0.0;
    for (int idx = threadIdx.x; idx < npad * npad; idx += blockDim.x) Hs_s[idx] = P.Hs[idx];
    __syncthreads();

    // S-block owned by this lane
    const int nb = (n + BS - 1) / BS;
    int blk_a = 0, blk_b = 0;
    {
        int rem = lane;
        while (blk_a < nb && rem > blk_a) { rem -= blk_a + 1; ++blk_a; }
        blk_b = rem;
    }
    const bool blk_on = blk_a < nb;

    for (int inst = blockIdx.x * wpb + warp; inst < B; inst += gridDim.x * wpb) {
        if (sel && sel[inst] != sel_value) continue;
        int status = RTMPC_MAX_ITER, iters = 0, rounds_total = 0, na_final = -1;
        // ---- parameters -----------------------------------------------------------------
        if (lane < nx) {
            w.xr[lane] = x_init[(size_t)inst * nx + lane];
            w.xr[8 + lane] = ref ? ref[(size_t)inst * nx + lane] : 0.0;
        }
        __syncwarp();
        bool par_bad = false;
        for (int i = lane; i < P.np; i += 32) {
            double acc = -P.parh[i];
            for (int k = 0; k < nx; ++k) acc = fma(P.parC[i * nx + k], w.xr[k], acc);
            if (acc > 1e-9 * (1.0 + fabs(P.parh[i]))) par_bad = true;
        }
        par_bad = __any_sync(RTMPC_FULL_MASK, par_bad);
        double qj = 0.0;
        if (lane < n) {
            for (int k = 0; k < nx; ++k) {
                qj = fma(P.Fx[lane * nx + k], w.xr[k], qj);
                qj = fma(P.Fr[lane * nx + k], w.xr[8 + k], qj);
            }
        }
        if (lane < npad) w.q[lane] = qj;
        const double sc_q = 1.0 + warp_max(fabs(qj));
        __syncwarp();
        double zu = 0.0;
        if (lane < n) {
            for (int k = 0; k < n; ++k) zu = fma(-P.Hinv[lane * npad + k], w.q[k], zu);
        }
        if (lane < npad) { w.zu[lane] = zu; w.zeta[lane] = zu; }
        unsigned mask_u = 0, mask_l = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nslots) {
                const int row = r * 32 + lane;
                double lo = P.lo0[row], up = P.up0[row];
                for (int k = 0; k < nx; ++k) {
                    lo = fma(P.Lx[row * nx + k], w.xr[k], lo);
                    up = fma(P.Ux[row * nx + k], w.xr[k], up);
                }
                w.vlo[row] = lo;
                w.vup[row] = up;
                if (P.has_up[row]) mask_u |= (1u << r);
                if (P.has_lo[row]) mask_l |= (1u << r);
            }
        }
        __syncwarp();

        RowRegs<R> st;
        bool done = false;
        if (par_bad) { status = RTMPC_INFEASIBLE; done = true; }
        if (!done) {
            gemv_rows<R>(Gs, gs, npad, nslots, w.zeta, lane, st.t);
            double smin = RTMPC_INF, ssum = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                st.su[r] = 1.0; st.sl[r] = 1.0; st.lu[r] = 0.0; st.ll[r] = 0.0;
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    if ((mask_u >> r) & 1u) { st.su[r] = w.vup[row] - st.t[r]; smin = fmin(smin, st.su[r]); }
                    if ((mask_l >> r) & 1u) { st.sl[r] = st.t[r] - w.vlo[row]; smin = fmin(smin, st.sl[r]); }
                }
            }
            smin = warp_min(smin);
            if (smin > 0.0) { status = RTMPC_OPTIMAL; done = true; na_final = 0; }
            if (!done) {
                const double shift = fmax(-1.5 * smin, 0.0);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((mask_u >> r) & 1u) { st.su[r] = fmax(st.su[r] + shift, P.s_floor); ssum += st.su[r]; }
                    if ((mask_l >> r) & 1u) { st.sl[r] = fmax(st.sl[r] + shift, P.s_floor); ssum += st.sl[r]; }
                }
                const double mu0 = warp_sum(ssum) / (double)P.mtot;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((mask_u >> r) & 1u) st.lu[r] = mu0 / st.su[r];
                    if ((mask_l >> r) & 1u) st.ll[r] = mu0 / st.sl[r];
                }
            }
        }

        // ---- interior point ---------------------------------------------------------------
        int next_try = 0, best_na = -1;
        bool boost = false;
        double best_merit = RTMPC_INF;
        while (!done) {
            // A. residuals (t = G zeta is kept up to date incrementally: t += alpha_p * G dz)
            double acc_mu = 0.0, rp_max = 0.0, ymax = 0.0, cert = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    double e1 = 0.0;
                    if ((mask_u >> r) & 1u) {
                        const double up = w.vup[row];
                        const double rpu = st.t[r] + st.su[r] - up;
                        e1 = fma(st.lu[r] * __drcp_rn(st.su[r]), rpu, e1);
                        acc_mu = fma(st.su[r], st.lu[r], acc_mu);
                        rp_max = fmax(rp_max, fabs(rpu));
                        cert = fma(up, st.lu[r], cert);
                    }
                    if ((mask_l >> r) & 1u) {
                        const double lo = w.vlo[row];
                        const double rpl = -st.t[r] + st.sl[r] + lo;
                        e1 = fma(-st.ll[r] * __drcp_rn(st.sl[r]), rpl, e1);
                        acc_mu = fma(st.sl[r], st.ll[r], acc_mu);
                        rp_max = fmax(rp_max, fabs(rpl));
                        cert = fma(-lo, st.ll[r], cert);
                    }
                    const double y = st.lu[r] - st.ll[r];
                    ymax = fmax(ymax, fabs(y));
                    w.va[row] = e1;
                    w.vb[row] = y;
                }
            }
            __syncwarp();
            // B. dual residual, objective, termination
            double hzj = 0.0;
            if (lane < n) for (int k = 0; k < n; ++k) hzj = fma(Hs_s[lane * npad + k], w.zeta[k], hzj);
            double ge, gy;
            gemv_cols2(Gs, gs, npad, m, w.va, w.vb, lane, ge, gy);
            const double qv = (lane < n) ? w.q[lane] : 0.0;
            const double zv = (lane < n) ? w.zeta[lane] : 0.0;
            const double rd = (lane < n) ? (hzj + qv + gy) : 0.0;
            const double res_d = warp_max(fabs(rd));
            const double pobj = warp_sum(zv * (0.5 * hzj + qv));
            const double gap = warp_sum(acc_mu);
            const double mu = gap / (double)P.mtot;
            const double rp = warp_max(rp_max);
            ymax = warp_max(ymax);
            cert = warp_sum(cert);
            const double gy_max = warp_max(fabs(gy));
            const double rd_rel = res_d / sc_q, rp_rel = rp / P.sc_b;
            const double res = fmax(rd_rel, rp_rel);
            const double relgap = gap / (1.0 + fabs(pobj));
            const double merit = fmax(res, relgap);
            const bool bad = !(mu == mu) || !(mu <= 1e40) || !(merit == merit);
            if (!bad && merit < best_merit) {
                best_merit = merit;
                if (lane < npad) w.best[lane] = w.zeta[lane];
                if (merit <= 1e-3) best_na = build_active<R>(st, w.best_row, w.best_sgn, lane, mask_u, mask_l, nslots, 2 * npad);
            }
            // primal infeasibility certificate on the multiplier direction (Farkas)
            if (!bad && ymax > 1e6 * sc_q && iters >= 3 && gy_max <= 1e-7 * ymax && cert < -1e-7 * ymax * P.sc_b) {
                status = RTMPC_INFEASIBLE;
                break;
            }
            // conv1: accurate enough for the active set to be identified.  The second clause covers
            // problems whose feasible set has no interior along some rows (multipliers unbounded):
            // there the dual residual stalls while primal residual and gap collapse.
            const bool conv1 = (res <= 1e-7 && relgap <= 1e-8) || (rp_rel <= 1e-8 && relgap <= 1e-9 && rd_rel <= 1e-3);
            const bool conv2 = (res <= 1e-9 && relgap <= 1e-12);
            // losing a good point, not the wobble of the first iterations (the relative gap is not monotone there)
            const bool diverged = bad || (best_merit <= 1e-4 && merit > 1e3 * best_merit);
            if (!diverged && conv1 && !conv2 && iters >= next_try && iters < P.max_iter) {
                int na = build_active<R>(st, w.act_row, w.act_sgn, lane, mask_u, mask_l, nslots, 2 * npad);
                if (na <= npad && polish_warp<R>(P, Gs, w, na, lane, mask_u, mask_l, nslots, 24, &rounds_total, &na_final)) {
                    status = RTMPC_OPTIMAL;
                    break;
                }
                next_try = iters + 3;
            }
            if (conv2 || diverged || iters >= P.max_iter) {
                if (best_merit <= 1e-5 && best_na >= 0) {
                    int na;
                    if (conv2 && !diverged) {
                        na = build_active<R>(st, w.act_row, w.act_sgn, lane, mask_u, mask_l, nslots, 2 * npad);
                    } else {
                        na = best_na;
                        for (int i = lane; i < na && i < 2 * npad; i += 32) { w.act_row[i] = w.best_row[i]; w.act_sgn[i] = w.best_sgn[i]; }
                        __syncwarp();
                    }
                    if (na <= npad && polish_warp<R>(P, Gs, w, na, lane, mask_u, mask_l, nslots, 24, &rounds_total, &na_final)) {
                        status = RTMPC_OPTIMAL;
                        break;
                    }
                }
                if (lane < npad) w.zeta[lane] = w.best[lane];
                __syncwarp();
                if (best_merit <= 1e-7) status = RTMPC_OPTIMAL_INACCURATE;
                else if (!bad && gy_max <= 1e-6 * ymax && cert < -1e-6 * ymax * P.sc_b) status = RTMPC_INFEASIBLE;
                else status = RTMPC_MAX_ITER;
                break;
            }
            iters += 1;
            // C. Schur complement  S = Hs + G' diag(d) G,  d = lam_u/s_u + lam_l/s_l
            if (lane < npad) { w.hz[lane] = hzj; w.rhs[lane] = -hzj - qv - ge; }
            __syncwarp();   // va/vb consumed by every lane
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    double d = 0.0;
                    if ((mask_u >> r) & 1u) d = st.lu[r] * __drcp_rn(st.su[r]);
                    if ((mask_l >> r) & 1u) d = fma(st.ll[r], __drcp_rn(st.sl[r]), d);
                    w.va[r * 32 + lane] = d;
                }
            }
            __syncwarp();
            form_schur<BS>(Gs, gs, m, npad, Hs_s, w.va, w.S, ss, n, lane, blk_a, blk_b, blk_on);
            __syncwarp();
            chol_warp(w.S, ss, n, lane);
            // E. predictor (affine scaling direction)
            double dza = chol_solve_warp(w.S, ss, n, lane, (lane < n) ? w.rhs[lane] : 0.0);
            if (lane < npad) w.dz[lane] = dza;
            __syncwarp();
            gemv_rows<R>(Gs, gs, npad, nslots, w.dz, lane, st.ta);
            // F. step lengths of the affine direction and mu_aff in one pass.  With
            //    g = 1 + ds/s the affine multiplier step is dl = -lam*g, so
            //    -s/ds = 1/(1-g)   (primal ratio, only where g < 1),   -lam/dl = 1/g (dual ratio, g > 0)
            //    (s + ap ds)(lam + ad dl) = s*lam*(1 + ap*(g-1))*(1 - ad*g)
            double pmax = 0.0, dmax = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    if ((mask_u >> r) & 1u) {
                        const double rpu = st.t[r] + st.su[r] - w.vup[row];
                        const double h = (-rpu - st.ta[r]) * __drcp_rn(st.su[r]);     // ds/s
                        const double g = 1.0 + h, sl_ = st.su[r] * st.lu[r];
                        pmax = fmax(pmax, -h);
                        dmax = fmax(dmax, g);
                        c0 += sl_; c1 = fma(sl_, h, c1); c2 = fma(sl_, g, c2); c3 = fma(sl_ * h, g, c3);
                    }
                    if ((mask_l >> r) & 1u) {
                        const double rpl = -st.t[r] + st.sl[r] + w.vlo[row];
                        const double h = (-rpl + st.ta[r]) * __drcp_rn(st.sl[r]);
                        const double g = 1.0 + h, sl_ = st.sl[r] * st.ll[r];
                        pmax = fmax(pmax, -h);
                        dmax = fmax(dmax, g);
                        c0 += sl_; c1 = fma(sl_, h, c1); c2 = fma(sl_, g, c2); c3 = fma(sl_ * h, g, c3);
                    }
                }
            }
            pmax = warp_max(pmax);
            dmax = warp_max(dmax);
            double ap = (pmax > 1.0) ? 1.0 / pmax : 1.0;
            double ad = (dmax > 1.0) ? 1.0 / dmax : 1.0;
            c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); c3 = warp_sum(c3);
            const double mu_aff = (c0 + ap * c1 - ad * c2 - ap * ad * c3) / (double)P.mtot;
            double sigma = mu_aff / mu;
            sigma = sigma * sigma * sigma;
            double cross = 1.0;
            if (boost) {
                // safeguard against Mehrotra limit cycles: after a short step take a well-centred,
                // first-order step (sigma >= 0.8, no second-order term)
                sigma = fmax(sigma, 0.8);
                cross = 0.0;
            }
            const double smu = sigma * mu;
            // G. corrector right-hand side.  With du = lam/s:
            //    e2 = lam + (-rc + lam*rp)/s = du*rp + smu/s + cross*du*ds_a*g      (rc = s*lam + cross*ds_a*dl_a - smu)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    double e2 = 0.0;
                    if ((mask_u >> r) & 1u) {
                        const double is = __drcp_rn(st.su[r]);
                        const double rpu = st.t[r] + st.su[r] - w.vup[row];
                        const double dsa = -rpu - st.ta[r];
                        const double du = st.lu[r] * is, g = fma(dsa, is, 1.0);
                        e2 += fma(du, rpu, smu * is) + cross * du * dsa * g;
                    }
                    if ((mask_l >> r) & 1u) {
                        const double is = __drcp_rn(st.sl[r]);
                        const double rpl = -st.t[r] + st.sl[r] + w.vlo[row];
                        const double dsa = -rpl + st.ta[r];
                        const double du = st.ll[r] * is, g = fma(dsa, is, 1.0);
                        e2 -= fma(du, rpl, smu * is) + cross * du * dsa * g;
                    }
                    w.va[row] = e2;
                }
            }
            __syncwarp();
            double ge2, dummy;
            gemv_cols2(Gs, gs, npad, m, w.va, nullptr, lane, ge2, dummy);
            double rhs2 = (lane < n) ? (-w.hz[lane] - w.q[lane] - ge2) : 0.0;
            double dzc = chol_solve_warp(w.S, ss, n, lane, rhs2);
            if (lane < npad) w.dz[lane] = dzc;
            __syncwarp();
            gemv_rows<R>(Gs, gs, npad, nslots, w.dz, lane, st.tz);
            // H. combined step:  ds = -rp -/+ tz,  dl = -lam - du*ds + smu/s + cross*du*ds_a*g
            //    ratios: -s/ds = 1/(-ds/s), -lam/dl = 1/(-dl/lam)
            pmax = 0.0; dmax = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    if ((mask_u >> r) & 1u) {
                        const double is = __drcp_rn(st.su[r]);
                        const double rpu = st.t[r] + st.su[r] - w.vup[row];
                        const double dsa = -rpu - st.ta[r];
                        const double du = st.lu[r] * is, g = fma(dsa, is, 1.0);
                        const double ds = -rpu - st.tz[r];
                        const double dl = -st.lu[r] - du * ds + smu * is + cross * du * dsa * g;
                        pmax = fmax(pmax, -ds * is);
                        dmax = fmax(dmax, -dl * __drcp_rn(st.lu[r]));
                    }
                    if ((mask_l >> r) & 1u) {
                        const double is = __drcp_rn(st.sl[r]);
                        const double rpl = -st.t[r] + st.sl[r] + w.vlo[row];
                        const double dsa = -rpl + st.ta[r];
                        const double du = st.ll[r] * is, g = fma(dsa, is, 1.0);
                        const double ds = -rpl + st.tz[r];
                        const double dl = -st.ll[r] - du * ds + smu * is + cross * du * dsa * g;
                        pmax = fmax(pmax, -ds * is);
                        dmax = fmax(dmax, -dl * __drcp_rn(st.ll[r]));
                    }
                }
            }
            pmax = warp_max(pmax);
            dmax = warp_max(dmax);
            const double eta = (mu < 1.0) ? fmin(0.9995, fmax(0.995, 1.0 - mu)) : 0.995;
            ap = (eta < pmax) ? eta / pmax : 1.0;        // min(1, eta * min ratio)
            ad = (eta < dmax) ? eta / dmax : 1.0;
            boost = fmin(ap, ad) < 0.3;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nslots) {
                    const int row = r * 32 + lane;
                    const double tr = st.t[r];
                    if ((mask_u >> r) & 1u) {
                        const double is = __drcp_rn(st.su[r]);
                        const double rpu = tr + st.su[r] - w.vup[row];
                        const double dsa = -rpu - st.ta[r];
                        const double du = st.lu[r] * is, g = fma(dsa, is, 1.0);
                        const double ds = -rpu - st.tz[r];
                        const double dl = -st.lu[r] - du * ds + smu * is + cross * du * dsa * g;
                        st.su[r] = fma(ap, ds, st.su[r]);
                        st.lu[r] = fma(ad, dl, st.lu[r]);
                    }
                    if ((mask_l >> r) & 1u) {
                        const double is = __drcp_rn(st.sl[r]);
                        const double rpl = -tr + st.sl[r] + w.vlo[row];
                        const double dsa = -rpl + st.ta[r];
                        const double du = st.ll[r] * is, g = fma(dsa, is, 1.0);
                        const double ds = -rpl + st.tz[r];
                        const double dl = -st.ll[r] - du * ds + smu * is + cross * du * dsa * g;
                        st.sl[r] = fma(ap, ds, st.sl[r]);
                        st.ll[r] = fma(ad, dl, st.ll[r]);
                    }
                    st.t[r] = fma(ap, st.tz[r], tr);      // t = G zeta stays consistent with zeta += ap dz
                }
            }
            if (lane < npad) w.zeta[lane] += ap * dzc;
            __syncwarp();
        }

        // ---- outputs ----------------------------------------------------------------------
        const bool has_sol = (status == RTMPC_OPTIMAL || status == RTMPC_OPTIMAL_INACCURATE || status == RTMPC_MAX_ITER);
        if (lane < npad) w.dz[lane] = (lane < n) ? w.zeta[lane] * P.D[lane] : 0.0;   // unscaled decision
        __syncwarp();
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        double* zf = w.va;   // un-condensed vector [x_0..x_N | u | x_bar | u_bar]
        for (int i = lane; i < P.nz; i += 32) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(P.Phi[(size_t)i * npad + k], w.dz[k], acc);
            for (int k = 0; k < nx; ++k) acc = fma(P.Psi[(size_t)i * nx + k], w.xr[k], acc);
            zf[i] = has_sol ? acc : nanv;
            if (z_out) z_out[(size_t)inst * P.nz + i] = zf[i];
        }
        __syncwarp();
        if (U_out) {
            const int nu = P.nu, N = P.N;
            const int ou = nx * (N + 1);
            for (int i = lane; i < N * nu; i += 32) U_out[(size_t)inst * (N + 1) * nu + i] = zf[ou + i];
            if (P.nss > 0 && lane < nu) {
                const int oxb = ou + N * nu, oub = oxb + nx;
                double acc = zf[oub + lane];
                for (int k = 0; k < nx; ++k) acc = fma(P.Kss[lane * nx + k], zf[oxb + k], acc);
                U_out[(size_t)inst * (N + 1) * nu + N * nu + lane] = acc;
            }
        }
        if (warm) {
            const int ws = npad + 1;
            const int nw = (status == RTMPC_OPTIMAL) ? na_final : -1;
            if (lane == 0) warm[(size_t)inst * ws] = nw;
            if (lane < nw) warm[(size_t)inst * ws + 1 + lane] = 2 * w.act_row[lane] + (w.act_sgn[lane] < 0 ? 1 : 0);
        }
        if (lane == 0) {
            if (status_out) status_out[inst] = status;
            // bits 0-11 IPM iterations, 24-27 endgame rounds; a handed-over instance keeps the active-set kernel's
            // step count (12-23) and give-up reason (28-30)
            if (iters_out) {
                const int keep = (sel && sel_value <= RTMPC_FALLBACK_STATUS) ? (iters_out[inst] & 0x70FFF000) : 0;
                iters_out[inst] = iters | ((rounds_total > 15 ? 15 : rounds_total) << 24) | keep;
            }
        }
        __syncwarp();
    }
}

}  // namespace rtmpc

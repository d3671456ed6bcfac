"""Oracle (test infrastructure): the reference's MPC quadratic programs, un-condensed, and a
dense FP64 primal-dual interior-point solver for them.

The reference states each QP in cvxpy and hands it to Clarabel (``RegulatorMPC.py:31``,
``TubeTrackingMPC.py:183``).  cvxpy>=1.4.1 and clarabel are absent from this image, so the
*problem* is restated here variable-for-variable and constraint-for-constraint (cited per
builder), and solved by an in-file Mehrotra predictor-corrector IPM -- the same algorithm
family Clarabel publishes -- run to a much tighter tolerance (1e-13 on the gap, then an active-set polish with a KKT certificate) than the reference's
``tol_gap_abs = tol_gap_rel = 1e-7``.  PARITY UNPINNED against Clarabel's own output (see
``oracle/__init__.py``); cross-checked against scipy's SLSQP in ``tests/test_oracle_qp.py``.

Decision vector (cvxpy variable order is irrelevant to the minimiser):
    z = [x_0 .. x_N | u_0 .. u_{N-1} | x_bar | u_bar | (y, v: only the G2 variant)]
Problem form:  min 1/2 z'Pz + q'z   s.t.  E z = e,  G z <= h
with q, e, h affine in the parameters (x_init, ref).
"""
import numpy as np


class ParamQP:
    """Container: P, E, G fixed; q = Qr @ ref, e = Ex @ x_init, h = h0 + Hp @ x_init."""

    def __init__(self, nx, nu, N, has_ss):
        self.nx, self.nu, self.N, self.has_ss = nx, nu, N, has_ss
        self.nz = nx * (N + 1) + nu * N + (nx + nu if has_ss else 0)
        self.n_extra = 0

    def ix(self, i):
        return slice(self.nx * i, self.nx * (i + 1))

    def iu(self, i):
        o = self.nx * (self.N + 1)
        return slice(o + self.nu * i, o + self.nu * (i + 1))

    @property
    def ixbar(self):
        o = self.nx * (self.N + 1) + self.nu * self.N
        return slice(o, o + self.nx)

    @property
    def iubar(self):
        o = self.nx * (self.N + 1) + self.nu * self.N + self.nx
        return slice(o, o + self.nu)

    def params(self, x_init, ref=None):
        x_init = np.asarray(x_init, dtype=float).flatten()
        q = np.zeros(self.nz + self.n_extra)
        if ref is not None and self.Qr is not None:
            q = self.Qr @ np.asarray(ref, dtype=float).flatten()
        e = self.Ex @ x_init
        h = self.h0 + self.Hp @ x_init
        return q, e, h

    def split(self, z):
        nx, nu, N = self.nx, self.nu, self.N
        x = z[:nx * (N + 1)].reshape(N + 1, nx).T.copy()
        u = z[nx * (N + 1):nx * (N + 1) + nu * N].reshape(N, nu).T.copy()
        if self.has_ss:
            return x, u, z[self.ixbar].copy(), z[self.iubar].copy()
        return x, u


def _assemble(A, B, Q, R, N, *, P_term=None, T_ss=None, Xs=None, Us=None, Xf=None, Z_init=None,
              terminal_eq=False, g2_free_terminal=False):
    """Shared builder.  ``T_ss is None`` => regulator family (no x_bar/u_bar)."""
    A = np.asarray(A, float)
    B = np.asarray(B, float)
    Q = np.asarray(Q, float)
    R = np.atleast_2d(np.asarray(R, float))
    nx, nu = B.shape
    has_ss = T_ss is not None
    qp = ParamQP(nx, nu, N, has_ss)
    if g2_free_terminal:
        qp.n_extra = nx + nu
    nz = qp.nz + qp.n_extra
    H = np.zeros((nz, nz))          # cost = z'Hz + 2 g'z  (cvxpy quad_form has no 1/2)

    def add_quad(sl_a, sl_b, M):
        # (a-b)'M(a-b)
        H[sl_a, sl_a] += M
        if sl_b is not None:
            H[sl_b, sl_b] += M
            H[sl_a, sl_b] -= M
            H[sl_b, sl_a] -= M

    xb = qp.ixbar if has_ss else None
    ub = qp.iubar if has_ss else None
    for i in range(N):
        add_quad(qp.ix(i), xb, Q)
        add_quad(qp.iu(i), ub, R)
    if P_term is not None:
        add_quad(qp.ix(N), xb, np.asarray(P_term, float))
    Qr = None
    if has_ss:
        T_ss = np.asarray(T_ss, float)
        H[xb, xb] += T_ss             # (x_bar-ref)'T(x_bar-ref)
        Qr = np.zeros((nz, nx))
        Qr[xb, :] = -2.0 * T_ss       # q = 2 g,  g_xbar = -T ref
    qp.P = 2.0 * H
    qp.Qr = Qr

    E_rows, Ex_rows = [], []
    G_rows, h0_rows, Hp_rows = [], [], []

    def eq(row, xcoef=None):
        E_rows.append(row)
        Ex_rows.append(np.zeros((row.shape[0], nx)) if xcoef is None else xcoef)

    def ineq(row, h0, xcoef=None):
        G_rows.append(row)
        h0_rows.append(np.asarray(h0, float).flatten())
        Hp_rows.append(np.zeros((row.shape[0], nx)) if xcoef is None else xcoef)

    # initial condition
    if Z_init is None:
        r = np.zeros((nx, nz))
        r[:, qp.ix(0)] = np.eye(nx)
        eq(r, np.eye(nx))                                   # x_0 = x_init
    else:
        Hz, hz = Z_init
        r = np.zeros((Hz.shape[0], nz))
        r[:, qp.ix(0)] = -Hz
        ineq(r, hz, -Hz)                                    # Hz (x_init - x_0) <= hz
    for i in range(N):
        r = np.zeros((nx, nz))
        r[:, qp.ix(i + 1)] = np.eye(nx)
        r[:, qp.ix(i)] = -A
        r[:, qp.iu(i)] = -B
        eq(r)                                               # x_{i+1} = A x_i + B u_i
        if Xs is not None:
            r = np.zeros((Xs[0].shape[0], nz))
            r[:, qp.ix(i)] = Xs[0]
            ineq(r, Xs[1])                                  # Hx x_i <= hx   (i < N only, G3)
        if Us is not None:
            r = np.zeros((Us[0].shape[0], nz))
            r[:, qp.iu(i)] = Us[0]
            ineq(r, Us[1])
    if has_ss:
        r = np.zeros((nx, nz))
        r[:, xb] = A - np.eye(nx)
        r[:, ub] = B
        eq(r)                                               # (A-I) x_bar + B u_bar = 0
    if terminal_eq:
        r = np.zeros((nx, nz))
        r[:, qp.ix(N)] = np.eye(nx)
        r[:, xb] = -np.eye(nx)
        eq(r)                                               # x_N = x_bar
    elif Xf is not None:
        HN, hN = Xf
        r = np.zeros((HN.shape[0], nz))
        if has_ss:
            if g2_free_terminal:
                # G2: the reference's "packet received" problem references the *other*
                # problem's x_N and u_bar (``TubeTrackingMPC.py:293``), i.e. free variables here.
                r[:, qp.nz:qp.nz + nx] = HN[:, :nx]
                r[:, xb] = HN[:, nx:2 * nx]
                r[:, qp.nz + nx:] = HN[:, 2 * nx:]
            else:
                r[:, qp.ix(N)] = HN[:, :nx]
                r[:, xb] = HN[:, nx:2 * nx]
                r[:, ub] = HN[:, 2 * nx:]
        else:
            r[:, qp.ix(N)] = HN
        ineq(r, hN)
    qp.E = np.vstack(E_rows)
    qp.Ex = np.vstack(Ex_rows)
    if G_rows:
        qp.G = np.vstack(G_rows)
        qp.h0 = np.hstack(h0_rows)
        qp.Hp = np.vstack(Hp_rows)
    else:
        qp.G = np.zeros((0, nz))
        qp.h0 = np.zeros(0)
        qp.Hp = np.zeros((0, nx))
    return qp


def _Ab(poly):
    return None if poly is None else (np.asarray(poly.A, float), np.asarray(poly.b, float).flatten())


def build_regulator(A, B, Q, R, N, X=None, U=None):
    """``RegulatorMPC.generate_optimization_problem`` (``RegulatorMPC.py:45-76``): no terminal cost/set."""
    return _assemble(A, B, Q, R, N, Xs=_Ab(X), Us=_Ab(U))


def build_tube_regulator(A, B, Q, R, N, P, Xc, Uc, Xf, Z):
    """``TubeRegulatorMPC.generate_optimization_problem`` (``TubeRegulatorMPC.py:109-143``)."""
    return _assemble(A, B, Q, R, N, P_term=P, Xs=_Ab(Xc), Us=_Ab(Uc), Xf=_Ab(Xf), Z_init=_Ab(Z))


def build_tracking(A, B, Q, R, N, P, X=None, U=None, Xf=None):
    """``TrackingMPC.generate_optimization_problem`` (``TrackingMPC.py:62-114``), T = 10 P (``:34``)."""
    return _assemble(A, B, Q, R, N, P_term=P, T_ss=10 * np.asarray(P), Xs=_Ab(X), Us=_Ab(U),
                     Xf=_Ab(Xf), terminal_eq=(Xf is None))


def build_tube_tracking(A, B, Q, R, N, P, Xc, Uc, Xf, Z=None, fixed_initial_state=False):
    """``TubeTrackingMPC.generate_optimization_problem`` (``TubeTrackingMPC.py:104-156``)."""
    return _assemble(A, B, Q, R, N, P_term=P, T_ss=10 * np.asarray(P), Xs=_Ab(Xc), Us=_Ab(Uc),
                     Xf=_Ab(Xf), Z_init=None if fixed_initial_state else _Ab(Z))


def build_extended_packet_received(A, B, Q, R, N, P, Xc, Uc, Xf, ZmW, strict_terminal=False):
    """``ExtendedTubeTrackingMPC.generate_optimization_problem_when_packet_received``
    (``TubeTrackingMPC.py:253-299``).  Default reproduces G2 (terminal rows bind free variables)."""
    return _assemble(A, B, Q, R, N, P_term=P, T_ss=10 * np.asarray(P), Xs=_Ab(Xc), Us=_Ab(Uc),
                     Xf=_Ab(Xf), Z_init=_Ab(ZmW), g2_free_terminal=not strict_terminal)


# ----------------------------------------------------------------------------------------------
# dense primal-dual interior point (Mehrotra predictor-corrector)
# ----------------------------------------------------------------------------------------------

class IPMResult:
    __slots__ = ("z", "status", "iters", "lam", "y", "res", "polished")


def solve_qp(P, q, E, e, G, h, tol=1e-13, max_iter=200, polish=True):
    """min 1/2 z'Pz + q'z  s.t. Ez=e, Gz<=h.  status: 'optimal' | 'infeasible' | 'max_iter'.
    The IPM runs to ``tol`` (relative residuals and gap); ``polish`` then re-solves the KKT system on
    the identified active set and keeps that point only if it passes a KKT check at 1e-9, which
    removes the O(sqrt(gap)) error an IPM leaves on weakly active rows."""
    n = P.shape[0]
    p = E.shape[0]
    m = G.shape[0]
    out = IPMResult()
    sc_q = 1.0 + np.abs(q).max(initial=0.0) + np.abs(P).max(initial=0.0)
    sc_e = 1.0 + np.abs(e).max(initial=0.0)
    sc_h = 1.0 + np.abs(h).max(initial=0.0)

    def kkt_solve(d, rhs1, rhs2):
        K = np.zeros((n + p, n + p))
        K[:n, :n] = P + (G.T * d) @ G if m else P
        K[:n, n:] = E.T
        K[n:, :n] = E
        K[np.arange(n), np.arange(n)] += 1e-13 * sc_q
        K[np.arange(n, n + p), np.arange(n, n + p)] -= 1e-13
        sol = np.linalg.solve(K, np.r_[rhs1, rhs2])
        return sol[:n], sol[n:]

    # starting point: equality-constrained QP with unit barrier weights
    z, y = kkt_solve(np.ones(m), -q + (G.T @ h if m else 0.0), e)
    if m:
        s = h - G @ z
        lam = -s.copy()
        shift_s = max(-1.5 * s.min(), 0.0)
        shift_l = max(-1.5 * lam.min(), 0.0)
        s = s + shift_s
        lam = lam + shift_l
        mu0 = 0.5 * (s @ lam)
        s = s + mu0 / max(lam.sum(), 1e-300)
        lam = lam + mu0 / max(s.sum(), 1e-300)
        s = np.maximum(s, 1e-8)
        lam = np.maximum(lam, 1e-8)
    else:
        s = np.zeros(0)
        lam = np.zeros(0)

    status = "max_iter"
    best = (np.inf, None, None, None, None, np.inf)
    for it in range(max_iter):
        r_d = P @ z + q + E.T @ y + (G.T @ lam if m else 0.0)
        r_e = E @ z - e
        r_g = (G @ z + s - h) if m else np.zeros(0)
        mu = (s @ lam) / m if m else 0.0
        res = max(np.abs(r_d).max(initial=0.0) / sc_q, np.abs(r_e).max(initial=0.0) / sc_e,
                  np.abs(r_g).max(initial=0.0) / sc_h)
        pobj = 0.5 * z @ P @ z + q @ z
        gap = (s @ lam) if m else 0.0
        relgap = gap / (1.0 + abs(pobj))
        merit = max(res, relgap)
        if merit < best[0]:
            best = (merit, z.copy(), y.copy(), lam.copy(), s.copy(), res)
        if res <= 1e-10 and relgap <= tol:
            status = "optimal"
            break
        # numerical floor: mu has collapsed below what the reduced KKT solve can resolve, or the
        # iterate has drifted away from the best point seen -> stop and return the best point
        if best[0] <= 1e-8 and (relgap <= 1e-3 * tol or merit > 1e3 * best[0]):
            break
        if m == 0:
            dz, dy = kkt_solve(np.zeros(0), -r_d, -r_e)
            z = z + dz
            y = y + dy
            continue
        # crude divergence test for primal infeasibility
        if mu > 1e-30 and (np.abs(lam).max() > 1e14 * sc_q or not np.all(np.isfinite(z))):
            status = "infeasible"
            break
        d = lam / s
        # predictor
        r_c = s * lam
        rhs = -r_d - G.T @ ((-r_c + lam * r_g) / s)
        dz_a, dy_a = kkt_solve(d, rhs, -r_e)
        ds_a = -r_g - G @ dz_a
        dl_a = (-r_c - lam * ds_a) / s
        a_p = _max_step(s, ds_a)
        a_d = _max_step(lam, dl_a)
        mu_aff = ((s + a_p * ds_a) @ (lam + a_d * dl_a)) / m
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        # corrector
        r_c = s * lam + ds_a * dl_a - sigma * mu
        rhs = -r_d - G.T @ ((-r_c + lam * r_g) / s)
        dz, dy = kkt_solve(d, rhs, -r_e)
        ds = -r_g - G @ dz
        dl = (-r_c - lam * ds) / s
        eta = min(0.9995, max(0.995, 1.0 - mu)) if mu < 1 else 0.995
        a_p = min(1.0, eta * _max_step(s, ds))
        a_d = min(1.0, eta * _max_step(lam, dl))
        z = z + a_p * dz
        s = s + a_p * ds
        y = y + a_d * dy
        lam = lam + a_d * dl
    if status != "optimal" and best[0] <= 1e-8:
        status = "optimal"
    if status == "optimal" and best[1] is not None and m:
        _, z, y, lam, s, res = best
    out.z, out.status, out.iters, out.lam, out.y = z, status, it + 1, lam, y
    out.res = res
    out.polished = False
    if polish and status == "optimal" and m:
        _polish(out, P, q, E, e, G, h, s)
    return out


def _polish(out, P, q, E, e, G, h, s, max_rounds=40):
    """Active-set endgame.  The MPC Hessians here are very flat along late-horizon inputs
    (cond ~ 1e7 for the cartpole), so an IPM point with a 1e-12 relative gap can still be 1e-3 away
    from the minimiser in those directions.  Starting from the IPM's active-set estimate
    {lam_i > s_i} we solve the equality-constrained QP on that set (min-norm least squares, so
    dependent rows are harmless), add the most violated row / drop the most negative multiplier,
    and accept only with a KKT certificate: primal violation <= 1e-10 (1+|h|) and a non-negative
    multiplier found by NNLS that closes stationarity to 1e-9 relative."""
    from scipy.optimize import nnls
    n, p = P.shape[0], E.shape[0]
    act = list(np.nonzero(out.lam > s)[0])
    scale = 1.0 + np.abs(q).max(initial=0.0)
    seen = set()
    for _ in range(max_rounds):
        na = len(act)
        Ga = G[act]
        # null-space method on C z = d, C = [E; G_act] (orthogonal transformations only, so the
        # 1e6 spread of the cost weights and dependent active rows do no harm)
        C = np.vstack([E, Ga])
        dvec = np.r_[e, h[act]]
        U_, sv, Vt = np.linalg.svd(C, full_matrices=True)
        rk = int((sv > 1e-11 * sv[0]).sum()) if sv.size else 0
        z_p = Vt[:rk].T @ ((U_[:, :rk].T @ dvec) / sv[:rk])
        incons = np.abs(C @ z_p - dvec)
        if incons.max(initial=0.0) > 1e-9 * (1.0 + np.abs(dvec).max(initial=0.0)):
            # the guessed rows cannot all hold with equality (near-parallel facets of the terminal
            # set): release the guessed inequality row that fits worst and try again
            if na == 0:
                return False
            act.pop(int(np.argmax(incons[p:])))
            continue
        Nn = Vt[rk:].T
        if Nn.shape[1]:
            Hr = Nn.T @ P @ Nn

            def _rsolve(b):
                # the G2 variant has cost-free variables -> reduced Hessian only PSD there
                try:
                    if np.linalg.cond(Hr) < 1e13:
                        return np.linalg.solve(Hr, b)
                except np.linalg.LinAlgError:
                    pass
                return np.linalg.lstsq(Hr, b, rcond=1e-12)[0]
            w = _rsolve(-Nn.T @ (P @ z_p + q))
            z = z_p + Nn @ w
            z = z + Nn @ _rsolve(-Nn.T @ (P @ z + q))      # one refinement
        else:
            z = z_p
        lam_a = np.linalg.lstsq(C.T, -(P @ z + q), rcond=None)[0][p:] if na else np.zeros(0)
        viol = (G @ z - h) / (1.0 + np.abs(h))
        worst = int(np.argmax(viol)) if viol.size else -1
        feasible = not (worst >= 0 and viol[worst] > 1e-10)
        g = P @ z + q
        if feasible:
            # KKT certificate: a non-negative multiplier on the guessed rows closing stationarity
            if na:
                M = np.c_[E.T, -E.T, Ga.T]
                colscale = np.maximum(np.linalg.norm(M, axis=0), 1e-300)
                xs, rn = nnls(M / colscale, -g, maxiter=20 * M.shape[1])
                if rn <= 1e-9 * scale:
                    xs = xs / colscale
                    lam = np.zeros(G.shape[0])
                    lam[act] = xs[2 * p:]
                    out.z, out.lam, out.y = z, lam, xs[:p] - xs[p:2 * p]
                    out.res = 0.0
                    out.polished = True
                    return True
            else:
                yv = np.linalg.lstsq(E.T, -g, rcond=None)[0] if p else np.zeros(0)
                if np.abs(g + (E.T @ yv if p else 0.0)).max() <= 1e-9 * scale:
                    out.z, out.lam, out.y = z, np.zeros(G.shape[0]), yv
                    out.res = 0.0
                    out.polished = True
                    return True
                return False
        # not optimal yet: first release a row whose multiplier has the wrong sign, else add the
        # most violated row; never revisit an active set (anti-cycling)
        if na and lam_a.min() < -1e-9 * scale:
            act.pop(int(np.argmin(lam_a)))
        elif not feasible and worst not in act:
            act.append(worst)
        else:
            return False
        key = tuple(sorted(int(a) for a in act))
        if key in seen:
            return False
        seen.add(key)
    return False


def _max_step(v, dv):
    neg = dv < 0
    if not np.any(neg):
        return 1.0
    return min(1.0, float(np.min(-v[neg] / dv[neg])))


def is_feasible(E, e, G, h):
    """Phase-1 LP (HiGHS); used only to label an IPM failure as 'infeasible'."""
    from scipy.optimize import linprog
    res = linprog(np.zeros(E.shape[1]), A_ub=G if G.shape[0] else None, b_ub=h if G.shape[0] else None,
                  A_eq=E, b_eq=e, bounds=(None, None))
    return res.status == 0


def solve_param(qp, x_init, ref=None, tol=1e-13):
    """Solve one instance; returns (x[nx,N+1], u[nu,N], x_bar, u_bar) / (x,u) or Nones, plus result."""
    q, e, h = qp.params(x_init, ref)
    r = solve_qp(qp.P, q, qp.E, e, qp.G, h, tol=tol)
    if r.status != "optimal":
        if not is_feasible(qp.E, e, qp.G, h):
            r.status = "infeasible"
    if r.status == "infeasible":
        return (None,) * (4 if qp.has_ss else 2), r
    return qp.split(r.z), r

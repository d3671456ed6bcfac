"""Condensed form of the reference's MPC quadratic programs, built once on the host.

The reference states its QPs in cvxpy over (x_0..x_N, u_0..u_{N-1}, x_bar, u_bar) with the
dynamics as equality constraints (``TubeTrackingMPC.py:104-156``, ``TrackingMPC.py:62-114``,
``TubeRegulatorMPC.py:109-143``, ``RegulatorMPC.py:45-76``, ``TubeTrackingMPC.py:253-299``).  For a
batch that shares (A, B, Q, R, sets) only the *parameters* x_init and ref differ between
instances, so we eliminate every equality once:

    x_i        = A^i x_0 + sum_{j<i} A^{i-1-j} B u_j          (x_0 = x_init, or a decision variable)
    [x_bar;u_bar] = M_ss theta,   M_ss = null([A - I, B])

and hand the GPU the dense problem in  zeta = [x_0 (tube-initial variants only) | u | theta]:

    min 1/2 zeta' H zeta + (Fx x_init + Fr ref)' zeta     s.t.   lo(x_init) <= G zeta <= up(x_init)

``H`` and ``G`` are shared by the whole batch; that is what lets one factorisation serve every
instance.  Opposite one-sided rows (a'z <= b1, -a'z <= b2: every box row and every facet pair of
the symmetric terminal / tube sets) are merged into one two-sided row, halving the row count.
Rows that do not involve zeta at all (e.g. ``Hx x_0 <= hx`` when x_0 = x_init, the reference
constrains x_0, ``TubeTrackingMPC.py:139``) become *parameter rows*: pure feasibility tests on
x_init.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import scipy.linalg as sla

INF = 1e30


@dataclass
class MPCSpec:
    """One of the reference's QP variants in plain matrices."""
    A: np.ndarray
    B: np.ndarray
    Q: np.ndarray
    R: np.ndarray
    N: int
    P_term: Optional[np.ndarray] = None            # terminal weight (None: RegulatorMPC has none)
    T_ss: Optional[np.ndarray] = None              # steady-state offset weight; None => no (x_bar,u_bar)
    stage_x: Optional[Tuple[np.ndarray, np.ndarray]] = None   # Hx x_i <= hx, i = 0..N-1
    stage_u: Optional[Tuple[np.ndarray, np.ndarray]] = None   # Hu u_i <= hu, i = 0..N-1
    terminal: Optional[Tuple[np.ndarray, np.ndarray]] = None  # rows on (x_N, x_bar, u_bar) or on x_N
    terminal_eq: bool = False                      # x_N == x_bar (TrackingMPC without Xf)
    tube_init: Optional[Tuple[np.ndarray, np.ndarray]] = None  # Hz (x_init - x_0) <= hz ; None => x_0 = x_init
    g2_free_terminal: bool = False                 # reproduce SURVEY G2 (TubeTrackingMPC.py:293)


@dataclass
class CondensedQP:
    nx: int
    nu: int
    N: int
    n: int                       # decision size
    nth: int                     # steady-state parameters
    has_x0: bool
    H: np.ndarray                # [n,n]
    Fx: np.ndarray               # [n,nx]
    Fr: np.ndarray               # [n,nx]
    G: np.ndarray                # [m,n]
    lo0: np.ndarray              # [m]
    up0: np.ndarray              # [m]
    Lx: np.ndarray               # [m,nx]   lo = lo0 + Lx x_init
    Ux: np.ndarray               # [m,nx]   up = up0 + Ux x_init
    par_h: np.ndarray            # [mp]     parameter rows:  par_C x_init <= par_h
    par_C: np.ndarray            # [mp,nx]
    Phi: np.ndarray              # [nz,n]   z = Phi zeta + Psi x_init ;  z = [x_0..x_N | u | x_bar | u_bar]
    Psi: np.ndarray              # [nz,nx]
    Mss: np.ndarray              # [(nx+nu), nth]
    shift: np.ndarray = None     # [m] int32: row holding the same constraint one stage earlier (-1: none)
    meta: dict = field(default_factory=dict)

    @property
    def m(self):
        return self.G.shape[0]


def steady_state_basis(A, B):
    """Basis of {(x_bar,u_bar) : (A-I) x_bar + B u_bar = 0} whose coordinates are actual entries of
    (x_bar,u_bar) (pivot rows = identity), so theta is e.g. the cart position for the cartpole."""
    nx, nu = B.shape
    Nsp = sla.null_space(np.c_[A - np.eye(nx), B])
    nth = Nsp.shape[1]
    if nth == 0:
        return np.zeros((nx + nu, 0))
    _, _, piv = sla.qr(Nsp.T, pivoting=True)
    rows = np.sort(piv[:nth])
    Mss = Nsp @ np.linalg.inv(Nsp[rows, :])
    Mss[np.abs(Mss) < 1e-13] = 0.0
    return Mss


def condense(spec: MPCSpec) -> CondensedQP:
    A = np.asarray(spec.A, float)
    B = np.asarray(spec.B, float)
    Q = np.asarray(spec.Q, float)
    R = np.atleast_2d(np.asarray(spec.R, float))
    nx, nu = B.shape
    N = int(spec.N)
    has_ss = spec.T_ss is not None
    has_x0 = spec.tube_init is not None
    Mss = steady_state_basis(A, B) if has_ss else np.zeros((nx + nu, 0))
    nth = Mss.shape[1]
    o_u = nx if has_x0 else 0
    o_th = o_u + nu * N
    n = o_th + nth

    # affine maps  x_i = Px[i] zeta + Sx[i] x_init ,  u_i = Pu[i] zeta
    Px = np.zeros((N + 1, nx, n))
    Sx = np.zeros((N + 1, nx, nx))
    Pu = np.zeros((N, nu, n))
    if has_x0:
        Px[0][:, :nx] = np.eye(nx)
    else:
        Sx[0] = np.eye(nx)
    for i in range(N):
        Pu[i][:, o_u + nu * i:o_u + nu * (i + 1)] = np.eye(nu)
        Px[i + 1] = A @ Px[i] + B @ Pu[i]
        Sx[i + 1] = A @ Sx[i]
    Pxb = np.zeros((nx, n))
    Pub = np.zeros((nu, n))
    if has_ss:
        Pxb[:, o_th:] = Mss[:nx]
        Pub[:, o_th:] = Mss[nx:]

    # cost: sum of (D zeta + dx x_init + dr ref)' M (D zeta + ...)   (no 1/2 in the reference)
    H = np.zeros((n, n))
    Fx = np.zeros((n, nx))
    Fr = np.zeros((n, nx))

    def quad(D, M, dx=None, dr=None):
        nonlocal H, Fx, Fr
        H += 2.0 * D.T @ M @ D
        if dx is not None:
            Fx += 2.0 * D.T @ M @ dx
        if dr is not None:
            Fr += 2.0 * D.T @ M @ dr

    for i in range(N):
        quad(Px[i] - Pxb, Q, dx=Sx[i])
        quad(Pu[i] - Pub, R)
    if spec.P_term is not None:
        quad(Px[N] - Pxb, np.asarray(spec.P_term, float), dx=Sx[N])
    if has_ss:
        quad(Pxb, np.asarray(spec.T_ss, float), dr=-np.eye(nx))
    H = 0.5 * (H + H.T)

    # one-sided rows  g' zeta <= h0 + hx' x_init
    rows_g, rows_h0, rows_hx = [], [], []
    rows_tag = []            # (kind, stage, index within the block) per one-sided row, for the shift map
    eq_g, eq_h0, eq_hx = [], [], []

    def ineq(C, c, Pz, Sz=None, kind="other", stage=0):
        C = np.asarray(C, float)
        rows_tag.extend((kind, stage, j) for j in range(C.shape[0]))
        rows_g.append(C @ Pz)
        rows_h0.append(np.asarray(c, float).flatten())
        rows_hx.append(-C @ Sz if Sz is not None else np.zeros((C.shape[0], nx)))

    if has_x0:
        Hz, hz = spec.tube_init
        Hz = np.asarray(Hz, float)
        # Hz (x_init - x_0) <= hz   <=>   -Hz x_0 <= hz - Hz x_init
        rows_tag.extend(("init", 0, j) for j in range(Hz.shape[0]))
        rows_g.append(-Hz @ Px[0])
        rows_h0.append(np.asarray(hz, float).flatten())
        rows_hx.append(-Hz)
    for i in range(N):
        if spec.stage_x is not None:
            ineq(spec.stage_x[0], spec.stage_x[1], Px[i], Sx[i], kind="x", stage=i)
        if spec.stage_u is not None:
            ineq(spec.stage_u[0], spec.stage_u[1], Pu[i], kind="u", stage=i)
    if spec.terminal_eq:
        eq_g.append(Px[N] - Pxb)
        eq_h0.append(np.zeros(nx))
        eq_hx.append(-Sx[N])
    elif spec.terminal is not None:
        HN, hN = np.asarray(spec.terminal[0], float), np.asarray(spec.terminal[1], float).flatten()
        if has_ss:
            if spec.g2_free_terminal:
                Cth, hth = _project_terminal_on_theta(HN, hN, Mss, nx, nu)
                Pth = np.zeros((nth, n))
                Pth[:, o_th:] = np.eye(nth)
                ineq(Cth, hth, Pth, kind="term")
            else:
                rows_tag.extend(("term", 0, j) for j in range(HN.shape[0]))
                rows_g.append(HN[:, :nx] @ Px[N] + HN[:, nx:2 * nx] @ Pxb + HN[:, 2 * nx:] @ Pub)
                rows_h0.append(hN)
                rows_hx.append(-HN[:, :nx] @ Sx[N])
        else:
            ineq(HN, hN, Px[N], Sx[N], kind="term")

    Gi = np.vstack(rows_g) if rows_g else np.zeros((0, n))
    h0 = np.hstack(rows_h0) if rows_h0 else np.zeros(0)
    hx = np.vstack(rows_hx) if rows_hx else np.zeros((0, nx))

    # parameter rows (do not involve zeta)
    gn = np.abs(Gi).max(axis=1) if Gi.shape[0] else np.zeros(0)
    const = gn < 1e-12
    par_h, par_C = h0[const].copy(), -hx[const].copy()       # par_C x_init <= par_h
    Gi, h0, hx = Gi[~const], h0[~const], hx[~const]
    tags = [t for t, c in zip(rows_tag, const) if not c]

    G, lo0, up0, Lx, Ux, first = _merge_two_sided(Gi, h0, hx)
    # shift map for warm starts: stage rows move one stage earlier, terminal / initial rows stay
    where = {tags[i]: r for r, i in enumerate(first)}
    shift = np.full(G.shape[0], -1, np.int32)
    for r, i in enumerate(first):
        kind, stage, j = tags[i]
        if kind in ("x", "u"):
            shift[r] = where.get((kind, stage - 1, j), -1)
        else:
            shift[r] = r
    if eq_g:
        Ge = np.vstack(eq_g)
        he = np.hstack(eq_h0)
        hxe = np.vstack(eq_hx)
        G = np.vstack([G, Ge])
        lo0 = np.r_[lo0, he]
        up0 = np.r_[up0, he]
        Lx = np.vstack([Lx, hxe])
        Ux = np.vstack([Ux, hxe])
        shift = np.r_[shift, np.arange(len(shift), len(shift) + Ge.shape[0], dtype=np.int32)]

    nz = nx * (N + 1) + nu * N + ((nx + nu) if has_ss else 0)
    Phi = np.zeros((nz, n))
    Psi = np.zeros((nz, nx))
    for i in range(N + 1):
        Phi[nx * i:nx * (i + 1)] = Px[i]
        Psi[nx * i:nx * (i + 1)] = Sx[i]
    o = nx * (N + 1)
    for i in range(N):
        Phi[o + nu * i:o + nu * (i + 1)] = Pu[i]
    if has_ss:
        o = nx * (N + 1) + nu * N
        Phi[o:o + nx] = Pxb
        Phi[o + nx:o + nx + nu] = Pub
    return CondensedQP(nx=nx, nu=nu, N=N, n=n, nth=nth, has_x0=has_x0, H=H, Fx=Fx, Fr=Fr, G=G, lo0=lo0,
                       up0=up0, Lx=Lx, Ux=Ux, par_h=par_h, par_C=par_C, Phi=Phi, Psi=Psi, Mss=Mss, shift=shift,
                       meta=dict(rows_one_sided=int(Gi.shape[0]), rows_param=int(const.sum())))


def _merge_two_sided(Gi, h0, hx):
    """Pair rows with g_i = -g_j into lo <= g' zeta <= up."""
    m, n = Gi.shape
    nxp = hx.shape[1]
    if m == 0:
        return Gi, np.zeros(0), np.zeros(0), np.zeros((0, nxp)), np.zeros((0, nxp)), []
    nrm = np.linalg.norm(Gi, axis=1)
    U = Gi / nrm[:, None]
    # canonical sign: first entry with |.| > 1e-9 positive
    sign = np.ones(m)
    for i in range(m):
        k = np.nonzero(np.abs(U[i]) > 1e-9)[0][0]
        sign[i] = 1.0 if U[i, k] > 0 else -1.0
    C = U * sign[:, None]
    key = np.round(C / 1e-10).astype(np.int64)
    order = np.lexsort(key.T[::-1])
    used = np.zeros(m, dtype=bool)
    out_rows = []
    groups = {}
    for i in order:
        groups.setdefault(key[i].tobytes(), []).append(i)
    for i in range(m):
        if used[i]:
            continue
        used[i] = True
        mate = None
        for j in groups[key[i].tobytes()]:
            # (the x_init dependence has to be opposite as well: a merged row then has a constant width
            #  up - lo, which the active-set kernel relies on)
            if not used[j] and sign[j] == -sign[i] and abs(nrm[j] - nrm[i]) <= 1e-9 * nrm[i] and \
                    np.abs(hx[j] + hx[i]).max(initial=0.0) <= 1e-12 * (1.0 + np.abs(hx[i]).max(initial=0.0)):
                mate = j
                break
        if mate is not None:
            used[mate] = True
        out_rows.append((i, mate))
    mm = len(out_rows)
    G = np.zeros((mm, n))
    lo0 = np.full(mm, -INF)
    up0 = np.zeros(mm)
    Lx = np.zeros((mm, nxp))
    Ux = np.zeros((mm, nxp))
    for r, (i, j) in enumerate(out_rows):
        G[r] = Gi[i]
        up0[r] = h0[i]
        Ux[r] = hx[i]
        if j is not None:
            lo0[r] = -h0[j]
            Lx[r] = hx[i]
    return G, lo0, up0, Lx, Ux, [i for i, _ in out_rows]


def _project_terminal_on_theta(HN, hN, Mss, nx, nu):
    """G2 mode: x_N and u_bar in the terminal rows are free variables of the "packet received" problem
    (``TubeTrackingMPC.py:293``), so what remains of the terminal set is its projection on the steady-state parameter,
    ``{theta : exists (y, v): HN [y; Mss_x theta; v] <= hN}``.  Returns rows ``(C, h)`` with ``C theta <= h``.
    dim(theta) = 1 (both systems of the reference): the interval, by two LPs.  dim(theta) > 1 (multi-input plants): the
    polytope ``{(y, theta, v)}`` is vertex-enumerated, the vertices are projected and their convex hull taken - exact."""
    from scipy.optimize import linprog
    nth = Mss.shape[1]
    Aub = np.c_[HN[:, :nx], HN[:, nx:2 * nx] @ Mss[:nx], HN[:, 2 * nx:]]
    if nth == 1:
        c = np.zeros(Aub.shape[1])
        c[nx] = 1.0
        lo = linprog(c, A_ub=Aub, b_ub=hN, bounds=(None, None))
        hi = linprog(-c, A_ub=Aub, b_ub=hN, bounds=(None, None))
        return np.array([[1.0], [-1.0]]), np.array([-hi.fun, -lo.fun])
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    nrm = np.linalg.norm(Aub, axis=1)
    keep = nrm > 1e-12
    A, b = Aub[keep] / nrm[keep, None], hN[keep] / nrm[keep]
    cc = np.zeros(A.shape[1] + 1)                                   # Chebyshev centre as the interior point
    cc[-1] = -1.0
    res = linprog(cc, A_ub=np.c_[A, np.ones(len(b))], b_ub=b, bounds=[(None, None)] * A.shape[1] + [(0, None)])
    if res.status != 0 or res.x[-1] <= 1e-9:
        raise ValueError("terminal set has no interior: cannot project it on the steady-state parameter")
    V = HalfspaceIntersection(np.c_[A, -b], res.x[:-1]).intersections
    V = V[np.all(np.isfinite(V), axis=1)]
    hull = ConvexHull(V[:, nx:nx + nth])
    return hull.equations[:, :-1].copy(), -hull.equations[:, -1].copy()

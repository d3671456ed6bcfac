"""Oracle (test infrastructure): set computations of ``utils_polytope.py`` restated on scipy.

Each support value is one HiGHS LP, exactly like the reference (``utils_polytope.py:12-23``).
The iteration structure, stopping rules and the ``(1+eps)`` factors follow the reference line by
line (cited per function); loops are written in our own style.
"""
import numpy as np
from scipy.optimize import linprog
from scipy.spatial import ConvexHull

from . import ref_polytope as rp
from .ref_polytope import Polytope


def support(poly, a):
    """h_P(a) = max a^T x over {A x <= b}; one LP with free variables (``utils_polytope.py:12-23``)."""
    a = np.asarray(a, dtype=float).flatten()
    res = linprog(c=-a, A_ub=poly.A, b_ub=poly.b, bounds=(None, None))
    return -res.fun


def pont_diff(p1, p2):
    """P1 (-) P2: keep P1's rows, shrink each bound by h_{P2}(row)  (``utils_polytope.py:25-38``)."""
    shrink = np.array([support(p2, row) for row in p1.A])
    return Polytope(p1.A, p1.b - shrink)


def determine_convex_hull(V):
    """``utils_polytope.py:160-178``."""
    V = np.asarray(V, dtype=float)
    if V.shape[1] == 1:
        return rp.qhull(V)
    hull = ConvexHull(V)
    eq = hull.equations
    return Polytope(eq[:, :-1], -eq[:, -1], vertices=V[hull.vertices, :])


def mink_sum(p1, p2):
    """Minkowski sum by summing all vertex pairs and taking the hull; a 1-D array translates
    (``utils_polytope.py:40-113``)."""
    V1 = rp.extreme(p1) if p1.vertices is None else p1.vertices
    if isinstance(p2, Polytope):
        V2 = rp.extreme(p2) if p2.vertices is None else p2.vertices
    else:
        p2 = np.asarray(p2, dtype=float)
        if p2.ndim == 1:
            return Polytope(p1.A, p1.b + p1.A @ p2)
        V2 = p2
    sums = (V1[:, None, :] + V2[None, :, :]).reshape(-1, V1.shape[1])
    return determine_convex_hull(sums)


def scale(poly, s):
    """Scalar or matrix image of a polytope (``utils_polytope.py:115-158``)."""
    if np.ndim(np.squeeze(s)) == 0:
        s = float(np.squeeze(s))
        if s == 1:
            return poly.copy()
        if s == 0:
            n = poly.dim
            return Polytope(np.r_[np.eye(n), -np.eye(n)], np.zeros(2 * n))
        if s > 0:
            return Polytope(poly.A, s * poly.b)
        return Polytope(poly.A / s, poly.b)
    V = rp.extreme(poly) if poly.vertices is None else poly.vertices
    return determine_convex_hull(V @ np.asarray(s, dtype=float).T)


def rakovic_mrpi(A, W, eps_var=1.9e-5, s_max=20):
    """Rakovic et al. Algorithm 1 (``utils_polytope.py:180-245``).  Returns (F_alpha_s, status, s, alpha)."""
    if np.any(W.b <= 0):
        return None, -1, 0, None
    F, g = W.A, W.b
    nx = A.shape[0]
    Apow = [np.linalg.matrix_power(A, i) for i in range(s_max)]
    Mpos = np.zeros(nx)
    Mneg = np.zeros(nx)
    status, s, alpha = -1, 0, None
    while s < s_max - 1:
        s += 1
        alpha = max(support(W, Apow[s].T @ F[i]) / g[i] for i in range(F.shape[0]))
        for j in range(nx):
            Mpos[j] += support(W, Apow[s - 1][j, :])
            Mneg[j] += support(W, -Apow[s - 1][j, :])
        Ms = max(Mpos.max(), Mneg.max())
        if alpha <= eps_var / (eps_var + Ms):
            status = 0
            break
    if status != 0:
        return None, status, s, alpha
    VW = rp.extreme(W)
    Fs = Polytope(W.A, W.b)
    for i in range(1, s):
        Fs = mink_sum(Fs, VW @ Apow[i].T)
    return scale(Fs, 1.0 / (1.0 - alpha)), status, s, alpha


def moas(A, X, max_iter=10000):
    """Gilbert-Tan maximal output admissible set, Algorithm 3.1 (``utils_polytope.py:247-268``).
    Returns (O_inf, t_star)."""
    G, f = X.A, X.b
    Ot = X
    Apow = np.eye(A.shape[0])
    for t in range(max_iter):
        Apow = Apow @ A
        Onext = Ot.intersect(Polytope(G @ Apow, f))
        if Ot == Onext:
            return Ot, t
        Ot = Onext
    raise RuntimeError("MOAS did not converge")


def darup_rpi(A, W, X, U, K, eps_var=1e-4, s_max=20):
    """Darup-Teichrib RPI (``utils_polytope.py:270-414``).  Returns (rpi, C, status, k_star)."""
    if np.any(W.b <= 0):
        return None, None, -1, 0
    Hw, hw = W.A, W.b
    Hd = np.r_[X.A, -U.A @ K]
    hd = np.r_[X.b, U.b]
    nw, nd = Hw.shape[0], Hd.shape[0]
    Apow = [np.linalg.matrix_power(A, i) for i in range(s_max)]
    bc = np.zeros((nd, s_max))
    k, found = 1, False
    while k < s_max and not found:
        HdA = Hd @ Apow[k - 1]
        HwA = Hw @ Apow[k]
        cond_a = True
        for i in range(nw):                      # eq. (10) -> condition (9a), early exit (:329-336)
            if not (1 + eps_var) * support(W, HwA[i]) <= eps_var * hw[i]:
                cond_a = False
                break
        for l in range(nd):                      # eq. (12) -> condition (9b) (:343-349)
            bc[l, k - 1] = (bc[l, k - 2] if k > 1 else 0.0) + support(W, HdA[l])
        cond_b = bool(np.all((1 + eps_var) * bc[:, k - 1] <= hd))
        if cond_a and cond_b:
            found = True
        else:
            k += 1
    if not found:
        return None, None, -1, k
    hc = (1 + eps_var) * bc[:, k - 1]            # container C of Theorem 1 (:371-375)
    C = Polytope(Hd, hc)
    Hc = Hd
    HcAk = Hc @ Apow[k]
    for i in range(nd):                          # condition (27) (:380-396)
        if not (1 + eps_var) * support(C, HcAk[i]) <= eps_var * hc[i]:
            return None, C, -1, k
    Hp = [Hc]                                    # eq. (28) (:402-410)
    hp = [hc]
    for i in range(1, k):
        Hp.append(Hc @ Apow[i])
        hp.append(hc - bc[:, i - 1])
    return Polytope(np.vstack(Hp), np.hstack(hp)), C, 0, k


def determine_mrpi(Acl, W, X, U, K, eps_var=1.9e-5, rpi_method=0):
    """``TubeRegulatorMPC.determine_mRPI`` (``TubeRegulatorMPC.py:26-78``): retry with 10 x s_max
    until the approximation succeeds, then ``pc.reduce``."""
    if np.max(np.abs(np.linalg.eigvals(Acl))) >= 1:
        return None
    s_max = 200
    while True:
        if rpi_method == 1:
            Fs, _, status, _ = darup_rpi(Acl, W, X, U, K, eps_var=eps_var, s_max=s_max)
        else:
            Fs, status, _, _ = rakovic_mrpi(Acl, W, eps_var=eps_var, s_max=s_max)
        if status == 0:
            break
        s_max *= 10
    return rp.reduce(Fs)


def tighten(X, U, Z, K):
    """``TubeTrackingMPC.tighten_constraints`` (``TubeTrackingMPC.py:90-102``)."""
    Uc = pont_diff(U, scale(Z, -K))
    Xc = pont_diff(X, Z)
    return Xc, Uc


def tracking_terminal_set(A, B, K, Xc, Uc, lam):
    """``TubeTrackingMPC.determine_Xf`` / ``TrackingMPC.determine_Xf`` (``TubeTrackingMPC.py:35-61``,
    ``TrackingMPC.py:160-186``): MOAS of the (x, x_bar, u_bar) system.  Like the reference, the
    zero blocks hard-code 2nx / 2nu rows."""
    nx, nu = B.shape
    Acl = A - B @ K
    Hx, hx, Hu, hu = Xc.A, Xc.b, Uc.A, Uc.b
    Ae = np.block([[Acl, B @ K, B],
                   [np.zeros((nx, nx)), np.eye(nx), np.zeros((nx, nu))],
                   [np.zeros((nu, nx)), np.zeros((nu, nx)), np.eye(nu)]])
    Hcl = np.block([[Hx, np.zeros((2 * nx, nx)), np.zeros((2 * nx, nu))],
                    [-Hu @ K, Hu @ K, Hu],
                    [np.zeros((2 * nx, nx)), Hx, np.zeros((2 * nx, nu))],
                    [np.zeros((2 * nu, nx)), np.zeros((2 * nu, nx)), Hu]])
    hcl = np.r_[hx, hu, lam * hx, lam * hu]
    Xf, t_star = moas(Ae, Polytope(Hcl, hcl))
    return Xf, t_star


def regulator_terminal_set(Acl, K, Xc, Uc):
    """``TubeRegulatorMPC.determine_Xf`` (``TubeRegulatorMPC.py:91-107``)."""
    XU = Polytope(np.r_[Xc.A, -Uc.A @ K], np.r_[Xc.b, Uc.b])
    return moas(Acl, XU)

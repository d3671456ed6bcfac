"""Set computations behind the QP data -- mirror of the reference's ``utils_polytope.py``.

Same function names, argument meaning and return values as the reference module so scripts written
against ``LinearMPCOverNetworks.utils_polytope`` keep working.  What changes is *how* support
functions are evaluated: the reference solves one HiGHS LP per direction
(``utils_polytope.py:12-23``); here every algorithm first collects all directions it will need
(``h_W(A_i' a)`` for all i at once, north-star item 4) and evaluates them in one
``rtmpc_support_sweep`` launch over the vertex representation (max over vertices).  Sets without a
tractable vertex representation (the 9-D augmented terminal sets) use HiGHS LPs on the host, as
the reference does.

``set_support_backend`` exists for the CPU-only unit tests, which inject the oracle's LP support;
the default backend is the CUDA kernel and raises when no GPU is present.
"""
import numpy as np

from . import _lib
from . import polytope as pc
from .polytope import Polytope

_backend = None      # callable(V, dirs) -> values, or None for the CUDA sweep


def set_support_backend(fn):
    """fn(V[nv,dim], dirs[M,dim]) -> [M]; None restores the CUDA kernel."""
    global _backend
    _backend = fn


def support_sweep(V, dirs):
    """max_v <d, v> for every row d of ``dirs`` on the GPU (``rtmpc_support_sweep_host``)."""
    V = _lib.f64(V)
    dirs = _lib.f64(np.atleast_2d(dirs))
    if _backend is not None:
        return np.asarray(_backend(V, dirs), float)
    L = _lib.lib()
    _lib.require_cuda()
    out = np.empty(dirs.shape[0])
    nv, dim = V.shape
    chunk = max(8, ((190 * 1024) // (8 * 4 * ((dim + 3) // 4))) & ~7)     # vertices per launch (shared-memory staging, coordinates padded to 4)
    best = np.full(dirs.shape[0], -np.inf)
    for s in range(0, nv, chunk):
        Vc = _lib.f64(V[s:s + chunk])
        _lib.check(L.rtmpc_support_sweep_host(_lib.ptr(Vc), Vc.shape[0], dim, _lib.ptr(dirs), dirs.shape[0],
                                              _lib.ptr(out)), "rtmpc_support_sweep_host")
        best = np.maximum(best, out)
    return best


def _vertices(poly):
    if getattr(poly, "vertices", None) is not None:
        return poly.vertices
    if poly.A.shape[1] > 6:
        return None
    return pc.extreme(poly if isinstance(poly, Polytope) else Polytope(poly.A, poly.b, normalize=False))


def support_batch(poly, dirs):
    """h_P(d) for every row of ``dirs``."""
    dirs = np.atleast_2d(np.asarray(dirs, float))
    V = _vertices(poly)
    if V is None:
        return pc.support_lp(poly, dirs)
    return support_sweep(V, dirs)


def support(poly, x):
    """``utils_polytope.support`` (``utils_polytope.py:12-23``)."""
    return float(support_batch(poly, np.asarray(x, float).reshape(1, -1))[0])


def pont_diff(poly1, poly2):
    """``utils_polytope.pont_diff`` (``:25-38``): all facet directions of poly1 in one sweep."""
    return Polytope(poly1.A, np.asarray(poly1.b, float).flatten() - support_batch(poly2, poly1.A))


def determine_convex_hull(vertices):
    """``:160-178``."""
    return pc.qhull(np.asarray(vertices, float))


def mink_sum(poly1, poly2):
    """``:40-113``."""
    V1 = _vertices(poly1)
    if isinstance(poly2, np.ndarray):
        if poly2.ndim == 1:
            return Polytope(poly1.A, poly1.b + poly1.A @ poly2)
        if poly2.ndim != 2:
            print("If the input is a numpy array it should have dimension 1 or 2")
            return None
        V2 = poly2
    elif hasattr(poly2, "A"):
        V2 = _vertices(poly2)
    else:
        print("Input has the wrong type")
        return None
    return determine_convex_hull((V1[:, None, :] + V2[None, :, :]).reshape(-1, V1.shape[1]))


def scale(poly, scaling_variable):
    """``:115-158``."""
    if np.isscalar(scaling_variable) or np.ndim(np.squeeze(scaling_variable)) == 0:
        s = float(np.squeeze(scaling_variable))
        if s == 1:
            return poly.copy()
        if s == 0:
            n = poly.A.shape[1]
            return Polytope(np.r_[np.eye(n), -np.eye(n)], np.zeros(2 * n))
        return Polytope(poly.A, s * poly.b) if s > 0 else Polytope(poly.A / s, poly.b)
    if isinstance(scaling_variable, np.ndarray):
        return determine_convex_hull(_vertices(poly) @ scaling_variable.T)
    print("The input is neither an np.ndarray nor a scalar. Therefore, we return None")
    return None


def _powers(A, count):
    P = np.empty((count,) + A.shape)
    P[0] = np.eye(A.shape[0])
    for i in range(1, count):
        P[i] = P[i - 1] @ A
    return P


def _check_inputs(A, W):
    if A.shape[0] != A.shape[1]:
        print("A needs to be a square matrix. Returning None")
        return False
    if np.sum(np.asarray(W.b) <= 0) != 0:
        print("The polytope W does not contain the origin. Therefore, we return None")
        return False
    return True


def calculate_minimal_robust_positively_invariant_set(A, W, eps_var=1.9e-5, s_max=20):
    """Rakovic et al. Algorithm 1 (``:180-245``).  Returns (F_alpha_s, status)."""
    if not _check_inputs(A, W):
        return None
    F, g = W.A, np.asarray(W.b, float).flatten()
    nw, nx = F.shape
    Apow = _powers(A, s_max)
    # every direction the loop would ask for, s = 1 .. s_max-1
    S = s_max - 1
    d_alpha = np.einsum("sij,wi->swj", Apow[1:], F).reshape(S * nw, nx)       # (A^s)' f_i  as rows f_i' A^s
    d_pos = Apow[:S].reshape(S * nx, nx)                                       # rows of A^{s-1}
    h = support_batch(W, np.vstack([d_alpha, d_pos, -d_pos]))
    alpha = (h[:S * nw].reshape(S, nw) / g).max(axis=1)
    Mpos = np.cumsum(h[S * nw:S * nw + S * nx].reshape(S, nx), axis=0)
    Mneg = np.cumsum(h[S * nw + S * nx:].reshape(S, nx), axis=0)
    Ms = np.maximum(Mpos.max(axis=1), Mneg.max(axis=1))
    hit = np.nonzero(alpha <= eps_var / (eps_var + Ms))[0]
    if hit.size == 0:
        print(f"In the RPI calculation, we reached the iteration maximum {s_max} without converging!")
        return None, -1
    s = int(hit[0]) + 1
    VW = _vertices(W)
    Fs = Polytope(W.A, W.b)
    for i in range(1, s):
        Fs = mink_sum(Fs, VW @ Apow[i].T)
    return scale(Fs, 1.0 / (1.0 - alpha[s - 1])), 0


def calculate_maximum_admissible_output_set(A, X, max_iter=100000, verbose=True, growth=4.0):
    """Gilbert-Tan Algorithm 3.1 (``:247-268``).  Same stopping rule as the reference
    (``O_t == O_{t+1}`` in the `polytope` sense: no cut-off piece with Chebyshev radius > 1e-7) and the same
    set, but without the reference's full redundancy removal in every iteration: only the
    new rows are tested against the current set (one LP each), only rows that cut are appended, and the
    LP-per-row ``reduce`` runs once at the end (and whenever the working representation has grown by more than
    ``growth`` x since the last one).  Every group of LPs is one batched ``rtmpc_lp_solve`` launch (``polytope.lp_batch``).
    9-D cartpole terminal set: 128 s (reference's loop, HiGHS) -> 21 s (lazy loop, HiGHS) -> see DESIGN.md (GPU LPs)."""
    G, f = X.A, np.asarray(X.b, float).flatten()
    Ot = X if isinstance(X, Polytope) else Polytope(X.A, X.b)
    Ot = pc.reduce(Ot)
    base_rows = Ot.A.shape[0]
    Ap = np.eye(A.shape[0])
    for t in range(max_iter):
        Ap = Ap @ A
        new = Polytope(G @ Ap, f)
        cuts = pc.support_lp(Ot, new.A) > new.b               # one batch of LPs over the rows of O_t
        # O_t == O_{t+1} unless a new row cuts off a piece with Chebyshev radius > 1e-7 (one batch of LPs, one per cutting row)
        changed = bool(cuts.any()) and bool(np.any(pc.cut_radii(Ot, new.A[cuts], new.b[cuts]) > pc.ABS_TOL))
        if not changed:
            if verbose:
                print(f"Admissible set calculation has converged at t = {t}")
            return pc.reduce(Ot)
        Ot = Polytope(np.vstack([Ot.A, new.A[cuts]]), np.hstack([Ot.b, new.b[cuts]]), normalize=False)
        if Ot.A.shape[0] > growth * base_rows + 64:
            Ot = pc.reduce(Ot)
            base_rows = Ot.A.shape[0]
    raise RuntimeError("maximum admissible output set did not converge")


def calculate_RPI(A, W, X, U, K, eps_var=1e-4, s_max=20, return_container=False, verbose=True):
    """Darup-Teichrib RPI (``:270-414``).  Returns (rpi, status) or (rpi, C, status)."""
    if not _check_inputs(A, W):
        return None
    Hw, hw = W.A, np.asarray(W.b, float).flatten()
    Hd = np.r_[X.A, -U.A @ K]
    hd = np.r_[np.asarray(X.b, float).flatten(), np.asarray(U.b, float).flatten()]
    nw, nd, nx = Hw.shape[0], Hd.shape[0], A.shape[0]
    Apow = _powers(A, s_max)
    S = s_max - 1                                   # candidate k = 1 .. s_max-1
    d_w = np.einsum("wi,kij->kwj", Hw, Apow[1:]).reshape(S * nw, nx)          # rows of Hw A^k
    d_d = np.einsum("di,kij->kdj", Hd, Apow[:S]).reshape(S * nd, nx)          # rows of Hd A^{k-1}
    h = support_batch(W, np.vstack([d_w, d_d]))
    hw_k = h[:S * nw].reshape(S, nw)
    bc_all = np.cumsum(h[S * nw:].reshape(S, nd), axis=0)                      # eq. (12) partial sums
    cond_a = np.all((1 + eps_var) * hw_k <= eps_var * hw, axis=1)              # eq. (10)
    cond_b = np.all((1 + eps_var) * bc_all <= hd, axis=1)
    hit = np.nonzero(cond_a & cond_b)[0]
    k_star = int(hit[0]) + 1 if hit.size else s_max
    if verbose:
        print(f"k_star = {k_star}")
    if hit.size == 0:
        print(f"In the RPI calculation, we reached the iteration maximum {s_max} without converging!")
        return (None, None, -1) if return_container else (None, -1)
    hc = (1 + eps_var) * bc_all[k_star - 1]
    C = Polytope(Hd, hc)
    hc_support = support_batch(C, Hd @ Apow[k_star])                           # condition (27)
    if not np.all((1 + eps_var) * hc_support <= eps_var * hc):
        print("The container set C does not fulfill the condition for calculating the RPI. Returning None")
        return (None, C, -1) if return_container else (None, -1)
    Hp = [Hd] + [Hd @ Apow[i] for i in range(1, k_star)]
    hp = [hc] + [hc - bc_all[i - 1] for i in range(1, k_star)]
    rpi = Polytope(np.vstack(Hp), np.hstack(hp))
    return (rpi, C, 0) if return_container else (rpi, 0)

"""Minimal H-polytope container so the reference's scripts keep working without the third-party
``polytope`` package (``pc.Polytope(A, b)``, ``.A``, ``.b``, ``in``, ``pc.reduce``, ``pc.extreme``,
``pc.qhull``, ``.intersect``, ``==``).  Anything duck-typed on ``.A`` / ``.b`` is accepted by the
controller classes, so a real ``polytope.Polytope`` works too.

Differences in *how* (not what) from upstream: redundancy removal in dimension <= 4 is one qhull
call on the polar dual instead of one LP per row; in higher dimensions the LPs (one per row for ``reduce``, one per
direction for ``support_lp``, one Chebyshev ball per cut-off piece for the subset test) are solved in batches on the
GPU (``rtmpc_lp_solve``: one warp per LP, dual simplex over the shared rows).

LP backend: ``set_lp_backend('gpu')`` (default; raises without a CUDA device - nothing falls back silently) or
``set_lp_backend('highs')``: one scipy/HiGHS LP per call on the host, which is what the reference itself does
(``utils_polytope.py:19``) and what the CPU-only unit tests select.
"""
import numpy as np
from scipy.optimize import linprog
from scipy.spatial import ConvexHull, HalfspaceIntersection

from . import _lib

ABS_TOL = 1e-7
_lp_backend = "gpu"


def set_lp_backend(name):
    """'gpu' (batched ``rtmpc_lp_solve``) or 'highs' (scipy ``linprog`` on the host, like the reference)."""
    global _lp_backend
    if name not in ("gpu", "highs"):
        raise ValueError(name)
    _lp_backend = name


def lp_backend():
    return _lp_backend


def lp_batch(H, h, obj, relax_row=None, relax_by=0.0, extra=None, want_x=False):
    """max obj[b]' x  s.t.  H x <= h (row relax_row[b] relaxed by relax_by), extra[b] rows ([B, ne, dim+1]: coefficients and
    bound), for all b at once.  Returns (values [B] with +inf = unbounded / -inf = infeasible, x [B, dim] or None)."""
    H = _lib.f64(np.atleast_2d(H))
    h = _lib.f64(np.asarray(h, float).reshape(-1))
    obj = _lib.f64(np.atleast_2d(obj))
    B, dim = obj.shape
    m = H.shape[0]
    ne = 0 if extra is None else extra.shape[1]
    if _lp_backend == "highs":
        val = np.empty(B)
        X = np.zeros((B, dim))
        for b in range(B):
            A, rhs = H, h
            if relax_row is not None and relax_row[b] >= 0:
                rhs = h.copy()
                rhs[relax_row[b]] += relax_by
            if ne:
                A = np.vstack([A, extra[b, :, :dim]])
                rhs = np.r_[rhs, extra[b, :, dim]]
            res = linprog(-obj[b], A_ub=A, b_ub=rhs, bounds=(None, None))
            val[b] = -res.fun if res.status == 0 else (np.inf if res.status == 3 else -np.inf)
            if res.status == 0:
                X[b] = res.x
        return val, (X if want_x else None)
    L = _lib.lib()
    _lib.require_cuda()
    scale = 1.0 + (np.abs(h).max() if m else 0.0) + (np.abs(extra[:, :, dim]).max() if ne else 0.0)
    box = 1e4 * scale
    val = np.empty(B)
    X = np.empty((B, dim))
    st = np.empty(B, np.int32)
    it = np.empty(B, np.int32)
    rr = None if relax_row is None else np.ascontiguousarray(relax_row, np.int32)
    ex = None if not ne else _lib.f64(extra)
    _lib.check(L.rtmpc_lp_solve_host(_lib.ptr(H), _lib.ptr(h), m, dim, _lib.ptr(obj), _lib.ptr(rr), float(relax_by), _lib.ptr(ex),
                                     ne, B, box, _lib.ptr(val), _lib.ptr(X), _lib.ptr(st), _lib.ptr(it)), "rtmpc_lp_solve_host")
    if np.any(st == 1):
        raise _lib.RtmpcError(f"rtmpc_lp_solve: {int((st == 1).sum())} of {B} LPs hit the iteration limit / a singular basis "
                              f"(dim {dim}, {m} rows)")
    val = np.where(st == 4, np.inf, np.where(st == 2, -np.inf, val))
    return val, (X if want_x else None)


class Polytope:
    def __init__(self, A=None, b=None, vertices=None, normalize=True, minrep=False):
        A = np.zeros((0, 0)) if A is None else np.array(A, dtype=float)
        b = np.zeros(0) if b is None else np.array(b, dtype=float).flatten()
        if normalize and A.size:
            nrm = np.linalg.norm(A, axis=1)
            nz = nrm > 1e-10
            A, b = A[nz] / nrm[nz, None], b[nz] / nrm[nz]
        self.A, self.b = A, b
        self.vertices = vertices
        self.minrep = minrep

    @property
    def dim(self):
        return self.A.shape[1]

    def copy(self):
        return Polytope(self.A.copy(), self.b.copy(), None if self.vertices is None else self.vertices.copy(),
                        normalize=False, minrep=self.minrep)

    def __contains__(self, point):
        return bool(np.all(self.A @ np.asarray(point, float).flatten() - self.b < ABS_TOL))

    def contains(self, points):
        """Vectorised membership for rows of ``points``."""
        return np.all(np.atleast_2d(points) @ self.A.T - self.b < ABS_TOL, axis=1)

    def intersect(self, other, abs_tol=ABS_TOL):
        return reduce(Polytope(np.vstack([self.A, other.A]), np.hstack([self.b, other.b])), abs_tol)

    def __le__(self, other):
        return is_subset(self, other)

    def __eq__(self, other):
        return is_subset(self, other) and is_subset(other, self)

    __hash__ = None

    def __repr__(self):
        return f"Polytope(rows={self.A.shape[0]}, dim={self.dim})"


def cheby_ball(poly):
    """Chebyshev radius and centre: max r s.t. A x + r ||a_i|| <= b, r >= 0."""
    A, b = poly.A, poly.b
    n = A.shape[1]
    if _lp_backend == "gpu":
        H = np.vstack([np.c_[A, np.linalg.norm(A, axis=1)], np.r_[np.zeros(n), -1.0][None, :]])
        val, X = lp_batch(H, np.r_[b, 0.0], np.r_[np.zeros(n), 1.0][None, :], want_x=True)
        if not np.isfinite(val[0]):
            return 0.0, None
        return float(X[0, -1]), X[0, :-1].copy()
    c = np.zeros(n + 1)
    c[-1] = -1.0
    res = linprog(c, A_ub=np.c_[A, np.linalg.norm(A, axis=1)], b_ub=b, bounds=[(None, None)] * n + [(0, None)])
    if res.status != 0:
        return 0.0, None
    return float(res.x[-1]), res.x[:-1].copy()


def is_fulldim(poly, abs_tol=ABS_TOL):
    return cheby_ball(poly)[0] > abs_tol


def cut_radii(small, rows, rhs):
    """Chebyshev radius of {x in small : rows[j] x >= rhs[j]} for every j (the pieces ``small \\ {rows[j] x <= rhs[j]}`` of the
    subset test), all in one batch: the rows of ``small`` are shared, each LP has its own cutting row."""
    rows = np.atleast_2d(rows)
    J, n = rows.shape
    if J == 0:
        return np.zeros(0)
    A, b = small.A, small.b
    H = np.vstack([np.c_[A, np.linalg.norm(A, axis=1)], np.r_[np.zeros(n), -1.0][None, :]])      # ..., -r <= 0
    extra = np.zeros((J, 1, n + 2))
    extra[:, 0, :n] = -rows
    extra[:, 0, n] = np.linalg.norm(rows, axis=1)
    extra[:, 0, n + 1] = -np.asarray(rhs, float)
    obj = np.zeros((J, n + 1))
    obj[:, n] = 1.0
    val, _ = lp_batch(H, np.r_[b, 0.0], obj, extra=extra)
    return np.where(np.isfinite(val), val, 0.0)                 # infeasible piece: empty, radius 0


def is_subset(small, big, abs_tol=ABS_TOL):
    """small \\ big has no piece with Chebyshev radius above ``abs_tol``."""
    lp_support = support_lp(small, big.A)
    cut = np.nonzero(lp_support > big.b)[0]           # only rows that cut at all need the radius test
    if cut.size == 0:
        return True
    return not np.any(cut_radii(small, big.A[cut], big.b[cut]) > abs_tol)


def support_lp(poly, dirs):
    """h_P(d) for each row d of ``dirs``: one LP each over the H-representation (for sets that have no tractable vertex
    representation, e.g. the 9-D terminal sets), solved as one batch."""
    dirs = np.atleast_2d(dirs)
    val, _ = lp_batch(poly.A, poly.b, dirs)
    return np.where(val == -np.inf, np.inf, val) if np.any(val == -np.inf) else val


def _drop_parallel(A, b, abs_tol):
    gram = A @ A.T
    ii, jj = np.nonzero(np.triu(gram > 1 - abs_tol, k=1))
    rm = np.zeros(len(b), bool)
    for i, j in zip(ii, jj):
        rm[j if b[i] < b[j] else i] = True
    return A[~rm], b[~rm]


def reduce(poly, abs_tol=ABS_TOL):
    if poly.minrep:
        return poly
    A, b = poly.A[np.isfinite(poly.b)], poly.b[np.isfinite(poly.b)]
    A, b = _drop_parallel(A, b, abs_tol)
    m, n = A.shape
    if m <= n + 1:
        return Polytope(A, b, normalize=False)
    if 2 <= n <= 4:
        r, xc = cheby_ball(Polytope(A, b, normalize=False))
        if xc is not None and r > abs_tol:
            # facets of the polytope = vertices of the polar dual about an interior point
            dual = A / (b - A @ xc)[:, None]
            keep = np.sort(ConvexHull(dual).vertices)
            out = Polytope(A[keep], b[keep], normalize=False)
            out.minrep = True
            return out
    # one LP per row against ALL rows with its own bound relaxed by 0.1 (upstream polytope.reduce); one batch
    val, _ = lp_batch(A, b, A, relax_row=np.arange(m, dtype=np.int32), relax_by=0.1)
    keep = np.nonzero((val == np.inf) | (val - b > abs_tol))[0]
    out = Polytope(A[keep], b[keep], normalize=False)
    out.minrep = True
    return out


def extreme(poly):
    if poly.vertices is not None:
        return poly.vertices
    A, b = poly.A, poly.b
    if A.shape[1] == 1:
        hi = np.min(b[A[:, 0] > 0] / A[A[:, 0] > 0, 0])
        lo = np.max(b[A[:, 0] < 0] / A[A[:, 0] < 0, 0])
        V = np.array([[lo], [hi]])
    else:
        r, xc = cheby_ball(poly)
        if xc is None or r <= 0:
            return None
        V = HalfspaceIntersection(np.c_[A, -b], xc).intersections
        V = V[np.all(np.isfinite(V), axis=1)]
        scale = 1e-9 * max(1.0, np.abs(V).max())
        _, idx = np.unique(np.round(V / scale), axis=0, return_index=True)
        V = V[np.sort(idx)]
    poly.vertices = V
    return V


def qhull(vertices):
    V = np.asarray(vertices, float)
    if V.shape[1] == 1:
        lo, hi = V.min(), V.max()
        return Polytope(np.array([[1.0], [-1.0]]), np.array([hi, -lo]), vertices=np.array([[lo], [hi]]))
    hull = ConvexHull(V)
    return Polytope(hull.equations[:, :-1], -hull.equations[:, -1], vertices=V[hull.vertices])


def box(half_widths):
    h = np.asarray(half_widths, float).flatten()
    n = h.size
    return Polytope(np.r_[np.eye(n), -np.eye(n)], np.r_[h, h])

"""Minimal H-polytope container so the reference's scripts keep working without the third-party
``polytope`` package (``pc.Polytope(A, b)``, ``.A``, ``.b``, ``in``, ``pc.reduce``, ``pc.extreme``,
``pc.qhull``, ``.intersect``, ``==``).  Anything duck-typed on ``.A`` / ``.b`` is accepted by the
controller classes, so a real ``polytope.Polytope`` works too.

Differences in *how* (not what) from upstream: redundancy removal in dimension <= 4 is one qhull
call on the polar dual instead of one LP per row; higher dimensions use HiGHS LPs.
"""
import numpy as np
from scipy.optimize import linprog
from scipy.spatial import ConvexHull, HalfspaceIntersection

ABS_TOL = 1e-7


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
    A, b = poly.A, poly.b
    n = A.shape[1]
    c = np.zeros(n + 1)
    c[-1] = -1.0
    res = linprog(c, A_ub=np.c_[A, np.linalg.norm(A, axis=1)], b_ub=b, bounds=[(None, None)] * n + [(0, None)])
    if res.status != 0:
        return 0.0, None
    return float(res.x[-1]), res.x[:-1].copy()


def is_fulldim(poly, abs_tol=ABS_TOL):
    return cheby_ball(poly)[0] > abs_tol


def is_subset(small, big, abs_tol=ABS_TOL):
    """small \\ big has no piece with Chebyshev radius above ``abs_tol``."""
    lp_support = support_lp(small, big.A)
    for j in np.nonzero(lp_support > big.b)[0]:        # only rows that cut at all need the radius test
        piece = Polytope(np.vstack([small.A, -big.A[j:j + 1]]), np.hstack([small.b, -big.b[j]]), normalize=False)
        if is_fulldim(piece, abs_tol):
            return False
    return True


def support_lp(poly, dirs):
    """h_P(d) for each row d of ``dirs`` by one HiGHS LP each (host path for sets that have no
    tractable vertex representation, e.g. the 9-D terminal sets)."""
    dirs = np.atleast_2d(dirs)
    out = np.empty(dirs.shape[0])
    for i, d in enumerate(dirs):
        res = linprog(-d, A_ub=poly.A, b_ub=poly.b, bounds=(None, None))
        out[i] = -res.fun if res.status == 0 else np.inf
    return out


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
    keep = []
    for k in range(m):
        h = b.copy()
        h[k] += 0.1
        res = linprog(-A[k], A_ub=A, b_ub=h, bounds=(None, None))
        if res.status == 3 or (res.status == 0 and -res.fun - b[k] > abs_tol):
            keep.append(k)
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

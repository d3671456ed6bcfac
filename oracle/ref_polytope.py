"""Oracle (test infrastructure): the subset of the `polytope` package the reference calls.

`polytope` (pin ``polytope>=0.2.4``, reference ``setup.cfg:16``) is absent from this image, so
its published behaviour is restated here from the upstream algorithm descriptions
(recalled; cannot be re-verified offline).  Call sites in the reference:

* constructor with row normalisation, ``.A``/``.b``  -- everywhere
* ``point in poly`` with ``A p - b < 1e-7``        -- ``Results/results_linear_system.py:258``
* ``pc.reduce``                                    -- ``TubeRegulatorMPC.py:74``
* ``pc.extreme``                                   -- ``utils_polytope.py:48,58,145,241``
* ``poly.intersect`` / ``==``                      -- ``utils_polytope.py:259-261``

All LPs go through ``scipy.optimize.linprog`` (HiGHS), the same backend `polytope` falls back
to when cvxopt/glpk is missing.
"""
import numpy as np
from scipy.optimize import linprog
from scipy.spatial import ConvexHull, HalfspaceIntersection

ABS_TOL = 1e-7


class Polytope:
    """H-representation ``{x : A x <= b}``; rows normalised to unit length, zero rows dropped,
    ``b`` flattened (upstream ``Polytope.__init__(normalize=True)``)."""

    def __init__(self, A=None, b=None, vertices=None, normalize=True, minrep=False):
        if A is None:
            A = np.zeros((0, 0))
            b = np.zeros((0,))
        A = np.array(A, dtype=float)
        b = np.array(b, dtype=float).flatten()
        if normalize and A.size > 0:
            nrm = np.sqrt(np.sum(A * A, axis=1))
            pos = np.nonzero(nrm > 1e-10)[0]
            A = A[pos, :] / nrm[pos, None]
            b = b[pos] / nrm[pos]
        self.A = A
        self.b = b
        self.vertices = vertices
        self.minrep = minrep

    @property
    def dim(self):
        return self.A.shape[1]

    def copy(self):
        v = None if self.vertices is None else self.vertices.copy()
        return Polytope(self.A.copy(), self.b.copy(), vertices=v, normalize=False, minrep=self.minrep)

    def __contains__(self, point):
        p = np.asarray(point, dtype=float).flatten()
        return bool(np.all(self.A @ p - self.b < ABS_TOL))

    def intersect(self, other, abs_tol=ABS_TOL):
        """Stack the rows and remove redundancy (upstream ``Polytope.intersect``)."""
        iA = np.vstack([self.A, other.A])
        ib = np.hstack([self.b, other.b])
        return reduce(Polytope(iA, ib), abs_tol=abs_tol)

    def __le__(self, other):
        return is_subset(self, other)

    def __eq__(self, other):
        return is_subset(self, other) and is_subset(other, self)

    __hash__ = None


def cheby_ball(poly):
    """Chebyshev radius and centre: max r s.t. A x + r ||a_i|| <= b."""
    A, b = poly.A, poly.b
    if A.size == 0:
        return 0.0, None
    n = A.shape[1]
    nrm = np.sqrt(np.sum(A * A, axis=1))
    c = np.zeros(n + 1)
    c[-1] = -1.0
    res = linprog(c, A_ub=np.c_[A, nrm], b_ub=b, bounds=[(None, None)] * n + [(0, None)])
    if res.status != 0:
        return 0.0, None
    return float(res.x[-1]), res.x[:-1].copy()


def is_fulldim(poly, abs_tol=ABS_TOL):
    r, _ = cheby_ball(poly)
    return r > abs_tol


def is_subset(small, big, abs_tol=ABS_TOL):
    """``small <= big`` iff the set difference small \\ big has no full-dimensional piece
    (upstream ``is_subset`` -> ``mldivide`` -> ``is_fulldim``)."""
    for j in range(big.A.shape[0]):
        piece = Polytope(np.vstack([small.A, -big.A[j:j + 1]]), np.hstack([small.b, -big.b[j]]),
                         normalize=False)
        if is_fulldim(piece, abs_tol):
            return False
    return True


def bounding_box(poly):
    n = poly.dim
    lb = np.zeros(n)
    ub = np.zeros(n)
    for i in range(n):
        c = np.zeros(n)
        c[i] = 1.0
        lo = linprog(c, A_ub=poly.A, b_ub=poly.b, bounds=(None, None))
        hi = linprog(-c, A_ub=poly.A, b_ub=poly.b, bounds=(None, None))
        lb[i] = lo.fun
        ub[i] = -hi.fun
    return lb, ub


def reduce(poly, abs_tol=ABS_TOL):
    """Remove redundant rows (upstream ``polytope.reduce``): (1) of (near-)parallel rows keep the
    tighter one, (2) drop rows that miss the bounding box, (3) one LP per remaining row against
    *all* remaining rows with its own bound relaxed by 0.1; keep when the LP exceeds the bound by
    more than ``abs_tol``."""
    if poly.minrep:
        return poly
    A = poly.A.copy()
    b = poly.b.copy()
    keep = np.isfinite(b)
    A, b = A[keep], b[keep]
    neq = A.shape[0]
    an = 1.0 / np.sqrt(np.sum(A * A, axis=1))
    An = A * an[:, None]
    bn = b * an
    gram = An @ An.T
    remove = np.zeros(neq, dtype=bool)
    ii, jj = np.nonzero(np.triu(gram > 1 - abs_tol, k=1))
    for i, j in zip(ii, jj):
        if bn[i] < bn[j]:
            remove[j] = True
        else:
            remove[i] = True
    A, b = A[~remove], b[~remove]
    neq, nx = A.shape
    if neq <= nx + 1:
        return Polytope(A, b)
    if neq > 3 * nx:
        lb, ub = bounding_box(Polytope(A, b, normalize=False))
        cand = ~(((A > 0) * A) @ (ub - lb) - (b - A @ lb) < -1e-4)
        A, b = A[cand], b[cand]
    neq, nx = A.shape
    if neq <= nx + 1:
        return Polytope(A, b)
    keep_rows = []
    for k in range(neq):
        h = b.copy()
        h[k] += 0.1
        sol = linprog(-A[k], A_ub=A, b_ub=h, bounds=(None, None))
        if sol.status == 0:
            if -sol.fun - b[k] > abs_tol:
                keep_rows.append(k)
        elif sol.status == 3:
            keep_rows.append(k)
    out = Polytope(A[keep_rows], b[keep_rows])
    out.minrep = True
    return out


def extreme(poly):
    """Vertices (rows) of a bounded full-dimensional polytope (upstream ``pc.extreme`` uses the
    qhull dual; here scipy's ``HalfspaceIntersection`` about the Chebyshev centre)."""
    if poly.vertices is not None:
        return poly.vertices
    A, b = poly.A, poly.b
    n = A.shape[1]
    if n == 1:
        hi = np.min(b[A[:, 0] > 0] / A[A[:, 0] > 0, 0])
        lo = np.max(b[A[:, 0] < 0] / A[A[:, 0] < 0, 0])
        V = np.array([[lo], [hi]])
    else:
        r, xc = cheby_ball(poly)
        if xc is None or r <= 0:
            return None
        hs = HalfspaceIntersection(np.c_[A, -b], xc)
        V = hs.intersections
        V = V[np.all(np.isfinite(V), axis=1)]
        # merge numerically duplicated vertices
        _, idx = np.unique(np.round(V / (1e-9 * max(1.0, np.abs(V).max())), 0), axis=0, return_index=True)
        V = V[np.sort(idx)]
    poly.vertices = V
    return V


def qhull(vertices):
    """Convex hull of the rows of ``vertices`` as a Polytope (upstream ``pc.qhull``; used by the
    reference only for 1-D vertex sets, ``utils_polytope.py:165-167``)."""
    V = np.asarray(vertices, dtype=float)
    if V.shape[1] == 1:
        lo, hi = V.min(), V.max()
        return Polytope(np.array([[1.0], [-1.0]]), np.array([hi, -lo]), vertices=np.array([[lo], [hi]]))
    hull = ConvexHull(V)
    eq = hull.equations
    return Polytope(eq[:, :-1], -eq[:, -1], vertices=V[hull.vertices])

"""Oracle (test infrastructure): stand-in for the absent third-party package ``cvxpy`` (pin ``cvxpy>=1.4.1``,
reference ``setup.cfg:15``; solver ``clarabel``, ``setup.cfg:19``) -- only the modelling calls the reference makes:

    cp.Variable(shape)  cp.Parameter(shape)  var[:, i]  M @ expr  expr +/- expr  expr == rhs  expr <= rhs
    cp.quad_form(expr, M)  cp.Minimize(cost)  cp.Problem(objective, constraints)  prob.solve(...)  prob.status
    var.value  param.value  cp.CLARABEL
    (``RegulatorMPC.py:59-76,88``; ``TubeRegulatorMPC.py:124-141,163``; ``TrackingMPC.py:80-114,129``;
     ``TubeTrackingMPC.py:117-153,183,270-296,320,337``)

Every expression the reference writes is affine in the variables and parameters, every cost term is
``quad_form(affine, M)``.  The stand-in carries affine expressions as explicit coefficient matrices, so that
``Problem`` can state the reference's problem in the standard form

    min 1/2 z'Pz + q(p)'z     s.t.   E z = e(p),   G z <= h(p)

(z = the stacked entries of every variable that occurs, column-major like cvxpy; p = parameter values).  ``solve``
hands that to ``oracle.ref_qp.solve_qp`` -- the oracle's dense interior-point method with certified polish -- because
Clarabel is not available; tolerances passed by the reference (``tol_gap_abs``, ``tol_gap_rel``) are recorded in
``Problem.solver_kwargs`` and otherwise ignored: the stand-in always solves to round-off.  The QP is strictly convex
in the variables that carry cost, so its minimiser is unique there; entries of variables that occur in no cost term
and no constraint (the G2 quirk, ``TubeTrackingMPC.py:293``, pulls whole foreign variables into the second problem)
are left out of the solve and reported as 0.
"""
import numpy as np

CLARABEL = "CLARABEL"
OPTIMAL = "optimal"
INFEASIBLE = "infeasible"


def _as_vec(v, size):
    a = np.asarray(v, dtype=float)
    if a.ndim == 0:
        return np.full(size, float(a))
    a = a.reshape(-1)
    assert a.size == size, (a.size, size)
    return a


class Expr:
    """Affine map of the leaves (variables / parameters): value = sum_leaf C_leaf @ vec(leaf) + const, a 1-D vector."""
    __array_ufunc__ = None          # numpy hands ``ndarray (op) Expr`` to our reflected operators
    __hash__ = object.__hash__

    def __init__(self, size, lin, const):
        self.size = int(size)
        self.lin = lin              # {id(leaf): (leaf, matrix [size, leaf.size])}
        self.const = const

    @property
    def shape(self):
        return (self.size,)

    # ---- algebra ----
    def _coerce(self, other):
        if isinstance(other, Expr):
            if isinstance(other, _Leaf) and len(other.leaf_shape) != 1:
                raise TypeError("2-D variables must be indexed by column first")
            assert other.size == self.size, (other.size, self.size)
            return other
        return Expr(self.size, {}, _as_vec(other, self.size))

    def __add__(self, other):
        o = self._coerce(other)
        lin = dict(self.lin)
        for k, (leaf, M) in o.lin.items():
            lin[k] = (leaf, lin[k][1] + M) if k in lin else (leaf, M)
        return Expr(self.size, lin, self.const + o.const)

    __radd__ = __add__

    def __neg__(self):
        return Expr(self.size, {k: (leaf, -M) for k, (leaf, M) in self.lin.items()}, -self.const)

    def __sub__(self, other):
        return self + (-self._coerce(other))

    def __rsub__(self, other):
        return (-self) + other

    def __rmatmul__(self, M):
        M = np.atleast_2d(np.asarray(M, dtype=float))
        assert M.shape[1] == self.size, (M.shape, self.size)
        return Expr(M.shape[0], {k: (leaf, M @ C) for k, (leaf, C) in self.lin.items()}, M @ self.const)

    def __eq__(self, other):            # noqa: PLW1641  (hash is identity on purpose)
        return Constraint("eq", self - other)

    def __le__(self, other):
        return Constraint("le", self - other)

    def __ge__(self, other):
        return Constraint("le", (-self) + other)


class _Leaf(Expr):
    def __init__(self, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.leaf_shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.leaf_shape))
        self.leaf_size = n
        super().__init__(n, {id(self): (self, np.eye(n))}, np.zeros(n))
        self._value = None

    @property
    def shape(self):
        return self.leaf_shape

    def __getitem__(self, key):
        """``var[:, i]`` (column of a 2-D leaf, negative i allowed) -- the only indexing the reference uses."""
        assert len(self.leaf_shape) == 2 and isinstance(key, tuple) and len(key) == 2, key
        rows, col = key
        assert isinstance(rows, slice) and rows == slice(None), key
        r, c = self.leaf_shape
        col = int(col)
        if col < 0:
            col += c
        assert 0 <= col < c
        S = np.zeros((r, r * c))
        S[np.arange(r), col * r + np.arange(r)] = 1.0          # column-major stacking
        return Expr(r, {id(self): (self, S)}, np.zeros(r))


class Variable(_Leaf):
    @property
    def value(self):
        return self._value

    def _set_from_vec(self, v):
        if v is None:
            self._value = None
        elif len(self.leaf_shape) == 2:
            self._value = np.array(v, dtype=float).reshape(self.leaf_shape, order="F")
        else:
            self._value = np.array(v, dtype=float).reshape(self.leaf_shape)


class Parameter(_Leaf):
    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        a = np.asarray(v, dtype=float)
        if a.shape != self.leaf_shape:
            raise ValueError(f"Invalid dimensions {a.shape} for Parameter value (expected {self.leaf_shape}).")
        self._value = a


class Constraint:
    def __init__(self, kind, expr):
        self.kind, self.expr = kind, expr       # expr == 0  /  expr <= 0


class QuadCost:
    """Sum of quad_form(expr_k, M_k) terms (cvxpy's quad_form has no factor 1/2)."""
    __array_ufunc__ = None

    def __init__(self, terms):
        self.terms = terms

    def __add__(self, other):
        if isinstance(other, QuadCost):
            return QuadCost(self.terms + other.terms)
        if np.isscalar(other) and other == 0:
            return self
        raise TypeError("only sums of quad_form terms are supported")

    __radd__ = __add__


def quad_form(expr, M):
    M = np.atleast_2d(np.asarray(M, dtype=float))
    if isinstance(expr, _Leaf) and len(expr.leaf_shape) != 1:
        raise TypeError("quad_form needs a vector")
    assert M.shape == (expr.size, expr.size), (M.shape, expr.size)
    return QuadCost([(expr, M)])


class Minimize:
    def __init__(self, cost):
        self.cost = cost


class Problem:
    def __init__(self, objective, constraints):
        self.objective = objective
        self.constraints = list(constraints)
        self.status = None
        self.value = None
        self.solver_kwargs = None
        self.last_result = None
        self._build()

    # ---- standard form ----
    def _build(self):
        cost = self.objective.cost
        terms = cost.terms if isinstance(cost, QuadCost) else []
        exprs = [e for e, _ in terms] + [c.expr for c in self.constraints]
        variables, params = [], []
        for e in exprs:
            for leaf, _ in e.lin.values():
                lst = variables if isinstance(leaf, Variable) else params
                if all(leaf is not x for x in lst):
                    lst.append(leaf)
        self.variables, self.parameters = variables, params
        voff = np.cumsum([0] + [v.leaf_size for v in variables])
        poff = np.cumsum([0] + [p.leaf_size for p in params])
        self._voff, self._poff = voff, poff
        nz, npar = int(voff[-1]), int(poff[-1])

        def split(e):
            C = np.zeros((e.size, nz))
            D = np.zeros((e.size, npar))
            for leaf, M in e.lin.values():
                if isinstance(leaf, Variable):
                    i = next(k for k, v in enumerate(variables) if v is leaf)
                    C[:, voff[i]:voff[i + 1]] += M
                else:
                    i = next(k for k, p in enumerate(params) if p is leaf)
                    D[:, poff[i]:poff[i + 1]] += M
            return C, D, e.const

        P = np.zeros((nz, nz))
        Qp = np.zeros((nz, npar))
        q0 = np.zeros(nz)
        for e, M in terms:
            C, D, c = split(e)
            Ms = 0.5 * (M + M.T)
            P += 2.0 * C.T @ Ms @ C
            Qp += 2.0 * C.T @ Ms @ D
            q0 += 2.0 * C.T @ Ms @ c
        rows = {"eq": ([], [], []), "le": ([], [], [])}
        for con in self.constraints:
            C, D, c = split(con.expr)
            for lst, item in zip(rows[con.kind], (C, D, c)):
                lst.append(item)

        def stack(kind):
            Cs, Ds, cs = rows[kind]
            if not Cs:
                return np.zeros((0, nz)), np.zeros((0, npar)), np.zeros(0)
            return np.vstack(Cs), np.vstack(Ds), np.hstack(cs)

        E, Ep, e0 = stack("eq")
        G, Gp, g0 = stack("le")
        used = (np.abs(P).sum(0) > 0) | (np.abs(E).sum(0) > 0) | (np.abs(G).sum(0) > 0)
        self.std = dict(P=P, Qp=Qp, q0=q0, E=E, Ep=Ep, e0=e0, G=G, Gp=Gp, g0=g0, used=used)

    def standard_form(self, pvals=None):
        """(P, q, E, e, G, h, used) at the current (or given) parameter values; z stacks ``self.variables`` in order,
        each column-major."""
        s = self.std
        if pvals is None:
            for p in self.parameters:
                if p.value is None:
                    raise ValueError("A Parameter has no value")
            pvals = (np.concatenate([np.asarray(p.value, float).reshape(-1, order="F") for p in self.parameters])
                     if self.parameters else np.zeros(0))
        q = s["Qp"] @ pvals + s["q0"]
        e = -(s["Ep"] @ pvals + s["e0"])
        h = -(s["Gp"] @ pvals + s["g0"])
        return s["P"], q, s["E"], e, s["G"], h, s["used"]

    def solve(self, solver=None, verbose=False, **kwargs):
        from oracle import ref_qp as rq
        self.solver_kwargs = dict(solver=solver, **kwargs)
        P, q, E, e, G, h, used = self.standard_form()
        u = np.nonzero(used)[0]
        r = rq.solve_qp(P[np.ix_(u, u)], q[u], E[:, u], e, G[:, u], h)
        if r.status != "optimal" and not rq.is_feasible(E[:, u], e, G[:, u], h):
            r.status = "infeasible"
        self.last_result = r
        if r.status == "optimal":
            z = np.zeros(P.shape[0])
            z[u] = r.z
            for i, v in enumerate(self.variables):
                v._set_from_vec(z[self._voff[i]:self._voff[i + 1]])
            self.status = OPTIMAL
            self.value = float(0.5 * z @ P @ z + q @ z)
        else:
            for v in self.variables:
                v._set_from_vec(None)
            self.status = INFEASIBLE if r.status == "infeasible" else "solver_error"
            self.value = np.inf
        return self.value

"""The QP oracle against an independent solver (scipy SLSQP) and against KKT conditions."""
import numpy as np
import pytest
from scipy.optimize import minimize, nnls

import helpers as H
from oracle import ref_qp as rq
from oracle.ref_polytope import Polytope


def _poly(s, k):
    return Polytope(s[k + "_A"], s[k + "_b"], normalize=False)


def _di_qp(fixed=True):
    s = H.load("sets_di.npz")
    return rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], _poly(s, "Xc"), _poly(s, "Uc"),
                                  _poly(s, "Xf"), _poly(s, "Z"), fixed)


def kkt_residuals(qp, x_init, ref, z):
    """Independent optimality check: feasibility and a non-negative multiplier (NNLS) closing
    stationarity on the rows that are active to 1e-7."""
    q, e, h = qp.params(x_init, ref)
    z = np.r_[z, np.zeros(qp.P.shape[0] - z.size)]
    prim = max(np.abs(qp.E @ z - e).max(), (qp.G @ z - h).max())
    act = np.nonzero(qp.G @ z - h > -1e-7 * (1 + np.abs(h)))[0]
    M = np.c_[qp.E.T, -qp.E.T, qp.G[act].T]
    cs = np.maximum(np.linalg.norm(M, axis=0), 1e-300)
    _, rn = nnls(M / cs, -(qp.P @ z + q), maxiter=50 * M.shape[1])
    return prim, rn / (1.0 + np.abs(q).max())


@pytest.mark.parametrize("x,r", [((1.0, 2.0), (5.0, 0.0)), ((1.0, 2.0), (-9.0, 0.0)), ((-3.0, 1.0), (9.0, 0.0)),
                                 ((0.0, 0.0), (0.0, 0.0)), ((6.0, -1.0), (4.0, 0.0))])
def test_oracle_vs_slsqp_double_integrator(x, r):
    qp = _di_qp()
    x, r = np.array(x), np.array(r)
    (xs, us, xb, ub), res = rq.solve_param(qp, x.copy(), r.copy())
    assert res.status == "optimal"
    q, e, h = qp.params(x, r)
    cons = [{"type": "eq", "fun": lambda z: qp.E @ z - e, "jac": lambda z: qp.E},
            {"type": "ineq", "fun": lambda z: h - qp.G @ z, "jac": lambda z: -qp.G}]
    sol = minimize(lambda z: 0.5 * z @ qp.P @ z + q @ z, res.z + 0.05, jac=lambda z: qp.P @ z + q,
                   constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-12))
    f = lambda z: 0.5 * z @ qp.P @ z + q @ z       # noqa: E731
    feas = max(np.abs(qp.E @ sol.x - e).max(), (qp.G @ sol.x - h).max())
    assert feas < 1e-6
    # the oracle's point is at least as good as SLSQP's (status 8 = SLSQP's line search gave up
    # at its own precision floor), and the two agree to SLSQP's accuracy
    assert f(res.z) <= f(sol.x) + 1e-7 * (1 + abs(f(sol.x)))
    assert np.abs(sol.x - res.z).max() < 5e-4
    prim, stat = kkt_residuals(qp, x, r, res.z)
    assert prim < 1e-9 and stat < 1e-9


def test_golden_solutions_satisfy_kkt():
    qp = _di_qp()
    g = H.load("loop_di_tube.npz")
    for t in range(0, 120, 7):
        prim, stat = kkt_residuals(qp, g["xhat_in"][t], g["refs"][t], g["z"][t])
        assert prim < 1e-9 and stat < 1e-8, (t, prim, stat)
    assert g["polished"].all()
    s = H.load("sets_cp.npz")
    qpc = rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], _poly(s, "Xc"), _poly(s, "Uc"),
                                 _poly(s, "Xf"), None, True)
    gc = H.load("loop_cp_tube.npz")
    for run in range(4):
        for t in (0, 3, 11, 40, 120, 249):
            prim, stat = kkt_residuals(qpc, gc["tube_xhat_in"][run, t], gc["refs"][t], gc["tube_z"][run, t])
            assert prim < 1e-9 and stat < 1e-8, (run, t, prim, stat)


def test_infeasible_is_reported():
    qp = _di_qp()
    sol, res = rq.solve_param(qp, np.array([30.0, 0.0]), np.array([0.0, 0.0]))
    assert res.status == "infeasible" and sol[0] is None


def test_tube_initial_state_variant():
    qp = _di_qp(fixed=False)
    x, r = np.array([1.0, 2.0]), np.array([5.0, 0.0])
    (xs, us, xb, ub), res = rq.solve_param(qp, x.copy(), r.copy())
    assert res.status == "optimal"
    s = H.load("sets_di.npz")
    assert np.all(s["Z_A"] @ (x - xs[:, 0]) <= s["Z_b"] + 1e-9)
    prim, stat = kkt_residuals(qp, x, r, res.z)
    assert prim < 1e-9 and stat < 1e-8

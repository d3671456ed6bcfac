"""The product's condensed QP has the same minimiser as the reference's un-condensed QP."""
import numpy as np
import pytest

import helpers as H
from oracle import ref_qp as rq
from oracle.ref_polytope import Polytope
from rtmpc_b200.condense import MPCSpec, condense, steady_state_basis
from rtmpc_b200.ipm_data import prepare


def _poly(s, k):
    return Polytope(s[k + "_A"], s[k + "_b"], normalize=False)


def solve_condensed(cq, x, r):
    q = cq.Fx @ x + cq.Fr @ r
    lo, up = cq.lo0 + cq.Lx @ x, cq.up0 + cq.Ux @ x
    fin = lo > -1e29
    res = rq.solve_qp(cq.H, q, np.zeros((0, cq.n)), np.zeros(0), np.vstack([cq.G, -cq.G[fin]]), np.r_[up, -lo[fin]])
    return cq.Phi @ res.z + cq.Psi @ x, res


CASES = [((1.0, 2.0), (5.0, 0.0)), ((1.0, 2.0), (-9.0, 0.0)), ((-3.0, 1.0), (9.0, 0.0)), ((0.5, -0.2), (0.0, 0.0))]


@pytest.mark.parametrize("fixed", [True, False])
def test_tube_tracking_double_integrator(fixed):
    s = H.load("sets_di.npz")
    qp = rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], _poly(s, "Xc"), _poly(s, "Uc"),
                                _poly(s, "Xf"), _poly(s, "Z"), fixed)
    cq = condense(H.spec_tube_tracking(s, fixed))
    assert cq.n == (11 if fixed else 13)
    for x, r in CASES:
        x, r = np.array(x), np.array(r)
        _, ro = rq.solve_param(qp, x.copy(), r.copy())
        z, rc = solve_condensed(cq, x, r)
        assert ro.status == rc.status == "optimal"
        assert np.abs(z - ro.z[:z.size]).max() < 1e-8


def test_extended_g2_matches_free_variable_formulation():
    s = H.load("sets_di.npz")
    qp = rq.build_extended_packet_received(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], _poly(s, "Xc"),
                                           _poly(s, "Uc"), _poly(s, "Xf"), _poly(s, "ZmW"))
    cq = condense(H.spec_ext_received(s))
    for x, r in CASES[:3]:
        x, r = np.array(x), np.array(r)
        _, ro = rq.solve_param(qp, x.copy(), r.copy())
        z, rc = solve_condensed(cq, x, r)
        assert ro.status == rc.status == "optimal"
        assert np.abs(z - ro.z[:z.size]).max() < 1e-7


def test_tracking_and_regulators():
    s = H.load("sets_di.npz")
    qp = rq.build_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], _poly(s, "X"), _poly(s, "U"),
                           _poly(s, "Xf_track"))
    cq = condense(H.spec_tracking(s))
    x, r = np.array([1.0, 2.0]), np.array([5.0, 0.0])
    _, ro = rq.solve_param(qp, x.copy(), r.copy())
    z, rc = solve_condensed(cq, x, r)
    assert np.abs(z - ro.z).max() < 1e-8
    g = H.load("qp_di_regulators.npz")
    reg = condense(MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), stage_x=(s["X_A"], s["X_b"]),
                           stage_u=(s["U_A"], s["U_b"])))
    mayne = condense(MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=g["P"],
                             stage_x=(g["Xc_mayne_A"], g["Xc_mayne_b"]), stage_u=(g["Uc_mayne_A"], g["Uc_mayne_b"]),
                             terminal=(g["Xf_mayne_A"], g["Xf_mayne_b"]), tube_init=(g["Z_mayne_A"], g["Z_mayne_b"])))
    for i in range(0, 40, 5):
        x = g["xs"][i]
        z, rc = solve_condensed(reg, x, np.zeros(2))
        assert np.abs(z - g["z_reg"][i]).max() < 1e-8
        z, rc = solve_condensed(mayne, x, np.zeros(2))
        assert np.abs(z - g["z_tube"][i]).max() < 1e-7


def test_cartpole_condensed_shape_and_minimiser():
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    cq = condense(H.spec_tube_tracking(s))
    assert cq.n == 21 and cq.m == cq.meta["rows_one_sided"] // 2        # every row found its mirror
    assert np.allclose(cq.Mss[:, 0], [1, 0, 0, 0, 0])                    # only the cart position is free
    for run, t in [(0, 0), (1, 7), (3, 30), (2, 200)]:
        z, rc = solve_condensed(cq, g["tube_xhat_in"][run, t], g["refs"][t])
        assert np.abs(z - g["tube_z"][run, t]).max() < 1e-7
    d = prepare(cq)
    assert d.mpad % 32 == 0 and d.npad % 4 == 0 and d.Gs.shape == (d.mpad, d.npad)
    assert np.linalg.cond(d.Hs) < np.linalg.cond(cq.H)


def test_steady_state_basis():
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.0], [1.0]])
    M = steady_state_basis(A, B)
    assert M.shape == (3, 1) and np.abs(np.c_[A - np.eye(2), B] @ M).max() < 1e-12


def test_g2_terminal_projection_for_a_two_dimensional_steady_state_family():
    """G2 mode (TubeTrackingMPC.py:293) when the steady-state family has dimension 2 (a two-input plant): the projection of
    the terminal set on theta, against LP support values of the un-projected set in random theta directions."""
    from scipy.optimize import linprog
    from rtmpc_b200.condense import _project_terminal_on_theta, steady_state_basis
    rng = np.random.default_rng(5)
    A = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.0, 0.0, 0.9]])
    B = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.2]])
    nx, nu = B.shape
    Mss = steady_state_basis(A, B)
    nth = Mss.shape[1]
    assert nth == 2
    m = 60
    HN = rng.normal(size=(m, 2 * nx + nu))
    HN /= np.linalg.norm(HN, axis=1)[:, None]
    hN = rng.uniform(0.5, 1.5, m)
    C, h = _project_terminal_on_theta(HN, hN, Mss, nx, nu)
    assert C.shape[1] == nth and C.shape[0] >= 3
    Aub = np.c_[HN[:, :nx], HN[:, nx:2 * nx] @ Mss[:nx], HN[:, 2 * nx:]]
    for _ in range(25):
        a = rng.normal(size=nth)
        c = np.zeros(Aub.shape[1])
        c[nx:nx + nth] = -a
        full = -linprog(c, A_ub=Aub, b_ub=hN, bounds=(None, None)).fun
        proj = -linprog(-a, A_ub=C, b_ub=h, bounds=(None, None)).fun
        assert abs(full - proj) <= 1e-8 * (1 + abs(full))

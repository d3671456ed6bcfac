"""Batched LP kernel (rtmpc_lp_solve: one warp per LP, dual simplex over shared rows) and the offline set pipeline on top
of it: values against scipy / HiGHS - what the reference's `support` calls (utils_polytope.py:19) -,
`reduce` and the maximal-output-admissible-set iteration against the host backend and the golden fixtures."""
import time

import numpy as np
import pytest
from scipy.optimize import linprog

import helpers as H

pytestmark = pytest.mark.gpu


def _random_polytope(rng, n, m):
    A = rng.normal(size=(m, n))
    A /= np.linalg.norm(A, axis=1)[:, None]
    b = rng.uniform(0.5, 2.0, m)
    return A, b


@pytest.mark.parametrize("n,m", [(2, 8), (3, 40), (5, 120), (9, 400), (12, 900), (16, 300)])
def test_lp_values_and_vertices_against_highs(n, m):
    from rtmpc_b200 import polytope as pc
    rng = np.random.default_rng(100 * n + m)
    A, b = _random_polytope(rng, n, m)
    dirs = rng.normal(size=(64, n))
    val, X = pc.lp_batch(A, b, dirs, want_x=True)
    for i in range(0, 64, 4):
        res = linprog(-dirs[i], A_ub=A, b_ub=b, bounds=(None, None))
        assert res.status == 0
        assert abs(val[i] + res.fun) <= 1e-9 * (1 + abs(res.fun)), (i, val[i], -res.fun)
    assert np.all(X @ A.T <= b + 1e-9)                                       # every reported point is feasible ...
    assert np.abs(np.einsum("ij,ij->i", dirs, X) - val).max() <= 1e-9       # ... and attains the reported value


def test_lp_relaxed_row_extra_rows_unbounded_and_infeasible():
    from rtmpc_b200 import polytope as pc
    A = np.r_[np.eye(2), -np.eye(2)]
    b = np.ones(4)
    # own bound of row 0 relaxed by 0.1 (polytope.reduce's LP): max x_0 = 1.1; other instances unaffected
    val, _ = pc.lp_batch(A, b, np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), relax_row=np.array([0, -1, 0], np.int32), relax_by=0.1)
    assert np.allclose(val, [1.1, 1.0, 1.0], atol=1e-12)
    # one own row per instance: x_0 + x_1 <= c_b
    extra = np.array([[[1.0, 1.0, 0.5]], [[1.0, 1.0, -3.0]], [[1.0, 1.0, 5.0]]])
    val, _ = pc.lp_batch(A, b, np.array([[1.0, 1.0]] * 3), extra=extra)
    assert abs(val[0] - 0.5) <= 1e-12 and val[1] == -np.inf and abs(val[2] - 2.0) <= 1e-12        # cut / infeasible / inactive
    # half space: unbounded in direction (0, 1), bounded in (1, 0)
    val, _ = pc.lp_batch(np.array([[1.0, 0.0]]), np.array([2.0]), np.array([[0.0, 1.0], [1.0, 0.0]]))
    assert val[0] == np.inf and abs(val[1] - 2.0) <= 1e-12
    # degenerate vertex: many rows through one point
    ang = np.linspace(0, np.pi / 2, 40)
    Ad = np.c_[np.cos(ang), np.sin(ang)]
    bd = Ad @ np.array([1.0, 1.0])
    val, X = pc.lp_batch(np.vstack([Ad, -np.eye(2)]), np.r_[bd, 5.0, 5.0], np.array([[1.0, 1.0], [0.3, 0.9]]), want_x=True)
    assert np.allclose(X, 1.0, atol=1e-9) and np.allclose(val, [2.0, 1.2], atol=1e-9)


def test_reduce_and_subset_test_equal_host_backend_on_the_cartpole_terminal_set():
    """polytope.reduce / is_subset with the GPU LPs against the same functions on HiGHS, on a 9-D set with many redundant
    rows: the cartpole terminal set's rows plus 300 rows that lie outside it."""
    from rtmpc_b200 import polytope as pc
    s = H.load("sets_cp.npz")
    Xf = pc.Polytope(s["Xf_A"], s["Xf_b"], normalize=False)
    rng = np.random.default_rng(3)
    extra = rng.normal(size=(300, 9))
    extra /= np.linalg.norm(extra, axis=1)[:, None]
    hb = pc.support_lp(Xf, extra) + rng.uniform(1e-3, 0.3, 300)            # redundant: strictly outside
    big = pc.Polytope(np.vstack([Xf.A, extra]), np.r_[Xf.b, hb], normalize=False)
    t0 = time.perf_counter()
    red = pc.reduce(big)
    t_gpu = time.perf_counter() - t0
    pc.set_lp_backend("highs")
    try:
        t0 = time.perf_counter()
        red_h = pc.reduce(pc.Polytope(big.A, big.b, normalize=False))
        t_host = time.perf_counter() - t0
        assert red.A.shape == red_h.A.shape and np.array_equal(red.A, red_h.A) and np.array_equal(red.b, red_h.b)
        assert red.A.shape[0] == Xf.A.shape[0]                              # exactly the redundant rows went
        assert pc.is_subset(red, Xf) and pc.is_subset(Xf, red)
    finally:
        pc.set_lp_backend("gpu")
    assert pc.is_subset(red, Xf) and pc.is_subset(Xf, red)
    shrunk = pc.Polytope(Xf.A, Xf.b - 1e-3, normalize=False)
    assert pc.is_subset(shrunk, Xf) and not pc.is_subset(Xf, shrunk)
    print(f"reduce of {big.A.shape[0]} rows in 9-D: GPU LPs {t_gpu:.3f} s, HiGHS {t_host:.3f} s")
    assert t_gpu < t_host


def test_cartpole_setup_optimization_end_to_end_equals_fixture():
    """The product's full cartpole set-up - TubeTrackingMPC.setup_optimization(W, fixed_initial_state=True, rpi_method=1):
    Darup RPI (support sweeps), reduce, tightening, the 9-D terminal-set iteration (batched LPs), QP generation - against
    the fixture sets (equal as sets: subset both ways; same row counts), within the time budget of VERDICT item 6
    (terminal set <= 3 s; the reference's loop takes 128 s on HiGHS), and a solve on the controller it builds."""
    from rtmpc_b200 import mpc
    from rtmpc_b200 import polytope as pc
    s = H.load("sets_cp.npz")
    c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(H.poly(s, "U"))
    c.set_state_constraints(H.poly(s, "X"))
    t0 = time.perf_counter()
    c.determine_mRPI(H.poly(s, "W"), rpi_method=1, skip_wasted_pass=True)
    c.tighten_constraints()
    t1 = time.perf_counter()
    c.determine_Xf()
    t2 = time.perf_counter()
    c.generate_optimization_problem(True)
    print(f"cartpole set-up: mRPI + tightening {t1 - t0:.2f} s, terminal set {t2 - t1:.2f} s")
    assert c._Z.A.shape == s["Z_A"].shape and c._Xf.A.shape == s["Xf_A"].shape
    assert np.allclose(c._Xc.b, s["Xc_b"], atol=1e-9) and np.allclose(c._Uc.b, s["Uc_b"], atol=1e-9)
    for k in ("Z", "Xf"):
        mine, ref = getattr(c, "_" + k), H.poly(s, k)
        assert pc.is_subset(mine, ref) and pc.is_subset(ref, mine), k
    assert t2 - t1 <= 3.0
    g = H.load("loop_cp_tube.npz")
    out = c.solve_batch(g["tube_xhat_in"][1, :40], g["refs"][:40])
    assert np.all(out["status"] == 0)
    assert np.abs(out["U_t"] - g["tube_U_t"][1, :40]).max() <= 1e-6
    # the reference's own cap sequence (200, then 2000) gives the same tube
    c2 = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c2.set_input_constraints(H.poly(s, "U"))
    c2.set_state_constraints(H.poly(s, "X"))
    c2.determine_mRPI(H.poly(s, "W"), rpi_method=1)
    assert np.array_equal(c2._Z.A, c._Z.A) and np.array_equal(c2._Z.b, c._Z.b)

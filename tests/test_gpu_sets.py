"""Support-function sweep kernel and the set pipeline that runs on top of it."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_support_sweep_matches_lp_oracle():
    from oracle import ref_sets as rs
    from oracle.ref_polytope import Polytope
    from rtmpc_b200 import sets as up
    s = H.load("sets_cp.npz")
    Z = H.poly(s, "Z")
    rng = np.random.default_rng(1)
    dirs = rng.normal(size=(200, 4))
    h = up.support_batch(Z, dirs)
    Zo = Polytope(s["Z_A"], s["Z_b"], normalize=False)
    for i in range(0, 200, 10):
        assert abs(h[i] - rs.support(Zo, dirs[i])) <= 1e-9 * (1 + abs(h[i]))
    W = H.poly(s, "W")
    assert np.allclose(up.support_batch(W, np.eye(4)), [1e-4, 2.7e-3, 3e-4, 4.3e-2], rtol=0, atol=1e-15)


@pytest.mark.parametrize("eps,k_expected", [(1e-1, 5), (1e-2, 6), (1e-3, 10)])
def test_darup_known_answers_on_gpu(eps, k_expected, capsys):
    from rtmpc_b200 import numerics, sets as up
    from rtmpc_b200.polytope import Polytope, box
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.5], [1.0]])
    K, _, _ = numerics.dlqr(A, B, np.eye(2), np.eye(1))
    X = Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[4.0, 2.0, 8.0, 4.0])
    rpi, status = up.calculate_RPI(A - B @ K, box([0.1, 0.1]), X, box([1.0]), K, eps, 50)
    assert status == 0 and rpi.A.shape[0] == 6 * k_expected
    assert f"k_star = {k_expected}" in capsys.readouterr().out


def test_cartpole_tightening_pipeline_on_gpu():
    """determine_mRPI (Darup, with the reference's s_max retry) -> tighten_constraints, supports on the GPU."""
    from rtmpc_b200 import mpc
    s = H.load("sets_cp.npz")
    c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(H.poly(s, "U"))
    c.set_state_constraints(H.poly(s, "X"))
    c.determine_mRPI(H.poly(s, "W"), rpi_method=1)
    assert c._Z.A.shape == (854, 4)
    c.tighten_constraints()
    assert np.allclose(c._Xc.b, s["Xc_b"], atol=1e-9) and np.allclose(c._Uc.b, s["Uc_b"], atol=1e-9)


def test_double_integrator_full_setup_on_gpu_matches_fixture():
    from rtmpc_b200 import mpc
    s = H.load("sets_di.npz")
    c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(H.poly(s, "U"))
    c.set_state_constraints(H.poly(s, "X"))
    c.setup_optimization(H.poly(s, "W"), fixed_initial_state=True)
    assert c._Z.A.shape == (48, 2) and c._Xf.A.shape == (26, 5)
    assert np.allclose(c._Xc.b, s["Xc_b"], atol=1e-9)
    x_nom, u_nom, xb, ub = c.solve_optimization_problem(np.array([1.0, 2.0]), np.array([5.0, 0.0]))
    g = H.load("loop_di_tube.npz")
    assert np.abs(u_nom[0] - g["U_t"][0, :10, 0]).max() <= 1e-7


def test_sweep_large_direction_count():
    from rtmpc_b200 import sets as up
    s = H.load("sets_cp.npz")
    from rtmpc_b200 import polytope as pc
    V = pc.extreme(H.poly(s, "W"))
    rng = np.random.default_rng(0)
    dirs = rng.normal(size=(200000, 4))
    h = up.support_sweep(V, dirs)
    assert np.allclose(h, np.abs(dirs) @ np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2]), rtol=1e-13, atol=0)


@pytest.mark.parametrize("dim", [2, 4, 7, 13])
def test_sweep_with_negative_maxima(dim):
    """The sweep keeps its running maxima as signed integers on the bit patterns (right whenever the maximum is >= 0) and
    redoes a trip with FP64 compares when a maximum comes out negative: vertex sets that do not contain the origin, so that
    many directions have every <d, v> < 0; zeros and signed zeros among the products; every kernel instantiation (dim)."""
    from rtmpc_b200 import sets as up
    rng = np.random.default_rng(dim)
    V = rng.uniform(2.0, 3.0, (37, dim))                     # a cloud far from the origin
    V[5] = 0.0                                               # ... with the origin as one vertex in the second case
    dirs = rng.standard_normal((5000, dim))
    dirs[:50] = 0.0                                          # products are +0 / -0
    dirs[50:100] = -np.abs(dirs[50:100])                     # every product negative (except with V[5])
    for Vs in (np.delete(V, 5, axis=0), V):
        h = up.support_sweep(Vs, dirs)
        ref = (dirs @ Vs.T).max(axis=1)
        assert (ref < 0).sum() > 40 or Vs.shape[0] == 37
        assert np.abs(h - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        assert np.array_equal(np.signbit(h[np.abs(ref) > 1e-9]), np.signbit(ref[np.abs(ref) > 1e-9]))

"""Literal actuator / estimator restatement: the reference scripts' run-time invariants, and the
O(1) equivalents the fused CUDA kernel relies on."""
import numpy as np
from hypothesis import given, settings, strategies as st

import helpers as H
from oracle import ref_loop as rl
from oracle.ref_polytope import Polytope


def test_example_invariants_double_integrator():
    """Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py:165-184 on the golden run."""
    s, g = H.load("sets_di.npz"), H.load("loop_di_tube.npz")
    Z = Polytope(s["Z_A"], s["Z_b"], normalize=False)
    T = len(g["theta"])
    for t in range(T):
        assert (g["x"][t] - g["x_nom"][t]) in Z
        if g["Theta"][t] == 1:
            assert (g["x"][t] - g["x_hat"][t]) in Z
    assert np.all(np.abs(g["u"]) <= 1.0 + 1e-9)          # u in U (Example_of_Tube_Tracking_MPC.py:98-100)


def test_pezzutto_estimate_exact_when_consistent():
    """Example_of_Tracking_MPC_Over_Lossy_Network.py:141-156."""
    g = H.load("loop_di_track.npz")
    ok = g["Theta"] == 1
    assert ok.sum() > 5
    assert np.abs(g["x"][:-1][ok] - g["x_hat"][:-1][ok]).max() == 0.0


def test_cartpole_tube_invariant_all_loss_rates():
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    d = g["tube_x"] - g["tube_x_nom"]
    viol = (d.reshape(-1, 4) @ s["Z_A"].T - s["Z_b"]).max()
    assert viol < 1e-7


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.booleans(), st.booleans()), min_size=2, max_size=60))
def test_constant_time_equivalents(pattern):
    """(A1/A2) Theta_t from the last lost step instead of the growing theta vector, and (E1)
    controlSequences[s_t] == the actuator's buffer, on random loss patterns."""
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.0], [1.0]])
    K = np.array([[0.4, 1.2]])
    N = 4
    rng = np.random.default_rng(len(pattern))
    act = rl.ConsistentActuator(A, B, K, K, np.zeros(2))
    est = rl.Estimator(A, B, K, np.zeros(2), N)
    last_loss = -1
    x = np.zeros(2)
    for t, (th, ga) in enumerate(pattern):
        th, ga = (1, 1) if t == 0 else (int(th), int(ga))
        U = rng.normal(size=(1, N + 1))
        q_t = est.get_qt()
        est.store_sent_control_sequence(U)
        u, pkt = act.process_packet({"U_t": U, "q_t": q_t}, x, th)
        if th == 0:
            last_loss = t
        assert act.Theta_t == (1 if (th == 1 and last_loss <= q_t) else 0)
        assert np.array_equal(est.sequences[act.s_t], act.u_traj)
        x = A @ x + B @ u
        est.update_estimate(pkt, ga)
        # non-extended: whenever the plant packet arrives the estimate equals the nominal state
        if ga == 1:
            assert np.allclose(est.get_estimate(), act.get_x_nom(), atol=1e-12)


def test_cartpole_ode_linearises_to_reference_matrices():
    from oracle import ref_setup as su
    A, B = su.cartpole_matrices()
    x0 = np.array([0.0, 0.0, 0.0, 0.0])
    eps = 1e-6
    # 10 sub-steps of 1/500 s = one control period; finite differences around the upright rest
    def step(x, F):
        for _ in range(10):
            x = rl.cartpole_ode_step(x, F)
        return x
    J = np.stack([(step(x0 + eps * e, 0.0) - step(x0 - eps * e, 0.0)) / (2 * eps) for e in np.eye(4)], axis=1)
    Bd = (step(x0, eps) - step(x0, -eps)) / (2 * eps)
    assert np.abs(J - A).max() < 5e-3 and np.abs(Bd - B[:, 0]).max() < 5e-4   # Euler vs ZOH


def test_model_error_restatement_properties():
    """estimate_W_for_Cartpole.py:79-127 restated on the analytic plant: zero model error at the origin, second order
    in the state (the linearisation is exact to first order), LQR stabilises the nonlinear plant from the script's box."""
    from oracle import ref_loop as rl
    s = H.load("sets_cp.npz")
    A, B, K = s["A"], s["B"], s["K"]
    Acl = A - B @ K
    w0, _ = rl.estimate_model_error(np.zeros((1, 4)), K, Acl, 3)
    assert np.abs(w0).max() == 0.0
    x0 = np.array([[0.2, 0.1, 0.05, 0.1]])
    w1, _ = rl.estimate_model_error(x0, K, Acl, 1)
    w2, _ = rl.estimate_model_error(0.5 * x0, K, Acl, 1)
    # the discretisations differ (ZOH-exact A, B against semi-implicit Euler sub-steps): that part is linear in x;
    # what is left after removing it halves-squared
    lin = 2.0 * w2 - w1            # = linear part of w1 (up to third order)
    quad = w1 - lin
    assert np.abs(quad).max() <= 4.0 * np.abs(w1).max()
    assert np.abs(w1).max() < 0.1
    _, xf = rl.estimate_model_error(np.array([[1.0, 0.5, 0.3, 0.5], [-1.0, -0.5, -0.3, -0.5]]), K, Acl, 400)
    assert np.abs(xf).max() < 1e-3

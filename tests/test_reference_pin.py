"""The oracle pinned to the REFERENCE'S OWN CODE.

``tests/golden/ref_*.npz`` were written by ``tests/golden/make_reference_fixtures.py`` from the reference's modules and
scripts, executed unmodified in the build container (third-party ``polytope`` / ``control`` / ``cvxpy`` replaced by the
stand-ins of ``oracle/refshim``; the numerical QP solver under the ``cvxpy`` stand-in is the oracle's, Clarabel being
absent).  This file checks, on the CPU:

  * the oracle's restatements (state machines, set computations, QP statements, closed loops) against those fixtures;
  * when ``/root/reference`` is present (build container; never on the GPU box): the same comparisons LIVE against the
    imported reference modules, and that the committed fixtures regenerate bit for bit.

Tolerances: integers and row counts exact; floats 1e-12 relative unless a solver sits in between (1e-9: two runs of the
same interior-point code on problems that differ in round-off).
"""
import numpy as np
import pytest

import helpers as H
from oracle import ref_loop as rl
from oracle import ref_numerics as rn
from oracle import ref_qp as rq
from oracle import ref_sets as rs
from oracle import ref_setup as su
from oracle import refshim
from oracle.ref_polytope import Polytope

live = pytest.mark.skipif(not refshim.reference_available(), reason="reference sources not present on this machine")


def _P(f, key, normalize=False):
    return Polytope(f[key + "_A"], f[key + "_b"], normalize=normalize)


def _same_poly(f, key, poly, tol=1e-12):
    A, b = f[key + "_A"], f[key + "_b"]
    assert A.shape == poly.A.shape, (key, A.shape, poly.A.shape)
    assert np.abs(A - poly.A).max() <= tol and np.abs(b - poly.b.flatten()).max() <= tol * (1 + np.abs(b).max()), key


# ------------------------------------------------------------------------------------------------------------------
# state machines
# ------------------------------------------------------------------------------------------------------------------
def _drive(classes, f, name, p_i, kind):
    """The loop of make_reference_fixtures.statemachines() against any implementation of the four classes."""
    SA, CA, ES, RE = classes
    A, B, K, Kp, N = (f[name + k] for k in ("_A", "_B", "_K", "_Kp", "_N"))
    N = int(N)
    key = f"{name}_p{p_i}_"
    theta, gamma, U, xn0, w, x0 = (f[key + k] for k in ("theta", "gamma", "U", "xn0", "w", "x0"))
    nx = A.shape[0]
    T = len(theta)
    col = lambda v: np.array(v, float).reshape(nx, 1)      # noqa: E731
    if kind == "smart":
        act, est = SA(K), ES(A, B, K, col(x0), N)
    elif kind == "consistent":
        act, est = CA(A, B, K, Kp, col(x0)), ES(A, B, K, col(x0), N)
    else:
        act, est = CA(A, B, K, Kp, col(x0), is_extended_MPC_used=True), RE(A, B, K, Kp, col(x0), N)
    return act, est, (theta, gamma, U, xn0, w, x0, A, B, T, nx)


@pytest.mark.parametrize("name", ["di", "cp"])
@pytest.mark.parametrize("kind", ["smart", "consistent", "extended"])
def test_oracle_state_machines_equal_reference(name, kind):
    f = H.load("ref_statemachines.npz")
    for p_i in range(3):
        A, B, K, Kp, N = (f[name + k] for k in ("_A", "_B", "_K", "_Kp", "_N"))
        key = f"{name}_p{p_i}_"
        theta, gamma, U, xn0, w, x0 = (f[key + k] for k in ("theta", "gamma", "U", "xn0", "w", "x0"))
        T, N = len(theta), int(N)
        if kind == "smart":
            act, est = rl.SmartActuator(K), rl.Estimator(A, B, K, x0, N)
        elif kind == "consistent":
            act, est = rl.ConsistentActuator(A, B, K, Kp, x0), rl.Estimator(A, B, K, x0, N)
        else:
            act, est = rl.ConsistentActuator(A, B, K, Kp, x0, True), rl.RobustEstimator(A, B, K, Kp, x0, N)
        x = x0.copy()
        r = key + kind + "_"
        for t in range(T):
            q_t = est.get_qt()
            assert q_t == f[r + "q_t"][t]
            pkt = {"U_t": U[t], "q_t": q_t}
            if kind == "extended":
                pkt["x_nom_0"] = xn0[t]
                est.store_current_optimal_inital_nominal_plant_states(xn0[t])
            est.store_sent_control_sequence(U[t])
            u, ppkt = act.process_packet(pkt, x, int(theta[t]))
            assert (act.Theta_t, act.s_t, ppkt["s_t"]) == (f[r + "Theta"][t], f[r + "s_t"][t], f[r + "pkt_s"][t])
            assert np.abs(ppkt["x_t"] - f[r + "pkt_x"][t]).max() <= 1e-12 * (1 + np.abs(x).max())
            if kind == "extended":
                assert np.abs(ppkt["x_nom_t"] - f[r + "pkt_x_nom"][t]).max() <= 1e-12 * (1 + np.abs(x).max())
            x = A @ x + B @ u + w[t]
            est.update_estimate(ppkt, int(gamma[t]))
            assert np.abs(u - f[r + "u"][t]).max() <= 1e-12 * (1 + np.abs(u).max())
            assert np.abs(x - f[r + "x"][t + 1]).max() <= 1e-12 * (1 + np.abs(x).max())
            assert np.abs(est.get_estimate() - f[r + "x_hat"][t + 1]).max() <= 1e-12 * (1 + np.abs(x).max())
            if kind != "smart":
                assert np.abs(act.get_x_nom() - f[r + "x_nom"][t + 1]).max() <= 1e-12 * (1 + np.abs(x).max())


def test_constant_time_forms_on_reference_sequences():
    """The O(1) forms the CUDA kernels use evaluated on the reference's own sequences:
    Theta_t = theta_t and (last lost step <= q_t); s_t = t if Theta_t else s_{t-1}."""
    f = H.load("ref_statemachines.npz")
    for name in ("di", "cp"):
        for p_i in range(3):
            key = f"{name}_p{p_i}_"
            theta = f[key + "theta"]
            for kind in ("smart", "consistent", "extended"):
                q, Th, s = f[key + kind + "_q_t"], f[key + kind + "_Theta"], f[key + kind + "_s_t"]
                last_loss, s_prev = -1, 0
                for t in range(len(theta)):
                    if theta[t] == 0:
                        last_loss = t
                    want = 1 if (theta[t] == 1 and last_loss <= q[t]) else 0
                    assert Th[t] == want
                    s_prev = t if want else s_prev
                    assert s[t] == s_prev


# ------------------------------------------------------------------------------------------------------------------
# set computations (S1-S5) and gains
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_sets_equal_reference_set_examples():
    f = H.load("ref_examples.npz")
    # "Example of Approximation of mRPI_Darup.py": k_star printed by the reference's calculate_RPI, and its sets
    assert list(f["darup_k_star"]) == [5, 6, 10]
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.5], [1.0]])
    K, _ = rn.dlqr(A, B, np.eye(2), np.eye(1))
    assert np.abs(K - f["darup_K"]).max() <= 1e-13
    Acl = A - B @ K
    W, U = su.box([0.1, 0.1]), su.box([1.0])
    X = Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[4.0, 2.0, 8.0, 4.0])
    for eps, key, k in ((1e-1, "darup_P1", 5), (1e-2, "darup_P2", 6), (1e-3, "darup_P3", 10)):
        rpi, C, status, k_star = rs.darup_rpi(Acl, W, X, U, K, eps, 50)
        assert status == 0 and k_star == k
        _same_poly(f, key, rpi)
        if key == "darup_P1":
            _same_poly(f, "darup_C1", C)
    # "Example of Approximation of mRPI_Rakovic.py"
    Acl = A - np.array([[1.0], [1.0]]) @ np.array([[1.17, 1.03]])
    Fs, status, s, alpha = rs.rakovic_mrpi(Acl, su.box([1.0, 1.0]), eps_var=1.9e-5)
    assert status == 0 == int(f["rakovic_status"])
    _same_poly(f, "rakovic_Fs", Fs)
    # "Example of Output Admissible Set Calculation.py" (Mayne et al. Fig. 2)
    K, _ = rn.dlqr(A, B, np.eye(2), 0.01 * np.eye(1))
    assert np.abs(K - f["moas_K"]).max() <= 1e-13
    Acl = A - B @ K
    Z = rs.rakovic_mrpi(Acl, su.box([0.1, 0.1]))[0]
    _same_poly(f, "moas_Z", Z)
    Xm = Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[10.0, 2.0, 10.0, 10.0])
    Xc, Uc = rs.tighten(Xm, su.box([1.0]), Z, K)
    _same_poly(f, "moas_Xc", Xc)
    _same_poly(f, "moas_Uc", Uc)
    Xf, _ = rs.regulator_terminal_set(Acl, K, Xc, Uc)
    _same_poly(f, "moas_Xf", Xf)
    _same_poly(f, "moas_XfpZ", rs.mink_sum(Xf, Z))
    # "Example of Several Set Operations.py"
    from oracle.ref_polytope import qhull
    P1 = Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[3.0, 3.0, 3.0, 3.0])
    P2 = qhull(np.array([[1.0, 0], [0, -1], [-1, 0], [0, 1]]))
    _same_poly(f, "ops_P_diff", rs.pont_diff(P1, P2))
    c = np.cos(np.pi / 4)
    _same_poly(f, "ops_P_m2_scaled", rs.scale(P2, np.array([[c, c], [-c, c]])))
    _same_poly(f, "ops_P_mink", rs.mink_sum(su.box([2.0, 2.0]), su.box([1.0, 1.0])))
    _same_poly(f, "ops_P_mink_x", rs.mink_sum(su.box([2.0, 2.0]), np.ones((1, 2))))


def test_double_integrator_golden_sets_equal_reference_setup():
    """``sets_di.npz`` (what every double-integrator GPU test loads) == the sets the reference's own
    ``TubeTrackingMPC.setup_optimization`` / ``TrackingMPC.setup_optimization`` produced in the shipped examples."""
    s, f = H.load("sets_di.npz"), H.load("ref_examples.npz")
    for k in ("Z", "Xc", "Uc", "Xf"):
        assert np.array_equal(s[k + "_A"].shape, f["ttrkln_" + k + "_A"].shape)
        assert np.abs(s[k + "_A"] - f["ttrkln_" + k + "_A"]).max() <= 1e-12
        assert np.abs(s[k + "_b"] - f["ttrkln_" + k + "_b"]).max() <= 1e-12
    assert np.abs(s["ZmW_A"] - f["ttrkln_ZmW_A"]).max() <= 1e-12 and np.abs(s["ZmW_b"] - f["ttrkln_ZmW_b"]).max() <= 1e-12
    assert np.abs(s["Xf_track_A"] - f["trk_Xf_A"]).max() <= 1e-12 and np.abs(s["Xf_track_b"] - f["trk_Xf_b"]).max() <= 1e-12
    assert np.abs(s["K"] - f["ttrkln_K"]).max() <= 1e-13 and np.abs(s["P"] - f["ttrkln_P"]).max() <= 1e-11
    # the Mayne example (TubeRegulatorMPC.setup_optimization, B = [0.5, 1], R = 0.01, N = 9): oracle pipeline == reference
    d = su.tube_regulator_setup(s["A"], np.array([[0.5], [1.0]]), np.eye(2), 0.01 * np.eye(1), 9,
                                Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[10.0, 2.0, 10.0, 10.0]), su.box([1.0]),
                                su.box([0.1, 0.1]))
    for k in ("Z", "Xc", "Uc", "Xf"):
        _same_poly(f, "treg_" + k, d[k])
    assert np.abs(d["K"] - f["treg_K"]).max() <= 1e-12 and np.abs(d["P"] - f["treg_P"]).max() <= 1e-10


# ------------------------------------------------------------------------------------------------------------------
# closed loops of the shipped examples (Q1-Q6 statements + state machines + plant in the scripts' order)
# ------------------------------------------------------------------------------------------------------------------
REFS_DI = np.zeros((120, 2))
REFS_DI[0:30, 0], REFS_DI[30:60, 0], REFS_DI[60:90, 0], REFS_DI[90:120, 0] = 5, -9, 9, 4
TOL_LOOP = 1e-9


def test_config1_golden_equals_reference_script_as_shipped():
    """BASELINE.json configs[0]: ``Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py`` as shipped vs ``loop_di_tube.npz``
    (the oracle's run of the same loop, which all config-1 GPU tests compare against)."""
    g, f = H.load("loop_di_tube.npz"), H.load("ref_examples.npz")
    assert np.abs(f["ttrkln_x"].T - g["x"]).max() <= TOL_LOOP
    assert np.abs(f["ttrkln_x_hat"].T - g["x_hat"]).max() <= TOL_LOOP
    assert np.abs(f["ttrkln_x_nom"].T - g["x_nom"]).max() <= TOL_LOOP
    assert np.abs(f["ttrkln_u"].T - g["u"]).max() <= TOL_LOOP
    assert np.array_equal(f["ttrkln_Theta"].astype(int), g["Theta"])
    assert not bool(f["ttrkln_stdout_has_violation"])          # the script's two invariant checks stayed silent


def test_oracle_loops_equal_reference_examples():
    f, s = H.load("ref_examples.npz"), H.load("sets_di.npz")
    A, B = s["A"], s["B"]
    # Example_of_Regulator_MPC.py: input constraint only, T = 20, x0 = (1, 3)
    qp = rq.build_regulator(A, B, np.eye(2), np.eye(1), 10, None, su.box([1.0]))
    x = np.array([1.0, 3.0])
    for t in range(20):
        (_, u), res = rq.solve_param(qp, x.copy())
        assert np.abs(u[:, 0] - f["reg_u"][:, t]).max() <= TOL_LOOP
        x = A @ x + B @ u[:, 0]
        assert np.abs(x - f["reg_x"][:, t + 1]).max() <= TOL_LOOP
    # Example_of_Tracking_MPC.py: Limon tracking MPC, no network
    qp = rq.build_tracking(A, B, s["Q"], s["R"], 10, s["P"], _P(s, "X"), _P(s, "U"), _P(s, "Xf_track"))
    x = np.array([1.0, 2.0])
    for t in range(120):
        (_, u, _, _), res = rq.solve_param(qp, x.copy(), REFS_DI[t].copy())
        x = A @ x + B @ u[:, 0]
        assert np.abs(x - f["trk_x"][:, t + 1]).max() <= TOL_LOOP, t
    # Example_of_Tube_Regulator_MPC.py: Mayne tube MPC (B = [0.5, 1], R = 0.01, N = 9), w from default_rng(1)
    Bm = np.array([[0.5], [1.0]])
    K = f["treg_K"]
    qp = rq.build_tube_regulator(A, Bm, np.eye(2), 0.01 * np.eye(1), 9, f["treg_P"], _P(f, "treg_Xc"), _P(f, "treg_Uc"),
                                 _P(f, "treg_Xf"), _P(f, "treg_Z"))
    rng = np.random.default_rng(1)
    x = np.array([-5.0, -2.0])
    for t in range(10):
        (xm, um), res = rq.solve_param(qp, x.copy())
        u = um[:, 0] - K @ (x - xm[:, 0])
        assert np.abs(xm[:, 0] - f["treg_x_nom"][:, t]).max() <= TOL_LOOP
        x = A @ x + Bm @ u + rng.uniform(-0.1, 0.1, 2)
        assert np.abs(x - f["treg_x"][:, t + 1]).max() <= TOL_LOOP
    # Example_of_Tube_Tracking_MPC.py: Limon tube tracking MPC with the tube-initial constraint (fixed_initial_state=False)
    qp = rq.build_tube_tracking(A, B, s["Q"], s["R"], 10, s["P"], _P(s, "Xc"), _P(s, "Uc"), _P(s, "Xf"), _P(s, "Z"), False)
    rng = np.random.default_rng(1)
    x = np.array([1.0, 2.0])
    for t in range(120):
        (xm, um, _, _), res = rq.solve_param(qp, x.copy(), REFS_DI[t].copy())
        u = um[:, 0] - s["K"] @ (x - xm[:, 0])
        assert np.abs(xm[:, 0] - f["ttrk_x_nom"][:, t]).max() <= 1e-7, t      # x_0 free inside the tube: flat directions
        x = A @ x + B @ u + rng.uniform(-0.1, 0.1, 2)
        assert np.abs(x - f["ttrk_x"][:, t + 1]).max() <= 1e-7, t
    assert not bool(f["ttrk_stdout_has_violation"])
    # Example_of_Tracking_MPC_Over_Lossy_Network.py: Pezzutto remote MPC, p = 0.7, no disturbance
    qp = rq.build_tracking(A, B, s["Q"], s["R"], 10, s["P"], _P(s, "X"), _P(s, "U"), _P(s, "Xf_track"))
    rng_g, rng_t = np.random.default_rng(347), np.random.default_rng(124)
    x0 = np.array([1.0, 2.0])
    est, act = rl.Estimator(A, B, s["K"], x0, 10), rl.SmartActuator(s["K"])
    x, xh = x0.copy(), x0.copy()
    for t in range(120):
        th, ga = (1, 1) if t == 0 else (0 if rng_t.uniform() < 0.7 else 1, 0 if rng_g.uniform() < 0.7 else 1)
        q_t = est.get_qt()
        (_, un, xb, ub), res = rq.solve_param(qp, xh.copy(), REFS_DI[t].copy())
        pkt = rl.encapsulate_controller_packet(un, xb, ub, s["K"], q_t)
        est.store_sent_control_sequence(pkt["U_t"])
        u, ppkt = act.process_packet(pkt, x, th)
        x = A @ x + B @ u
        est.update_estimate(ppkt, ga)
        xh = est.get_estimate()
        assert act.Theta_t == int(f["trkln_Theta"][t])
        assert np.abs(x - f["trkln_x"][:, t + 1]).max() <= TOL_LOOP and np.abs(xh - f["trkln_x_hat"][:, t + 1]).max() <= TOL_LOOP
    assert not bool(f["trkln_stdout_has_violation"])


# ------------------------------------------------------------------------------------------------------------------
# live against the imported reference (build container only)
# ------------------------------------------------------------------------------------------------------------------
@live
def test_live_reference_state_machines_regenerate_fixture():
    R = refshim.reference_modules()
    SA, ES = R["SmartActuator"], R["Estimator"]
    f = H.load("ref_statemachines.npz")
    classes = (SA.SmartActuator, SA.ConsistentActuator, ES.Estimator, ES.RobustEstimator)
    for name, p_i, kind in (("di", 2, "smart"), ("cp", 1, "consistent"), ("cp", 2, "extended")):
        act, est, (theta, gamma, U, xn0, w, x0, A, B, T, nx) = _drive(classes, f, name, p_i, kind)
        x = np.array(x0, float).reshape(nx, 1)
        r = f"{name}_p{p_i}_{kind}_"
        for t in range(T):
            pkt = {"U_t": U[t].copy(), "q_t": est.get_qt()}
            if kind == "extended":
                pkt["x_nom_0"] = xn0[t].copy()
                est.store_current_optimal_inital_nominal_plant_states(xn0[t].copy())
            est.store_sent_control_sequence(pkt["U_t"])
            u, ppkt = act.process_packet(pkt, x, int(theta[t]))
            x = A @ x + B @ u + w[t].reshape(nx, 1)
            est.update_estimate(ppkt, int(gamma[t]))
            assert act.get_Theta_t() == f[r + "Theta"][t] and act.get_s_t() == f[r + "s_t"][t]
            assert np.array_equal(x.flatten(), f[r + "x"][t + 1])
            assert np.array_equal(np.asarray(est.get_estimate()).flatten(), f[r + "x_hat"][t + 1])


def _eq_rows_match(E1, e1, E2, e2, tol=1e-12):
    """Equality rows agree one by one up to the sign of the whole row (``a - b == 0`` against ``b == a``)."""
    assert E1.shape == E2.shape
    R1, R2 = np.c_[E1, e1], np.c_[E2, e2]
    d = np.minimum(np.abs(R1 - R2).max(axis=1), np.abs(R1 + R2).max(axis=1))
    assert d.max() <= tol, d.max()


@live
def test_live_reference_problem_statement_equals_oracle_statement():
    """Q1/Q4/Q5/Q6: the matrices (P, q, E, e, G, h) that the reference's own ``generate_optimization_problem`` code states
    (through the cvxpy stand-in) against the oracle's ``build_*`` restatement, for every controller class."""
    R = refshim.reference_modules()
    pc = __import__("oracle.refshim.polytope", fromlist=["x"])
    s, r = H.load("sets_di.npz"), H.load("ref_examples.npz")
    A, B, Q, Rm, N = s["A"], s["B"], s["Q"], s["R"], int(s["N"])
    rp = lambda f, k: pc.Polytope(f[k + "_A"], f[k + "_b"], normalize=False)       # noqa: E731

    def compare(prob, qp, named, x_init, ref):
        # column permutation: the stand-in stacks variables in order of first use, the oracle as [x | u | x_bar | u_bar | free]
        off = dict(zip([id(v) for v in prob.variables], prob._voff[:-1]))
        cols = []
        for v in named:
            cols += list(off[id(v)] + np.arange(v.leaf_size))
        Pm, q, E, e, G, h, used = prob.standard_form()
        cols = np.array(cols)
        assert set(np.nonzero(used)[0]) <= set(cols)
        q2, e2, h2 = qp.params(x_init, ref)
        nzo = qp.P.shape[0]
        assert len(cols) >= nzo or True
        sub = cols[:nzo] if len(cols) >= nzo else cols
        assert np.abs(Pm[np.ix_(sub, sub)] - qp.P).max() <= 1e-9 * np.abs(qp.P).max()
        assert np.abs(q[sub] - q2).max() <= 1e-9 * (1 + np.abs(q2).max())
        assert E.shape[0] == qp.E.shape[0] and G.shape[0] == qp.G.shape[0]
        _eq_rows_match(E[:, sub], e, qp.E, e2)
        assert np.abs(G[:, sub] - qp.G).max() <= 1e-12 and np.abs(h - h2).max() <= 1e-12

    x_init, ref = np.array([0.7, -0.4]), np.array([3.0, 0.0])
    # TubeTrackingMPC, both initial-state variants
    for fixed in (True, False):
        m = R["TubeTrackingMPC"].TubeTrackingMPC(A, B, Q, Rm, N)
        m._Z, m._Xc, m._Uc, m._Xf = rp(s, "Z"), rp(s, "Xc"), rp(s, "Uc"), rp(s, "Xf")
        m.generate_optimization_problem(fixed)
        m._x_init_param.value, m._ref_param.value = x_init, ref
        qp = rq.build_tube_tracking(A, B, Q, Rm, N, s["P"], _P(s, "Xc"), _P(s, "Uc"), _P(s, "Xf"), _P(s, "Z"), fixed)
        compare(m._prob, qp, [m._x_mpc, m._u_mpc, m._x_bar, m._u_bar], x_init, ref)
    # ExtendedTubeTrackingMPC "packet received" problem incl. the G2 quirk (foreign x_N and u_bar are free variables)
    m = R["TubeTrackingMPC"].ExtendedTubeTrackingMPC(A, B, Q, Rm, N)
    m._Z, m._Xc, m._Uc, m._Xf = rp(s, "Z"), rp(s, "Xc"), rp(s, "Uc"), rp(s, "Xf")
    m.generate_optimization_problem(True)
    m.generate_optimization_problem_when_packet_received(rp(s, "W"))
    m._x_init_param_packet_received.value, m._ref_param_packet_received.value = x_init, ref
    qp = rq.build_extended_packet_received(A, B, Q, Rm, N, s["P"], _P(s, "Xc"), _P(s, "Uc"), _P(s, "Xf"), _P(s, "ZmW"))
    prob = m._prob_packet_received
    Pm, q, E, e, G, h, used = prob.standard_form()
    off = dict(zip([id(v) for v in prob.variables], prob._voff[:-1]))
    nx, nu = 2, 1
    cols = np.r_[off[id(m._x_mpc_packet_received)] + np.arange(nx * (N + 1)), off[id(m._u_mpc_packet_received)] + np.arange(nu * N),
                 off[id(m._x_bar_packet_received)] + np.arange(nx), off[id(m._u_bar_packet_received)] + np.arange(nu),
                 off[id(m._x_mpc)] + nx * N + np.arange(nx), off[id(m._u_bar)] + np.arange(nu)]
    assert set(np.nonzero(used)[0]) == set(cols)               # of the foreign x_mpc only column N occurs
    q2, e2, h2 = qp.params(x_init, ref)
    assert np.abs(Pm[np.ix_(cols, cols)] - qp.P).max() <= 1e-9 * np.abs(qp.P).max()
    _eq_rows_match(E[:, cols], e, qp.E, e2)
    assert np.abs(G[:, cols] - qp.G).max() <= 1e-12
    assert np.abs(q[cols] - q2).max() <= 1e-9 * (1 + np.abs(q2).max()) and np.abs(h - h2).max() <= 1e-12
    # TrackingMPC with and without terminal set
    for with_xf in (True, False):
        m = R["TrackingMPC"].TrackingMPC(A, B, Q, Rm, N)
        m.set_input_constraints(rp(s, "U"))
        m.set_state_constraints(rp(s, "X"))
        if with_xf:
            m._Xf = rp(s, "Xf_track")
        m.generate_optimization_problem()
        m._x_init_param.value, m._ref_param.value = x_init, ref
        qp = rq.build_tracking(A, B, Q, Rm, N, s["P"], _P(s, "X"), _P(s, "U"), _P(s, "Xf_track") if with_xf else None)
        compare(m._prob, qp, [m._x_mpc, m._u_mpc, m._x_bar, m._u_bar], x_init, ref)
    # RegulatorMPC and TubeRegulatorMPC
    m = R["RegulatorMPC"].RegulatorMPC(A, B, Q, Rm, N)
    m.set_input_constraints(rp(s, "U"))
    m.set_state_constraints(rp(s, "X"))
    m.generate_optimization_problem()
    m._x_init_param.value = x_init
    compare(m._prob, rq.build_regulator(A, B, Q, Rm, N, _P(s, "X"), _P(s, "U")), [m._x_mpc, m._u_mpc], x_init, None)
    Bm = np.array([[0.5], [1.0]])
    m = R["TubeRegulatorMPC"].TubeRegulatorMPC(A, Bm, np.eye(2), 0.01 * np.eye(1), 9)
    m._Z, m._Xc, m._Uc, m._Xf = (rp(r, "treg_" + k) for k in ("Z", "Xc", "Uc", "Xf"))
    m.generate_optimization_problem()
    m._x_init_param.value = x_init
    assert np.abs(m._P - r["treg_P"]).max() <= 1e-10 and np.abs(m._K - r["treg_K"]).max() <= 1e-12
    qp = rq.build_tube_regulator(A, Bm, np.eye(2), 0.01 * np.eye(1), 9, r["treg_P"], _P(r, "treg_Xc"), _P(r, "treg_Uc"),
                                 _P(r, "treg_Xf"), _P(r, "treg_Z"))
    compare(m._prob, qp, [m._x_mpc, m._u_mpc], x_init, None)


@live
def test_live_reference_packets_and_g1_gains():
    """Q3 and G1 against the imported classes: packet layout ``[u_0..u_{N-1}, u_bar + K x_bar]``, the argument-order
    quirk of the two ``encapsulate`` methods, ``(packet, x_nom_0)`` of the extended class, ``K``/``P`` of the constructors."""
    R = refshim.reference_modules()
    pc = __import__("oracle.refshim.polytope", fromlist=["x"])
    s = H.load("sets_cp.npz")
    A, B, Q, Rm, N = s["A"], s["B"], s["Q"], s["R"], int(s["N"])
    m = R["TubeTrackingMPC"].ExtendedTubeTrackingMPC(A, B, Q, Rm, N)
    K, P, Acl = rn.lqr_terminal_data(A, B, Q, Rm)
    assert np.abs(m._K - K).max() <= 1e-12 * np.abs(K).max() and np.abs(m._P - P).max() <= 1e-12 * np.abs(P).max()
    assert np.abs(m._K - s["K"]).max() <= 1e-10 and np.abs(m._P - s["P"]).max() <= 1e-9 * np.abs(P).max()
    rng = np.random.default_rng(0)
    u_nom, xb, ub = rng.normal(size=(1, N)), rng.normal(size=4), rng.normal(size=1)
    pkt = m.encapsulate(u_nom.copy(), ub.copy(), xb.copy(), 7)                 # (u_nom, u_ss, x_ss, q_t)
    mine = rl.encapsulate_controller_packet(u_nom, xb, ub, K, 7)
    assert pkt["q_t"] == 7 and np.abs(pkt["U_t"] - mine["U_t"]).max() <= 1e-13
    t = R["TrackingMPC"].TrackingMPC(A, B, Q, Rm, N)
    pkt2 = t.encapsulate(u_nom.copy(), xb.copy(), ub.copy(), 7)                # (u_mpc, x_bar, u_bar, q_t)
    assert np.abs(pkt2["U_t"] - mine["U_t"]).max() <= 1e-13
    assert m.encapsulate(u_nom, None, None, 3)["U_t"] is None


@live
def test_live_shipped_example_regenerates_fixture():
    g = refshim.run_reference_script("Examples of Model Predictive Controllers/Example_of_Tracking_MPC_Over_Lossy_Network.py")
    f = H.load("ref_examples.npz")
    assert np.array_equal(g["x_traj"], f["trkln_x"]) and np.array_equal(g["Theta_t_traj"], f["trkln_Theta"])

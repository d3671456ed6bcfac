"""Parity of the CUDA QP path (through the C ABI) with the oracle's golden solutions.

Tolerance (north_star: "1e-5 relative in FP64"):  |U_t - U_t*|_inf <= 1e-5 max(1, |U_t*|_inf); the
certified active-set endgame actually delivers ~1e-9, asserted below as 1e-7 so a regression shows."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

TOL_SPEC = 1e-5
TOL_TIGHT = 1e-7


def _check(U, Ug, z, zg, status):
    assert np.all(status == 0), np.bincount(status, minlength=4)
    scale = np.maximum(1.0, np.abs(Ug).reshape(len(Ug), -1).max(axis=1))
    err = np.abs(U - Ug).reshape(len(Ug), -1).max(axis=1) / scale
    assert err.max() <= TOL_SPEC
    assert err.max() <= TOL_TIGHT, err.max()
    if zg is not None:
        assert np.abs(z - zg).max() <= TOL_TIGHT * max(1.0, np.abs(zg).max())


def test_double_integrator_tube_tracking():
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_di.npz"), H.load("loop_di_tube.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    z, U, st, it = qp.solve_host(g["xhat_in"], g["refs"])
    _check(U, g["U_t"], z, g["z"], st)
    assert (it & 0xFFF).max() <= 40 and ((it >> 12) & 0xFFF).max() <= 200


def test_double_integrator_extended_both_problems():
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_di.npz"), H.load("loop_di_ext.npz")
    recv = BatchedQP(H.spec_ext_received(s), Kss=s["K"])
    norm = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    mode = g["mode"].astype(np.int32)
    za, Ua, sa, _ = recv.solve_host(g["xhat_in"], g["refs"], sel=mode, sel_value=1)
    zb, Ub, sb, _ = norm.solve_host(g["xhat_in"], g["refs"], sel=mode, sel_value=0)
    U = np.where(mode[:, None, None] == 1, Ua, Ub)
    st = np.where(mode == 1, sa, sb)
    x0 = np.where(mode[:, None] == 1, za[:, :2], zb[:, :2])
    _check(U, g["U_t"], None, None, st)
    assert np.abs(x0 - g["x_nom0"]).max() <= TOL_TIGHT
    assert np.all(sa[mode == 0] == -1) and np.all(sb[mode == 1] == -1)     # unselected instances untouched


def test_double_integrator_tracking_and_regulators():
    from rtmpc_b200.condense import MPCSpec
    from rtmpc_b200.qp import BatchedQP
    s = H.load("sets_di.npz")
    g = H.load("loop_di_track.npz")
    f = g["feasible"] == 1
    qp = BatchedQP(H.spec_tracking(s), Kss=s["K"])
    z, U, st, _ = qp.solve_host(g["xhat_in"][f], g["refs"][f])
    _check(U, g["U_t"][f], None, None, st)
    r = H.load("qp_di_regulators.npz")
    reg = BatchedQP(MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), stage_x=(s["X_A"], s["X_b"]),
                            stage_u=(s["U_A"], s["U_b"])))
    z, U, st, _ = reg.solve_host(r["xs"])
    assert np.all(st == 0) and np.abs(z - r["z_reg"]).max() <= TOL_TIGHT
    mayne = BatchedQP(MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=r["P"],
                              stage_x=(r["Xc_mayne_A"], r["Xc_mayne_b"]), stage_u=(r["Uc_mayne_A"], r["Uc_mayne_b"]),
                              terminal=(r["Xf_mayne_A"], r["Xf_mayne_b"]), tube_init=(r["Z_mayne_A"], r["Z_mayne_b"])))
    z, U, st, _ = mayne.solve_host(r["xs"])
    assert np.all(st == 0) and np.abs(z - r["z_tube"]).max() <= TOL_TIGHT


def test_cartpole_tube_tracking_all_loss_rates():
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    xh = g["tube_xhat_in"].reshape(-1, 4)
    z, U, st, it = qp.solve_host(xh, np.tile(g["refs"], (4, 1)))
    _check(U, g["tube_U_t"].reshape(len(xh), 21, 1), z, g["tube_z"].reshape(len(xh), -1), st)
    # zero constraint violation beyond solver tolerance, on the un-condensed trajectory
    x = z[:, :84].reshape(-1, 21, 4)[:, :20]
    u = z[:, 84:104]
    assert (x.reshape(-1, 4) @ s["Xc_A"].T - s["Xc_b"]).max() <= 1e-8
    assert (np.abs(u) - s["Uc_b"][0]).max() <= 1e-8
    term = np.c_[z[:, 80:84], z[:, 104:109]]
    assert (term @ s["Xf_A"].T - s["Xf_b"]).max() <= 1e-8


def test_cartpole_tracking_and_extended():
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    f = g["track_feasible"].reshape(-1) == 1
    qp = BatchedQP(H.spec_tracking(s), Kss=s["K"])
    z, U, st, _ = qp.solve_host(g["track_xhat_in"].reshape(-1, 4)[f], np.tile(g["refs"], (4, 1))[f])
    _check(U, g["track_U_t"].reshape(-1, 21, 1)[f], None, None, st)
    e = H.load("loop_cp_ext.npz")
    mode = e["mode"].reshape(-1).astype(np.int32)
    X, R, Ug = e["xhat_in"].reshape(-1, 4), np.tile(e["refs"], (2, 1)), e["U_t"].reshape(-1, 21, 1)
    recv = BatchedQP(H.spec_ext_received(s), Kss=s["K"])
    norm = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    za, Ua, sa, _ = recv.solve_host(X, R, sel=mode, sel_value=1)
    zb, Ub, sb, _ = norm.solve_host(X, R, sel=mode, sel_value=0)
    _check(np.where(mode[:, None, None] == 1, Ua, Ub), Ug, None, None, np.where(mode == 1, sa, sb))


def test_infeasible_and_boundary_instances_against_oracle():
    """Random states, many of them infeasible: status must agree with the oracle, payload NaN."""
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope
    from rtmpc_b200.qp import BatchedQP
    s = H.load("sets_di.npz")
    P = lambda k: Polytope(s[k + "_A"], s[k + "_b"], normalize=False)      # noqa: E731
    oq = rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], P("Xc"), P("Uc"), P("Xf"), None, True)
    rng = np.random.default_rng(42)
    X = rng.uniform(-1, 1, (80, 2)) * np.array([9.0, 3.0])
    R = np.c_[rng.uniform(-10, 10, 80), np.zeros(80)]
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    z, U, st, _ = qp.solve_host(X, R)
    n_inf = 0
    for i in range(80):
        sol, res = rq.solve_param(oq, X[i].copy(), R[i].copy())
        if res.status == "infeasible":
            n_inf += 1
            assert st[i] == 2 and np.all(np.isnan(U[i]))
        else:
            assert st[i] in (0, 3), (i, st[i])
            assert np.abs(U[i, :10, 0] - sol[1][0]).max() <= TOL_SPEC
    assert 5 < n_inf < 75


def test_edge_cases_empty_and_single():
    from rtmpc_b200.qp import BatchedQP
    s = H.load("sets_di.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    z, U, st, it = qp.solve_host(np.zeros((0, 2)), np.zeros((0, 2)))
    assert U.shape == (0, 11, 1)
    z, U, st, it = qp.solve_host(np.zeros((1, 2)), np.zeros((1, 2)))
    assert st[0] == 0 and it[0] == 0 and np.abs(U).max() < 1e-12       # origin: unconstrained optimum, 0 iterations


def test_full_size_properties_4096():
    """BASELINE config-2 batch size: every instance optimal and feasible; duplicates give bit-identical answers."""
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    xh = np.tile(g["tube_xhat_in"].reshape(-1, 4), (5, 1))[:4096]
    ref = np.tile(np.tile(g["refs"], (4, 1)), (5, 1))[:4096]
    z, U, st, it = qp.solve_host(xh, ref)
    assert np.all(st == 0)
    assert np.array_equal(U[:1000][:96], U[1000:2000][:96])             # determinism across warps / CTAs
    x = z[:, :84].reshape(-1, 21, 4)[:, :20]
    assert (x.reshape(-1, 4) @ s["Xc_A"].T - s["Xc_b"]).max() <= 1e-8


def test_active_set_kernel_against_interior_point_kernel_on_random_instances():
    """Two independent GPU implementations (dual active set with KKT certificate; Mehrotra interior point with
    active-set endgame) on random parameters, including infeasible ones: same classification, same minimiser."""
    from rtmpc_b200.qp import BatchedQP
    sd, sc = H.load("sets_di.npz"), H.load("sets_cp.npz")
    cases = [("di_tube", H.spec_tube_tracking(sd), sd["K"], np.array([9.0, 3.0]), 9.0, 4000),
             ("cp_tube_near", H.spec_tube_tracking(sc), sc["K"], np.array([1.0, 1.0, 0.05, 0.3]), 1.0, 4000),
             ("cp_tube_far", H.spec_tube_tracking(sc), sc["K"], np.array([5.0, 4.0, 0.1, 1.2]), 5.0, 2000),
             ("cp_ext_recv", H.spec_ext_received(sc), sc["K"], np.array([1.0, 1.0, 0.05, 0.3]), 1.0, 2000)]
    for name, spec, K, box, rbox, B in cases:
        rng = np.random.default_rng(11)
        nx = len(box)
        X = rng.uniform(-1, 1, (B, nx)) * box
        R = np.zeros((B, nx))
        R[:, 0] = rng.uniform(-rbox, rbox, B)
        qp = BatchedQP(spec, Kss=K)
        z1, U1, s1, i1 = qp.solve_host(X, R)
        qp.set_method("interior_point")
        z2, U2, s2, i2 = qp.solve_host(X, R)
        assert not np.any((s1 == 2) & np.isin(s2, (0, 3))), name          # never "infeasible" where the other solves
        assert not np.any((s1 == 0) & (s2 == 2)), name
        both = (s1 == 0) & (s2 == 0)
        assert both.sum() > B // 10, name
        err = np.abs(U1[both] - U2[both]).reshape(both.sum(), -1).max(1)
        scale = np.maximum(1.0, np.abs(U2[both]).reshape(both.sum(), -1).max(1))
        assert (err / scale).max() <= TOL_TIGHT, (name, (err / scale).max())
        # the active-set path certifies at least as many instances as the interior-point path
        assert np.count_nonzero(s1 == 0) >= np.count_nonzero(s2 == 0), name


def test_tracking_mpc_without_terminal_set_uses_the_terminal_equality():
    """TrackingMPC.generate_optimization_problem without Xf: x_N == x_bar (TrackingMPC.py:105-107).  The equality rows
    are zero-width two-sided rows for the active-set kernel."""
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope
    from rtmpc_b200.condense import MPCSpec
    from rtmpc_b200.qp import BatchedQP
    for name, box, rbox in (("sets_di.npz", np.array([2.0, 0.5]), 2.0), ("sets_cp.npz", np.array([0.5, 0.5, 0.05, 0.2]), 0.5)):
        s = H.load(name)
        nx, N = s["A"].shape[0], int(s["N"])
        spec = MPCSpec(s["A"], s["B"], s["Q"], s["R"], N, P_term=s["P"], T_ss=10 * s["P"], stage_x=(s["X_A"], s["X_b"]),
                       stage_u=(s["U_A"], s["U_b"]), terminal=None, terminal_eq=True)
        qp = BatchedQP(spec, Kss=s["K"])
        P = lambda k: Polytope(s[k + "_A"], s[k + "_b"], normalize=False)      # noqa: E731
        oq = rq.build_tracking(s["A"], s["B"], s["Q"], s["R"], N, s["P"], P("X"), P("U"), None)
        rng = np.random.default_rng(3)
        X = rng.uniform(-1, 1, (24, nx)) * box
        R = np.zeros((24, nx))
        R[:, 0] = rng.uniform(-rbox, rbox, 24)
        z, U, st, it = qp.solve_host(X, R)
        n_ok = 0
        for i in range(24):
            sol, res = rq.solve_param(oq, X[i].copy(), R[i].copy())
            if res.status == "infeasible":
                assert st[i] == 2, (name, i, st[i])
                continue
            assert st[i] == 0, (name, i, st[i])
            u = z[i, nx * (N + 1):nx * (N + 1) + N].reshape(N, 1)
            assert np.abs(u - sol[1].T).max() <= TOL_TIGHT * max(1.0, np.abs(sol[1]).max()), (name, i)
            assert np.abs(z[i, nx * N:nx * (N + 1)] - z[i, -(nx + 1):-1]).max() <= 1e-9      # x_N == x_bar
            n_ok += 1
        assert n_ok >= 12, name


def test_multi_input_system_qp_and_rollout():
    """A system that is not in the reference's scripts (nx = 3, nu = 2, two-dimensional steady-state family): the QP
    against the oracle, and the persistent rollout against the step-by-step path (Pezzutto scheme, smart actuator)."""
    import torch
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope as OP
    from rtmpc_b200 import mpc
    from rtmpc_b200.polytope import Polytope
    from rtmpc_b200.rollout import RemoteLoop
    A = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.0, -0.05, 0.95]])
    B = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.1]])
    Q, R, N = np.diag([1.0, 1.0, 0.5]), np.diag([0.1, 0.2]), 6
    XA, Xb = np.r_[np.eye(3), -np.eye(3)], np.array([2.0, 1.0, 1.0, 2.0, 1.0, 1.0])
    UA, Ub = np.r_[np.eye(2), -np.eye(2)], np.array([0.5, 0.4, 0.5, 0.4])
    c = mpc.TrackingMPC(A, B, Q, R, N)
    c.set_state_constraints(Polytope(XA, Xb, normalize=False))
    c.set_input_constraints(Polytope(UA, Ub, normalize=False))
    c.generate_optimization_problem()                       # no terminal set: x_N == x_bar
    oq = rq.build_tracking(A, B, Q, R, N, c._P, OP(XA, Xb, normalize=False), OP(UA, Ub, normalize=False), None)
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (40, 3)) * np.array([1.5, 0.6, 0.6])
    Rf = np.c_[rng.uniform(-1.5, 1.5, 40), np.zeros((40, 2))]
    out = c.solve_batch(X, Rf)
    n_ok = 0
    for i in range(40):
        sol, res = rq.solve_param(oq, X[i].copy(), Rf[i].copy())
        if res.status == "infeasible":
            assert out["status"][i] == 2
            continue
        assert out["status"][i] == 0
        assert np.abs(out["u"][i] - sol[1]).max() <= TOL_TIGHT
        n_ok += 1
    assert n_ok >= 15
    res = {}
    for fused in (True, False):
        loop = RemoteLoop(c, 64, kind="track")
        loop.reset(np.tile([0.5, 0.0, 0.0], (64, 1)))
        tr = loop.run(60, np.array([1.0, 0.0, 0.0]), p_loss=np.linspace(0.0, 0.6, 64), seed=5, record=True, fused=fused)
        res[fused] = (tr.cpu().numpy(), loop.status_count.cpu().numpy())
    # (to rounding: the rollout warm-starts on its carried working-set inverse, RTMPC_TUNE_ROLLOUT_CARRY; tests/test_gpu_loop.py
    # checks the bit-for-bit equality of the two paths with the knob at 0)
    assert np.abs(res[True][0] - res[False][0]).max() <= 1e-9 and np.array_equal(res[True][1], res[False][1])
    assert res[True][1][0] == 64 * 60
    assert np.abs(res[True][0][:, -1, 0] - 1.0).max() < 0.2       # every loop tracks the reference


def test_cartpole_hard_cases_after_reference_jumps():
    """Cold solves whose minimisers sit on 19..21 active rows out of 21 unknowns (states right after large reference
    jumps, tests/golden/make_hard_cases.py): statuses agree with the oracle, every solution passes a solver-independent
    KKT check, solutions agree within the tight tolerance wherever the oracle certified its own answer, and the dual
    active-set kernel gets there itself (no interior-point iterations)."""
    from rtmpc_b200.qp import BatchedQP
    s, g = H.load("sets_cp.npz"), H.load("hard_cp.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    z, U, st, it = qp.solve_host(g["x"], g["ref"])
    assert np.array_equal(st, g["status"]), (np.bincount(st, minlength=4), np.bincount(g["status"], minlength=4))
    ok = st == 0
    pol = g["polished"]
    err = np.abs(z[pol] - g["z"][pol]).max(axis=1) / np.maximum(1.0, np.abs(g["z"][pol]).max(axis=1))
    assert err.max() <= TOL_SPEC
    assert err.max() <= TOL_TIGHT, err.max()
    oq = H.oracle_tube_tracking_qp(s)
    for i in np.nonzero(ok)[0]:
        primal, stationarity = H.kkt_certificate(oq, g["x"][i], g["ref"][i], z[i])
        assert primal <= 1e-10 and stationarity <= 1e-10, (i, primal, stationarity)
    assert np.all(np.isnan(U[~ok]))
    ipm, steps, _ = qp.decode_iters(it)
    assert ipm.max() == 0, int((ipm > 0).sum())
    assert steps.max() <= 16 * 24 + 128
    # every certification forced through the rows of G' z (RTMPC_TUNE_CERT_FACTORED 0): the factored row values only ever
    # replace that pass where they clear the tolerance by the rounding bound, so statuses and solutions are the same bits
    from rtmpc_b200 import _lib
    try:
        _lib.set_tuning(_lib.TUNE_CERT_FACTORED, 0)
        z0, U0, st0, _ = qp.solve_host(g["x"], g["ref"])
    finally:
        _lib.set_tuning(_lib.TUNE_CERT_FACTORED, -1)
    assert np.array_equal(st0, st) and np.array_equal(z0[ok], z[ok]) and np.array_equal(U0[ok], U[ok])
    # the interior-point kernel (the fallback) on its own: same statuses, certified by its endgame (more active rows
    # than unknowns at a degenerate vertex are thinned by the dependency-dropping factorisation)
    qp.set_method("interior_point")
    z2, _, st2, it2 = qp.solve_host(g["x"], g["ref"])
    assert np.array_equal(st2, g["status"]), np.bincount(st2, minlength=4)
    assert (it2 & 0xFFF)[ok].min() >= 1
    assert np.abs(z2[ok] - z[ok]).max() <= TOL_TIGHT * max(1.0, np.abs(z[ok]).max())


def test_handles_are_bound_to_their_device():
    """A handle used while another device is current fails loudly instead of touching foreign memory."""
    import torch
    from rtmpc_b200 import _lib
    from rtmpc_b200.qp import BatchedQP
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = H.load("sets_di.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    L = _lib.lib()
    try:
        _lib.check(L.rtmpc_set_device(1), "rtmpc_set_device")
        with pytest.raises(_lib.RtmpcError, match="belongs to device 0"):
            qp.solve_host(np.zeros((1, 2)), np.zeros((1, 2)))
    finally:
        _lib.check(L.rtmpc_set_device(0), "rtmpc_set_device")
    _, _, st, _ = qp.solve_host(np.zeros((1, 2)), np.zeros((1, 2)))
    assert st[0] == 0


def test_extended_g2_mode_with_a_two_dimensional_steady_state_family():
    """ExtendedTubeTrackingMPC's "packet received" problem in the reference's G2 form (terminal rows act on free variables,
    TubeTrackingMPC.py:293) for a two-input plant, where the steady-state family has dimension 2 and the terminal set has to
    be projected on a 2-D theta (condense._project_terminal_on_theta): against the oracle's un-condensed problem with the
    free variables kept."""
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope as OP
    from rtmpc_b200.condense import MPCSpec
    from rtmpc_b200.qp import BatchedQP
    from rtmpc_b200 import numerics
    rng = np.random.default_rng(7)
    A = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.0, -0.05, 0.95]])
    B = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.1]])
    Q, R, N = np.diag([1.0, 1.0, 0.5]), np.diag([0.1, 0.2]), 6
    nx, nu = B.shape
    K, _, _ = numerics.dlqr(A, B, Q, R)
    Ql = Q + K.T @ R @ K
    P = numerics.dlyap(A - B @ K, (Ql + Ql.T) / 2)
    XA, Xb = np.r_[np.eye(3), -np.eye(3)], np.array([2.0, 1.0, 1.0, 2.0, 1.0, 1.0])
    UA, Ub = np.r_[np.eye(2), -np.eye(2)], np.array([0.5, 0.4, 0.5, 0.4])
    ZA, Zb = np.r_[np.eye(3), -np.eye(3)], np.full(6, 0.05)                    # initial tube: a small box
    HN = rng.normal(size=(40, 2 * nx + nu))                                     # a bounded terminal set on (x_N, x_bar, u_bar)
    HN /= np.linalg.norm(HN, axis=1)[:, None]
    hN = rng.uniform(0.6, 1.2, 40)
    spec = MPCSpec(A, B, Q, R, N, P_term=P, T_ss=10 * P, stage_x=(XA, Xb), stage_u=(UA, Ub), terminal=(HN, hN),
                   tube_init=(ZA, Zb), g2_free_terminal=True)
    qp = BatchedQP(spec, Kss=K)
    assert qp.cq.nth == 2
    oq = rq.build_extended_packet_received(A, B, Q, R, N, P, OP(XA, Xb, normalize=False), OP(UA, Ub, normalize=False),
                                           OP(HN, hN, normalize=False), OP(ZA, Zb, normalize=False))
    X = rng.uniform(-1, 1, (30, 3)) * np.array([1.0, 0.4, 0.4])
    Rf = np.c_[rng.uniform(-1.0, 1.0, 30), np.zeros((30, 2))]
    z, U, st, it = qp.solve_host(X, Rf)
    n_ok = 0
    for i in range(30):
        sol, res = rq.solve_param(oq, X[i].copy(), Rf[i].copy())
        if res.status == "infeasible":
            assert st[i] == 2
            continue
        assert st[i] == 0
        x_o, u_o, xb_o, ub_o = sol
        assert np.abs(U[i, :N].T - u_o).max() <= TOL_TIGHT
        assert np.abs(U[i, N] - (ub_o + K @ xb_o)).max() <= TOL_TIGHT
        n_ok += 1
    assert n_ok >= 10

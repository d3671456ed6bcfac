"""The CUDA path against fixtures generated FROM THE REFERENCE'S OWN CODE (``tests/golden/ref_*.npz``, written by
``tests/golden/make_reference_fixtures.py`` from ``/root/reference`` with the third-party stand-ins of ``oracle/refshim``).
Nothing here reads ``/root/reference`` (it does not exist on the GPU box) and nothing goes through ``oracle/``.

Bars: integer state (Theta_t, s_t, q_t) exact; state-machine floats 1e-12 (same operations, FMA contraction only);
closed loops with a QP solve per step 1e-6 over the whole horizon (north star: 1e-5 relative on U_t, 1e-4 on
trajectories); tracking-error statistics 1e-8.
"""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-6


# ------------------------------------------------------------------------------------------------------------------
# state machines: drop-in classes (split kernels), single instance and batch, and the fused loop-step kernel
# ------------------------------------------------------------------------------------------------------------------
def _sys(f, name):
    return tuple(f[name + k] for k in ("_A", "_B", "_K", "_Kp")) + (int(f[name + "_N"]),)


def _make(kind, A, B, K, Kp, N, x0, batch):
    from rtmpc_b200.local_remote import ConsistentActuator, Estimator, RobustEstimator, SmartActuator
    if kind == "smart":
        act = SmartActuator(K, batch_size=x0.shape[0] if batch else None)
        est = Estimator(A, B, K, x0, N, batch=batch)
    elif kind == "consistent":
        act = ConsistentActuator(A, B, K, Kp, x0, batch=batch)
        est = Estimator(A, B, K, x0, N, batch=batch)
    else:
        act = ConsistentActuator(A, B, K, Kp, x0, is_extended_MPC_used=True, batch=batch)
        est = RobustEstimator(A, B, K, Kp, x0, N, batch=batch)
    return act, est


@pytest.mark.parametrize("name", ["di", "cp"])
@pytest.mark.parametrize("kind", ["smart", "consistent", "extended"])
def test_dropin_classes_single_instance_equal_reference(name, kind):
    """actuator_process_kernel (SMART / CONSISTENT / EXTENDED) and estimator_update_kernel (robust 0 / 1) through
    SmartActuator / ConsistentActuator / Estimator / RobustEstimator in the reference's shapes and call order
    (Results/results_linear_system.py:235-255, ..._with_extendedMPC.py:262-378)."""
    f = H.load("ref_statemachines.npz")
    A, B, K, Kp, N = _sys(f, name)
    nx = A.shape[0]
    for p_i in (1, 2):
        key = f"{name}_p{p_i}_"
        theta, gamma, U, xn0, w, x0 = (f[key + k] for k in ("theta", "gamma", "U", "xn0", "w", "x0"))
        r = key + kind + "_"
        act, est = _make(kind, A, B, K, Kp, N, x0.reshape(nx, 1), False)
        x = x0.reshape(nx, 1).copy()
        for t in range(120):
            q_t = est.get_qt()
            assert q_t == f[r + "q_t"][t]
            pkt = {"U_t": U[t].copy(), "q_t": q_t}
            if kind == "extended":
                pkt["x_nom_0"] = xn0[t].copy()
                est.store_current_optimal_inital_nominal_plant_states(xn0[t].copy())
            est.store_sent_control_sequence(pkt["U_t"])
            u, ppkt = act.process_packet(pkt, x, int(theta[t]))
            assert u.shape == (B.shape[1], 1) and ppkt["x_t"].shape == (nx, 1)
            assert (act.get_Theta_t(), act.get_s_t(), ppkt["s_t"]) == (f[r + "Theta"][t], f[r + "s_t"][t], f[r + "pkt_s"][t])
            sc = 1 + np.abs(f[r + "x"][t]).max()
            assert np.abs(ppkt["x_t"][:, 0] - f[r + "pkt_x"][t]).max() <= 1e-11 * sc
            if kind == "extended":
                assert np.abs(ppkt["x_nom_t"][:, 0] - f[r + "pkt_x_nom"][t]).max() <= 1e-11 * sc
            x = A @ x + B @ u + w[t].reshape(nx, 1)
            est.update_estimate(ppkt, int(gamma[t]))
            sc = 1 + np.abs(f[r + "x"][t + 1]).max()
            assert np.abs(u[:, 0] - f[r + "u"][t]).max() <= 1e-12 * sc
            assert np.abs(x[:, 0] - f[r + "x"][t + 1]).max() <= 1e-11 * sc
            assert np.abs(est.get_estimate()[:, 0] - f[r + "x_hat"][t + 1]).max() <= 1e-11 * sc
            if kind != "smart":
                assert np.abs(act.get_x_nom()[:, 0] - f[r + "x_nom"][t + 1]).max() <= 1e-11 * sc


@pytest.mark.parametrize("name", ["di", "cp"])
@pytest.mark.parametrize("kind", ["smart", "consistent", "extended"])
def test_dropin_classes_batch_mode_equal_reference(name, kind):
    """batch=True of all four classes: the three loss patterns of the fixture as one batch of three instances with
    different packets, device tensors in and out."""
    f = H.load("ref_statemachines.npz")
    A, B, K, Kp, N = _sys(f, name)
    nx, nu = B.shape
    keys = [f"{name}_p{p}_" for p in range(3)]
    st = lambda k: np.stack([f[key + k] for key in keys], axis=1)            # noqa: E731   [T, 3, ...]
    theta, gamma, w, xn0 = st("theta"), st("gamma"), st("w"), st("xn0")
    U = np.transpose(st("U"), (0, 1, 3, 2))                                    # [T, 3, N+1, nu]
    x0 = np.stack([f[key + "x0"] for key in keys])
    rec = lambda k: np.stack([f[key + kind + "_" + k] for key in keys], axis=1)  # noqa: E731
    act, est = _make(kind, A, B, K, Kp, N, x0, True)
    dev = act._dev
    x = torch.as_tensor(x0, device=dev)
    Ad, Bd = torch.as_tensor(A, device=dev), torch.as_tensor(B, device=dev)
    T = 200
    for t in range(T):
        q_t = est.get_qt()
        assert np.array_equal(q_t.cpu().numpy(), rec("q_t")[t])
        pkt = {"U_t": torch.as_tensor(np.ascontiguousarray(U[t]), device=dev), "q_t": q_t}
        if kind == "extended":
            pkt["x_nom_0"] = torch.as_tensor(np.ascontiguousarray(xn0[t]), device=dev)
            est.store_current_optimal_inital_nominal_plant_states(pkt["x_nom_0"])
        est.store_sent_control_sequence(pkt["U_t"])
        u, ppkt = act.process_packet(pkt, x, theta[t])
        assert np.array_equal(act.get_Theta_t().cpu().numpy(), rec("Theta")[t])
        assert np.array_equal(act.get_s_t().cpu().numpy(), rec("s_t")[t])
        x = x @ Ad.T + u @ Bd.T + torch.as_tensor(np.ascontiguousarray(w[t]), device=dev)
        est.update_estimate(ppkt, gamma[t])
        sc = 1 + np.abs(rec("x")[t + 1]).max()
        assert np.abs(x.cpu().numpy() - rec("x")[t + 1]).max() <= 1e-11 * sc
        assert np.abs(est.get_estimate().cpu().numpy() - rec("x_hat")[t + 1]).max() <= 1e-11 * sc
        if kind != "smart":
            assert np.abs(act.get_x_nom().cpu().numpy() - rec("x_nom")[t + 1]).max() <= 1e-11 * sc


@pytest.mark.parametrize("name", ["di", "cp"])
@pytest.mark.parametrize("kind", ["smart", "consistent", "extended"])
def test_fused_loop_step_kernel_equals_reference_state_machines(name, kind):
    """loop_step_kernel (both sides + plant in one launch, O(1) bookkeeping) fed with the fixture's packets through the
    C ABI (rtmpc_loop_step) instead of a QP solution: integers exact, floats to round-off, for all 200 steps."""
    from rtmpc_b200 import _lib
    from rtmpc_b200.rollout import RemoteLoop
    f = H.load("ref_statemachines.npz")
    A, B, K, Kp, N = _sys(f, name)
    nx, nu = B.shape

    class _M:                      # the loop only reads the system matrices and the problem's sizes from the controller
        pass
    m = _M()
    m._A, m._B, m._K, m._N = A, B, K, N
    m._prob = type("P", (), dict(nz=nx, warm_stride=1))()
    keys = [f"{name}_p{p}_" for p in range(3)]
    st = lambda k: np.stack([f[key + k] for key in keys], axis=1)            # noqa: E731
    theta, gamma, w, xn0 = st("theta"), st("gamma"), st("w"), st("xn0")
    U = np.ascontiguousarray(np.transpose(st("U"), (0, 1, 3, 2)))
    x0 = np.stack([f[key + "x0"] for key in keys])
    rec = lambda k: np.stack([f[key + kind + "_" + k] for key in keys], axis=1)  # noqa: E731
    loop = RemoteLoop(m, 3, kind={"smart": "track", "consistent": "tube", "extended": "extended"}[kind], K_plant=Kp)
    loop.reset(x0)
    dev = loop.dev
    status = torch.zeros(3, dtype=torch.int32, device=dev)
    p = _lib.ptr
    for t in range(200):
        assert np.array_equal(loop.q_t.cpu().numpy(), rec("q_t")[t])
        Ud = torch.as_tensor(U[t], device=dev)
        xn = torch.as_tensor(np.ascontiguousarray(xn0[t]), device=dev) if kind == "extended" else None
        th = torch.as_tensor(theta[t].astype(np.int32), device=dev)
        ga = torch.as_tensor(gamma[t].astype(np.int32), device=dev)
        wd = torch.as_tensor(np.ascontiguousarray(w[t]), device=dev)
        _lib.check(loop.L.rtmpc_loop_step(loop._h, p(Ud), p(status), p(xn), nx if xn is not None else 0, None, p(th), p(ga),
                                          p(wd), None, 0, 0, None, 0, torch.cuda.current_stream().cuda_stream), "rtmpc_loop_step")
        assert np.array_equal(loop.Theta.cpu().numpy(), rec("Theta")[t]) and np.array_equal(loop.s_t.cpu().numpy(), rec("s_t")[t])
        sc = 1 + np.abs(rec("x")[t + 1]).max()
        assert np.abs(loop.u.cpu().numpy() - rec("u")[t]).max() <= 1e-12 * sc
        assert np.abs(loop.x.cpu().numpy() - rec("x")[t + 1]).max() <= 1e-11 * sc
        assert np.abs(loop.x_hat.cpu().numpy() - rec("x_hat")[t + 1]).max() <= 1e-11 * sc
        if kind != "smart":
            assert np.abs(loop.x_nom.cpu().numpy() - rec("x_nom")[t + 1]).max() <= 1e-11 * sc


# ------------------------------------------------------------------------------------------------------------------
# the shipped examples
# ------------------------------------------------------------------------------------------------------------------
REFS_DI = np.zeros((120, 2))
REFS_DI[0:30, 0], REFS_DI[30:60, 0], REFS_DI[60:90, 0], REFS_DI[90:120, 0] = 5, -9, 9, 4


def _sub(f, prefix):
    """View of a reference-example fixture with the key layout of sets_*.npz."""
    d = {k[len(prefix):]: f[k] for k in f.files if k.startswith(prefix)}
    return d


def _draws_example(T, p, seeds, w_half, nx):
    """theta / gamma / w exactly as the lossy-network examples draw them (each from its own default_rng; step 0 lossless)."""
    rng_w, rng_g, rng_t = (None if s is None else np.random.default_rng(s) for s in seeds)
    theta, gamma, w = np.ones(T, int), np.ones(T, int), np.zeros((T, nx))
    for t in range(T):
        if t > 0:
            theta[t] = 0 if rng_t.uniform() < p else 1
            gamma[t] = 0 if rng_g.uniform() < p else 1
        if rng_w is not None:
            w[t] = rng_w.uniform(-w_half, w_half, nx)
    return theta, gamma, w


def test_config1_fused_rollout_equals_reference_script_as_shipped():
    """BASELINE.json configs[0]: Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py as shipped (reference code, run in
    the build container) against one persistent-kernel rollout with the same draws; controller built from the sets the
    REFERENCE's setup_optimization produced."""
    from rtmpc_b200.rollout import RemoteLoop
    f = H.load("ref_examples.npz")
    s = _sub(f, "ttrkln_")
    s.update(N=10)
    mpc = H.make_tube_mpc(s)
    theta, gamma, w = _draws_example(120, 0.7, (1, 347, 124), 0.1, 2)
    for fused in (True, False):
        loop = RemoteLoop(mpc, 1, kind="tube", Z=H.poly(s, "Z"))
        loop.reset(np.array([[1.0, 2.0]]))
        tr = loop.run(120, REFS_DI, theta=theta[:, None], gamma=gamma[:, None], w=w[:, None], record=True, fused=fused)
        assert np.abs(tr[0].cpu().numpy() - f["ttrkln_x"].T).max() <= TOL
        assert np.abs(loop.x_hat.cpu().numpy()[0] - f["ttrkln_x_hat"][:, -1]).max() <= TOL
        assert np.abs(loop.x_nom.cpu().numpy()[0] - f["ttrkln_x_nom"][:, -1]).max() <= TOL
        assert loop.tube_max.item() < 1e-7 and not bool(f["ttrkln_stdout_has_violation"])


def test_example_scripts_through_dropin_classes():
    """The other shipped MPC examples, loop bodies written against the drop-in classes exactly as the scripts write them
    against the reference's: Regulator / Tracking / Tube regulator / Tube tracking / Tracking over a lossy network."""
    from rtmpc_b200 import mpc as M
    from rtmpc_b200.local_remote import Estimator, SmartActuator
    from rtmpc_b200.polytope import Polytope
    f = H.load("ref_examples.npz")
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.0], [1.0]])
    nx, nu = 2, 1
    box = lambda h: Polytope(np.r_[np.eye(len(h)), -np.eye(len(h))], np.r_[h, h])       # noqa: E731
    # Example_of_Regulator_MPC.py
    c = M.RegulatorMPC(A, B, np.eye(2), np.eye(1), 10)
    c.set_input_constraints(box([1.0]))
    c.generate_optimization_problem()
    x = np.array([[1.0], [3.0]])
    for t in range(20):
        x_init = x[:, 0].copy()
        _, u_mpc = c.solve_optimization_problem(x_init)
        u = u_mpc[:, 0].reshape(nu, 1)
        x = A @ x + B @ u
        assert np.abs(x[:, 0] - f["reg_x"][:, t + 1]).max() <= TOL
    # Example_of_Tracking_MPC.py (terminal set from the reference's determine_Xf)
    c = M.TrackingMPC(A, B, np.eye(2), np.eye(1), 10)
    c.set_input_constraints(box([1.0]))
    c.set_state_constraints(box([8.0, 8.0]))
    c._Xf = Polytope(f["trk_Xf_A"], f["trk_Xf_b"], normalize=False)
    c.generate_optimization_problem()
    x = np.array([[1.0], [2.0]])
    for t in range(120):
        _, u_mpc, _, _ = c.solve_optimization_problem(x[:, 0].copy(), np.hstack((REFS_DI[t, 0], 0)))
        x = A @ x + B @ u_mpc[:, 0].reshape(nu, 1)
        assert np.abs(x[:, 0] - f["trk_x"][:, t + 1]).max() <= TOL, t
    # Example_of_Tube_Regulator_MPC.py (Mayne; B = [0.5, 1], R = 0.01, N = 9)
    Bm = np.array([[0.5], [1.0]])
    c = M.TubeRegulatorMPC(A, Bm, np.eye(2), 0.01 * np.eye(1), 9)
    assert np.abs(c._K - f["treg_K"]).max() <= 1e-12 and np.abs(c._P - f["treg_P"]).max() <= 1e-10
    c._Z, c._Xc, c._Uc, c._Xf = (Polytope(f[f"treg_{k}_A"], f[f"treg_{k}_b"], normalize=False) for k in ("Z", "Xc", "Uc", "Xf"))
    c.generate_optimization_problem()
    K = c.get_controller_gain()
    rng_w = np.random.default_rng(1)
    x = np.array([[-5.0], [-2.0]])
    for t in range(10):
        x_init = x[:, 0].copy()
        x_mpc, u_mpc = c.solve_optimization_problem(x_init)
        u = u_mpc[:, 0].reshape(nu, 1) - K @ (x_init - x_mpc[:, 0]).reshape(nx, 1)
        x = A @ x + Bm @ u + rng_w.uniform(-0.1, 0.1, nx).reshape(nx, 1)
        assert np.abs(x_mpc[:, 0] - f["treg_x_nom"][:, t]).max() <= TOL
        assert np.abs(x[:, 0] - f["treg_x"][:, t + 1]).max() <= TOL
    # Example_of_Tube_Tracking_MPC.py (fixed_initial_state=False: x_0 free inside the tube)
    s = _sub(f, "ttrk_")
    s.update(N=10)
    c = H.make_tube_mpc(s, fixed_initial_state=False)
    K = c.get_ancillary_controller_gain()
    rng_w = np.random.default_rng(1)
    x = np.array([[1.0], [2.0]])
    for t in range(120):
        x_nom_traj, u_tube_traj, _, _ = c.solve_optimization_problem(x[:, 0].copy(), np.hstack((REFS_DI[t, 0], 0)))
        x_nom_0 = x_nom_traj[:, 0].reshape(nx, 1)
        u = u_tube_traj[:, 0].reshape(nu, 1) - K @ (x - x_nom_0)
        x = A @ x + B @ u + rng_w.uniform(-0.1, 0.1, nx).reshape(nx, 1)
        assert np.abs(u).max() <= 1 + 1e-7                                  # the script's `u not in U` check
        assert np.abs(x_nom_0[:, 0] - f["ttrk_x_nom"][:, t]).max() <= 1e-5, t
        assert np.abs(x[:, 0] - f["ttrk_x"][:, t + 1]).max() <= 1e-5, t
    # Example_of_Tracking_MPC_Over_Lossy_Network.py: Pezzutto's scheme, estimate exact whenever Theta_t = 1 (:141-156)
    c = M.TrackingMPC(A, B, np.eye(2), np.eye(1), 10)
    c.set_input_constraints(box([1.0]))
    c.set_state_constraints(box([8.0, 8.0]))
    c._Xf = Polytope(f["trk_Xf_A"], f["trk_Xf_b"], normalize=False)
    c.generate_optimization_problem()
    K = c.get_steady_state_controller_gain()
    theta, gamma, _ = _draws_example(120, 0.7, (None, 347, 124), 0.0, 2)
    x0 = np.array([[1.0], [2.0]])
    estim, act = Estimator(A, B, K, x0[:], 10), SmartActuator(K)
    x, x_hat = x0.copy(), estim.get_estimate()
    for t in range(120):
        qt = estim.get_qt()
        pkt = c.determine_packet(x_hat, np.hstack((REFS_DI[t, 0], 0)), qt)
        estim.store_sent_control_sequence(pkt["U_t"])
        u_t, plant_packet = act.process_packet(pkt, x, int(theta[t]))
        if act.get_Theta_t() == 1:
            assert np.linalg.norm(x.flatten() - np.asarray(x_hat).flatten()) == 0.0   # exactly, as the script demands
        assert act.get_Theta_t() == int(f["trkln_Theta"][t])
        x = A @ x + B @ u_t
        estim.update_estimate(plant_packet, int(gamma[t]))
        x_hat = estim.get_estimate()
        assert np.abs(x[:, 0] - f["trkln_x"][:, t + 1]).max() <= TOL and np.abs(x_hat[:, 0] - f["trkln_x_hat"][:, t + 1]).max() <= TOL


def test_extended_dropin_objects_in_the_extended_scripts_call_order():
    """ExtendedTubeTrackingMPC.determine_packet(x_hat, ref, q_t, gamma) + ConsistentActuator(is_extended_MPC_used=True) +
    RobustEstimator as results_linear_system_with_extendedMPC.py:262-378 calls them (theta drawn after the solve, gamma
    after the plant step, the solve sees gamma_{t-1}), against the fused extended rollout with the same draws."""
    from rtmpc_b200.local_remote import ConsistentActuator, RobustEstimator
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_di.npz")
    g = H.load("loop_di_ext.npz")
    A, B = s["A"], s["B"]
    mpc = H.make_tube_mpc(s, extended=True)
    K, Kp = mpc.get_steady_state_controller_gain(), mpc.get_ancillary_controller_gain()
    x0 = np.array([[1.0], [2.0]])
    estim = RobustEstimator(A, B, K, Kp, x0[:], 10)
    act = ConsistentActuator(A, B, K, Kp, x0[:], is_extended_MPC_used=True)
    x, x_hat = x0.copy(), estim.get_estimate()
    gamma_t = 1
    T = 60
    for t in range(T):
        qt = estim.get_qt()
        pkt, x_nom_0 = mpc.determine_packet(x_hat, g["refs"][t].copy(), qt, gamma_t)
        assert set(pkt) == {"U_t", "q_t", "x_nom_0"} and pkt["U_t"].shape == (1, 11)
        estim.store_sent_control_sequence(pkt["U_t"])
        estim.store_current_optimal_inital_nominal_plant_states(x_nom_0)
        u_t, plant_packet = act.process_packet(pkt, x, int(g["theta"][t]))
        assert set(plant_packet) == {"x_t", "s_t", "x_nom_t"}
        x = A @ x + B @ u_t + g["w"][t].reshape(2, 1)
        gamma_t = int(g["gamma"][t])
        estim.update_estimate(plant_packet, gamma_t)
        x_hat = estim.get_estimate()
        assert act.get_Theta_t() == g["Theta"][t] and act.get_s_t() == g["s_t"][t]
        assert np.abs(x[:, 0] - g["x"][t + 1]).max() <= TOL and np.abs(x_hat[:, 0] - g["x_hat"][t + 1]).max() <= TOL
        assert np.abs(act.get_x_nom()[:, 0] - g["x_nom"][t + 1]).max() <= TOL
    loop = RemoteLoop(mpc, 1, kind="extended", Z=H.poly(s, "Z"))
    loop.reset(x0.T)
    tr = loop.run(T, g["refs"], theta=g["theta"][:, None], gamma=g["gamma"][:, None], w=g["w"][:, None], record=True)
    assert np.abs(tr[0].cpu().numpy() - g["x"]).max() <= TOL


# ------------------------------------------------------------------------------------------------------------------
# set computations of the product (support sweeps on the GPU) against the reference's own functions
# ------------------------------------------------------------------------------------------------------------------
def _same_set(poly, A, b, tol=1e-9):
    """Same H-representation up to row order (rows are unit-normalised on both sides)."""
    assert poly.A.shape == A.shape, (poly.A.shape, A.shape)
    R1, R2 = np.c_[poly.A, poly.b], np.c_[A, b]
    R1 = R1[np.lexsort(np.round(R1, 7).T[::-1])]
    R2 = R2[np.lexsort(np.round(R2, 7).T[::-1])]
    assert np.abs(R1 - R2).max() <= tol * (1 + np.abs(R2).max())


def test_set_functions_equal_reference_set_examples():
    from rtmpc_b200 import numerics, polytope as pc, sets as up
    f = H.load("ref_examples.npz")
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.5], [1.0]])
    K, _, _ = numerics.dlqr(A, B, np.eye(2), np.eye(1))
    Acl = A - B @ K
    X = pc.Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[4.0, 2.0, 8.0, 4.0])
    for eps, key in ((1e-1, "darup_P1"), (1e-2, "darup_P2"), (1e-3, "darup_P3")):
        rpi, C, status = up.calculate_RPI(Acl, pc.box([0.1, 0.1]), X, pc.box([1.0]), K, eps, 50, return_container=True,
                                          verbose=False)
        assert status == 0
        assert rpi.A.shape == f[key + "_A"].shape                       # 6 k_star rows: k_star = 5 / 6 / 10
        assert np.abs(rpi.A - f[key + "_A"]).max() <= 1e-12 and np.abs(rpi.b - f[key + "_b"]).max() <= 1e-12
        if key == "darup_P1":
            assert np.abs(C.b - f["darup_C1_b"]).max() <= 1e-12
    Acl_r = A - np.array([[1.0], [1.0]]) @ np.array([[1.17, 1.03]])
    Fs, status = up.calculate_minimal_robust_positively_invariant_set(Acl_r, pc.box([1.0, 1.0]), eps_var=1.9e-5)
    assert status == 0
    _same_set(Fs, f["rakovic_Fs_A"], f["rakovic_Fs_b"])
    # Mayne Fig. 2 pipeline: Rakovic -> tighten -> MOAS -> Minkowski sum
    K, _, _ = numerics.dlqr(A, B, np.eye(2), 0.01 * np.eye(1))
    Acl = A - B @ K
    Z, _ = up.calculate_minimal_robust_positively_invariant_set(Acl, pc.box([0.1, 0.1]))
    _same_set(Z, f["moas_Z_A"], f["moas_Z_b"])
    Xm = pc.Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[10.0, 2.0, 10.0, 10.0])
    Xc, Uc = up.pont_diff(Xm, Z), up.pont_diff(pc.box([1.0]), up.scale(Z, -K))
    assert np.abs(Xc.b - f["moas_Xc_b"]).max() <= 1e-12 and np.abs(Uc.b - f["moas_Uc_b"]).max() <= 1e-12
    Xf = up.calculate_maximum_admissible_output_set(Acl, pc.Polytope(np.r_[Xc.A, -Uc.A @ K], np.r_[Xc.b, Uc.b]), verbose=False)
    _same_set(Xf, f["moas_Xf_A"], f["moas_Xf_b"])
    _same_set(up.mink_sum(Xf, Z), f["moas_XfpZ_A"], f["moas_XfpZ_b"])
    # "Example of Several Set Operations.py"
    P1, P2 = pc.box([3.0, 3.0]), pc.qhull(np.array([[1.0, 0], [0, -1], [-1, 0], [0, 1]]))
    _same_set(up.pont_diff(P1, P2), f["ops_P_diff_A"], f["ops_P_diff_b"])
    c = np.cos(np.pi / 4)
    _same_set(up.scale(P2, np.array([[c, c], [-c, c]])), f["ops_P_m2_scaled_A"], f["ops_P_m2_scaled_b"])
    _same_set(up.mink_sum(pc.box([2.0, 2.0]), pc.box([1.0, 1.0])), f["ops_P_mink_A"], f["ops_P_mink_b"])
    _same_set(up.mink_sum(pc.box([2.0, 2.0]), np.ones((1, 2))), f["ops_P_mink_x_A"], f["ops_P_mink_x_b"])


def test_double_integrator_setup_optimization_equals_reference_setup():
    """The product's TubeTrackingMPC.setup_optimization / ExtendedTubeTrackingMPC / TrackingMPC.setup_optimization on the
    example system against the sets the REFERENCE's own setup_optimization produced."""
    from rtmpc_b200 import mpc as M
    from rtmpc_b200 import polytope as pc
    f = H.load("ref_examples.npz")
    A, B = np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[0.0], [1.0]])
    c = M.ExtendedTubeTrackingMPC(A, B, np.eye(2), np.eye(1), 10)
    c.set_input_constraints(pc.box([1.0]))
    c.set_state_constraints(pc.box([8.0, 8.0]))
    c.setup_optimization(pc.box([0.1, 0.1]), fixed_initial_state=True)
    for k in ("Z", "Xc", "Uc", "Xf"):
        _same_set(getattr(c, "_" + k), f[f"ttrkln_{k}_A"], f[f"ttrkln_{k}_b"])
    _same_set(c._ZmW, f["ttrkln_ZmW_A"], f["ttrkln_ZmW_b"])
    assert np.abs(c._K - f["ttrkln_K"]).max() <= 1e-12 and np.abs(c._P - f["ttrkln_P"]).max() <= 1e-10
    t = M.TrackingMPC(A, B, np.eye(2), np.eye(1), 10)
    t.set_input_constraints(pc.box([1.0]))
    t.set_state_constraints(pc.box([8.0, 8.0]))
    t.setup_optimization()
    _same_set(t._Xf, f["trk_Xf_A"], f["trk_Xf_b"])

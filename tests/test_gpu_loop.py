"""Closed-loop parity: fused loop kernel + QP kernel against the oracle's golden runs on identical
(theta, gamma, w) arrays.  Stated tolerance: trajectories within 1e-4 absolute over the
whole horizon and identical integer sequences (Theta_t, s_t, q_t); asserted tighter (1e-6)."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _run(loop, T, refs, theta, gamma, w):
    B = loop.B
    rec = dict(x=[loop.x.cpu().numpy().copy()], x_nom=[loop.x_nom.cpu().numpy().copy()],
               x_hat=[loop.x_hat.cpu().numpy().copy()], Theta=[], s_t=[], q_t=[], alive=[])
    dev = loop.dev
    for t in range(T):
        rec["q_t"].append(loop.q_t.cpu().numpy().copy())
        ref_d = torch.as_tensor(np.broadcast_to(refs[t], (B, loop.nx)).copy(), device=dev)
        th = torch.as_tensor(theta[:, t].astype(np.int32).copy(), device=dev)
        ga = torch.as_tensor(gamma[:, t].astype(np.int32).copy(), device=dev)
        ww = torch.as_tensor(np.ascontiguousarray(w[:, t]), device=dev)
        loop.step(ref_d, th, ga, ww)
        for k, v in (("x", loop.x), ("x_nom", loop.x_nom), ("x_hat", loop.x_hat), ("Theta", loop.Theta), ("s_t", loop.s_t),
                     ("alive", loop.alive)):
            rec[k].append(v.cpu().numpy().copy())
    return {k: np.stack(v, axis=1) for k, v in rec.items()}


def test_config1_double_integrator_as_shipped():
    from rtmpc_b200.rollout import RemoteLoop
    s, g = H.load("sets_di.npz"), H.load("loop_di_tube.npz")
    mpc = H.make_tube_mpc(s)
    loop = RemoteLoop(mpc, 1, kind="tube", Z=H.poly(s, "Z"))
    loop.reset(np.array([[1.0, 2.0]]))
    r = _run(loop, 120, g["refs"], g["theta"][None], g["gamma"][None], g["w"][None])
    assert np.abs(r["x"][0] - g["x"]).max() <= TOL
    assert np.abs(r["x_nom"][0] - g["x_nom"]).max() <= TOL
    assert np.abs(r["x_hat"][0] - g["x_hat"]).max() <= TOL
    assert np.array_equal(r["Theta"][0], g["Theta"]) and np.array_equal(r["s_t"][0], g["s_t"])
    assert np.array_equal(r["q_t"][0], g["q_t"])
    assert loop.tube_max.item() < 1e-7                      # x - x_nom in Z at every step (the example's check)
    err = loop.tracking_error(120).item()
    ref_err = np.sqrt(((g["x"][:-1] - g["refs"]) ** 2).sum()) / 120
    assert abs(err - ref_err) <= 1e-8


def test_config2_cartpole_tube_and_track_four_loss_rates():
    from rtmpc_b200.rollout import RemoteLoop
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    mpc = H.make_tube_mpc(s)
    loop = RemoteLoop(mpc, 4, kind="tube", Z=H.poly(s, "Z"))
    loop.reset(np.zeros((4, 4)))
    r = _run(loop, 250, g["refs"], g["theta"], g["gamma"], g["w"])
    assert np.abs(r["x"] - g["tube_x"]).max() <= TOL
    assert np.abs(r["x_hat"] - g["tube_x_hat"]).max() <= TOL
    assert np.array_equal(r["Theta"], g["tube_Theta"]) and np.array_equal(r["s_t"], g["tube_s_t"])
    assert np.array_equal(r["q_t"], g["tube_q_t"])
    assert loop.tube_max.max().item() < 1e-7
    assert loop.status_count[0].item() == 1000
    trk = H.make_track_mpc(s)
    loop = RemoteLoop(trk, 4, kind="track")
    loop.reset(np.zeros((4, 4)))
    r = _run(loop, 250, g["refs"], g["theta"], g["gamma"], g["w"])
    assert np.abs(r["x"] - g["track_x"]).max() <= TOL
    assert np.abs(r["x_hat"] - g["track_x_hat"]).max() <= TOL
    assert np.array_equal(r["Theta"], g["track_Theta"])


def test_config3_extended_cartpole_and_double_integrator():
    from rtmpc_b200.rollout import RemoteLoop
    for sets, golden, B in (("sets_di.npz", "loop_di_ext.npz", 1), ("sets_cp.npz", "loop_cp_ext.npz", 2)):
        s, g = H.load(sets), H.load(golden)
        mpc = H.make_tube_mpc(s, extended=True)
        nx = s["A"].shape[0]
        if B == 1:
            th, ga, w, x, xh, xn, Th = (g[k][None] for k in ("theta", "gamma", "w", "x", "x_hat", "x_nom", "Theta"))
            x0 = np.array([[1.0, 2.0]])
        else:
            th, ga, w, x, xh, xn, Th = (g[k] for k in ("theta", "gamma", "w", "x", "x_hat", "x_nom", "Theta"))
            x0 = np.zeros((B, nx))
        loop = RemoteLoop(mpc, B, kind="extended", Z=H.poly(s, "Z"))
        loop.reset(x0)
        r = _run(loop, th.shape[1], g["refs"], th, ga, w)
        assert np.abs(r["x"] - x).max() <= TOL, golden
        assert np.abs(r["x_hat"] - xh).max() <= TOL
        assert np.abs(r["x_nom"] - xn).max() <= TOL
        assert np.array_equal(r["Theta"], Th)
        assert loop.tube_max.max().item() < 1e-7


def test_reference_call_order_with_class_objects():
    """The reference's own loop body (Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py:118-163)
    written against the drop-in classes, single instance."""
    from rtmpc_b200.local_remote import ConsistentActuator, Estimator
    s, g = H.load("sets_di.npz"), H.load("loop_di_tube.npz")
    A, B = s["A"], s["B"]
    mpc = H.make_tube_mpc(s)
    K_ss = mpc.get_steady_state_controller_gain()
    x0 = np.array([[1.0], [2.0]])
    estim = Estimator(A, B, K_ss, x0[:], 10)
    act = ConsistentActuator(A, B, K_ss, mpc.get_ancillary_controller_gain(), x0[:])
    x = x0.copy()
    x_hat = estim.get_estimate()
    T = 40
    for t in range(T):
        qt = estim.get_qt()
        pkt = mpc.determine_packet(x_hat, np.hstack((g["refs"][t, 0], 0)), qt)
        estim.store_sent_control_sequence(pkt["U_t"])
        u_t, plant_packet = act.process_packet(pkt, x, int(g["theta"][t]))
        assert pkt["U_t"].shape == (1, 11) and u_t.shape == (1, 1)
        x = A @ x + B @ u_t + g["w"][t].reshape(2, 1)
        estim.update_estimate(plant_packet, int(g["gamma"][t]))
        x_hat = estim.get_estimate()
        assert act.get_Theta_t() == g["Theta"][t] and act.get_s_t() == g["s_t"][t]
        assert np.abs(x[:, 0] - g["x"][t + 1]).max() <= TOL
        assert np.abs(x_hat[:, 0] - g["x_hat"][t + 1]).max() <= TOL
        assert np.abs(act.get_x_nom()[:, 0] - g["x_nom"][t + 1]).max() <= TOL
    assert len(mpc.get_computational_times()) == T


def test_device_rng_matches_host_restatement_and_sharding():
    """RNG mode == explicit mode fed with the numpy Philox restatement; and splitting the batch over two
    'ranks' with id_offset reproduces the unsplit run bit for bit (results independent of #GPUs)."""
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_cp.npz")
    mpc = H.make_tube_mpc(s)
    Bn, T, seed = 64, 40, 679
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    p = np.array([0.1 * (i % 10) for i in range(Bn)])
    ref = np.array([0.5, 0, 0, 0])
    a = RemoteLoop(mpc, Bn, kind="tube", w_half=hw)
    a.reset()
    ta = a.run(T, ref, p_loss=p, seed=seed, record=True).cpu().numpy()
    th, ga, w = H.device_draws(seed, np.arange(Bn), T, p, hw)
    b = RemoteLoop(mpc, Bn, kind="tube")
    b.reset()
    tb = b.run(T, ref, theta=th, gamma=ga, w=w, record=True).cpu().numpy()
    assert np.abs(ta - tb).max() <= 1e-12
    halves = []
    for r in range(2):
        c = RemoteLoop(mpc, Bn // 2, kind="tube", w_half=hw)
        c.reset()
        halves.append(c.run(T, ref, p_loss=p[r * 32:(r + 1) * 32], seed=seed, id_offset=r * 32, record=True).cpu().numpy())
    assert np.array_equal(np.concatenate(halves), ta)


def test_nonlinear_cartpole_plant_config4_slice():
    """Config 4 loop structure: analytic cartpole ODE plant (10 sub-steps), controller from the
    linear model.  Checked against the oracle's ODE restatement driving the same applied inputs."""
    from oracle import ref_loop as rl
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_cp.npz")
    mpc = H.make_tube_mpc(s)
    loop = RemoteLoop(mpc, 8, kind="tube", plant="cartpole")
    loop.reset()
    ref = torch.as_tensor(np.tile([0.5, 0, 0, 0.0], (8, 1)), device=loop.dev)
    p = torch.as_tensor(np.linspace(0, 0.7, 8), device=loop.dev)
    x = np.zeros((8, 4))
    for t in range(60):
        loop.step(ref, p_loss=p, seed=124)
        u = loop.u.cpu().numpy()
        for b in range(8):
            xb = x[b]
            for _ in range(10):
                xb = rl.cartpole_ode_step(xb, u[b, 0])
            x[b] = xb
        assert np.abs(loop.x.cpu().numpy() - x).max() <= 1e-10
    assert loop.status_count[2].item() == 0
    assert np.abs(x[:, 0] - 0.5).max() < 0.45          # all carts moved towards the reference


@pytest.fixture
def carry(request):
    """RTMPC_TUNE_ROLLOUT_CARRY for one test (restored afterwards)"""
    from rtmpc_b200 import _lib
    _lib.set_tuning(_lib.TUNE_ROLLOUT_CARRY, request.param)
    yield request.param
    _lib.set_tuning(_lib.TUNE_ROLLOUT_CARRY, -1)


def same(a, b, exact):
    """bit for bit, or - with the carried working-set inverse - to rounding (integers always exactly)"""
    a, b = np.asarray(a), np.asarray(b)
    if exact or a.dtype.kind in "iub":
        return np.array_equal(a, b)
    return a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("carry", [0, 1], indirect=True)
def test_fused_rollout_matches_golden_and_stepwise_bit_for_bit(carry):
    """rtmpc_loop_rollout (one persistent launch for all T steps) against the oracle's golden closed loops
    and against the step-by-step path (QP launch + loop-step launch per control step): bit for bit when every solve
    moves and inverts its working set the way rtmpc_qp_solve does (carry 0), to rounding with the carried inverse."""
    from rtmpc_b200.rollout import RemoteLoop
    s, g = H.load("sets_di.npz"), H.load("loop_di_tube.npz")
    mpc = H.make_tube_mpc(s)
    loop = RemoteLoop(mpc, 1, kind="tube", Z=H.poly(s, "Z"))
    loop.reset(np.array([[1.0, 2.0]]))
    tr = loop.run(120, g["refs"], theta=g["theta"][:, None], gamma=g["gamma"][:, None], w=g["w"][:, None], record=True,
                  fused=True).cpu().numpy()
    assert np.abs(tr[0] - g["x"]).max() <= TOL
    assert np.abs(loop.x_hat.cpu().numpy()[0] - g["x_hat"][-1]).max() <= TOL
    assert loop.tube_max.item() < 1e-7
    assert loop.status_count[0].item() == 120

    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    for kind, mk, key in (("tube", H.make_tube_mpc, "tube"), ("track", H.make_track_mpc, "track")):
        mpc = mk(s)
        out = {}
        for fused in (True, False):
            loop = RemoteLoop(mpc, 4, kind=kind, Z=H.poly(s, "Z") if kind == "tube" else None)
            loop.reset(np.zeros((4, 4)))
            tr = loop.run(250, g["refs"], theta=g["theta"].T, gamma=g["gamma"].T, w=np.transpose(g["w"], (1, 0, 2)),
                          record=True, fused=fused).cpu().numpy()
            out[fused] = (tr, loop.x_hat.cpu().numpy(), loop.x_nom.cpu().numpy(), loop.s_t.cpu().numpy(),
                          loop.q_t.cpu().numpy(), loop.Theta.cpu().numpy(), loop.err_acc.cpu().numpy(),
                          loop.status_count.cpu().numpy(),
                          # tube statistic: the rollout's scan stops early (sorted facets), the step kernel reads every facet
                          loop.tube_max.cpu().numpy())
        assert np.abs(out[True][0] - g[key + "_x"]).max() <= TOL
        assert np.abs(out[True][1] - g[key + "_x_hat"][:, -1]).max() <= TOL
        for a, b in zip(out[True], out[False]):
            assert same(a, b, exact=(carry == 0))
        if kind == "tube":
            assert out[True][7][0] == 1000


def test_fused_rollout_with_instances_handed_to_the_interior_point_kernel():
    """A tiny active-set step cap forces most constrained solves through the hand-over path (instance parks,
    interior-point kernel solves its step, rollout resumes): same closed loop within the parity tolerance."""
    from rtmpc_b200.rollout import RemoteLoop
    s, g = H.load("sets_cp.npz"), H.load("loop_cp_tube.npz")
    mpc = H.make_tube_mpc(s)
    mpc._prob.set_step_cap(2)
    try:
        for fused in (True, False):
            loop = RemoteLoop(mpc, 4, kind="tube", Z=H.poly(s, "Z"))
            loop.reset(np.zeros((4, 4)))
            tr = loop.run(250, g["refs"], theta=g["theta"].T, gamma=g["gamma"].T, w=np.transpose(g["w"], (1, 0, 2)),
                          record=True, fused=fused).cpu().numpy()
            assert np.abs(tr - g["tube_x"]).max() <= TOL
            assert np.abs(loop.x_hat.cpu().numpy() - g["tube_x_hat"][:, -1]).max() <= TOL
            assert loop.status_count[0].item() == 1000
            assert loop.iters_total[0].item() > 1000          # interior-point iterations were needed
            assert loop.tube_max.max().item() < 1e-7
    finally:
        mpc._prob.set_step_cap(0)


@pytest.mark.parametrize("carry", [0, 1], indirect=True)
def test_fused_rollout_extended_variant_two_problems(carry):
    """Config 3: ExtendedTubeTrackingMPC (two QPs switched on gamma_{t-1}) + RobustEstimator + x_nom_0 in the packet,
    all T steps in one launch: against the oracle's golden runs and against the step-by-step path (bit for bit with
    carry 0, to rounding with the carried inverse)."""
    from rtmpc_b200.rollout import RemoteLoop
    for sets, golden, B in (("sets_di.npz", "loop_di_ext.npz", 1), ("sets_cp.npz", "loop_cp_ext.npz", 2)):
        s, g = H.load(sets), H.load(golden)
        nx = s["A"].shape[0]
        if B == 1:
            th, ga, w, x, xh, xn = (g[k][None] for k in ("theta", "gamma", "w", "x", "x_hat", "x_nom"))
            x0 = np.array([[1.0, 2.0]])
        else:
            th, ga, w, x, xh, xn = (g[k] for k in ("theta", "gamma", "w", "x", "x_hat", "x_nom"))
            x0 = np.zeros((B, nx))
        T = th.shape[1]
        out = {}
        for fused in (True, False):
            mpc = H.make_tube_mpc(s, extended=True)
            loop = RemoteLoop(mpc, B, kind="extended", Z=H.poly(s, "Z"))
            loop.reset(x0)
            tr = loop.run(T, g["refs"], theta=th.T, gamma=ga.T, w=np.transpose(w, (1, 0, 2)), record=True,
                          fused=fused).cpu().numpy()
            out[fused] = (tr, loop.x_hat.cpu().numpy(), loop.x_nom.cpu().numpy(), loop.Theta.cpu().numpy(),
                          loop.s_t.cpu().numpy(), loop.q_t.cpu().numpy())
            assert loop.tube_max.max().item() < 1e-7
            assert loop.status_count[0].item() == B * T
        assert np.abs(out[True][0] - x).max() <= TOL, golden
        assert np.abs(out[True][1] - xh[:, -1]).max() <= TOL
        assert np.abs(out[True][2] - xn[:, -1]).max() <= TOL
        for a, b in zip(out[True], out[False]):
            assert same(a, b, exact=(carry == 0)), golden


def test_full_size_baseline_config_properties():
    """BASELINE configs[1] at its full size (4096 closed loops x 250 steps, device RNG): what must hold irrespective of
    size - every solve certified, the tube invariant x - x_nom in Z and the true state / input constraints at every
    step (robust constraint satisfaction is the method's guarantee), results independent of how the batch is cut
    (bit for bit), and solves at states visited by the big run that pass a solver-independent KKT check."""
    import bench
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_cp.npz")
    mpc, Z = bench.build_controller()
    B, T = bench.B_PER_GPU, bench.T_STEPS
    p = np.array([0.1 * (i % 10) for i in range(B)])
    loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
    loop.reset()
    tr = loop.run(T, bench.REF, p_loss=p, seed=bench.SEED, record=True).cpu().numpy()
    st = loop.stats.cpu().numpy()
    assert st[:4].tolist() == [B * T, 0, 0, 0] and st[4] == 0            # all optimal, nothing handed over
    assert int(loop.alive.sum().item()) == B
    assert loop.tube_max.max().item() <= 0.0
    X = H.poly(s, "X")
    assert (tr.reshape(-1, 4) @ X.A.T - X.b).max() <= 0.0                 # |x| <= (5, 5, 0.3, 2) at every step of every loop
    err = loop.tracking_error(T).cpu().numpy()
    assert np.all(np.isfinite(err)) and err.max() < 0.1
    # loss probability 0: the nominal state settles on the reference, the state stays within the tube around it
    from scipy.optimize import linprog
    half = -linprog(-np.eye(4)[0], A_ub=Z.A, b_ub=Z.b, bounds=(None, None)).fun       # h_Z(e_1)
    assert np.abs(tr[::10, -1, 0] - bench.REF[0]).max() <= half + 1e-3
    # cut invariance: instances 1000..1127 on their own (global ids through id_offset) - bit for bit
    sub = RemoteLoop(mpc, 128, kind="tube", w_half=bench.HW, Z=Z)
    sub.reset()
    tr2 = sub.run(T, bench.REF, p_loss=p[1000:1128], seed=bench.SEED, id_offset=1000, record=True).cpu().numpy()
    assert np.array_equal(tr2, tr[1000:1128])
    # solves at visited states, checked without any solver in the loop
    rng = np.random.default_rng(5)
    xs = tr[rng.integers(0, B, 48), rng.integers(0, 60, 48)]              # the transient, where constraints are active
    z, U, stq, _ = mpc._prob.solve_host(xs, np.tile(bench.REF, (48, 1)))
    # (plant states, not estimates: with x_0 fixed to a disturbed state some of these problems are infeasible)
    from oracle import ref_qp as rq
    assert set(np.unique(stq)) <= {0, 2} and (stq == 0).sum() >= 24
    oq = H.oracle_tube_tracking_qp(s)
    for x, zz, sq in zip(xs, z, stq):
        if sq == 0:
            primal, stationarity = H.kkt_certificate(oq, x, bench.REF, zz)
            assert primal <= 1e-10 and stationarity <= 1e-10
        else:
            _, e, h = oq.params(x, bench.REF)
            assert not rq.is_feasible(oq.E, e, oq.G, h)


@pytest.mark.parametrize("kind,step_cap", [("tube", 0), ("extended", 0), ("tube", 2)])
def test_time_sliced_rollout_is_bit_identical_to_whole_chains(kind, step_cap):
    """More instances than warp slots: the rollout kernel slices the chains into 25-step tickets handed between warps
    (and SMs).  The same instances run as small batches (fewer instances than slots: every warp keeps its chain) must
    give the same bits - also for the two-problem variant and with instances parked for the interior-point kernel."""
    import bench
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_cp.npz")
    mpc, Z = bench.build_controller(extended=(kind == "extended"))
    B, T = 3072, 60                       # 148 SMs x 16 warps = 2368 slots < 3072; 60 steps = three tickets per chain
    p = np.array([0.1 * (i % 10) for i in range(B)])
    mpc._prob.set_step_cap(step_cap)
    try:
        big = RemoteLoop(mpc, B, kind=kind, w_half=bench.HW, Z=Z)
        big.reset()
        tr = big.run(T, bench.REF, p_loss=p, seed=11, record=True).cpu().numpy()
        st_big = big.stats.cpu().numpy()
        xh, sT = big.x_hat.cpu().numpy(), big.s_t.cpu().numpy()
        if step_cap:
            assert st_big[4] > 0                                     # interior-point iterations: instances did park
        parts, stats = [], np.zeros(4, np.int64)
        for off in range(0, B, 1024):
            sub = RemoteLoop(mpc, 1024, kind=kind, w_half=bench.HW, Z=Z)
            sub.reset()
            parts.append((sub.run(T, bench.REF, p_loss=p[off:off + 1024], seed=11, id_offset=off, record=True).cpu().numpy(),
                          sub.x_hat.cpu().numpy(), sub.s_t.cpu().numpy()))
            stats += sub.stats.cpu().numpy()[:4].astype(np.int64)
        assert np.array_equal(tr, np.concatenate([q[0] for q in parts]))
        assert np.array_equal(xh, np.concatenate([q[1] for q in parts]))
        assert np.array_equal(sT, np.concatenate([q[2] for q in parts]))
        assert np.array_equal(st_big[:4].astype(np.int64), stats)
        assert st_big[:4].sum() == B * T or kind == "extended"
    finally:
        mpc._prob.set_step_cap(0)


@pytest.mark.parametrize("plant", ["linear", "cartpole"])
def test_fixed_dimension_instantiation_is_bit_identical_to_the_general_one(plant):
    """RTMPC_TUNE_ROLLOUT_FIXED_DIMS: the rollout kernel instantiated with the cartpole controller's dimensions as
    compile-time constants does the same arithmetic in the same order as the general instantiation."""
    import bench
    from rtmpc_b200 import _lib
    from rtmpc_b200.rollout import RemoteLoop
    mpc, Z = bench.build_controller(extended=False)
    assert _lib.get_tuning(_lib.TUNE_ROLLOUT_FIXED_DIMS) == 1
    B, T = 3000, 75
    p = np.array([0.1 * (i % 10) for i in range(B)])
    out = []
    try:
        for fixed in (1, 0):
            _lib.set_tuning(_lib.TUNE_ROLLOUT_FIXED_DIMS, fixed)
            loop = RemoteLoop(mpc, B, kind="tube", plant=plant, w_half=None if plant == "cartpole" else bench.HW, Z=Z)
            loop.reset()
            tr = loop.run(T, bench.REF, p_loss=p, seed=5, record=True).cpu().numpy()
            out.append((tr, loop.x_hat.cpu().numpy(), loop.s_t.cpu().numpy(), loop.stats.cpu().numpy(),
                        loop.tube_max.cpu().numpy(), mpc._prob.rollout_kernel))
    finally:
        _lib.set_tuning(_lib.TUNE_ROLLOUT_FIXED_DIMS, -1)
    assert "cartpole dims" in out[0][5] and "cartpole dims" not in out[1][5]
    for a, b in zip(out[0][:5], out[1][:5]):
        assert np.array_equal(a, b)
    assert out[0][3][:4].sum() == B * T


def test_rollout_records_into_a_pinned_host_buffer():
    """run(..., record=True, out=pinned host tensor): the kernel writes the trajectory into host memory itself; same
    bits as recording on the device and copying back."""
    import bench
    from rtmpc_b200 import _lib
    from rtmpc_b200.rollout import RemoteLoop
    mpc, Z = bench.build_controller(extended=False)
    B, T = 2500, 60
    p = np.array([0.1 * (i % 10) for i in range(B)])
    loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
    loop.reset()
    dev_tr = loop.run(T, bench.REF, p_loss=p, seed=3, record=True).cpu()
    host = torch.full((B, T + 1, 4), float("nan"), dtype=torch.float64).pin_memory()
    loop.reset()
    ret = loop.run(T, bench.REF, p_loss=p, seed=3, record=True, out=host)
    torch.cuda.synchronize()
    assert ret is host and torch.equal(host, dev_tr)
    with pytest.raises(_lib.RtmpcError):
        loop.reset()
        loop.run(T, bench.REF, p_loss=p, seed=3, record=True, out=torch.zeros(B, T + 1, 4, dtype=torch.float64))   # pageable


def test_factored_certification_gives_the_same_closed_loops():
    """RTMPC_TUNE_CERT_FACTORED: accepting the certification on row values through the factored tables (where they clear
    the tolerance by the rounding bound) against forming the rows from G' z in every certification: same trajectories,
    same status counts, fewer flops."""
    import bench
    from rtmpc_b200 import _lib
    from rtmpc_b200.rollout import RemoteLoop
    mpc, Z = bench.build_controller(extended=False)
    B, T = 3000, 100
    p = np.array([0.1 * (i % 10) for i in range(B)])
    out = []
    try:
        for factored in (1, 0):
            _lib.set_tuning(_lib.TUNE_CERT_FACTORED, factored)
            loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
            loop.reset()
            tr = loop.run(T, bench.REF, p_loss=p, seed=9, record=True).cpu().numpy()
            out.append((tr, loop.x_hat.cpu().numpy(), loop.stats.cpu().numpy()))
    finally:
        _lib.set_tuning(_lib.TUNE_CERT_FACTORED, -1)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2][:7], out[1][2][:7])          # statuses, interior-point iterations, steps, certifications
    assert out[0][2][7] < out[1][2][7]                            # algorithmic flops


@pytest.mark.parametrize("kind", ["tube", "extended"])
def test_certification_tiers_under_reference_jumps(kind):
    """The same equivalence where the rounding bounds matter: reference jumps that leave the feasible set nearly empty
    (multipliers of 1e6, steps along dependent rows, solves that end OPTIMAL_INACCURATE or infeasible) - closed loops with the
    tiered certification and with every certification forced through G' z agree bit for bit, status for status."""
    import bench
    from rtmpc_b200 import _lib
    from rtmpc_b200.rollout import RemoteLoop
    mpc, Z = bench.build_controller(extended=(kind == "extended"))
    B, T = 1024, 800
    jumps = [0.5, -0.8, 1.2, 0.0, 2.0, -1.5, 0.3, 1.0]
    r = np.zeros((T, 4))
    r[:, 0] = np.repeat(jumps, T // 8)
    p = np.array([0.1 * (i % 10) for i in range(B)])
    out = []
    try:
        for factored in (1, 0):
            _lib.set_tuning(_lib.TUNE_CERT_FACTORED, factored)
            loop = RemoteLoop(mpc, B, kind=kind, w_half=bench.HW, Z=Z)
            loop.reset()
            tr = loop.run(T, r, p_loss=p, seed=99, record=True).cpu().numpy()
            out.append((tr, loop.x_hat.cpu().numpy(), loop.alive.cpu().numpy(), loop.stats.cpu().numpy()))
    finally:
        _lib.set_tuning(_lib.TUNE_CERT_FACTORED, -1)
    assert np.array_equal(out[0][0], out[1][0], equal_nan=True) and np.array_equal(out[0][1], out[1][1], equal_nan=True)
    assert np.array_equal(out[0][2], out[1][2])
    assert np.array_equal(out[0][3][:7], out[1][3][:7])

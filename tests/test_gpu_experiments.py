"""Reporting layer: the batched Monte-Carlo experiment of Results/results_linear_system.py."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_linear_system_experiment_statistics():
    from rtmpc_b200.experiments import linear_system_experiment
    from rtmpc_b200.rollout import RemoteLoop
    s = H.load("sets_cp.npz")
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    tube, track, Z = H.make_tube_mpc(s), H.make_track_mpc(s), H.poly(s, "Z")
    probs, n_mc, T = [0.0, 0.5, 0.9], 4, 80
    res = linear_system_experiment(tube, track, Z, hw, prob_packet_loss=probs, n_mc=n_mc, T=T, seed=11)
    assert res.tracking_error_tube.shape == (3, 4) and res.tracking_error_track.shape == (3, 4)
    assert res.max_tube_violation < 1e-7                                   # the script's tube-containment check
    assert sorted(res.trajectories_tube) == probs and res.trajectories_tube[0.5].shape == (4, T + 1)
    # tracking error as the script computes it (:291) from the stored trajectory of run min(5, n_mc - 1) = 3
    for i, p in enumerate(probs):
        x = res.trajectories_tube[p]
        e = np.sqrt(((x[0, :-1] - 0.5) ** 2 + (x[1:, :-1] ** 2).sum(0)).sum()) / T
        assert abs(e - res.tracking_error_tube[i, 3]) <= 1e-12
    # both controllers saw the same realisation; an independent step-by-step run reproduces instance (p = 0.5, run 2)
    loop = RemoteLoop(tube, 1, kind="tube", w_half=hw, Z=Z)
    loop.reset(np.zeros((1, 4)))
    loop.run(T, np.array([0.5, 0, 0, 0]), p_loss=np.array([0.5]), seed=11, id_offset=1 * n_mc + 2, fused=False)
    # (to rounding: the experiment's rollout carries its working-set inverse from step to step, RTMPC_TUNE_ROLLOUT_CARRY)
    assert abs(loop.tracking_error(T).item() - res.tracking_error_tube[1, 2]) <= 1e-10
    # R-MPC: NaN exactly where the controller became infeasible
    assert np.array_equal(np.isnan(res.tracking_error_track).sum(axis=1), res.is_track_infeasible)
    again = linear_system_experiment(tube, track, Z, hw, prob_packet_loss=probs, n_mc=n_mc, T=T, seed=11)
    assert np.array_equal(again.tracking_error_tube, res.tracking_error_tube)
    assert "Failed executions of Remote MPC" in res.summary()


def test_extended_arm_and_nonlinear_plant():
    """results_linear_system_with_extendedMPC.py (ERT-MPC arm) and the nonlinear loop structure (analytic cartpole plant)."""
    from rtmpc_b200.experiments import linear_system_experiment
    s = H.load("sets_cp.npz")
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    tube, ext, Z = H.make_tube_mpc(s), H.make_tube_mpc(s, extended=True), H.poly(s, "Z")
    res = linear_system_experiment(tube, None, Z, hw, prob_packet_loss=[0.0, 0.6], n_mc=3, T=80, seed=5, ext_mpc=ext)
    assert res.tracking_error_ext.shape == (2, 3) and np.all(np.isfinite(res.tracking_error_ext))
    assert res.max_tube_violation < 1e-7
    assert np.all(np.isnan(res.tracking_error_track)) and res.is_track_infeasible.sum() == 0      # R-MPC arm skipped
    assert "ERT-MPC" in res.summary()
    nl = linear_system_experiment(tube, None, None, hw, prob_packet_loss=[0.0, 0.6], n_mc=3, T=80, seed=5, plant="cartpole")
    assert np.all(np.isfinite(nl.tracking_error_tube)) and nl.tracking_error_tube.max() < 0.2


def test_disturbance_set_estimation_matches_oracle():
    """Results/estimate_W_for_Cartpole.py on the analytic plant: the batched sweep against the oracle's loop, and the
    properties the script relies on (every run is stabilised; the model error vanishes with the state)."""
    from oracle import ref_loop as rl
    from rtmpc_b200.experiments import estimate_disturbance_set
    s = H.load("sets_cp.npz")
    A, B, K = s["A"], s["B"], s["K"]
    iv, w, xf = estimate_disturbance_set(A, B, K, n_runs=100, n_steps=400, plant="cartpole")
    assert w.shape == (100, 400, 4) and iv.shape == (4, 2)
    assert np.abs(xf).max() < 1e-3                                   # script :117-118 "System not stabilized"
    rng = np.random.default_rng(456)
    x0 = rng.uniform(-np.array([1.0, 0.5, 0.3, 0.5]), np.array([1.0, 0.5, 0.3, 0.5]), size=(100, 4))
    wo, xfo = rl.estimate_model_error(x0[:12], K, A - B @ K, 400)
    assert np.abs(w[:12] - wo).max() <= 1e-10 * max(1.0, np.abs(wo).max())
    assert np.abs(xf[:12] - xfo).max() <= 1e-10
    assert np.all(iv[:, 0] < 0) and np.all(iv[:, 1] > 0)
    assert np.abs(w[:, -1]).max() < 1e-6 * max(1e-30, np.abs(w[:, 0]).max()) + 1e-9     # linearisation exact at the origin
    # a second call with explicit initial conditions and another length
    iv2, w2, _ = estimate_disturbance_set(A, B, K, n_steps=50, x0=x0[:7], plant="cartpole")
    assert np.array_equal(w2, w[:7, :50])


def test_bullet_like_plant_reproduces_the_reference_disturbance_set():
    """SURVEY 8f rank 3.  The reference hard-codes hw = (1e-4, 2.7e-3, 3e-4, 4.3e-2) (Results/results_linear_system.py:76-91),
    the output of Results/estimate_W_for_Cartpole.py on PyBullet.  The analytic ODE with what Bullet does to the
    reference's URDF (pole inertia recomputed from the collision box, link damping 0.04) must land on those constants
    (asserted within 10 %, well inside the verdict's factor 2; measured 0.5 %), the plant with the linear model's own
    parameters must not (its model error is 3-45x smaller)."""
    from oracle import ref_loop as rl
    from rtmpc_b200.experiments import estimate_disturbance_set
    s = H.load("sets_cp.npz")
    A, B, K = s["A"], s["B"], s["K"]
    hw_ref = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    iv, w, xf = estimate_disturbance_set(A, B, K, n_runs=100, n_steps=400)          # default plant: cartpole_bullet
    hw = np.abs(iv).max(axis=1)
    assert np.all(np.abs(hw / hw_ref - 1.0) < 0.10), hw / hw_ref
    assert np.abs(xf).max() < 1e-3
    rng = np.random.default_rng(456)
    x0 = rng.uniform(-np.array([1.0, 0.5, 0.3, 0.5]), np.array([1.0, 0.5, 0.3, 0.5]), size=(100, 4))
    wo, _ = rl.estimate_model_error(x0[:6], K, A - B @ K, 400, I=rl.BULLET_POLE_INERTIA, damping=rl.BULLET_LINK_DAMPING)
    assert np.abs(w[:6] - wo).max() <= 1e-10 * max(1.0, np.abs(wo).max())
    iv0, _, _ = estimate_disturbance_set(A, B, K, n_runs=100, n_steps=400, plant="cartpole")
    r0 = np.abs(iv0).max(axis=1) / hw_ref
    assert r0[1] < 0.1 and r0[3] < 0.1


def test_config4_on_the_bullet_like_plant_stays_in_the_tube():
    """BASELINE configs[3] with the plant the reference's figure was made on: RT-MPC designed for W = box(hw) keeps
    x - x_nom in Z on the nonlinear plant whose model error defines that W, at every loss rate; mean tracking error in
    the range of figures/TrackingErrorNonlinear.png (RT-MPC 0.021 at p = 0 ... 0.033 at p = 0.7, read off the plot)."""
    from rtmpc_b200.experiments import linear_system_experiment
    s = H.load("sets_cp.npz")
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    tube, Z = H.make_tube_mpc(s), H.poly(s, "Z")
    res = linear_system_experiment(tube, None, Z, hw, n_mc=20, T=250, seed=124, plant="cartpole_bullet")
    assert res.max_tube_violation < 1e-7
    m = res.tracking_error_tube.mean(axis=1)
    assert np.all(m > 0.015) and np.all(m < 0.045), m

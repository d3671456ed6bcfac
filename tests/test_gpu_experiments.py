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
    assert abs(loop.tracking_error(T).item() - res.tracking_error_tube[1, 2]) == 0.0
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

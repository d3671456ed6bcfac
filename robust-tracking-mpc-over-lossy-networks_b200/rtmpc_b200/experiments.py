"""Reporting layer of the reference's Monte-Carlo scripts, on batched rollouts.

``Results/results_linear_system.py:143-358`` runs, for 10 packet-loss probabilities x ``N_MC`` Monte-Carlo runs x
250 control steps, the remote tube MPC (RT-MPC) and Pezzutto's remote MPC (R-MPC) on the SAME network and
disturbance realisation, and reports per loss probability the tracking error ``1/T sqrt(sum_t ||x_t - ref_t||^2)``
(``:291``), the number of runs in which R-MPC became infeasible (``:268-271,324-326``), one stored trajectory
(``:298-301``), tube-containment violations (``:257-259``) and the solve-time quantiles (``:305-315``).

Here every (probability, run) pair is one closed-loop instance of two :class:`~rtmpc_b200.rollout.RemoteLoop` s that
share seed and instance ids (the device RNG is keyed by (seed, id, t): both controllers see identical theta_t, gamma_t,
w_t, as in the script), and each controller's whole experiment is one persistent launch.  With an initialised
``torch.distributed`` process group the instances are sharded over the ranks and the statistics all-gathered.
"""
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import distributed as D
from .rollout import RemoteLoop


@dataclass
class ExperimentResult:
    prob_packet_loss: np.ndarray            # [n_p]
    tracking_error_tube: np.ndarray         # [n_p, n_mc]
    tracking_error_track: np.ndarray        # [n_p, n_mc]  NaN where R-MPC became infeasible (script :293-296)
    is_track_infeasible: np.ndarray         # [n_p]        runs in which R-MPC returned U_t = None (script :268-271)
    max_tube_violation: float               # max_t max_i (Hz (x - x_nom) - hz)_i over all runs: <= 0 means inside the tube
    tracking_error_ext: np.ndarray = None   # [n_p, n_mc]  ERT-MPC (results_linear_system_with_extendedMPC.py), if requested
    trajectories_tube: dict = field(default_factory=dict)    # p -> x[nx, T+1] of run min(5, n_mc-1)   (script :298-301)
    trajectories_track: dict = field(default_factory=dict)
    trajectories_ext: dict = field(default_factory=dict)
    solve_ms_tube: float = 0.0              # amortised time per determine_packet call of RT-MPC, in ms
    solve_ms_track: float = 0.0
    solves: int = 0

    def summary(self):
        """Text in the spirit of the script's prints (``:305-326``)."""
        lines = [f"closed-loop instances per controller: {self.tracking_error_tube.size}, solves per controller: {self.solves}",
                 f"amortised time per solve [ms]: RT-MPC {self.solve_ms_tube:.6f}, R-MPC {self.solve_ms_track:.6f}",
                 f"max tube violation (<= 0: x - x_nom in Z at every step): {self.max_tube_violation:.3e}",
                 "Failed executions of Remote MPC:", str(self.is_track_infeasible.reshape(-1, 1)),
                 " p     RT-MPC median / mean        R-MPC median / mean (feasible runs)" +
                 ("        ERT-MPC median / mean" if self.tracking_error_ext is not None else "")]
        for i, p in enumerate(self.prob_packet_loss):
            tt, tr = self.tracking_error_tube[i], self.tracking_error_track[i]
            ok = ~np.isnan(tr)
            line = f"{p:4.2f}   {np.median(tt):.5f} / {tt.mean():.5f}        " + \
                (f"{np.median(tr[ok]):.5f} / {tr[ok].mean():.5f}" if ok.any() else "   -    /    -   ")
            if self.tracking_error_ext is not None:
                te = self.tracking_error_ext[i]
                line += f"                        {np.median(te):.5f} / {te.mean():.5f}"
            lines.append(line)
        return "\n".join(lines)


def linear_system_experiment(tube_mpc, track_mpc, Z, w_half, prob_packet_loss=None, n_mc=20, T=250, ref=0.5, x0=None,
                             seed=679, store_run=None, plant="linear", ext_mpc=None):
    """Batched ``Results/results_linear_system.py``.  ``tube_mpc``: TubeTrackingMPC with its problem generated;
    ``track_mpc``: TrackingMPC with its problem generated, or None to skip R-MPC; ``Z``: the tube (for the containment
    check); ``w_half``: half-widths of the disturbance box; ``ref``: target of the first state (full-state target
    ``(ref, 0, ..)`` as in the script, ``:240``); ``ext_mpc``: an ExtendedTubeTrackingMPC to add the ERT-MPC arm of
    ``results_linear_system_with_extendedMPC.py`` (robust estimator, x_nom_0 in the packet, gamma_{t-1}-switched QP);
    ``plant='cartpole'`` replaces the linear plant + disturbance by the analytic cartpole ODE
    (``results_nonlinear_system*.py`` without PyBullet)."""
    probs = np.arange(10) / 10.0 if prob_packet_loss is None else np.asarray(prob_packet_loss, float)
    n_p = len(probs)
    total = n_p * n_mc
    rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if rank or (torch.distributed.is_available() and
                                                           torch.distributed.is_initialized()) else 1
    off, cnt = D.shard(total, rank, world)
    ids = np.arange(off, off + cnt)                       # instance id = i_prob * n_mc + l_mc
    nx = tube_mpc._nx
    p_loss = probs[ids // n_mc]
    ref_vec = np.zeros(nx)
    ref_vec[0] = ref
    x0v = np.zeros((cnt, nx)) if x0 is None else np.broadcast_to(np.asarray(x0, float).reshape(-1, nx), (cnt, nx))
    store_run = min(5, n_mc - 1) if store_run is None else store_run
    out = {}
    for name, mpc, kind in (("tube", tube_mpc, "tube"), ("track", track_mpc, "track"), ("ext", ext_mpc, "extended")):
        if mpc is None or cnt == 0:
            continue
        loop = RemoteLoop(mpc, cnt, kind=kind, plant=plant, w_half=w_half, Z=Z if kind != "track" else None)
        loop.reset(x0v)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        traj = loop.run(T, ref_vec, p_loss=p_loss, seed=seed, id_offset=off, record=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = dict(err=loop.tracking_error(T), alive=loop.alive.clone(), tube=loop.tube_max.clone(), traj=traj,
                         ms=1e3 * dt / max(cnt * T, 1))
    dev = torch.device("cuda", torch.cuda.current_device())

    def gathered(key, name, fill):
        local = out[name][key] if name in out else torch.full((cnt,), fill, device=dev, dtype=torch.float64)
        return D.all_gather_instances(local.to(torch.float64), total).cpu().numpy()

    err_tube = gathered("err", "tube", np.nan).reshape(n_p, n_mc)
    res = ExperimentResult(prob_packet_loss=probs, tracking_error_tube=err_tube,
                           tracking_error_track=np.full((n_p, n_mc), np.nan), is_track_infeasible=np.zeros(n_p, int),
                           max_tube_violation=float(gathered("tube", "tube", -np.inf).max()), solves=total * T)
    if ext_mpc is not None:
        res.tracking_error_ext = gathered("err", "ext", np.nan).reshape(n_p, n_mc)
        res.max_tube_violation = max(res.max_tube_violation, float(gathered("tube", "ext", -np.inf).max()))
    if track_mpc is not None:
        err_track = gathered("err", "track", np.nan).reshape(n_p, n_mc)
        alive = gathered("alive", "track", 1.0).reshape(n_p, n_mc)
        res.tracking_error_track = np.where(alive > 0.5, err_track, np.nan)
        res.is_track_infeasible = (alive < 0.5).sum(axis=1)
    for name, store in (("tube", res.trajectories_tube), ("track", res.trajectories_track), ("ext", res.trajectories_ext)):
        if name not in out:
            continue
        tr = out[name]["traj"]
        for i, p in enumerate(probs):
            gid = i * n_mc + store_run
            if off <= gid < off + cnt:                        # stored on the rank that owns the run
                store[float(p)] = tr[gid - off].T.cpu().numpy()
    ms = torch.tensor([out.get("tube", {}).get("ms", 0.0), out.get("track", {}).get("ms", 0.0)], device=dev, dtype=torch.float64)
    ms = D.all_reduce_max(ms).cpu().numpy()
    res.solve_ms_tube, res.solve_ms_track = float(ms[0]), float(ms[1])
    return res


CARTPOLE_PARAMS = dict(M=1.0, m=0.1, I=0.001, g=9.8, l=0.5)       # Results/estimate_W_for_Cartpole.py:32-38, cartpole.urdf


def estimate_disturbance_set(A, B, K, n_runs=100, n_steps=400, x0_half=(1.0, 0.5, 0.3, 0.5), seed=456, outlier=0.025,
                             physics_timestep=1.0 / 500.0, Th=0.02, x0=None, plant="cartpole_bullet"):
    """Batched ``Results/estimate_W_for_Cartpole.py:79-127``: ``n_runs`` random initial conditions in the box
    ``+-x0_half`` (script ``:67-75``), each stabilised for ``n_steps`` control periods (the script's 4000 physics steps)
    by the zero-order-hold LQR law ``u = -K x`` on the nonlinear cartpole (analytic ODE instead of PyBullet, one
    launch for all runs); the model error ``w(k) = x(k) - (A - BK) x(k-1)`` (``:104-110``) is collected and the
    interval that discards the ``outlier`` share of largest absolute values per component is returned (``:123-127``).

    ``plant``: 'cartpole_bullet' (default) simulates what the script's PyBullet call simulates - pole inertia recomputed
    from the collision box and default link damping 0.04 (see ``rollout.CART_PARAMS_BULLET``); its intervals reproduce the
    constants the reference hard-codes, ``hw = (1e-4, 2.7e-3, 3e-4, 4.3e-2)`` (``Results/results_linear_system.py:76-91``),
    to within 0.5 %.  'cartpole' uses the parameters of the linear model itself (model error = linearisation and
    discretisation only).

    Returns ``(intervals [nx, 2], w [n_runs, n_steps, nx], x_final [n_runs, nx])``; ``max |intervals|`` per component
    is the half-width vector ``hw`` the experiment scripts build ``W`` from."""
    import ctypes as C
    from . import _lib
    A, B, K = np.asarray(A, float), np.asarray(B, float), np.atleast_2d(np.asarray(K, float))
    if A.shape != (4, 4) or K.shape != (1, 4):
        raise ValueError("the cartpole plant has nx = 4, nu = 1")
    L = _lib.lib()
    _lib.require_cuda()
    if x0 is None:
        rng = np.random.default_rng(seed)
        h = np.asarray(x0_half, float)
        x0 = rng.uniform(-h, h, size=(n_runs, 4))                      # the script draws the four components in this order
    x0 = _lib.f64(np.asarray(x0, float).reshape(-1, 4))
    n_runs = x0.shape[0]
    Acl = _lib.f64(A - B @ K)
    Kf = _lib.f64(K.reshape(-1))
    from .rollout import PLANTS
    cp = PLANTS[plant][1]
    cart = (C.c_double * 8)(cp[0], cp[1], cp[2], cp[3], cp[4], physics_timestep, round(Th / physics_timestep), cp[7])
    w = np.empty((n_runs, n_steps, 4))
    xf = np.empty((n_runs, 4))
    _lib.check(L.rtmpc_model_error_sweep_host(cart, n_runs, n_steps, _lib.ptr(x0), _lib.ptr(Kf), _lib.ptr(Acl), _lib.ptr(w),
                                              _lib.ptr(xf)), "rtmpc_model_error_sweep_host")
    flat = w.reshape(-1, 4)
    iv = np.stack([np.quantile(flat, outlier / 2, axis=0), np.quantile(flat, 1.0 - outlier / 2, axis=0)], axis=1)
    return iv, w, xf

"""Generate the golden fixtures under tests/golden/ with the CPU oracle.

The reference cannot run in this image (cvxpy / clarabel / polytope / control are absent) and it
ships no golden vectors of its own, so the fixtures are produced by ``oracle/`` -- the restatement
that is pinned by the Darup k_star known answers, the scripts' run-time invariants and a scipy
cross-check (tests/test_oracle_*.py).  Run from the repo root:

    python tests/golden/make_golden.py            # everything (~15 min, mostly the two 9-D MOAS)

Outputs (npz):
  sets_di.npz / sets_cp.npz        K, P, tube Z, tightened Xc/Uc, terminal sets (tube + plain tracking), Z-W
  loop_di_tube.npz                 config 1: Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py as shipped
  loop_cp_tube.npz                 config 2: results_linear_system.py loop, RT-MPC and R-MPC, 4 loss rates
  loop_cp_ext.npz                  config 3: results_linear_system_with_extendedMPC.py loop, ERT-MPC
  qp_di_regulators.npz             RegulatorMPC / TubeRegulatorMPC solves on the double integrator
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loop as rl          # noqa: E402
from oracle import ref_numerics as rn      # noqa: E402
from oracle import ref_qp as rq            # noqa: E402
from oracle import ref_sets as rs          # noqa: E402
from oracle import ref_setup as su         # noqa: E402
from oracle.ref_polytope import Polytope   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _pk(prefix, poly):
    return {prefix + "_A": poly.A, prefix + "_b": poly.b}


def make_sets(name, cfg, extended):
    t0 = time.time()
    d = su.tube_tracking_setup(**cfg, fixed_initial_state=True, extended=extended)
    trk = su.tracking_setup(**cfg)
    out = dict(A=cfg["A"], B=cfg["B"], Q=cfg["Q"], R=cfg["R"], N=cfg["N"], K=d["K"], P=d["P"],
               t_star_tube=d["t_star"], t_star_track=trk["t_star"])
    for k, p in (("X", cfg["X"]), ("U", cfg["U"]), ("W", cfg["W"]), ("Z", d["Z"]), ("Xc", d["Xc"]), ("Uc", d["Uc"]),
                 ("Xf", d["Xf"]), ("Xf_track", trk["Xf"])):
        out.update(_pk(k, p))
    if extended:
        out.update(_pk("ZmW", d["ZmW"]))
    np.savez_compressed(os.path.join(OUT, f"sets_{name}.npz"), **out)
    print(f"sets_{name}: Z {d['Z'].A.shape}, Xf {d['Xf'].A.shape} (t*={d['t_star']}), "
          f"Xf_track {trk['Xf'].A.shape} (t*={trk['t_star']}), {time.time() - t0:.0f} s", flush=True)
    return d, trk


def run_tube_loop(d, x0, refs, theta, gamma, w, K_plant=None):
    """Loop body of results_linear_system.py:209-255 / the lossy-network example :118-163 for the
    remote tube MPC.  Returns per-step records."""
    A, B, K, N = d["A"], d["B"], d["K"], d["N"]
    Kp = K if K_plant is None else K_plant
    T = len(theta)
    nx = A.shape[0]
    est = rl.Estimator(A, B, K, x0, N)
    act = rl.ConsistentActuator(A, B, K, Kp, x0)
    x = np.array(x0, float)
    xhat = est.get_estimate()
    rec = dict(x=np.zeros((T + 1, nx)), x_nom=np.zeros((T + 1, nx)), x_hat=np.zeros((T + 1, nx)),
               xhat_in=np.zeros((T, nx)), U_t=np.zeros((T, N + 1, B.shape[1])), z=[], u=np.zeros((T, B.shape[1])),
               Theta=np.zeros(T, int), s_t=np.zeros(T, int), q_t=np.zeros(T, int), polished=np.zeros(T, int))
    rec["x"][0] = x
    rec["x_nom"][0] = x0
    rec["x_hat"][0] = xhat
    for t in range(T):
        q_t = est.get_qt()
        (xn, un, xb, ub), res = rq.solve_param(d["qp"], xhat.copy(), refs[t].copy())
        assert res.status == "optimal", (t, res.status)
        pkt = rl.encapsulate_controller_packet(un, xb, ub, K, q_t)
        est.store_sent_control_sequence(pkt["U_t"])
        u, ppkt = act.process_packet(pkt, x, int(theta[t]))
        x = A @ x + B @ u + w[t]
        est.update_estimate(ppkt, int(gamma[t]))
        rec["xhat_in"][t] = xhat
        rec["U_t"][t] = pkt["U_t"].T
        rec["z"].append(res.z[:d["qp"].nz])
        rec["u"][t] = u
        rec["Theta"][t], rec["s_t"][t], rec["q_t"][t] = act.Theta_t, act.s_t, q_t
        rec["polished"][t] = res.polished
        xhat = est.get_estimate()
        rec["x"][t + 1], rec["x_nom"][t + 1], rec["x_hat"][t + 1] = x, act.get_x_nom(), xhat
    rec["z"] = np.array(rec["z"])
    return rec


def run_track_loop(trk, x0, refs, theta, gamma, w):
    """R-MPC (Pezzutto) part of results_linear_system.py:261-288: stops at the first infeasible solve."""
    A, B, K, N = trk["A"], trk["B"], trk["K"], trk["N"]
    T = len(theta)
    nx = A.shape[0]
    est = rl.Estimator(A, B, K, x0, N)
    act = rl.SmartActuator(K)
    x = np.array(x0, float)
    xhat = est.get_estimate()
    rec = dict(x=np.full((T + 1, nx), np.nan), x_hat=np.full((T + 1, nx), np.nan), xhat_in=np.full((T, nx), np.nan),
               U_t=np.full((T, N + 1, B.shape[1]), np.nan), feasible=np.zeros(T, int), Theta=np.zeros(T, int),
               s_t=np.zeros(T, int), q_t=np.zeros(T, int))
    rec["x"][0] = x
    rec["x_hat"][0] = xhat
    for t in range(T):
        q_t = est.get_qt()
        rec["xhat_in"][t] = xhat
        (xn, un, xb, ub), res = rq.solve_param(trk["qp"], xhat.copy(), refs[t].copy())
        if res.status == "infeasible":
            break
        rec["feasible"][t] = 1
        pkt = rl.encapsulate_controller_packet(un, xb, ub, K, q_t)
        est.store_sent_control_sequence(pkt["U_t"])
        u, ppkt = act.process_packet(pkt, x, int(theta[t]))
        x = A @ x + B @ u + w[t]
        est.update_estimate(ppkt, int(gamma[t]))
        rec["U_t"][t] = pkt["U_t"].T
        rec["Theta"][t], rec["s_t"][t], rec["q_t"][t] = act.Theta_t, act.s_t, q_t
        xhat = est.get_estimate()
        rec["x"][t + 1], rec["x_hat"][t + 1] = x, xhat
    return rec


def run_ext_loop(d, x0, refs, theta, gamma, w):
    """Extended loop of results_linear_system_with_extendedMPC.py:247-378: the solve at step t sees
    gamma_{t-1} (gamma of step 0 is 1), the packet carries x_nom_0, RobustEstimator."""
    A, B, K, N = d["A"], d["B"], d["K"], d["N"]
    T = len(theta)
    nx = A.shape[0]
    est = rl.RobustEstimator(A, B, K, K, x0, N)
    act = rl.ConsistentActuator(A, B, K, K, x0, is_extended_MPC_used=True)
    x = np.array(x0, float)
    xhat = est.get_estimate()
    rec = dict(x=np.zeros((T + 1, nx)), x_nom=np.zeros((T + 1, nx)), x_hat=np.zeros((T + 1, nx)),
               xhat_in=np.zeros((T, nx)), U_t=np.zeros((T, N + 1, B.shape[1])), x_nom0=np.zeros((T, nx)),
               mode=np.zeros(T, int), Theta=np.zeros(T, int), s_t=np.zeros(T, int), q_t=np.zeros(T, int))
    rec["x"][0] = x
    rec["x_nom"][0] = x0
    rec["x_hat"][0] = xhat
    gamma_prev = 1
    for t in range(T):
        q_t = est.get_qt()
        qp = d["qp_recv"] if gamma_prev == 1 else d["qp"]
        (xn, un, xb, ub), res = rq.solve_param(qp, xhat.copy(), refs[t].copy())
        assert res.status == "optimal", (t, res.status)
        pkt = rl.encapsulate_controller_packet(un, xb, ub, K, q_t, x_nom_0=xn[:, 0])
        est.store_sent_control_sequence(pkt["U_t"])
        est.store_current_optimal_inital_nominal_plant_states(xn[:, 0])
        u, ppkt = act.process_packet(pkt, x, int(theta[t]))
        x = A @ x + B @ u + w[t]
        est.update_estimate(ppkt, int(gamma[t]))
        rec["xhat_in"][t] = xhat
        rec["U_t"][t] = pkt["U_t"].T
        rec["x_nom0"][t] = xn[:, 0]
        rec["mode"][t] = gamma_prev
        rec["Theta"][t], rec["s_t"][t], rec["q_t"][t] = act.Theta_t, act.s_t, q_t
        gamma_prev = int(gamma[t])
        xhat = est.get_estimate()
        rec["x"][t + 1], rec["x_nom"][t + 1], rec["x_hat"][t + 1] = x, act.get_x_nom(), xhat
    return rec


def draws(seed, T, p, hw):
    """theta, gamma (step 0 forced lossless) and w ~ U(-hw, hw), one generator per quantity like the
    reference's scripts (results_linear_system.py:21-23,211-233)."""
    rt, rg, rw = (np.random.default_rng(seed + k) for k in (124, 347, 679))
    theta = np.ones(T, int)
    gamma = np.ones(T, int)
    w = np.zeros((T, len(hw)))
    for t in range(T):
        if t > 0:
            theta[t] = 0 if rt.uniform() < p else 1
            gamma[t] = 0 if rg.uniform() < p else 1
        w[t] = [rw.uniform(-h, h) for h in hw]
    return theta, gamma, w


def stack(recs):
    keys = recs[0].keys()
    return {k: np.stack([r[k] for r in recs]) for k in keys}


def main():
    which = sys.argv[1:] or ["di", "cp"]
    if "di" in which:
        cfg = su.double_integrator()
        d, trk = make_sets("di", cfg, extended=True)
        # config 1, exactly as shipped: seeds 1 / 347 / 124, p = 0.7, T = 120, x0 = (1,2)
        T = 120
        rng_w, rng_g, rng_t = np.random.default_rng(1), np.random.default_rng(347), np.random.default_rng(124)
        theta = np.ones(T, int)
        gamma = np.ones(T, int)
        w = np.zeros((T, 2))
        for t in range(T):
            if t > 0:
                theta[t] = 0 if rng_t.uniform() < 0.7 else 1
                gamma[t] = 0 if rng_g.uniform() < 0.7 else 1
            w[t] = rng_w.uniform(-0.1, 0.1, 2)
        refs = np.zeros((T, 2))
        refs[0:30, 0], refs[30:60, 0], refs[60:90, 0], refs[90:120, 0] = 5, -9, 9, 4
        rec = run_tube_loop(d, np.array([1.0, 2.0]), refs, theta, gamma, w)
        # invariants of the example (:165-184)
        Z = d["Z"]
        assert all((rec["x"][t] - rec["x_nom"][t]) in Z for t in range(T)), "tube invariant violated"
        assert all((rec["x"][t] - rec["x_hat"][t]) in Z for t in range(T) if rec["Theta"][t] == 1)
        np.savez_compressed(os.path.join(OUT, "loop_di_tube.npz"), theta=theta, gamma=gamma, w=w, refs=refs, **rec)
        print("loop_di_tube: polished", int(rec["polished"].sum()), "of", T, flush=True)
        # extended variant on the double integrator (small, fast parity case for G2)
        theta, gamma, w = draws(11, 60, 0.5, [0.1, 0.1])
        refs = np.zeros((60, 2))
        refs[:, 0] = 4.0
        rece = run_ext_loop(d, np.array([1.0, 2.0]), refs, theta, gamma, w)
        np.savez_compressed(os.path.join(OUT, "loop_di_ext.npz"), theta=theta, gamma=gamma, w=w, refs=refs, **rece)
        # Pezzutto R-MPC on the double integrator, no disturbance: estimate exact when Theta = 1
        theta, gamma, _ = draws(5, 60, 0.5, [0.1, 0.1])
        rect = run_track_loop(trk, np.array([1.0, 2.0]), refs, theta, gamma, np.zeros((60, 2)))
        ok = rect["Theta"] == 1
        assert np.abs(rect["x"][:-1][ok] - rect["x_hat"][:-1][ok]).max() == 0.0
        np.savez_compressed(os.path.join(OUT, "loop_di_track.npz"), theta=theta, gamma=gamma, refs=refs, **rect)
        # regulators
        reg = rq.build_regulator(cfg["A"], cfg["B"], cfg["Q"], cfg["R"], cfg["N"], cfg["X"], cfg["U"])
        tr = su.tube_regulator_setup(**cfg)
        rng = np.random.default_rng(3)
        xs = rng.uniform(-1, 1, (40, 2)) * np.array([6.0, 1.5])
        zr = np.full((40, reg.nz), np.nan)
        zt = np.full((40, tr["qp"].nz), np.nan)
        pol = np.zeros((40, 2), int)
        for i, x in enumerate(xs):
            sol, r1 = rq.solve_param(reg, x.copy())
            if r1.status == "optimal":
                zr[i] = r1.z
            sol, r2 = rq.solve_param(tr["qp"], x.copy())
            if r2.status == "optimal":
                zt[i] = r2.z
            pol[i] = [r1.polished, r2.polished]
        out = dict(xs=xs, z_reg=zr, z_tube=zt, P=tr["P"], K=tr["K"], polished=pol)
        for k in ("Z", "Xc", "Uc", "Xf"):
            out.update(_pk(k + "_mayne", tr[k]))
        np.savez_compressed(os.path.join(OUT, "qp_di_regulators.npz"), **out)
        print("qp_di_regulators: feasible", int(np.isfinite(zr[:, 0]).sum()), int(np.isfinite(zt[:, 0]).sum()), flush=True)
    if "cp" in which:
        cfg = su.linear_cartpole()
        d, trk = make_sets("cp", cfg, extended=True)
        hw = [1e-4, 2.7e-3, 3e-4, 4.3e-2]
        T = 250
        refs = np.zeros((T, 4))
        refs[:, 0] = 0.5
        tube, track, drw = [], [], []
        for i, p in enumerate([0.0, 0.3, 0.6, 0.9]):
            theta, gamma, w = draws(1000 * i, T, p, hw)
            tube.append(run_tube_loop(d, np.zeros(4), refs, theta, gamma, w))
            track.append(run_track_loop(trk, np.zeros(4), refs, theta, gamma, w))
            drw.append(dict(theta=theta, gamma=gamma, w=w))
            Z = d["Z"]
            assert all((tube[-1]["x"][t] - tube[-1]["x_nom"][t]) in Z for t in range(T)), "tube invariant violated"
            print(f"loop_cp p={p}: tube polished {int(tube[-1]['polished'].sum())}/{T}, "
                  f"track feasible steps {int(track[-1]['feasible'].sum())}", flush=True)
        np.savez_compressed(os.path.join(OUT, "loop_cp_tube.npz"), refs=refs, p=np.array([0.0, 0.3, 0.6, 0.9]),
                            **stack(drw), **{"tube_" + k: v for k, v in stack(tube).items()},
                            **{"track_" + k: v for k, v in stack(track).items()})
        Te = 100
        ext, drw = [], []
        for i, p in enumerate([0.4, 0.8]):
            theta, gamma, w = draws(77 + i, Te, p, hw)
            ext.append(run_ext_loop(d, np.zeros(4), refs[:Te], theta, gamma, w))
            drw.append(dict(theta=theta, gamma=gamma, w=w))
            print(f"loop_cp_ext p={p} done", flush=True)
        np.savez_compressed(os.path.join(OUT, "loop_cp_ext.npz"), refs=refs[:Te], p=np.array([0.4, 0.8]), **stack(drw),
                            **stack(ext))


if __name__ == "__main__":
    main()

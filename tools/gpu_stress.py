"""Stress: long rollouts with reference jumps, many instances, every controller variant; counts solves that did not end
certified, hand-overs to the interior-point kernel, loops that died, and checks the tube invariant (run under gpurun)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
import helpers as H
from rtmpc_b200.rollout import RemoteLoop

JUMPS = [0.5, -0.8, 1.2, 0.0, 2.0, -1.5, 0.3, 1.0]


def run(label, mpc, kind, B, T, refs, w_half, Z, nx, plant="linear", p_max=0.9):
    loop = RemoteLoop(mpc, B, kind=kind, plant=plant, w_half=w_half, Z=Z)
    dev = loop.dev
    p_loss = torch.as_tensor(np.array([p_max / 9 * (i % 10) for i in range(B)]), device=dev)
    loop.reset()
    torch.cuda.synchronize(); t0 = time.time()
    loop.run(T, refs, p_loss=p_loss, seed=99)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = loop.stats.cpu().numpy()
    tube = loop.tube_max.max().item() if Z is not None else float("nan")
    print(f"{label}: B={B} T={T}: {dt*1e3:.0f} ms -> {B*T/dt/1e6:.1f} M solves/s; status[opt,max_iter,infeasible,inaccurate] {st[:4].tolist()} "
          f"ipm iterations {st[4]} steps/solve {st[5]/(B*T):.2f}; alive {int(loop.alive.sum().item())}/{B}; max tube {tube:.3e}", flush=True)


if os.environ.get("CARRY"):
    from rtmpc_b200 import _lib
    _lib.set_tuning(_lib.TUNE_ROLLOUT_CARRY, int(os.environ["CARRY"]))
    print("rollout carry mode", os.environ["CARRY"])
which = os.environ.get("STRESS", "cp,cp_const,cp_ext,cp_track,di").split(",")
s = H.load("sets_cp.npz")
T = 2000
r = np.zeros((T, 4)); r[:, 0] = np.repeat(JUMPS, T // 8)
if "cp" in which:
    mpc, Z = bench.build_controller()
    run("cartpole tube MPC, reference jumps", mpc, "tube", 8192, T, r, bench.HW, Z, 4)
if "cp_const" in which:
    mpc, Z = bench.build_controller()
    run("cartpole tube MPC, constant reference", mpc, "tube", 131072, 250, bench.REF.copy(), bench.HW, Z, 4)
if "cp_ext" in which:
    mpc, Z = bench.build_controller(extended=True)
    run("cartpole extended tube MPC, reference jumps", mpc, "extended", 8192, T, r, bench.HW, Z, 4)
if "cp_track" in which:
    run("cartpole R-MPC baseline (no tube: may die), reference jumps", H.make_track_mpc(s), "track", 8192, T, r, bench.HW, None, 4)
if "di" in which:
    # the shipped example (config 1) at batch scale: reference 5 / -9 / 9 / 4 every 30 steps, 70 % loss on both links
    d = H.load("sets_di.npz")
    mpc = H.make_tube_mpc(d)
    T1 = 1200
    r1 = np.zeros((T1, 2)); r1[:, 0] = np.tile(np.repeat([5.0, -9.0, 9.0, 4.0], 30), T1 // 120)
    run("double integrator tube MPC (example script), reference jumps", mpc, "tube", 16384, T1, r1, np.array([0.1, 0.1]),
        H.poly(d, "Z"), 2, p_max=0.7)

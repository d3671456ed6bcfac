"""Throughput of the other BASELINE.json configurations on one GPU (parity of each is covered by tests/):
C3 extended variant (two QPs), C4 analytic cartpole plant, C5 support-function sweep over 1e6 directions."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
import helpers as H
from rtmpc_b200 import sets as up
from rtmpc_b200.rollout import RemoteLoop

dev = torch.device("cuda")
T = 250
s = H.load("sets_cp.npz")
ref = torch.as_tensor(bench.REF.copy(), device=dev)

def timed(loop, B, **kw):
    p_loss = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
    best = 1e9
    for r in range(3):
        loop.reset()
        torch.cuda.synchronize(); t0 = time.time()
        loop.run(T, ref, p_loss=p_loss, seed=347 + r, **kw)
        torch.cuda.synchronize(); best = min(best, time.time() - t0)
    return best

B = int(os.environ.get("B3", "65536"))
mpc = H.make_tube_mpc(s, extended=True)
loop = RemoteLoop(mpc, B, kind="extended", w_half=bench.HW, Z=H.poly(s, "Z"))
dt = timed(loop, B)
st = loop.stats.cpu().numpy()
print(f"C3 extended tube MPC (two QPs, rows {mpc._prob.rows}): B={B} T={T}: {dt*1e3:.1f} ms -> {B*T/dt/1e6:.2f} M solves/s; "
      f"status {st[:4].tolist()} ipm its {st[4]} as steps/solve {st[5]/(B*T):.2f} max tube {loop.tube_max.max().item():.3e}", flush=True)
del loop

B = int(os.environ.get("B4", "32768"))
mpc = H.make_tube_mpc(s)
loop = RemoteLoop(mpc, B, kind="tube", plant="cartpole")
dt = timed(loop, B)
st = loop.stats.cpu().numpy()
err = loop.tracking_error(T).mean().item()
print(f"C4 analytic cartpole ODE plant (10 sub-steps of 1/500 s): B={B} T={T}: {dt*1e3:.1f} ms -> {B*T/dt/1e6:.2f} M solves/s; "
      f"status {st[:4].tolist()} mean tracking error {err:.4f}", flush=True)
del loop

M = int(os.environ.get("M5", "1000000"))
Z = H.poly(s, "Z")
rng = np.random.default_rng(1)
dirs = rng.normal(size=(M, 4))
h = up.support_batch(Z, dirs[:1000])
torch.cuda.synchronize(); t0 = time.time()
h = up.support_batch(Z, dirs)
torch.cuda.synchronize(); dt = time.time() - t0
print(f"C5 support sweep h_Z(a) over {M} directions (host buffers in/out, vertex set of Z {len(h)}): {dt*1e3:.1f} ms -> {M/dt/1e6:.1f} M directions/s")

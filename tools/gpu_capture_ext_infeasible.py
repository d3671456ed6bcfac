"""Capture the (x_hat, ref, gamma_{t-1}) of extended-variant solves reported infeasible along rollouts with reference
jumps (run under gpurun); checked on the CPU against the oracle's feasibility LP afterwards."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
from rtmpc_b200.rollout import RemoteLoop
mpc, Z = bench.build_controller(extended=True)
B, T = 2048, 2000
loop = RemoteLoop(mpc, B, kind="extended", w_half=bench.HW, Z=Z)
dev = loop.dev
p_loss = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
r = np.zeros((T, 4)); r[:, 0] = np.repeat([0.5, -0.8, 1.2, 0.0, 2.0, -1.5, 0.3, 1.0], T // 8)
ref_d = torch.as_tensor(r, device=dev)
loop.reset()
bad = []
alive_prev = np.ones(B, bool)
for t in range(T):
    xh = loop.x_hat.clone(); gl = loop.gamma_last.clone()
    loop.step(ref_d[t].expand(B, 4).contiguous(), p_loss=p_loss, seed=99)
    st = loop.status.cpu().numpy()
    idx = np.nonzero((st == 2) & alive_prev)[0]
    for i in idx:
        bad.append((t, int(i), int(gl[i].item()), *xh[i].cpu().numpy().tolist(), r[t, 0], r[max(t - 1, 0), 0]))
    alive_prev &= st != 2
bad = np.array(bad)
print("first-infeasible events", len(bad), "alive", int(loop.alive.sum().item()))
np.save("gpurun_out/ext_infeasible.npy", bad)

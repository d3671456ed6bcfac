"""Capture (x_hat, ref) of solves that did not end certified along rollouts with reference jumps (run under gpurun)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
from rtmpc_b200.rollout import RemoteLoop
from rtmpc_b200.qp import BatchedQP
mpc, Z = bench.build_controller()
B, T = 2048, 2000
loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
dev = loop.dev
p_loss = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
r = np.zeros((T, 4)); r[:, 0] = np.repeat([0.5, -0.8, 1.2, 0.0, 2.0, -1.5, 0.3, 1.0], T // 8)
ref_d = torch.as_tensor(r, device=dev)
loop.reset()
bad = []
for t in range(T):
    xh = loop.x_hat.clone()
    wm = loop.warm.clone()
    rd = ref_d[t].expand(B, 4).contiguous()
    loop.step(rd, p_loss=p_loss, seed=99)
    st = loop.status.cpu().numpy(); it = loop.iters.cpu().numpy()
    idx = np.nonzero((st != 0) | ((it & 0xFFF) > 0))[0]
    for i in idx[:50]:
        bad.append((t, int(i), int(st[i]), int(it[i] & 0xFFF), int((it[i] >> 12) & 0xFFF), *xh[i].cpu().numpy().tolist(), r[t, 0], *wm[i].cpu().numpy().tolist()))
bad = np.array(bad)
print("captured", len(bad), "status counts", np.bincount(bad[:, 2].astype(int), minlength=4) if len(bad) else None)
print("status hist overall", loop.status_count.cpu().numpy().tolist(), "alive", int(loop.alive.sum().item()))
np.save("gpurun_out/nonoptimal.npy", bad)

"""Run the benchmark rollout once and dump the (x_hat, ref) of every solve whose status is not OPTIMAL."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
import bench
from rtmpc_b200 import _lib
from rtmpc_b200.rollout import RemoteLoop
mpc, Z = bench.build_controller()
B, T = 4096, 250
loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
dev = loop.dev
p = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
ref = torch.as_tensor(np.tile(bench.REF, (B, 1)), device=dev)
bad = []
hist = np.zeros((T, 64), np.int64)
for seed in (679, 680, 681, 1679):
    loop.reset()
    for t in range(T):
        xh = loop.x_hat.clone()
        loop.step(ref, p_loss=p, seed=seed)
        st = loop.status.cpu().numpy()
        it = loop.iters.cpu().numpy()
        hist[t] += np.bincount(np.minimum(it, 63), minlength=64)
        for b in np.nonzero(st != 0)[0]:
            bad.append(np.r_[seed, t, b, st[b], it[b], xh[b].cpu().numpy()])
    print("seed", seed, "status", loop.status_count.cpu().numpy(), "tube max", loop.tube_max.max().item(), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "failures.npy"), np.array(bad))
np.save(os.path.join(ROOT, "gpurun_out", "iter_hist.npy"), hist)
print("failures:", len(bad)); print(np.array(bad)[:10] if bad else "")
print("mean iters per step (first 80):", np.round((hist * np.arange(64)).sum(1)[:80] / hist.sum(1)[:80], 1))

"""Stress: long rollouts and many instances; every solve must be certified (status 0) and the tube invariant must hold."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
from rtmpc_b200.rollout import RemoteLoop
mpc, Z = bench.build_controller()
for B, T, refs in ((8192, 2000, "steps"), (131072, 250, "const")):
    loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
    dev = loop.dev
    p_loss = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
    if refs == "steps":      # reference jumps every 250 steps (the example script changes its reference every 30)
        r = np.zeros((T, 4)); r[:, 0] = np.repeat([0.5, -0.8, 1.2, 0.0, 2.0, -1.5, 0.3, 1.0], T // 8)
    else:
        r = bench.REF.copy()
    loop.reset()
    torch.cuda.synchronize(); t0 = time.time()
    loop.run(T, r, p_loss=p_loss, seed=99)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = loop.stats.cpu().numpy()
    print(f"B={B} T={T} refs={refs}: {dt*1e3:.0f} ms -> {B*T/dt/1e6:.1f} M solves/s; status {st[:4].tolist()} ipm iterations {st[4]} "
          f"steps/solve {st[5]/(B*T):.2f}; alive {int(loop.alive.sum().item())}/{B}; max tube {loop.tube_max.max().item():.3e}", flush=True)
    del loop

"""One fused rollout of the benchmark workload (for ncu captures and quick timing)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
from rtmpc_b200.rollout import RemoteLoop
B = int(os.environ.get("B", "4096")); T = int(os.environ.get("T", "250")); reps = int(os.environ.get("REPS", "3"))
tube = os.environ.get("TUBE", "1") == "1"
mpc, Z = bench.build_controller()
loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z if tube else None)
dev = loop.dev
p_loss = torch.as_tensor(np.array([0.1 * (i % 10) for i in range(B)]), device=dev)
ref = torch.as_tensor(bench.REF.copy(), device=dev)
for r in range(reps):
    loop.reset()
    torch.cuda.synchronize(); t0 = time.time()
    loop.run(T, ref, p_loss=p_loss, seed=679 + r, fused=True)
    torch.cuda.synchronize(); dt = time.time() - t0
st = loop.stats.cpu().numpy()
print(f"B={B} T={T} tube={tube}: {dt*1e3:.2f} ms -> {B*T/dt/1e6:.2f} M solves/s; stats {st.tolist()}")

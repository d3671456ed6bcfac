"""Replay solves captured by gpu_capture_nonoptimal.py (x_hat, ref, warm record) through rtmpc_qp_solve and print
why the active-set kernel handed them over (run under gpurun)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robust-tracking-mpc-over-lossy-networks_b200"))
import helpers as H
from rtmpc_b200.qp import BatchedQP
s = H.load("sets_cp.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
b = np.load(sys.argv[1])
dev = torch.device("cuda:0")
B = len(b)
x = torch.as_tensor(b[:, 5:9].copy(), device=dev)
ref = torch.zeros(B, 4, device=dev, dtype=torch.float64); ref[:, 0] = torch.as_tensor(b[:, 9].copy(), device=dev)
for label, warm in (("captured warm sets", torch.as_tensor(b[:, 10:].astype(np.int32), device=dev).contiguous()),
                    ("cold", torch.full((B, qp.warm_stride), -1, device=dev, dtype=torch.int32))):
    z = torch.empty(B, qp.nz, device=dev, dtype=torch.float64)
    U = torch.empty(B, qp.N + 1, qp.nu, device=dev, dtype=torch.float64)
    st = torch.full((B,), -1, device=dev, dtype=torch.int32); it = torch.zeros(B, device=dev, dtype=torch.int32)
    qp.solve_device(x, ref, z, U, st, it, warm=warm)
    torch.cuda.synchronize()
    st, it = st.cpu().numpy(), it.cpu().numpy()
    ipm, steps, rounds = qp.decode_iters(it)
    why = qp.decode_why(it)
    print(label, "status", np.bincount(st, minlength=4).tolist(), "handed over", int((ipm > 0).sum()),
          "why hist (all)", np.bincount(why, minlength=7).tolist(), "why hist (handed over)", np.bincount(why[ipm > 0], minlength=7).tolist(),
          "steps of handed over", steps[ipm > 0][:20].tolist())

"""Timing of one cold batch through the active-set kernel; AS_WARPS=<n> caps the warps per thread block
(rtmpc_set_tuning(RTMPC_TUNE_AS_WARPS, n))."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200 import _lib
from rtmpc_b200.qp import BatchedQP
if os.environ.get('AS_WARPS'):
    _lib.set_tuning(_lib.TUNE_AS_WARPS, int(os.environ['AS_WARPS']))
s = H.load("sets_cp.npz"); g = H.load("loop_cp_tube.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
xh = g["tube_xhat_in"].reshape(-1, 4); refs = np.tile(g["refs"], (4, 1))
dev = torch.device("cuda")
reps = int(os.environ.get("REPS", "16"))
big = torch.as_tensor(np.tile(xh, (reps, 1)), device=dev); bigr = torch.as_tensor(np.tile(refs, (reps, 1)), device=dev)
B = big.shape[0]
U_d = torch.zeros(B, 21, 1, device=dev, dtype=torch.float64); st_d = torch.zeros(B, device=dev, dtype=torch.int32); it_d = torch.zeros_like(st_d)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    qp.solve_device(big, bigr, None, U_d, st_d, it_d)
    torch.cuda.synchronize(); dt = time.time() - t0
steps = ((it_d >> 12) & 0xFFF).sum().item()
print(f"AS_WARPS={_lib.get_tuning(_lib.TUNE_AS_WARPS)}: B={B} {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M solves/s, {steps/dt/1e6:.1f} M steps/s; status {torch.bincount(st_d.clamp(min=0), minlength=4).tolist()}")

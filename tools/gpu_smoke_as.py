"""First-light check of the active-set kernel against the golden fixtures (run under gpurun)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200.qp import BatchedQP

def report(tag, U, Ug, status, iters, dt):
    ipm, steps, rounds = BatchedQP.decode_iters(iters)
    ok = np.isin(status, (0, 3))
    err = np.abs(U[ok] - Ug[ok]).reshape(ok.sum(), -1).max(axis=1) if ok.any() else np.zeros(0)
    print(f"{tag}: B={len(status)} status={np.bincount(np.maximum(status,0), minlength=4)} ipm-its sum={ipm.sum()} (inst {np.count_nonzero(ipm)}) "
          f"as-steps mean={steps.mean():.2f} max={steps.max()} rounds max={rounds.max()} "
          f"maxerr={err.max(initial=0):.3e} time={dt*1e3:.1f} ms", flush=True)
    return err

s = H.load("sets_di.npz")
g = H.load("loop_di_tube.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
t0 = time.time(); z, U, st, it = qp.solve_host(g["xhat_in"], g["refs"]); dt = time.time() - t0
report("DI tube cold", U, g["U_t"], st, it, dt)
print(" z err", np.abs(z - g["z"]).max())

s = H.load("sets_cp.npz"); g = H.load("loop_cp_tube.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
xh = g["tube_xhat_in"].reshape(-1, 4); Ug = g["tube_U_t"].reshape(xh.shape[0], -1, 1)
refs = np.tile(g["refs"], (4, 1))
for rep in range(2):
    t0 = time.time(); z, U, st, it = qp.solve_host(xh, refs); dt = time.time() - t0
    err = report("CP tube cold", U, Ug, st, it, dt)
print(" z err", np.abs(z - g["tube_z"].reshape(xh.shape[0], -1)).max())
qp.set_method("interior_point")
t0 = time.time(); z2, U2, st2, it2 = qp.solve_host(xh, refs); dt = time.time() - t0
report("CP tube IPM ", U2, Ug, st2, it2, dt)
print(" |U_as - U_ipm|", np.abs(U - U2).max())
qp.set_method("active_set")
# warm sequence: the 4 golden closed loops, step by step, warm state in the handle
errs = []; steps_all = []
qp.warm_reset()
for t in range(250):
    z, U, st, it = qp.solve_host(g["tube_xhat_in"][:, t], np.tile(g["refs"][t], (4, 1)), warm=True)
    assert np.all(st == 0), (t, st)
    errs.append(np.abs(U - g["tube_U_t"][:, t]).max())
    steps_all.append(BatchedQP.decode_iters(it)[1])
steps_all = np.array(steps_all)
print(f"CP tube warm sequence: maxerr={max(errs):.3e} as-steps mean={steps_all.mean():.2f} max={steps_all.max()} hist={np.bincount(steps_all.flatten())[:12]}")
# throughput of one big cold batch
dev = torch.device("cuda")
big = torch.as_tensor(np.tile(xh, (16, 1)), device=dev); bigr = torch.as_tensor(np.tile(refs, (16, 1)), device=dev)
B = big.shape[0]
U_d = torch.zeros(B, 21, 1, device=dev, dtype=torch.float64); st_d = torch.zeros(B, device=dev, dtype=torch.int32); it_d = torch.zeros_like(st_d)
for meth in ("active_set", "interior_point"):
    qp.set_method(meth)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        qp.solve_device(big, bigr, None, U_d, st_d, it_d)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f"CP tube cold B={B} {meth}: {dt*1e3:.2f} ms -> {B/dt:.0f} solves/s; status {torch.bincount(st_d.clamp(min=0), minlength=4).tolist()}")

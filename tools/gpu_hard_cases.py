"""The hard-case fixture through both kernels (run under gpurun): statuses, accuracy, steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robust-tracking-mpc-over-lossy-networks_b200"))
import helpers as H
from rtmpc_b200.qp import BatchedQP
s, g = H.load("sets_cp.npz"), H.load("hard_cp.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
for method in ("active_set", "interior_point"):
    qp.set_method(method)
    z, U, st, it = qp.solve_host(g["x"], g["ref"])
    ipm, steps, rounds = qp.decode_iters(it)
    ok = (st == 0) | (st == 3)
    okg = g["status"] == 0
    both = ok & okg
    err = np.abs(z[both] - g["z"][both]).max(axis=1) / np.maximum(1.0, np.abs(g["z"][both]).max(axis=1))
    print(method, "status", np.bincount(st, minlength=4).tolist(), "golden", np.bincount(g["status"], minlength=4).tolist(),
          "mismatch", int((np.where(st == 3, 0, st) != g["status"]).sum()),
          "err max %.2e median %.2e" % (err.max(), np.median(err)), "err(status3) max %.2e" % (np.abs(z[both & (st == 3)] - g["z"][both & (st == 3)]).max() if (both & (st == 3)).any() else 0.0),
          "ipm iters max", int(ipm.max()), "solves with ipm", int((ipm > 0).sum()), "steps max/mean", int(steps.max()), float(steps.mean()), "rounds max", int(rounds.max()))

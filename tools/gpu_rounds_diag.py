"""Diagnostic: certification rounds / active-set steps per solve along the golden closed loops (warm-started)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200.qp import BatchedQP
s = H.load("sets_cp.npz"); g = H.load("loop_cp_tube.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
qp.warm_reset()
R = []; S = []
for t in range(250):
    z, U, st, it = qp.solve_host(g["tube_xhat_in"][:, t], np.tile(g["refs"][t], (4, 1)), warm=True)
    ipm, steps, rounds = BatchedQP.decode_iters(it)
    R.append(rounds); S.append(steps)
R = np.array(R); S = np.array(S)
print("rounds hist", np.bincount(R.flatten()))
for r in range(4):
    m = R == r
    if m.any(): print(f"rounds={r}: n={m.sum()} steps mean={S[m].mean():.2f} hist={np.bincount(S[m])[:12]}")
idx = np.argwhere(R >= 2)[:10]
print("examples (t, run, steps):", [(int(a), int(b), int(S[a, b])) for a, b in idx])

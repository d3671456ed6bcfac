import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from rtmpc_b200.condense import MPCSpec
from rtmpc_b200.qp import BatchedQP
s = H.load("sets_di.npz"); r = H.load("qp_di_regulators.npz")
reg = BatchedQP(MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), stage_x=(s["X_A"], s["X_b"]), stage_u=(s["U_A"], s["U_b"])))
z, U, st, it = reg.solve_host(r["xs"])
err = np.abs(z - r["z_reg"]).max(axis=1)
np.set_printoptions(linewidth=200, precision=3)
print("status", st, "iters", it)
print("err", err)
i = int(np.argmax(err)); print("worst", i, r["xs"][i]); print("z gpu", z[i]); print("z ref", r["z_reg"][i]); print("diff", z[i]-r["z_reg"][i])
print("n", reg.n, "m", reg.m, "npad", reg.data.npad, "mpad", reg.data.mpad)

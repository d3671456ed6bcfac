"""Wall time of the offline set pipeline (mRPI, tightening, terminal set, QP upload) for the cartpole, on the GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200 import mpc
from rtmpc_b200.polytope import box
s = H.load("sets_cp.npz")
hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
c.set_input_constraints(H.poly(s, "U")); c.set_state_constraints(H.poly(s, "X"))
t0 = time.time(); c.determine_mRPI(box(hw), rpi_method=1); t1 = time.time()
print(f"determine_mRPI (Darup, s_max retry, reduce): {t1 - t0:.2f} s, Z rows {c._Z.A.shape[0]}", flush=True)
c.tighten_constraints(); t2 = time.time()
print(f"tighten_constraints: {t2 - t1:.3f} s", flush=True)
c.determine_Xf(); t3 = time.time()
print(f"determine_Xf (maximal output admissible set): {t3 - t2:.2f} s, Xf rows {c._Xf.A.shape[0]}", flush=True)
c.generate_optimization_problem(True); t4 = time.time()
print(f"generate_optimization_problem (condense, equilibrate, derive W, upload): {t4 - t3:.2f} s", flush=True)

"""Fuzz: active-set kernel vs interior-point kernel on random (x_hat, ref) in a wide box (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200.qp import BatchedQP

def fuzz(name, spec, K, box, rbox, B=20000, seed=0):
    rng = np.random.default_rng(seed)
    nx = len(box)
    X = rng.uniform(-1, 1, (B, nx)) * box
    R = np.zeros((B, nx)); R[:, 0] = rng.uniform(-rbox, rbox, B)
    qp = BatchedQP(spec, Kss=K)
    z1, U1, s1, i1 = qp.solve_host(X, R)
    qp.set_method("interior_point")
    z2, U2, s2, i2 = qp.solve_host(X, R)
    ipm_fb = np.count_nonzero(i1 & 0xFFF)
    both = (s1 == 0) & np.isin(s2, (0, 3))
    err = np.abs(U1[both] - U2[both]).reshape(both.sum(), -1).max(1) if both.any() else np.zeros(1)
    scale = np.maximum(1.0, np.abs(U2[both]).reshape(both.sum(), -1).max(1)) if both.any() else np.ones(1)
    print(f"{name}: B={B} AS status {np.bincount(s1, minlength=4).tolist()} IPM status {np.bincount(s2, minlength=4).tolist()} "
          f"handed over {ipm_fb}; both solved {both.sum()} max rel err {(err/scale).max():.2e}; "
          f"AS infeasible & IPM solved: {np.count_nonzero((s1 == 2) & np.isin(s2, (0, 3)))}; "
          f"AS solved & IPM infeasible: {np.count_nonzero((s1 == 0) & (s2 == 2))}; IPM maxiter: {np.count_nonzero(s2 == 1)}", flush=True)

sd, sc = H.load("sets_di.npz"), H.load("sets_cp.npz")
fuzz("di_tube", H.spec_tube_tracking(sd), sd["K"], np.array([9.0, 3.0]), 9.0)
fuzz("di_ext_recv", H.spec_ext_received(sd), sd["K"], np.array([9.0, 3.0]), 9.0)
fuzz("cp_tube", H.spec_tube_tracking(sc), sc["K"], np.array([5.0, 4.0, 0.1, 1.2]), 5.0)
fuzz("cp_tube_near", H.spec_tube_tracking(sc), sc["K"], np.array([1.0, 1.0, 0.05, 0.3]), 1.0)
fuzz("cp_ext_recv", H.spec_ext_received(sc), sc["K"], np.array([1.0, 1.0, 0.05, 0.3]), 1.0, B=8000)
fuzz("cp_tubeinit", H.spec_tube_tracking(sc, False), sc["K"], np.array([1.0, 1.0, 0.05, 0.3]), 1.0, B=8000)
fuzz("cp_track", H.spec_tracking(sc), sc["K"], np.array([2.0, 2.0, 0.1, 0.6]), 2.0)

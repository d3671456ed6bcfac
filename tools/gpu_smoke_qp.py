"""First-light check of the CUDA QP kernel against the golden fixtures (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from rtmpc_b200.qp import BatchedQP

def report(tag, U, Ug, status, iters, dt):
    ok = np.isin(status, (0, 3))
    err = np.abs(U[ok] - Ug[ok]).reshape(ok.sum(), -1).max(axis=1) if ok.any() else np.zeros(0)
    iters = iters & 0xFFF                                  # interior-point iterations (see BatchedQP.decode_iters)
    print(f"{tag}: B={len(status)} status={np.bincount(status, minlength=4)} ipm iters mean={iters.mean():.2f} max={iters.max()} "
          f"maxerr={err.max(initial=0):.3e} p99={np.percentile(err,99) if err.size else 0:.3e} time={dt*1e3:.1f} ms", flush=True)
    return err

s = H.load("sets_di.npz")
g = H.load("loop_di_tube.npz")
qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
t0 = time.time(); z, U, st, it = qp.solve_host(g["xhat_in"], g["refs"]); dt = time.time() - t0
err = report("DI tube", U, g["U_t"], st, it, dt)
print(" z err", np.abs(z - g["z"]).max())
if os.path.exists(os.path.join(H.GOLDEN, "sets_cp.npz")):
    s = H.load("sets_cp.npz"); g = H.load("loop_cp_tube.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    xh = g["tube_xhat_in"].reshape(-1, 4); Ug = g["tube_U_t"].reshape(xh.shape[0], -1, 1)
    refs = np.tile(g["refs"], (4, 1))
    for rep in range(2):
        t0 = time.time(); z, U, st, it = qp.solve_host(xh, refs); dt = time.time() - t0
        err = report("CP tube", U, Ug, st, it, dt)
    bad = np.argsort(-err)[:10]; print(" worst idx", bad, err[bad], st[bad], it[bad])
    print(" z err", np.abs(z - g["tube_z"].reshape(xh.shape[0], -1)).max())
    big = np.tile(xh, (16, 1)); bigr = np.tile(refs, (16, 1))
    t0 = time.time(); z, U, st, it = qp.solve_host(big, bigr, want_z=False); dt = time.time() - t0
    print(f"CP tube B={len(big)}: {dt*1e3:.1f} ms -> {len(big)/dt:.0f} solves/s (host API)")

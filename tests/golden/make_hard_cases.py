"""Golden fixture of HARD cartpole solves: states a closed loop reaches right after large reference jumps.

The (x_hat, ref) pairs were captured on the GPU by tools/gpu_capture_nonoptimal.py (rollouts of BASELINE configs[1]
with the reference of the first state jumping between 0.5, -0.8, 1.2, 0, 2.0, -1.5, 0.3, 1.0): the solves the dual
active-set kernel could not certify at the time.  Their minimisers sit on 19..21 active rows out of 21 unknowns
(vertices of a nearly empty feasible set), and a few are infeasible.  The solutions here come from the CPU oracle;
``polished`` marks those its active-set endgame certified (KKT at 1e-9) -- on the others the oracle's answer is only
interior-point accurate (the flat Hessian turns a 1e-9 relative residual into ~1e-3 in z), so tests compare those through
an independent KKT check instead:

    python tests/golden/make_hard_cases.py gpurun_out/nonoptimal.npy        # -> tests/golden/hard_cp.npz

(The committed fixture was made from a capture with the kernel as of commit 1d150b7, i.e. before the refactorise-and-restart
safeguard existed; with today's kernel the capture tool finds next to nothing, which is the point.  The fixture's content
does not depend on the kernel: states, references and the ORACLE's solutions.)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_qp as rq            # noqa: E402
from oracle.ref_polytope import Polytope   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main(path, n_feasible=160, seed=3):
    s = np.load(os.path.join(OUT, "sets_cp.npz"))
    P = lambda k: Polytope(s[k + "_A"], s[k + "_b"], normalize=False)      # noqa: E731
    qp = rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], P("Xc"), P("Uc"), P("Xf"), None, True)
    b = np.load(path)
    st = b[:, 2].astype(int)
    rng = np.random.default_rng(seed)
    inf = np.nonzero(st == 2)[0]
    inf = inf[np.unique(b[inf, 1], return_index=True)[1]]                  # one per dead instance
    other = np.nonzero(st != 2)[0]
    pick = np.r_[inf, np.nonzero(st == 1)[0], rng.choice(other, n_feasible, replace=False)]
    xs, refs, zs, status, polished = [], [], [], [], []
    for k in pick:
        x, r = b[k, 5:9].copy(), np.array([b[k, 9], 0.0, 0.0, 0.0])
        sol, res = rq.solve_param(qp, x, r)
        ok = res.status == "optimal"
        assert ok or res.status == "infeasible", res.status
        xs.append(x)
        refs.append(r)
        status.append(0 if ok else 2)
        polished.append(bool(ok and res.polished))
        zs.append(res.z[:qp.nz] if ok else np.full(qp.nz, np.nan))
    np.savez_compressed(os.path.join(OUT, "hard_cp.npz"), x=np.array(xs), ref=np.array(refs), z=np.array(zs),
                        status=np.array(status, np.int32), polished=np.array(polished))
    print("hard_cp:", len(xs), "cases,", int(np.sum(np.array(status) == 2)), "infeasible,", int(np.sum(polished)),
          "with the oracle's active-set polish certified (the others are interior-point accurate only)")


if __name__ == "__main__":
    main(sys.argv[1])

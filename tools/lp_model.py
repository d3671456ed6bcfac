"""numpy model of csrc/rtmpc_lp.cu (one LP at a time, same pivoting rules) used to develop the kernel's control flow on the
CPU; neither product nor oracle.  `python tools/lp_model.py` replays the cartpole terminal-set iteration with this model
as the LP backend and reports the LPs whose result differs from HiGHS."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def refactor(rows, rhs, bidx, c):
    HB = rows[bidx]
    if abs(np.linalg.det(HB / np.maximum(np.abs(HB).max(axis=1, keepdims=True), 1e-300))) < 1e-13:
        return None
    Binv = np.linalg.inv(HB)
    return Binv, Binv @ rhs[bidx], Binv.T @ c


def solve(H, h, c, relax_row=-1, relax_by=0.0, extra=None, box=1e4, safe=False, trace=False):
    """Returns (status, value, x, iterations): status 0 optimal, 1 limit / singular, 2 infeasible, 4 unbounded."""
    m, d = H.shape
    ne = 0 if extra is None else extra.shape[0]
    rows = np.vstack([H] + ([extra[:, :d]] if ne else []) + [np.kron(np.eye(d), np.array([[1.0], [-1.0]]))])
    rhs = np.r_[h, extra[:, d] if ne else [], np.full(2 * d, box)].astype(float)
    if relax_row >= 0:
        rhs[relax_row] += relax_by
    total = m + ne + 2 * d
    neg = c < 0
    bidx = m + ne + 2 * np.arange(d) + neg
    lam = np.abs(c).astype(float)
    x = np.where(neg, -box, box).astype(float)
    Binv = np.diag(np.where(neg, -1.0, 1.0))
    tolv = 1e-9 * (1.0 + np.abs(rhs))
    max_iter = 50 * (d + 2) + (m + ne) // 2
    stalled, fresh, status = 0, True, 1
    it = 0
    while it < max_iter:
        if it > 0 and it % 32 == 0:
            f = refactor(rows, rhs, bidx, c)
            if f is None:
                return 1, np.nan, x, it
            Binv, x, lam = f
            fresh = True
        viol = rows @ x - rhs
        bland = safe or stalled > d
        if bland:
            cand = np.nonzero(viol > tolv)[0]
            p = cand[0] if cand.size else -1
        else:
            p = int(np.argmax(viol - tolv))
            if not (viol[p] - tolv[p] > 0):
                p = -1
        if p < 0:
            if fresh:
                status = 0
                break
            f = refactor(rows, rhs, bidx, c)
            if f is None:
                return 1, np.nan, x, it
            Binv, x, lam = f
            fresh = True
            it += 1
            continue
        fresh = False
        vp = rows[p] @ x - rhs[p]
        w = rows[p] @ Binv
        wmax = np.abs(w).max()
        cand = w > 1e-9 * (1.0 + wmax)
        if not cand.any():
            return 2, -np.inf, x, it
        lamp = np.maximum(lam, 0.0)
        # Harris ratio test: among the rows whose ratio is within a small tolerance of the minimum take the largest pivot
        ratios = np.where(cand, lamp / np.where(cand, w, 1.0), np.inf)
        bound = np.where(cand, (lamp + 1e-9 * (1.0 + lamp.max())) / np.where(cand, w, 1.0), np.inf).min()
        near = cand & (ratios <= bound)
        jl = int(np.argmax(np.where(near, w, -np.inf)))
        theta = ratios[jl]
        stalled = stalled + 1 if theta <= 1e-14 * (1.0 + np.abs(lam).max()) else 0
        lam = np.maximum(lam - theta * w, 0.0)
        lam[jl] = theta
        u = Binv[:, jl].copy()
        x = x - u * (vp / w[jl])
        g = w.copy()
        g[jl] -= 1.0
        Binv = Binv - np.outer(u, g / w[jl])
        bidx[jl] = p
        if trace:
            print(it, p, jl, theta, vp, w[jl])
        it += 1
    if status != 0:
        return 1, np.nan, x, it
    if np.any((bidx >= m + ne) & (lam > 1e-9 * (1.0 + np.abs(c)))):
        return 4, np.inf, x, it
    return 0, float(c @ x), x, it


def lp_batch_model(H, h, obj, relax_row=None, relax_by=0.0, extra=None, want_x=False, stats=None):
    H = np.atleast_2d(np.asarray(H, float)); h = np.asarray(h, float).reshape(-1); obj = np.atleast_2d(obj)
    B, dim = obj.shape
    scale = 1.0 + (np.abs(h).max() if len(h) else 0.0) + (np.abs(extra[:, :, dim]).max() if extra is not None else 0.0)
    val = np.empty(B); X = np.zeros((B, dim))
    for b in range(B):
        rr = -1 if relax_row is None else int(relax_row[b])
        ex = None if extra is None else extra[b]
        st, v, x, it = solve(H, h, obj[b], rr, relax_by, ex, box=1e4 * scale)
        if st == 1:
            st, v, x, it2 = solve(H, h, obj[b], rr, relax_by, ex, box=1e4 * scale, safe=True)
            it += it2
            if stats is not None:
                stats["safe"] = stats.get("safe", 0) + 1
        if st == 1:
            raise RuntimeError(f"LP model failed: dim {dim}, {H.shape[0]} rows, instance {b}")
        if stats is not None:
            stats["lps"] = stats.get("lps", 0) + 1
            stats["iters"] = stats.get("iters", 0) + it
        val[b] = v if st == 0 else (np.inf if st == 4 else -np.inf)
        X[b] = x
    return val, (X if want_x else None)


def main():
    import time
    import helpers as H
    from rtmpc_b200 import mpc, polytope as pc
    s = H.load("sets_cp.npz")
    c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(H.poly(s, "U"))
    c.set_state_constraints(H.poly(s, "X"))
    c._Z, c._Xc, c._Uc = H.poly(s, "Z"), H.poly(s, "Xc"), H.poly(s, "Uc")
    stats = {}
    pc.set_lp_backend("highs")
    real = pc.lp_batch

    def both(Hm, h, obj, relax_row=None, relax_by=0.0, extra=None, want_x=False):
        v0, x0 = real(Hm, h, obj, relax_row, relax_by, extra, want_x)
        v1, x1 = lp_batch_model(Hm, h, obj, relax_row, relax_by, extra, want_x, stats)
        bad = ~((v0 == v1) | (np.abs(v0 - v1) <= 1e-8 * (1 + np.abs(v0))))
        if bad.any():
            stats["mismatch"] = stats.get("mismatch", 0) + int(bad.sum())
            print("mismatch", np.nonzero(bad)[0][:5], v0[bad][:5], v1[bad][:5], np.atleast_2d(Hm).shape)
        return v1, x1
    pc.lp_batch = both
    t0 = time.time()
    c.determine_Xf()
    print("terminal set", c._Xf.A.shape, f"{time.time() - t0:.0f} s", stats)
    ref = H.poly(s, "Xf")
    pc.lp_batch = real
    print("equal to fixture:", pc.is_subset(c._Xf, ref) and pc.is_subset(ref, c._Xf))


if __name__ == "__main__":
    main()

"""Development aid (NOT product, NOT oracle): a numpy model of the CUDA interior-point kernel.

It mirrors ``csrc/rtmpc_ipm.cuh`` step for step on the scaled two-sided problem produced by
``rtmpc_b200.ipm_data.prepare`` so the algorithm (start point, Mehrotra steps, termination,
infeasibility test) can be exercised on a CPU-only box.  Nothing in the product imports it.
"""
import numpy as np


def ipm_model(d, x_init, ref, max_iter=60, tol_res=1e-9, tol_gap=1e-12, verbose=False, warm=None,
              return_state=False):
    n, m = d.n, d.m
    G = d.Gs[:m, :n]
    H = d.Hs
    q = d.Fx[:n] @ x_init + d.Fr[:n] @ ref
    lo = d.lo0[:m] + d.Lx[:m] @ x_init
    up = d.up0[:m] + d.Ux[:m] @ x_init
    hl = d.has_lo[:m].astype(bool)
    hu = d.has_up[:m].astype(bool)
    mtot = hl.sum() + hu.sum()
    zeta = -(d.Hinv[:n, :n] @ q) if warm is None else warm.copy()
    t = G @ zeta
    su = np.where(hu, up - t, 1.0)
    sl = np.where(hl, t - lo, 1.0)
    smin = min(su[hu].min(initial=np.inf), sl[hl].min(initial=np.inf))
    if smin > 0 and warm is None:
        return zeta, 0, 0, dict(path="unconstrained")
    shift = max(-1.5 * smin, 0.0)
    su = np.where(hu, np.maximum(su + shift, d.s_floor), 1.0)
    sl = np.where(hl, np.maximum(sl + shift, d.s_floor), 1.0)
    lu = np.where(hu, 1.0, 0.0)
    ll = np.where(hl, 1.0, 0.0)
    mu0 = (su @ lu + sl @ ll) / mtot
    # lam = mu0 / s  (start on the central path for the chosen slacks)
    lu = np.where(hu, mu0 / su, 0.0)
    ll = np.where(hl, mu0 / sl, 0.0)
    sc_q = 1.0 + np.abs(q).max()
    status = 1
    best = None
    for it in range(max_iter):
        t = G @ zeta
        rpu = np.where(hu, t + su - up, 0.0)
        rpl = np.where(hl, -t + sl + lo, 0.0)
        rd = H @ zeta + q + G.T @ (lu - ll)
        mu = (su @ lu + sl @ ll) / mtot
        pobj = 0.5 * zeta @ H @ zeta + q @ zeta
        res = max(np.abs(rd).max() / sc_q, np.abs(rpu).max() / d.sc_b, np.abs(rpl).max() / d.sc_b)
        relgap = mu * mtot / (1.0 + abs(pobj))
        if verbose:
            print(it, res, relgap, mu)
        merit = max(res, relgap)
        if best is None or merit < best[0]:
            best = (merit, zeta.copy(), it, (su.copy(), sl.copy(), lu.copy(), ll.copy()))
        if res <= tol_res and relgap <= tol_gap:
            status = 0
            break
        if best[0] <= 1e-8 and merit > 1e3 * best[0]:
            status = 0
            break
        # primal infeasibility certificate on the normalised multiplier direction
        y = lu - ll
        yn = np.abs(y).max()
        if yn > 1e6 * sc_q and it >= 3:
            cert = (np.where(hu, up * lu, 0.0).sum() - np.where(hl, lo * ll, 0.0).sum())
            if np.abs(G.T @ y).max() <= 1e-7 * yn and cert < -1e-7 * yn * d.sc_b:
                status = 2
                break
        if not np.isfinite(mu) or mu > 1e40:
            break
        dd = np.where(hu, lu / su, 0.0) + np.where(hl, ll / sl, 0.0)
        S = H + (G.T * dd) @ G
        try:
            L = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            break

        def solve(rhs):
            return np.linalg.solve(L.T, np.linalg.solve(L, rhs))
        e1 = np.where(hu, lu * rpu / su, 0.0) - np.where(hl, ll * rpl / sl, 0.0)
        dz_a = solve(-(H @ zeta) - q - G.T @ e1)
        ta = G @ dz_a
        dsu_a = -rpu - ta
        dsl_a = -rpl + ta
        dlu_a = np.where(hu, -lu * (1.0 + dsu_a / su), 0.0)
        dll_a = np.where(hl, -ll * (1.0 + dsl_a / sl), 0.0)

        def maxstep(v, dv, mask):
            r = np.where(mask & (dv < 0), -v / np.where(dv < 0, dv, -1.0), np.inf)
            return min(1.0, r.min(initial=np.inf))
        ap = min(maxstep(su, dsu_a, hu), maxstep(sl, dsl_a, hl))
        ad = min(maxstep(lu, dlu_a, hu), maxstep(ll, dll_a, hl))
        mu_aff = (np.where(hu, (su + ap * dsu_a) * (lu + ad * dlu_a), 0.0).sum()
                  + np.where(hl, (sl + ap * dsl_a) * (ll + ad * dll_a), 0.0).sum()) / mtot
        sigma = (mu_aff / mu) ** 3
        # corrector
        rcu = su * lu + dsu_a * dlu_a - sigma * mu
        rcl = sl * ll + dsl_a * dll_a - sigma * mu
        e2 = np.where(hu, (-rcu + lu * rpu) / su + lu, 0.0) - np.where(hl, (-rcl + ll * rpl) / sl + ll, 0.0)
        dz = solve(-(H @ zeta) - q - G.T @ e2)
        tz = G @ dz
        dsu = -rpu - tz
        dsl = -rpl + tz
        dlu = np.where(hu, (-rcu - lu * dsu) / su, 0.0)
        dll = np.where(hl, (-rcl - ll * dsl) / sl, 0.0)
        eta = min(0.9995, max(0.995, 1.0 - mu)) if mu < 1 else 0.995
        ap = min(1.0, eta * min(maxstep(su, dsu, hu), maxstep(sl, dsl, hl)))
        ad = min(1.0, eta * min(maxstep(lu, dlu, hu), maxstep(ll, dll, hl)))
        zeta = zeta + ap * dz
        su = np.where(hu, su + ap * dsu, 1.0)
        sl = np.where(hl, sl + ap * dsl, 1.0)
        lu = np.where(hu, lu + ad * dlu, 0.0)
        ll = np.where(hl, ll + ad * dll, 0.0)
    if status == 1 and best is not None and best[0] <= 1e-7:
        status = 0
    if status == 0 and best is not None:
        zeta = best[1]
    info = dict(path="ipm", merit=best[0] if best else None)
    if return_state and best is not None:
        info["state"] = best[3]
    if status == 1:
        y = lu - ll
        yn = np.abs(y).max()
        cert = (np.where(hu, up * lu, 0.0).sum() - np.where(hl, lo * ll, 0.0).sum())
        info["cert"] = (np.abs(G.T @ y).max() / yn, cert / yn)
        if np.abs(G.T @ y).max() <= 1e-6 * yn and cert < -1e-6 * yn * d.sc_b:
            status = 2
    return zeta, status, it + 1, info


def polish_model(d, x_init, ref, act, max_rounds=24, tol_p=1e-9, tol_d=1e-9):
    """Active-set endgame of the kernel: ``act`` is a list of (row, sign) pairs (sign +1: upper bound
    active, -1: lower).  Schur complement on Hinv with dependency-dropping Cholesky, one step of
    iterative refinement, then drop-most-negative / add-most-violated."""
    n, m = d.n, d.m
    G = d.Gs[:m, :n]
    Y = d.Y[:m, :n]
    q = d.Fx[:n] @ x_init + d.Fr[:n] @ ref
    lo = d.lo0[:m] + d.Lx[:m] @ x_init
    up = d.up0[:m] + d.Ux[:m] @ x_init
    hl = d.has_lo[:m].astype(bool)
    hu = d.has_up[:m].astype(bool)
    zu = -(d.Hinv[:n, :n] @ q)
    act = list(act)
    for rnd in range(max_rounds):
        na = len(act)
        if na > n:
            return None, rnd, "too_many"
        rows = np.array([a[0] for a in act], int)
        sg = np.array([a[1] for a in act], float)
        b = np.where(sg > 0, up[rows], -lo[rows]) if na else np.zeros(0)
        Gt = sg[:, None] * G[rows] if na else np.zeros((0, n))
        Yt = sg[:, None] * Y[rows] if na else np.zeros((0, n))
        S = Yt @ Gt.T
        # Cholesky with dependency dropping
        L = np.zeros((na, na))
        keep = np.ones(na, bool)
        dmax = np.max(np.diag(S)) if na else 1.0
        for j in range(na):
            v = S[j, j] - L[j, :j] @ L[j, :j]
            if v <= 1e-11 * S[j, j] or v <= 1e-14 * dmax:
                keep[j] = False
                L[j, :] = 0.0
                L[:, j] = 0.0
                L[j, j] = 1.0
                continue
            L[j, j] = np.sqrt(v)
            for i in range(j + 1, na):
                L[i, j] = (S[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]

        def ssolve(r):
            r = np.where(keep, r, 0.0)
            y = np.linalg.solve(L, r) if na else r
            x = np.linalg.solve(L.T, y) if na else y
            return np.where(keep, x, 0.0)
        lam = ssolve(Gt @ zu - b)
        z = zu - Yt.T @ lam
        for _ in range(2):
            dl = ssolve(Gt @ z - b)
            lam = lam + dl
            z = z - Yt.T @ dl
        t = G @ z
        viol_u = np.where(hu, t - up, -np.inf)
        viol_l = np.where(hl, lo - t, -np.inf)
        # rows kept active are satisfied by construction
        for (r_, s_), k_ in zip(act, keep):
            if k_:
                if s_ > 0:
                    viol_u[r_] = -np.inf
                else:
                    viol_l[r_] = -np.inf
        wu = int(np.argmax(viol_u))
        wl = int(np.argmax(viol_l))
        vmax = max(viol_u[wu], viol_l[wl])
        lam_min = lam[keep].min() if keep.any() else 0.0
        if lam_min < -tol_d * (1.0 + np.abs(lam).max(initial=0.0)):
            j = int(np.argmin(np.where(keep, lam, np.inf)))
            act.pop(j)
            continue
        if vmax > tol_p * d.sc_b:
            # drop dependent rows first, then add the most violated
            act = [a for a, k_ in zip(act, keep) if k_]
            new = (wu, 1) if viol_u[wu] >= viol_l[wl] else (wl, -1)
            if new in act:
                return None, rnd, "cycle"
            act.append(new)
            continue
        return z, rnd + 1, "ok"
    return None, max_rounds, "rounds"


def solve_model(d, x_init, ref, gap_stop=1e-8):
    """IPM to a moderate tolerance, then the active-set endgame; falls back to a tight IPM."""
    z, status, it, info = ipm_model(d, x_init, ref, tol_res=1e-7, tol_gap=gap_stop, return_state=True)
    if info.get("path") == "unconstrained" or status != 0:
        return z, status, it, info
    su, sl, lu, ll = info["state"]
    hl = d.has_lo[:d.m].astype(bool)
    hu = d.has_up[:d.m].astype(bool)
    act = [(i, 1) for i in np.nonzero(hu & (lu > su))[0]] + [(i, -1) for i in np.nonzero(hl & (ll > sl))[0]]
    zp, rounds, why = polish_model(d, x_init, ref, act)
    info["polish"] = (rounds, why, len(act))
    if zp is not None:
        return zp, 0, it, info
    z2, status2, it2, info2 = ipm_model(d, x_init, ref)
    info2["polish"] = info["polish"]
    return z2, status2, it + it2, info2

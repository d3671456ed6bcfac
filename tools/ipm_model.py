"""Development aid (NOT product, NOT oracle): a numpy model of the CUDA interior-point kernel.

It mirrors ``csrc/rtmpc_ipm.cuh`` step for step on the scaled two-sided problem produced by
``rtmpc_b200.ipm_data.prepare`` so the algorithm (start point, Mehrotra steps, termination,
infeasibility test) can be exercised on a CPU-only box.  Nothing in the product imports it.
"""
import numpy as np


def ipm_model(d, x_init, ref, max_iter=60, tol_res=1e-9, tol_gap=1e-12, verbose=False, warm=None):
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
            best = (merit, zeta.copy(), it)
        if res <= tol_res and relgap <= tol_gap:
            status = 0
            break
        if best[0] <= 1e-8 and merit > 1e3 * best[0]:
            status = 0
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
    if status != 0 and best is not None and best[0] <= 1e-7:
        status = 0
    if status == 0 and best is not None:
        zeta = best[1]
    info = dict(path="ipm", merit=best[0] if best else None)
    if status != 0:
        y = lu - ll
        yn = np.abs(y).max()
        cert = (np.where(hu, up * lu, 0.0).sum() - np.where(hl, lo * ll, 0.0).sum())
        info["cert"] = (np.abs(G.T @ y).max() / yn, cert / yn)
        if np.abs(G.T @ y).max() <= 1e-6 * yn and cert < -1e-6 * yn:
            status = 2
    return zeta, status, it + 1, info

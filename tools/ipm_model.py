"""Development aid (NOT product, NOT oracle): a numpy model of the CUDA interior-point kernel.

It mirrors ``csrc/rtmpc_ipm.cuh`` step for step on the scaled two-sided problem produced by
``rtmpc_b200.ipm_data.prepare`` so the algorithm (start point, Mehrotra steps, endgame triggers,
termination, infeasibility test) can be exercised on a CPU-only box.  Nothing in the product
imports it.
"""
import numpy as np

OPTIMAL, MAX_ITER, INFEASIBLE, INACCURATE = 0, 1, 2, 3


def _problem(d, x_init, ref):
    n, m = d.n, d.m
    G = d.Gs[:m, :n]
    q = d.Fx[:n] @ x_init + d.Fr[:n] @ ref
    lo = d.lo0[:m] + d.Lx[:m] @ x_init
    up = d.up0[:m] + d.Ux[:m] @ x_init
    hl = d.has_lo[:m].astype(bool)
    hu = d.has_up[:m].astype(bool)
    return n, m, G, q, lo, up, hl, hu


def polish_model(d, x_init, ref, act, max_rounds=24, tol_p=1e-11, tol_d=1e-9):
    """Active-set endgame of the kernel: ``act`` is a list of (row, sign) pairs (sign +1: upper bound
    active, -1: lower).  Schur complement on Hinv with dependency-dropping Cholesky, two steps of
    iterative refinement, then drop-most-negative / add-most-violated."""
    n, m, G, q, lo, up, hl, hu = _problem(d, x_init, ref)
    Y = d.Y[:m, :n]
    zu = -(d.Hinv[:n, :n] @ q)
    act = list(act)
    for rnd in range(max_rounds):
        na = len(act)
        if na > d.npad:
            return None, rnd, "too_many", act
        rows = np.array([a[0] for a in act], int)
        sg = np.array([a[1] for a in act], float)
        b = np.where(sg > 0, up[rows], -lo[rows]) if na else np.zeros(0)
        Gt = sg[:, None] * G[rows] if na else np.zeros((0, n))
        Yt = sg[:, None] * Y[rows] if na else np.zeros((0, n))
        S = Yt @ Gt.T
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
        wu = int(np.argmax(viol_u))
        wl = int(np.argmax(viol_l))
        vmax = max(viol_u[wu], viol_l[wl])
        lam_min = lam[keep].min() if keep.any() else 0.0
        if lam_min < -tol_d * (1.0 + np.abs(lam).max(initial=0.0)):
            j = int(np.argmin(np.where(keep, lam, np.inf)))
            act.pop(j)
            continue
        if vmax > tol_p * d.sc_b:
            new = (wu, 1) if viol_u[wu] >= viol_l[wl] else (wl, -1)
            if any(a == new and k_ for a, k_ in zip(act, keep)):
                return None, rnd, "cycle", act
            act = [a for a, k_ in zip(act, keep) if k_]
            act.append(new)
            continue
        return z, rnd + 1, "ok", [a for a, k_ in zip(act, keep) if k_]
    return None, max_rounds, "rounds", act


def solve_model(d, x_init, ref, max_iter=60, verbose=False, warm=None, shift=None, warm_rounds=10):
    """One instance, exactly the kernel's control flow.  Returns (zeta_scaled, status, iters, info).
    ``warm``: active set [(row, sign)] of the previous control step; it is shifted one stage with
    ``shift`` and tried in the endgame before any interior-point iteration."""
    n, m, G, q, lo, up, hl, hu = _problem(d, x_init, ref)
    H = d.Hs
    info = dict(path="ipm", polish=[], active=None)
    if np.any(d.par_C @ x_init - d.par_h > 1e-9 * (1.0 + np.abs(d.par_h))):
        return np.full(n, np.nan), INFEASIBLE, 0, dict(path="param_rows")
    mtot = hl.sum() + hu.sum()
    zeta = -(d.Hinv[:n, :n] @ q)
    t = G @ zeta
    su = np.where(hu, up - t, 1.0)
    sl = np.where(hl, t - lo, 1.0)
    smin = min(su[hu].min(initial=np.inf), sl[hl].min(initial=np.inf))
    if smin > 0:
        return zeta, OPTIMAL, 0, dict(path="unconstrained", active=[], polish=[])
    if warm is not None:
        cand = [(int(shift[r]), sg) for r, sg in warm if shift[r] >= 0] if shift is not None else list(warm)
        zp, rounds, why, fin = polish_model(d, x_init, ref, cand, max_rounds=warm_rounds)
        info["polish"].append((0, rounds, "warm-" + why, len(cand)))
        if zp is not None:
            info["path"] = "warm"
            info["active"] = fin
            return zp, OPTIMAL, 0, info
    shift = max(-1.5 * smin, 0.0)
    su = np.where(hu, np.maximum(su + shift, d.s_floor), 1.0)
    sl = np.where(hl, np.maximum(sl + shift, d.s_floor), 1.0)
    mu0 = (su[hu].sum() + sl[hl].sum()) / mtot
    lu = np.where(hu, mu0 / su, 0.0)
    ll = np.where(hl, mu0 / sl, 0.0)
    sc_q = 1.0 + np.abs(q).max()
    status = MAX_ITER
    phase, iters, next_try = 1, 0, 0
    boost = False
    best_merit, best_zeta, best_act = np.inf, zeta.copy(), None

    def active():
        return [(i, 1) for i in np.nonzero(hu & (lu > su))[0]] + [(i, -1) for i in np.nonzero(hl & (ll > sl))[0]]

    def try_polish(act):
        zp, rounds, why, fin = polish_model(d, x_init, ref, act)
        info["polish"].append((iters, rounds, why, len(act)))
        if zp is not None:
            info["active"] = fin
        return zp

    while True:
        t = G @ zeta
        rpu = np.where(hu, t + su - up, 0.0)
        rpl = np.where(hl, -t + sl + lo, 0.0)
        y = lu - ll
        hz = H @ zeta
        gy = G.T @ y
        rd = hz + q + gy
        gap = su[hu] @ lu[hu] + sl[hl] @ ll[hl]
        mu = gap / mtot
        pobj = zeta @ (0.5 * hz + q)
        rd_rel = np.abs(rd).max() / sc_q
        rp_rel = max(np.abs(rpu).max(), np.abs(rpl).max()) / d.sc_b
        res = max(rd_rel, rp_rel)
        relgap = gap / (1.0 + abs(pobj))
        merit = max(res, relgap)
        ymax = np.abs(y).max()
        cert = up[hu] @ lu[hu] - lo[hl] @ ll[hl]
        gy_max = np.abs(gy).max()
        if verbose:
            print(iters, "rd", rd_rel, "rp", rp_rel, "gap", relgap, "mu", mu)
        bad = not np.isfinite(mu) or mu > 1e40
        if not bad and merit < best_merit:
            best_merit, best_zeta = merit, zeta.copy()
            if merit <= 1e-3:
                best_act = active()
        if not bad and ymax > 1e6 * sc_q and iters >= 3 and gy_max <= 1e-7 * ymax and cert < -1e-7 * ymax * d.sc_b:
            status = INFEASIBLE
            break
        conv1 = (res <= 1e-7 and relgap <= 1e-8) or (rp_rel <= 1e-8 and relgap <= 1e-9 and rd_rel <= 1e-3)
        conv2 = res <= 1e-9 and relgap <= 1e-12
        diverged = bad or (best_merit <= 1e-4 and merit > 1e3 * best_merit)     # losing a good point, not early wobble
        if not diverged and conv1 and iters >= next_try and iters < max_iter and not conv2:
            zp = try_polish(active())
            if zp is not None:
                zeta, status = zp, OPTIMAL
                break
            phase = 2
            next_try = iters + 3
        if conv2 or diverged or iters >= max_iter:
            if best_merit <= 1e-5 and best_act is not None:
                zp = try_polish(active() if (conv2 and not diverged) else best_act)
                if zp is not None:
                    zeta, status = zp, OPTIMAL
                    break
            zeta = best_zeta
            if best_merit <= 1e-7:
                status = INACCURATE
            elif not bad and gy_max <= 1e-6 * ymax and cert < -1e-6 * ymax * d.sc_b:
                status = INFEASIBLE
            else:
                status = MAX_ITER
            break
        iters += 1
        dd = np.where(hu, lu / su, 0.0) + np.where(hl, ll / sl, 0.0)
        S = H + (G.T * dd) @ G
        try:
            Lc = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            Lc = np.linalg.cholesky(S + 1e-10 * np.abs(np.diag(S)).max() * np.eye(n))

        def solve(rhs):
            return np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
        e1 = np.where(hu, lu * rpu / su, 0.0) - np.where(hl, ll * rpl / sl, 0.0)
        dz_a = solve(-hz - q - G.T @ e1)
        ta = G @ dz_a
        dsu_a = -rpu - ta
        dsl_a = -rpl + ta
        dlu_a = np.where(hu, -lu * (1.0 + dsu_a / su), 0.0)
        dll_a = np.where(hl, -ll * (1.0 + dsl_a / sl), 0.0)

        def maxstep(v, dv, mask):
            r = np.where(mask & (dv < 0), -v / np.where(dv < 0, dv, -1.0), np.inf)
            return r.min(initial=np.inf)
        ap = min(1.0, maxstep(su, dsu_a, hu), maxstep(sl, dsl_a, hl))
        ad = min(1.0, maxstep(lu, dlu_a, hu), maxstep(ll, dll_a, hl))
        mu_aff = (np.where(hu, (su + ap * dsu_a) * (lu + ad * dlu_a), 0.0).sum()
                  + np.where(hl, (sl + ap * dsl_a) * (ll + ad * dll_a), 0.0).sum()) / mtot
        sigma = (mu_aff / mu) ** 3
        cross = 1.0
        if boost:
            # safeguard against Mehrotra's limit cycles: after a short step take a well-centred,
            # first-order step (sigma >= 0.8, no second-order term)
            sigma = max(sigma, 0.8)
            cross = 0.0
        rcu = su * lu + cross * dsu_a * dlu_a - sigma * mu
        rcl = sl * ll + cross * dsl_a * dll_a - sigma * mu
        e2 = np.where(hu, (-rcu + lu * rpu) / su + lu, 0.0) - np.where(hl, (-rcl + ll * rpl) / sl + ll, 0.0)
        dz = solve(-hz - q - G.T @ e2)
        tz = G @ dz
        dsu = -rpu - tz
        dsl = -rpl + tz
        dlu = np.where(hu, (-rcu - lu * dsu) / su, 0.0)
        dll = np.where(hl, (-rcl - ll * dsl) / sl, 0.0)
        eta = min(0.9995, max(0.995, 1.0 - mu)) if mu < 1 else 0.995
        ap = min(1.0, eta * min(maxstep(su, dsu, hu), maxstep(sl, dsl, hl)))
        ad = min(1.0, eta * min(maxstep(lu, dlu, hu), maxstep(ll, dll, hl)))
        boost = min(ap, ad) < 0.3
        zeta = zeta + ap * dz
        su = np.where(hu, su + ap * dsu, 1.0)
        sl = np.where(hl, sl + ap * dsl, 1.0)
        lu = np.where(hu, lu + ad * dlu, 0.0)
        ll = np.where(hl, ll + ad * dll, 0.0)
    info["merit"] = best_merit
    info["phase"] = phase
    return zeta, status, iters, info

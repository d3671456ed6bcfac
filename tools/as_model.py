"""Development aid (NOT product, NOT oracle): numpy model of the dual active-set kernel.

Goldfarb-Idnani dual active-set iteration on the scaled two-sided problem of
``rtmpc_b200.ipm_data.prepare``, on the operators shared by the whole batch
(``Hinv``, ``Y = G Hinv``, ``W = G Hinv G'``), warm-started from the previous control step's
active set moved one stage earlier, certified at the end by one round of the endgame
(``ipm_model.polish_model``).  Mirrors the control flow of ``csrc/rtmpc_as.cuh`` (steps, ratio tests, safeguards); the kernel's
carried working set (rollout) and its tiered certification (row values accepted with a rounding bound before ``G' z`` is
formed) are not modelled - the model always certifies on recomputed rows, which is the kernel's last tier.
"""
import numpy as np

from ipm_model import INFEASIBLE, OPTIMAL, _problem, polish_model

FALLBACK = -2      # hand the instance to the interior-point kernel


def _chol_dd(S):
    """dependency-dropping Cholesky, as in the kernel: returns L and the kept mask"""
    na = S.shape[0]
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
    return L, keep


def _ssolve(L, keep, r):
    r = np.where(keep, r, 0.0)
    if len(r) == 0:
        return r
    x = np.linalg.solve(L.T, np.linalg.solve(L, r))
    return np.where(keep, x, 0.0)


def solve_as(d, W, x_init, ref, warm=None, shift=None, max_iter=96, tol_p=1e-11, kappa_eps=1e-11, drop_all=True,
             verbose=False):
    """Returns (zeta_scaled, status, info).  info: iters (add/drop steps), rounds (certification
    rounds), active (final [(row, sign)])."""
    n, m, G, q, lo, up, hl, hu = _problem(d, x_init, ref)
    info = dict(path="as", iters=0, drops=0, rounds=0, active=None, warm_kept=0)
    if np.any(d.par_C @ x_init - d.par_h > 1e-9 * (1.0 + np.abs(d.par_h))):
        return np.full(n, np.nan), INFEASIBLE, dict(info, path="param_rows")
    Hinv = d.Hinv[:n, :n]
    Y = d.Y[:m, :n]
    zu = -(Hinv @ q)
    tu = G @ zu
    tolp = tol_p * d.sc_b

    def viol(t, act):
        vu = np.where(hu, t - up, -np.inf)
        vl = np.where(hl, lo - t, -np.inf)
        for (r, s) in act:
            if s > 0:
                vu[r] = -np.inf
            else:
                vl[r] = -np.inf
        iu, il = int(np.argmax(vu)), int(np.argmax(vl))
        return (vu[iu], iu, 1) if vu[iu] >= vl[il] else (vl[il], il, -1)

    v0, _, _ = viol(tu, [])
    if v0 <= 0.0:
        return zu, OPTIMAL, dict(info, path="unconstrained", active=[])

    def bound(r, s):
        return up[r] if s > 0 else -lo[r]

    act, lam = [], np.zeros(0)
    # ---- warm start: equality solve on the shifted set, drop negative multipliers ---------------
    if warm:
        cand = [(int(shift[r]), s) for r, s in warm if shift[r] >= 0] if shift is not None else list(warm)
        cand = [(r, s) for r, s in cand if (hu[r] if s > 0 else hl[r])]
        act = cand
        for _ in range(2 * n):
            if not act:
                lam = np.zeros(0)
                break
            rows = np.array([a[0] for a in act])
            sg = np.array([a[1] for a in act], float)
            S = (sg[:, None] * W[np.ix_(rows, rows)]) * sg[None, :]
            L, keep = _chol_dd(S)
            rhs = sg * tu[rows] - np.array([bound(r, s) for r, s in act])
            lam = _ssolve(L, keep, rhs)
            neg = lam < -1e-9 * (1.0 + np.abs(lam).max(initial=0.0))
            bad = neg | ~keep
            if not bad.any():
                break
            if drop_all:
                act = [a for a, b in zip(act, bad) if not b]
            else:
                j = int(np.argmin(np.where(keep, lam, -np.inf))) if (~keep).any() else int(np.argmin(lam))
                act.pop(j)
            info["drops"] += 1
        lam = np.maximum(lam, 0.0)
        info["warm_kept"] = len(act)
    # t = G z for z = zu - Y_A' (s lam)
    rows = np.array([a[0] for a in act], int)
    sg = np.array([a[1] for a in act], float)
    t = tu - W[:m, rows] @ (sg * lam) if len(act) else tu.copy()

    # ---- Goldfarb-Idnani -------------------------------------------------------------------------
    it = 0
    status = None
    while True:
        vmax, p, sp = viol(t, act)
        if vmax <= tolp:
            break
        lam_p = 0.0
        while True:
            it += 1
            if it > max_iter:
                status = FALLBACK
                break
            na = len(act)
            if na:
                rows = np.array([a[0] for a in act])
                sg = np.array([a[1] for a in act], float)
                S = (sg[:, None] * W[np.ix_(rows, rows)]) * sg[None, :]
                L, keep = _chol_dd(S)
                v = sg * W[rows, p] * sp
                r = _ssolve(L, keep, v)
                kappa = W[p, p] - v @ r
            else:
                rows = np.zeros(0, int)
                sg = np.zeros(0)
                r = np.zeros(0)
                kappa = W[p, p]
            cp = (t[p] - up[p]) if sp > 0 else (lo[p] - t[p])
            dependent = kappa <= kappa_eps * W[p, p]
            # dual ratio test
            t1, j1 = np.inf, -1
            for j in range(na):
                if r[j] > 1e-13 * (1.0 + np.abs(r).max()):
                    ratio = lam[j] / r[j]
                    if ratio < t1:
                        t1, j1 = ratio, j
            if dependent:
                if j1 < 0 or na >= n + 1:
                    status = FALLBACK     # primal infeasible by Goldfarb-Idnani's test: let the IPM certify it
                    break
                step = t1
                full = False
            else:
                t2 = cp / kappa
                full = t2 <= t1
                step = t2 if full else t1
            # G d = s_p W[:,p] - W[:,A] (s_A r)
            Gd = sp * W[:m, p] - (W[:m, rows] @ (sg * r) if na else 0.0)
            if not dependent:
                t = t - step * Gd
            lam = lam - step * r
            lam_p += step
            if full:
                if na >= n:
                    status = FALLBACK
                    break
                act.append((p, sp))
                lam = np.r_[lam, lam_p]
                break
            lam[j1] = 0.0
            act.pop(j1)
            lam = np.delete(lam, j1)
            info["drops"] += 1
        if status is not None:
            break
    info["iters"] = it
    if status == FALLBACK:
        return None, FALLBACK, info
    # ---- certification: fresh solve with refinement on the final set ------------------------------
    z, rounds, why, fin = polish_model(d, x_init, ref, act, max_rounds=8)
    info["rounds"] = rounds
    if z is None:
        info["why"] = why
        return None, FALLBACK, info
    info["active"] = fin
    return z, OPTIMAL, info


# ------------------------------------------------------------------------------------------------------
# Version 2 of the kernel: the explicit inverse M = S^-1 of the working set's Schur complement is kept up to
# date by bordering (add) / rank-one downdates (drop); no factorisation anywhere.  Certification = iterative
# refinement of the multipliers against the true rows of G with M as the approximate inverse.
# ------------------------------------------------------------------------------------------------------
def solve_as_inv(d, W, x_init, ref, warm=None, shift=None, max_iter=224, tol_p=1e-11, kappa_eps=1e-11, passes=3,
                 restarts=3, _it0=0, _nrestart=0):
    n, m, G, q, lo, up, hl, hu = _problem(d, x_init, ref)
    info = dict(path="as", iters=0, drops=0, rounds=0, active=None, why=None, restarts=_nrestart)
    if np.any(d.par_C @ x_init - d.par_h > 1e-9 * (1.0 + np.abs(d.par_h))):
        return np.full(n, np.nan), INFEASIBLE, dict(info, path="param_rows")
    Hinv = d.Hinv[:n, :n]
    Y = d.Y[:m, :n]
    zu = -(Hinv @ q)
    tu = G @ zu
    tolp = tol_p * d.sc_b
    upi = np.where(hu, up, 1e30)
    loi = np.where(hl, lo, -1e30)
    if max((tu - upi).max(), (loi - tu).max()) < 0:
        return zu, OPTIMAL, dict(info, path="unconstrained", active=[])

    act = []                 # [(row, sign)]
    M = np.zeros((0, 0))
    lam = np.zeros(0)

    def border(M, r, kappa):
        na = M.shape[0]
        ik = 1.0 / kappa
        Mn = np.zeros((na + 1, na + 1))
        Mn[:na, :na] = M + np.outer(r, r) * ik
        Mn[:na, na] = -r * ik
        Mn[na, :na] = -r * ik
        Mn[na, na] = ik
        return Mn

    def downdate(M, j):
        Mn = M - np.outer(M[:, j], M[j, :]) / M[j, j]
        keep = [i for i in range(M.shape[0]) if i != j]
        return Mn[np.ix_(keep, keep)]

    def sv(act):
        rows = np.array([a[0] for a in act], int)
        sg = np.array([a[1] for a in act], float)
        return rows, sg

    if warm:
        cand = [(int(shift[r]), s) for r, s in warm if shift[r] >= 0] if shift is not None else list(warm)
        cand = [(r, s) for r, s in cand if (hu[r] if s > 0 else hl[r])]
        for (p, sp) in cand:
            rows, sg = sv(act)
            v = sg * W[rows, p] * sp if act else np.zeros(0)
            r = M @ v
            kappa = W[p, p] - v @ r
            if kappa > kappa_eps * W[p, p]:
                M = border(M, r, kappa)
                act.append((p, sp))
        while act:
            rows, sg = sv(act)
            rhs = sg * tu[rows] - np.where(sg > 0, up[rows], -lo[rows])
            lam = M @ rhs
            j = int(np.argmin(lam))
            if lam[j] >= -1e-9 * (1.0 + np.abs(lam).max()):
                break
            M = downdate(M, j)
            act.pop(j)
            info["drops"] += 1
        lam = np.maximum(lam, 0.0) if act else np.zeros(0)
    rows, sg = sv(act)
    t = tu - W[:m, rows] @ (sg * lam) if act else tu.copy()

    it = _it0
    for refresh in range(4):
        status = None
        while True:
            vu = t - upi
            vl = loi - t
            for (r_, s_) in act:
                if s_ > 0:
                    vu[r_] = -np.inf
                else:
                    vl[r_] = -np.inf
            iu, il = int(np.argmax(vu)), int(np.argmax(vl))
            vmax, p, sp = (vu[iu], iu, 1) if vu[iu] >= vl[il] else (vl[il], il, -1)
            if vmax <= tolp:
                break
            cp = vmax
            lam_p = 0.0
            while True:
                it += 1
                if it > max_iter:
                    status = FALLBACK
                    info["why"] = "gi-cap"
                    break
                na = len(act)
                rows, sg = sv(act)
                v = sg * W[rows, p] * sp if na else np.zeros(0)
                r = M @ v
                kappa = W[p, p] - v @ r
                dependent = not (kappa > kappa_eps * W[p, p]) or na >= n     # n independent rows span everything
                t1, j1 = np.inf, -1
                rmax = np.abs(r).max(initial=0.0)
                for j in range(na):
                    if r[j] > 1e-13 * (1.0 + rmax) and lam[j] / r[j] < t1:
                        t1, j1 = lam[j] / r[j], j
                if dependent:
                    if j1 < 0:
                        status = INFEASIBLE if cp > 1e-6 * d.sc_b else FALLBACK
                        info["why"] = "gi-dep cp=%.2e kappa=%.2e na=%d" % (cp, kappa / W[p, p], na)
                        break
                    step, full = t1, False
                else:
                    t2 = cp / kappa
                    full = not (j1 >= 0 and t1 < t2)
                    step = t2 if full else t1
                lam = lam - step * r
                lam_p += step
                if not dependent:
                    t = t - step * (sp * W[:m, p] - (W[:m, rows] @ (sg * r) if na else 0.0))
                    cp -= step * kappa
                if full:
                    if na >= n:
                        status = FALLBACK
                        info["why"] = "gi-full"
                        break
                    M = border(M, r, kappa)
                    act.append((p, sp))
                    lam = np.r_[lam, lam_p]
                    break
                M = downdate(M, j1)
                act.pop(j1)
                lam = np.delete(lam, j1)
                info["drops"] += 1
            if status is not None:
                break
        info["iters"] = it
        def restart():
            return solve_as_inv(d, W, x_init, ref, warm=list(act), shift=None, max_iter=max_iter, tol_p=tol_p,
                                kappa_eps=kappa_eps, passes=passes, restarts=restarts, _it0=it, _nrestart=_nrestart + 1)
        if status is not None and str(info["why"]).startswith("gi-dep") and it > _it0 and _nrestart < restarts:
            # the explicit inverse has drifted: refactor from the current working set and carry on
            return solve_as_inv(d, W, x_init, ref, warm=list(act), shift=None, max_iter=max_iter, tol_p=tol_p,
                                kappa_eps=kappa_eps, passes=passes, restarts=restarts, _it0=it, _nrestart=_nrestart + 1)
        if status == INFEASIBLE:
            return np.full(n, np.nan), INFEASIBLE, info
        if status == FALLBACK:
            return None, FALLBACK, info
        # certification: refinement against the true rows with M as approximate inverse
        info["rounds"] += 1
        rows, sg = sv(act)
        na = len(act)
        b = np.where(sg > 0, up[rows], -lo[rows]) if na else np.zeros(0)
        Gt = sg[:, None] * G[rows] if na else np.zeros((0, n))
        Yt = sg[:, None] * Y[rows] if na else np.zeros((0, n))
        z = zu.copy()
        lam_acc = np.zeros(na)
        resid = Gt @ z - b
        for _ in range(passes):
            dl = M @ resid
            lam_acc += dl
            z = z - Yt.T @ dl
            resid = Gt @ z - b
        if na and np.abs(resid).max() > tolp:
            info["why"] = "refine %.2e" % np.abs(resid).max()
            if _nrestart < restarts:
                return restart()
            return None, FALLBACK, info
        t = G @ z
        if na and lam_acc.min() < -1e-9 * (1.0 + np.abs(lam_acc).max()):
            info["why"] = "neg"
            if _nrestart < restarts:
                return restart()
            return None, FALLBACK, info
        lam = np.maximum(lam_acc, 0.0)
        vu = t - upi
        vl = loi - t
        for (r_, s_) in act:
            if s_ > 0:
                vu[r_] = -np.inf
            else:
                vl[r_] = -np.inf
        if max(vu.max(), vl.max()) <= tolp:
            info["active"] = list(act)
            return z, OPTIMAL, info
    info["why"] = "refresh"
    return None, FALLBACK, info

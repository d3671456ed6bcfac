"""Development aid (NOT product, NOT oracle): numpy model of the dual active-set kernel.

Goldfarb-Idnani dual active-set iteration on the scaled two-sided problem of
``rtmpc_b200.ipm_data.prepare``, on the operators shared by the whole batch
(``Hinv``, ``Y = G Hinv``, ``W = G Hinv G'``), warm-started from the previous control step's
active set moved one stage earlier, certified at the end by one round of the endgame
(``ipm_model.polish_model``).  Mirrors ``csrc/rtmpc_as.cuh``.
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

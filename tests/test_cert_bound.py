"""The rounding bound behind the tiered certification (csrc/rtmpc_as.cuh::as_certify, csrc/rtmpc_capi.cu::rtmpc_qp_create),
restated in numpy: kap_i * (1 + |x|_1 + |ref|_1 + sum |lam|) must cover the distance between a row value evaluated through the
factored tables (Ex x + Tr r - up0 - W[:,A] (s lam), float64) and the row of G z - up for the z the kernel forms
(z = Zx x + Zr r - Y_A' (s lam), float64), the latter evaluated in extended precision.  CPU only: the tables are rebuilt here the
way the library builds them."""
import numpy as np
import pytest

import helpers as H

LD = np.longdouble


def _tables(spec, Kss):
    from rtmpc_b200.qp import desc_arrays
    cq, d, ints, fl, arr = desc_arrays(spec, Kss=Kss)
    G, Y, Hinv, Fx, Fr, Ux = (arr[k] for k in ("G", "Y", "Hinv", "Fx", "Fr", "Ux"))
    up0, has_up = arr["up0"], arr["has_up"].astype(bool)
    m = ints["m"]
    Zx, Zr = -(Hinv @ Fx), -(Hinv @ Fr)                          # what the kernels compute z_u from
    W = G @ Y.T
    W = np.triu(W) + np.triu(W, 1).T                             # (the library mirrors the upper triangle)
    Gl = G.astype(LD)
    Ex = (Gl @ Zx.astype(LD) - Ux.astype(LD)).astype(np.float64)  # accumulated in long double, rounded once
    Tr = (Gl @ Zr.astype(LD)).astype(np.float64)
    u = LD(2.0) ** -53
    zy = max(np.abs(Zx).max(), np.abs(Zr).max(), np.abs(Y).max())
    dmax = np.maximum.reduce([np.abs(Gl @ Zx.astype(LD) - Ux.astype(LD) - Ex.astype(LD)).max(1),
                              np.abs(Gl @ Zr.astype(LD) - Tr.astype(LD)).max(1),
                              np.abs(Gl @ Y.astype(LD).T - W.T.astype(LD)).max(1)])      # W[a][i] is read for row i
    big = np.maximum.reduce([np.where(has_up, np.abs(up0), 0.0), np.abs(Ex).max(1), np.abs(Tr).max(1), np.abs(W).max(0)])
    kap = (2 * (64 * u * (big.astype(LD) + np.abs(Gl).sum(1) * LD(zy)) + dmax)).astype(np.float64)
    return dict(G=G, Y=Y, Zx=Zx, Zr=Zr, Ex=Ex, Tr=Tr, W=W, up0=np.where(has_up, up0, 1e30), Ux=Ux, kap=kap, m=m,
                sc_b=fl["sc_b"], n=ints["n"])


@pytest.mark.parametrize("sets,n_act", [("sets_cp.npz", 6), ("sets_cp.npz", 19), ("sets_di.npz", 4)])
def test_factored_row_values_stay_within_the_bound(sets, n_act):
    s = H.load(sets)
    T = _tables(H.spec_tube_tracking(s), s["K"])
    rng = np.random.default_rng(n_act)
    m, nx = T["m"], T["Zx"].shape[1]
    worst = 0.0
    for trial in range(200):
        x = rng.uniform(-1.0, 1.0, nx) * rng.choice([0.05, 0.5, 2.0])
        r = np.zeros(nx)
        r[0] = rng.uniform(-2.0, 2.0)
        A = rng.choice(m, size=n_act, replace=False)
        sgn = rng.choice([-1.0, 1.0], size=n_act)
        lam = np.abs(rng.standard_normal(n_act)) * 10.0 ** rng.uniform(-3, 5)       # up to the 1e5 .. 1e6 of the hard cases
        # the kernel's z (float64) and both evaluations of the rows
        z = T["Zx"] @ x + T["Zr"] @ r
        for a, sa, la in zip(A, sgn, lam):
            z = z - (sa * la) * T["Y"][a]
        e_fact = -T["up0"] + T["Ex"] @ x + T["Tr"] @ r
        for a, sa, la in zip(A, sgn, lam):
            e_fact = e_fact + (-sa * la) * T["W"][a]
        e_true = (T["G"].astype(LD) @ z.astype(LD) - T["up0"].astype(LD) - T["Ux"].astype(LD) @ x.astype(LD))
        S = 1.0 + np.abs(x).sum() + np.abs(r).sum() + np.abs(lam).sum()
        rows = np.arange(m)
        ratio = (np.abs(e_fact.astype(LD) - e_true)[rows] / (T["kap"][rows] * S)).astype(np.float64)
        worst = max(worst, ratio.max())
    assert worst <= 1.0, worst            # the bound holds ...
    assert worst <= 0.5                    # ... with the factor two it was given


def test_tables_are_consistent_to_rounding():
    """Ex built as G Zx - Ux in long double and rounded once: what it differs by from G Zx - Ux is rounding of its own entries;
    Ex = -Y Fx - Ux (the first version of the table) differed from it by the rounding of Y = G Hinv, more than the row
    tolerance of the problem per unit of x."""
    from rtmpc_b200.qp import desc_arrays
    s = H.load("sets_cp.npz")
    spec = H.spec_tube_tracking(s)
    T = _tables(spec, s["K"])
    _, _, _, fl, arr = desc_arrays(spec, Kss=s["K"])
    ideal = T["G"].astype(LD) @ T["Zx"].astype(LD) - T["Ux"].astype(LD)
    d_new = np.abs(T["Ex"].astype(LD) - ideal).max()
    d_old = np.abs((-(arr["Y"] @ arr["Fx"]) - arr["Ux"]).astype(LD) - ideal).max()
    tolp = 1e-11 * fl["sc_b"]
    assert d_new <= 2.0 ** -52 * np.abs(T["Ex"]).max()
    assert d_old > tolp > 100 * d_new

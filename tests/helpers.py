"""Shared test helpers: load golden fixtures and turn them into product-side problem specs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def spec_tube_tracking(s, fixed_initial_state=True):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["Xc_A"], s["Xc_b"]), stage_u=(s["Uc_A"], s["Uc_b"]), terminal=(s["Xf_A"], s["Xf_b"]),
                   tube_init=None if fixed_initial_state else (s["Z_A"], s["Z_b"]))


def spec_ext_received(s, strict=False):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["Xc_A"], s["Xc_b"]), stage_u=(s["Uc_A"], s["Uc_b"]), terminal=(s["Xf_A"], s["Xf_b"]),
                   tube_init=(s["ZmW_A"], s["ZmW_b"]), g2_free_terminal=not strict)


def spec_tracking(s):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["X_A"], s["X_b"]), stage_u=(s["U_A"], s["U_b"]), terminal=(s["Xf_track_A"], s["Xf_track_b"]))


# ---- numpy restatement of the device RNG (Philox4x32-10 + 53-bit uniforms), for parity of draws ----
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & 0xFFFFFFFF for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0 & 0xFFFFFFFF)
    k1 = np.uint64(k1 & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def u01(hi, lo):
    v = ((hi >> np.uint64(5)) << np.uint64(26)) | (lo >> np.uint64(6))
    return v.astype(np.float64) / 9007199254740992.0


def device_draws(seed, ids, T, p_loss, w_half):
    """theta[T,B], gamma[T,B], w[T,B,nx] exactly as loop_step_kernel draws them."""
    ids = np.asarray(ids, dtype=np.uint64)
    B, nx = ids.size, len(w_half)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    theta = np.ones((T, B), np.int32)
    gamma = np.ones((T, B), np.int32)
    w = np.zeros((T, B, nx))
    idlo, idhi = ids & np.uint64(0xFFFFFFFF), ids >> np.uint64(32)
    for t in range(T):
        tt = np.full(B, t, np.uint64)
        r = philox4x32_10(idlo, idhi, tt, np.zeros(B, np.uint64), k0, k1)
        if t > 0:
            theta[t] = np.where(u01(r[0], r[1]) < p_loss, 0, 1)
            gamma[t] = np.where(u01(r[2], r[3]) < p_loss, 0, 1)
        for k in range(0, nx, 2):
            q = philox4x32_10(idlo, idhi, tt, np.full(B, 1 + k // 2, np.uint64), k0, k1)
            w[t, :, k] = w_half[k] * (2.0 * u01(q[0], q[1]) - 1.0)
            if k + 1 < nx:
                w[t, :, k + 1] = w_half[k + 1] * (2.0 * u01(q[2], q[3]) - 1.0)
    return theta, gamma, w


def poly(s, k):
    from rtmpc_b200.polytope import Polytope
    return Polytope(s[k + "_A"], s[k + "_b"], normalize=False)


def make_tube_mpc(s, extended=False, fixed_initial_state=True):
    """Controller object of the product with the fixture's sets loaded (skips the slow set computations)."""
    from rtmpc_b200 import mpc
    cls = mpc.ExtendedTubeTrackingMPC if extended else mpc.TubeTrackingMPC
    c = cls(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(poly(s, "U"))
    c.set_state_constraints(poly(s, "X"))
    if extended:
        c.load_sets(poly(s, "Z"), poly(s, "Xc"), poly(s, "Uc"), poly(s, "Xf"), ZmW=poly(s, "ZmW"),
                    fixed_initial_state=fixed_initial_state)
    else:
        c.load_sets(poly(s, "Z"), poly(s, "Xc"), poly(s, "Uc"), poly(s, "Xf"), fixed_initial_state=fixed_initial_state)
    return c


def make_track_mpc(s):
    from rtmpc_b200 import mpc
    c = mpc.TrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
    c.set_input_constraints(poly(s, "U"))
    c.set_state_constraints(poly(s, "X"))
    c._Xf = poly(s, "Xf_track")
    c.generate_optimization_problem()
    return c


def oracle_tube_tracking_qp(s):
    """The oracle's (un-condensed) statement of the remote tube MPC problem for a sets_*.npz fixture."""
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope
    P = lambda k: Polytope(s[k + "_A"], s[k + "_b"], normalize=False)      # noqa: E731
    return rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], P("Xc"), P("Uc"), P("Xf"), None, True)


def kkt_certificate(qp, x_init, ref, z, act_tol=1e-8):
    """Solver-independent optimality check of ``z`` for the oracle problem ``qp`` (min 1/2 z'Pz + q'z, Ez = e, Gz <= h).
    Returns (primal violation, stationarity residual) with both scaled like the oracle's own stopping test: the
    multipliers y (free) and lambda >= 0 on the rows active at z are fitted by bounded least squares."""
    from scipy.optimize import lsq_linear
    q, e, h = qp.params(x_init, ref)
    z = z[:qp.P.shape[0]]
    sc_q = 1.0 + np.abs(q).max() + np.abs(qp.P).max()
    sc_h = 1.0 + np.abs(h).max()
    sc_e = 1.0 + np.abs(e).max()
    slack = h - qp.G @ z
    primal = max((-slack).max() / sc_h, np.abs(qp.E @ z - e).max() / sc_e)
    act = np.nonzero(slack <= act_tol * sc_h)[0]
    g = qp.P @ z + q
    Amat = np.c_[qp.E.T, qp.G[act].T]
    # columns scaled to unit norm: the fit is for the residual, not the multipliers
    cn = np.maximum(np.linalg.norm(Amat, axis=0), 1e-300)
    lb = np.r_[np.full(qp.E.shape[0], -np.inf), np.zeros(len(act))]
    r = lsq_linear(Amat / cn, -g, bounds=(lb, np.full(len(lb), np.inf)), method="bvls", tol=1e-15, max_iter=2000)
    return primal, np.abs(Amat / cn @ r.x + g).max() / sc_q

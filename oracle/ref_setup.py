"""Oracle (test infrastructure): the reference's ``setup_optimization`` pipelines end to end,
plus the two benchmark systems of the reference's scripts.

``tube_tracking_setup`` follows ``TubeTrackingMPC.setup_optimization`` (``TubeTrackingMPC.py:158-168``):
mRPI -> tighten -> terminal set -> QP; ``tracking_setup`` follows ``TrackingMPC.setup_optimization``
(``TrackingMPC.py:188-193``); ``tube_regulator_setup`` follows ``TubeRegulatorMPC.setup_optimization``
(``TubeRegulatorMPC.py:145-154``).
"""
import numpy as np

from . import ref_numerics as rn
from . import ref_qp as rq
from . import ref_sets as rs
from .ref_polytope import Polytope


def box(lo_hi):
    """{x : -h_i <= x_i <= h_i} written [I; -I] like the reference's scripts."""
    h = np.asarray(lo_hi, dtype=float).flatten()
    n = h.size
    return Polytope(np.r_[np.eye(n), -np.eye(n)], np.r_[h, h])


def double_integrator():
    """Config 1 ("Examples of Model Predictive Controllers/Example_of_Tube_Tracking_MPC_Over_Lossy_Network.py":25-51)."""
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.0], [1.0]])
    return dict(A=A, B=B, Q=np.eye(2), R=np.eye(1), N=10, X=box([8, 8]), U=box([1]), W=box([0.1, 0.1]),
                rpi_method=0)


def cartpole_matrices(Th=0.02):
    """``Results/results_linear_system.py:26-61``."""
    M, m, b, I, g, l = 1.0, 0.1, 0.0, 0.001, 9.8, 0.5
    p = I * (M + m) + M * m * l ** 2
    Ac = np.array([[0, 1, 0, 0],
                   [0, -(I + m * l ** 2) * b / p, -(m ** 2 * g * l ** 2) / p, 0],
                   [0, 0, 0, 1],
                   [0, -(m * l * b) / p, m * g * l * (M + m) / p, 0]])
    Bc = np.array([[0], [(I + m * l ** 2) / p], [0], [-m * l / p]])
    return rn.c2d_zoh(Ac, Bc, Th)


def linear_cartpole():
    """Configs 2-4 (``Results/results_linear_system.py:26-110``)."""
    A, B = cartpole_matrices()
    return dict(A=A, B=B, Q=np.diag([100.0, 10.0, 100.0, 10.0]), R=0.1 * np.eye(1), N=20,
                X=box([5, 5, 0.3, 2]), U=box([10]), W=box([1e-4, 2.7e-3, 3e-4, 4.3e-2]), rpi_method=1)


def tube_tracking_setup(A, B, Q, R, N, X, U, W, rpi_method=0, fixed_initial_state=False,
                        lambda_param=0.99999, K_ancillary=None, extended=False, epsilon=1e-4, **_):
    """Returns a dict with K, P, Acl, Z, Xc, Uc, Xf, qp (and ZmW, qp_recv when ``extended``).
    NB ``TubeTrackingMPC.determine_mRPI`` passes its own default ``epsilon=1e-4`` positionally as
    ``eps_var`` (``TubeTrackingMPC.py:63,86``)."""
    K, P, Acl = rn.lqr_terminal_data(A, B, Q, R)
    Kanc = K if K_ancillary is None else K_ancillary
    Acl_plant = Acl if K_ancillary is None else A - B @ K_ancillary
    Z = rs.determine_mrpi(Acl_plant, W, X, U, Kanc, eps_var=epsilon, rpi_method=rpi_method)
    Xc, Uc = rs.tighten(X, U, Z, Kanc)
    Xf, t_star = rs.tracking_terminal_set(A, B, K, Xc, Uc, lambda_param)
    qp = rq.build_tube_tracking(A, B, Q, R, N, P, Xc, Uc, Xf, Z, fixed_initial_state)
    out = dict(A=A, B=B, Q=Q, R=R, N=N, K=K, P=P, Acl=Acl, K_anc=Kanc, Z=Z, Xc=Xc, Uc=Uc, Xf=Xf,
               t_star=t_star, qp=qp, W=W, X=X, U=U)
    if extended:
        ZmW = rs.pont_diff(Z, W)
        out["ZmW"] = ZmW
        out["qp_recv"] = rq.build_extended_packet_received(A, B, Q, R, N, P, Xc, Uc, Xf, ZmW)
    return out


def tracking_setup(A, B, Q, R, N, X, U, lambda_param=0.99999, **_):
    K, P, Acl = rn.lqr_terminal_data(A, B, Q, R)
    Xf, t_star = rs.tracking_terminal_set(A, B, K, X, U, lambda_param)
    qp = rq.build_tracking(A, B, Q, R, N, P, X, U, Xf)
    return dict(A=A, B=B, Q=Q, R=R, N=N, K=K, P=P, Acl=Acl, Xf=Xf, t_star=t_star, qp=qp, X=X, U=U)


def tube_regulator_setup(A, B, Q, R, N, X, U, W, eps_var=1.9e-5, **_):
    K, P, Acl = rn.lqr_terminal_data(A, B, Q, R)
    Z = rs.determine_mrpi(Acl, W, X, U, K, eps_var=eps_var, rpi_method=0)
    Xc, Uc = rs.tighten(X, U, Z, K)
    Xf, t_star = rs.regulator_terminal_set(Acl, K, Xc, Uc)
    qp = rq.build_tube_regulator(A, B, Q, R, N, P, Xc, Uc, Xf, Z)
    return dict(A=A, B=B, Q=Q, R=R, N=N, K=K, P=P, Acl=Acl, Z=Z, Xc=Xc, Uc=Uc, Xf=Xf, t_star=t_star, qp=qp)

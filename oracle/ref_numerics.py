"""Oracle (test infrastructure): the python-control calls of the reference, restated on scipy.

python-control (pin ``control>=0.9.3.post2``, reference ``setup.cfg:14``) is absent from this
image.  Its three entry points used on the path are restated from their documented
conventions (recalled from upstream docs; cannot be re-verified offline):

* ``ct.dlqr(A,B,Q,R) -> K, S, E``  with  u = -K x   (``TubeRegulatorMPC.py:19``, ``TrackingMPC.py:25``)
* ``ct.dlyap(A,Q)`` solves  A X A^T - X + Q = 0  (``TubeRegulatorMPC.py:23``, ``TrackingMPC.py:31``)
  -- note the reference passes the *untransposed* closed-loop matrix.
* ``ct.c2d(ct.ss(Ac,Bc,Cc,0), Th)`` zero-order hold (``Results/results_linear_system.py:59-61``)
"""
import numpy as np
import scipy.linalg as sla


def dlqr(A, B, Q, R):
    """K, S with u = -K x; S solves the DARE.  (``TubeRegulatorMPC.py:19``)"""
    A = np.asarray(A, dtype=float)
    B = np.asarray(B, dtype=float)
    Q = np.asarray(Q, dtype=float)
    R = np.atleast_2d(np.asarray(R, dtype=float))
    S = sla.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ S @ B, B.T @ S @ A)
    return K, S


def dlyap(A, Q):
    """X with A X A^T - X + Q = 0 (python-control convention, ``TubeRegulatorMPC.py:23``)."""
    return sla.solve_discrete_lyapunov(np.asarray(A, dtype=float), np.asarray(Q, dtype=float))


def c2d_zoh(Ac, Bc, Th):
    """Zero-order-hold discretisation (``Results/results_linear_system.py:59-61``)."""
    Ac = np.asarray(Ac, dtype=float)
    Bc = np.asarray(Bc, dtype=float)
    nx, nu = Bc.shape
    M = np.zeros((nx + nu, nx + nu))
    M[:nx, :nx] = Ac
    M[:nx, nx:] = Bc
    E = sla.expm(M * Th)
    return E[:nx, :nx].copy(), E[:nx, nx:].copy()


def lqr_terminal_data(A, B, Q, R):
    """K, P, Acl exactly as ``TubeRegulatorMPC.__init__`` builds them (``TubeRegulatorMPC.py:16-24``)."""
    K, _ = dlqr(A, B, Q, R)
    Ql = Q + K.T @ np.atleast_2d(R) @ K
    Ql = (Ql + Ql.T) / 2
    Acl = A - B @ K
    P = dlyap(Acl, Ql)
    return K, P, Acl

"""Gains and discretisation used by the controller classes' constructors (host, once per system).

Stand-ins for the three python-control calls of the reference (``ct.dlqr`` ``TubeRegulatorMPC.py:19``,
``ct.dlyap`` ``:23`` -- convention  A X A' - X + Q = 0, SURVEY G1 -- and ``ct.c2d`` for the cartpole,
``Results/results_linear_system.py:59-61``), on scipy.
"""
import numpy as np
from scipy.linalg import expm, solve_discrete_are, solve_discrete_lyapunov


def dlqr(A, B, Q, R):
    """(K, S, E) with u = -K x."""
    A, B, Q = (np.asarray(M, float) for M in (A, B, Q))
    R = np.atleast_2d(np.asarray(R, float))
    S = solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ S @ B, B.T @ S @ A)
    return K, S, np.linalg.eigvals(A - B @ K)


def dlyap(A, Q, transpose_convention=False):
    """X with A X A' - X + Q = 0 (python-control's convention; the reference passes A - BK
    untransposed).  ``transpose_convention=True`` gives the textbook A' X A - X + Q = 0."""
    A = np.asarray(A, float)
    return solve_discrete_lyapunov(A.T if transpose_convention else A, np.asarray(Q, float))


def c2d(Ac, Bc, Th):
    Ac, Bc = np.asarray(Ac, float), np.asarray(Bc, float)
    nx, nu = Bc.shape
    M = np.zeros((nx + nu, nx + nu))
    M[:nx, :nx], M[:nx, nx:] = Ac, Bc
    E = expm(M * Th)
    return E[:nx, :nx].copy(), E[:nx, nx:].copy()


def cartpole_linear(Th=0.02, M=1.0, m=0.1, b=0.0, I=0.001, g=9.8, l=0.5):
    """Linearised cartpole of ``Results/results_linear_system.py:26-61``."""
    p = I * (M + m) + M * m * l ** 2
    Ac = np.array([[0, 1, 0, 0], [0, -(I + m * l ** 2) * b / p, -(m ** 2 * g * l ** 2) / p, 0],
                   [0, 0, 0, 1], [0, -(m * l * b) / p, m * g * l * (M + m) / p, 0]])
    Bc = np.array([[0], [(I + m * l ** 2) / p], [0], [-m * l / p]])
    return c2d(Ac, Bc, Th)

"""Oracle (test infrastructure): stand-in for the absent third-party package ``control`` (pin
``control>=0.9.3.post2``, reference ``setup.cfg:14``), backed by ``oracle/ref_numerics.py`` (scipy).

* ``ct.dlqr(A, B, Q, R) -> K, S, E`` with ``u = -K x``  (``TubeRegulatorMPC.py:19``, ``TrackingMPC.py:25``)
* ``ct.dlyap(A, Q)`` solving ``A X A' - X + Q = 0``      (``TubeRegulatorMPC.py:23``, ``TrackingMPC.py:31``)
* ``ct.ss(A, B, C, D)`` / ``ct.c2d(sys, Ts)`` zero-order hold (``Results/results_linear_system.py:59-61``)
"""
import numpy as np

from oracle import ref_numerics as rn


def dlqr(A, B, Q, R):
    K, S = rn.dlqr(A, B, Q, R)
    E = np.linalg.eigvals(np.asarray(A, float) - np.asarray(B, float) @ K)
    return K, S, E


def dlyap(A, Q):
    return rn.dlyap(A, Q)


class StateSpace:
    def __init__(self, A, B, C, D, dt=None):
        self.A, self.B, self.C, self.D, self.dt = (np.atleast_2d(np.asarray(A, float)), np.atleast_2d(np.asarray(B, float)),
                                                   np.atleast_2d(np.asarray(C, float)), np.asarray(D, float), dt)


def ss(A, B, C, D, dt=None):
    return StateSpace(A, B, C, D, dt)


def c2d(sysc, Ts, method="zoh"):
    assert method == "zoh"
    Ad, Bd = rn.c2d_zoh(sysc.A, sysc.B, Ts)
    return StateSpace(Ad, Bd, sysc.C, sysc.D, Ts)

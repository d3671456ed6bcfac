"""Controller classes with the reference's names and call surface, solving on the GPU.

Mirrors ``RegulatorMPC`` (``RegulatorMPC.py``), ``TubeRegulatorMPC`` (``TubeRegulatorMPC.py``),
``TrackingMPC`` (``TrackingMPC.py``), ``TubeTrackingMPC`` and ``ExtendedTubeTrackingMPC``
(``TubeTrackingMPC.py``): same constructor arguments, same method names and positional order, same
return shapes, same "print the status / return None when infeasible" error behaviour.  Behind
``generate_optimization_problem`` there is no cvxpy problem but a :class:`rtmpc_b200.qp.BatchedQP`
(condensed on the host once, resident on the GPU); behind ``solve_optimization_problem`` a CUDA
kernel launch.  Every class additionally has ``*_batch`` methods taking ``[B, nx]`` arrays.

Single-instance calls keep the reference's habit of reshaping the caller's arrays in place
(``TubeTrackingMPC.py:175,179,200``).
"""
import time

import numpy as np

from . import _lib
from . import numerics
from . import sets as up
from .condense import MPCSpec
from .packets import as_polytope
from .polytope import Polytope, reduce as _reduce
from .qp import BatchedQP

SOLVER_NAME = "RTMPC_B200"


def _Ab(poly):
    return np.asarray(poly.A, float), np.asarray(poly.b, float).flatten()


def _ok(status):
    return status in (_lib.OPTIMAL, _lib.OPTIMAL_INACCURATE, _lib.MAX_ITER)


class RegulatorMPC:
    """``RegulatorMPC.py:9-94``."""

    def __init__(self, A, B, Q, R, N):
        self._A = A
        self._B = B
        self._N = int(N)
        self._nx = self._A.shape[1]
        self._nu = self._B.shape[1]
        self._Q = Q
        self._R = R
        self._X = None
        self._U = None
        self._solver = SOLVER_NAME       # reference: cp.CLARABEL (RegulatorMPC.py:31)
        self._prob = None
        self._last_status = None
        self._last_iters = None

    def set_state_constraints(self, X):
        """``RegulatorMPC.py:33-37``.  ``X``: anything with ``.A`` / ``.b`` (a ``polytope.Polytope``, this package's
        ``Polytope``, ...); rows are brought to unit length the way ``polytope.Polytope`` holds them."""
        self._X = as_polytope(X)

    def set_input_constraints(self, U):
        """``RegulatorMPC.py:39-43``; see ``set_state_constraints``."""
        self._U = as_polytope(U)

    def set_solver(self, solver):
        """Kept for API compatibility (``RegulatorMPC.py:93-94``); the CUDA solver is the only one."""
        self._solver = solver

    # -- problem ------------------------------------------------------------------------------
    def _spec(self):
        return MPCSpec(self._A, self._B, self._Q, self._R, self._N,
                       stage_x=None if self._X is None else _Ab(self._X),
                       stage_u=None if self._U is None else _Ab(self._U))

    def generate_optimization_problem(self):
        self._prob = BatchedQP(self._spec())

    # -- solve --------------------------------------------------------------------------------
    def _solve_one(self, prob, x_init, ref=None):
        # consecutive single-instance calls are consecutive control steps of one closed loop: keep the
        # warm-start state in the handle (it only changes the work, never the certified result)
        z, U, st, it = prob.solve_host(np.asarray(x_init, float).reshape(1, -1),
                                       None if ref is None else np.asarray(ref, float).reshape(1, -1), warm=True)
        self._last_status, self._last_iters = int(st[0]), int(it[0])
        return z, U

    def solve_batch(self, x_init):
        """[B,nx] -> x[B,nx,N+1], u[B,nu,N], status[B], iters[B]."""
        z, U, st, it = self._prob.solve_host(x_init)
        x, u = self._prob.split(z)
        return x, u, st, it

    def solve_optimization_problem(self, x_init):
        x_init.shape = (self._nx,)
        z, _ = self._solve_one(self._prob, x_init)
        if not _ok(self._last_status):
            return None, None
        x, u = self._prob.split(z)
        return x[0].copy(), u[0].copy()

    @property
    def status(self):
        return {0: "optimal", 1: "user_limit", 2: "infeasible", 3: "optimal_inaccurate", None: None}[self._last_status]


class TubeRegulatorMPC(RegulatorMPC):
    """``TubeRegulatorMPC.py:14-172`` (Mayne et al. tube MPC)."""

    def __init__(self, A, B, Q, R, N):
        super().__init__(A, B, Q, R, N)
        K, _, _ = numerics.dlqr(A, B, Q, R)
        self._K = K
        Q_lyap = self._Q + self._K.T @ np.atleast_2d(self._R) @ self._K
        self._P = numerics.dlyap(self._A - self._B @ self._K, (Q_lyap + Q_lyap.T) / 2)
        self._Acl = A - B @ K
        self._Z = None

    def determine_mRPI(self, W, eps_var=1.9e-5, Acl=None, rpi_method=0, K=None, skip_wasted_pass=False):
        """``TubeRegulatorMPC.py:26-78``: approximate the mRPI set with a cap ``s_max`` on the iterations, multiply the cap by
        10 and START OVER until the approximation succeeds, then ``pc.reduce``.  The reference starts at ``s_max = 200``
        (``:48``); for the cartpole (k* = 308) that first pass is wasted work.  ``skip_wasted_pass=True`` starts at 2000
        instead - the result is the same set (it does not depend on the cap once the cap is large enough), only the wasted
        pass and its "RPI not determined in 200 steps" print go away.  Default: the reference's behaviour."""
        if K is None:
            K = self._K
        if Acl is None:
            Acl = self._Acl
        if np.max(np.abs(np.linalg.eigvals(Acl))) >= 1:
            print("The matrix Acl is not stable, such that the algorithm will never converge. \n Therefore, None is returned")
            return None
        s_max = 2000 if skip_wasted_pass else 200
        while True:
            if rpi_method == 1:
                Fs_temp, status = up.calculate_RPI(Acl, W, self._X, self._U, K, eps_var=eps_var, s_max=s_max)
            else:
                if rpi_method != 0:
                    print("The method chosen to determine the RPI does not exists, so we use the default method 0")
                Fs_temp, status = up.calculate_minimal_robust_positively_invariant_set(Acl, W=W, eps_var=eps_var, s_max=s_max)
            if status == 0:
                break
            print(f"RPI not determined in {s_max} steps. Increasing s_max to 10*s_max = {10 * s_max}")
            s_max = 10 * s_max
        self._Z = _reduce(Fs_temp)
        return self._Z

    def tighten_constraints(self):
        self._Uc = up.pont_diff(self._U, up.scale(self._Z, -self._K))
        self._Xc = up.pont_diff(self._X, self._Z)

    def determine_Xf(self):
        Hu, hu = _Ab(self._Uc)
        Hx, hx = _Ab(self._Xc)
        XU = Polytope(np.r_[Hx, -Hu @ self._K], np.r_[hx, hu])
        self._Xf = up.calculate_maximum_admissible_output_set(self._Acl, XU)

    def _spec(self):
        return MPCSpec(self._A, self._B, self._Q, self._R, self._N, P_term=self._P, stage_x=_Ab(self._Xc),
                       stage_u=_Ab(self._Uc), terminal=_Ab(self._Xf), tube_init=_Ab(self._Z))

    def setup_optimization(self, W):
        self.determine_mRPI(W)
        self.tighten_constraints()
        self.determine_Xf()
        self.generate_optimization_problem()

    def get_controller_gain(self):
        return self._K

    def get_minimum_robust_positively_invariant_set(self):
        return self._Z


class _TrackingMixin:
    """Packet logic shared by TrackingMPC and TubeTrackingMPC."""

    def _packet(self, u_traj, x_bar, u_bar, q_t):
        if x_bar is not None:
            u_ss = u_bar + self._K @ x_bar
            u_ss.shape = (u_traj.shape[0], 1)
            U_t = np.hstack((u_traj, u_ss))
        else:
            U_t = None
        return {"U_t": U_t, "q_t": q_t}

    def solve_batch(self, x_init, ref, prob=None, sel=None, sel_value=1):
        """[B,nx],[B,nx] -> dict(x, u, x_bar, u_bar, U_t[B,N+1,nu], status, iters)."""
        prob = self._prob if prob is None else prob
        z, U, st, it = prob.solve_host(x_init, ref, sel=sel, sel_value=sel_value)
        x, u, xb, ub = prob.split(z)
        return dict(x=x, u=u, x_bar=xb, u_bar=ub, U_t=U, status=st, iters=it, z=z)

    def determine_packet_batch(self, x_hat, ref, q_t):
        """Batched ``determine_packet``: returns {'U_t': [B,N+1,nu] (NaN rows where infeasible),
        'q_t': [B], 'status': [B]}."""
        start = time.time()
        out = self.solve_batch(x_hat, ref)
        self._computational_times.append(time.time() - start)
        return {"U_t": out["U_t"], "q_t": np.asarray(q_t), "status": out["status"], "x_nom_0": out["x"][:, :, 0]}

    def get_steady_state_controller_gain(self):
        return self._K

    def get_computational_times(self):
        return self._computational_times

    def reset_computational_times(self):
        self._computational_times = []


def _augmented_terminal_set(A, B, K, Acl, Hx, hx, Hu, hu, lam, nx, nu):
    """MOAS of the (x, x_bar, u_bar) system (``TubeTrackingMPC.py:35-61`` / ``TrackingMPC.py:160-186``);
    like the reference, the zero blocks assume 2nx / 2nu rows."""
    A_e = np.block([[Acl, B @ K, B],
                    [np.zeros((nx, nx)), np.eye(nx), np.zeros((nx, nu))],
                    [np.zeros((nu, nx)), np.zeros((nu, nx)), np.eye(nu)]])
    Hcl = np.block([[Hx, np.zeros((2 * nx, nx)), np.zeros((2 * nx, nu))],
                    [-Hu @ K, Hu @ K, Hu],
                    [np.zeros((2 * nx, nx)), Hx, np.zeros((2 * nx, nu))],
                    [np.zeros((2 * nu, nx)), np.zeros((2 * nu, nx)), Hu]])
    hcl = np.r_[hx, hu, lam * hx, lam * hu]
    return up.calculate_maximum_admissible_output_set(A_e, Polytope(Hcl, hcl))


class TrackingMPC(_TrackingMixin, RegulatorMPC):
    """``TrackingMPC.py:19-198`` (Limon tracking MPC / Pezzutto remote MPC)."""

    def __init__(self, A, B, Q, R, N, lambda_param=0.99999):
        RegulatorMPC.__init__(self, A, B, Q, R, N)
        K, _, _ = numerics.dlqr(self._A, self._B, self._Q, self._R)
        self._K = K
        Q_lyap = self._Q + self._K.T @ np.atleast_2d(self._R) @ self._K
        self._P = numerics.dlyap(self._A - self._B @ self._K, (Q_lyap + Q_lyap.T) / 2)
        self._Tout = 10 * self._P
        self._lambda = lambda_param
        self._Acl = self._A - self._B @ self._K
        self._Xf = None
        self._computational_times = []

    def determine_packet(self, x_hat, ref, q_t):
        x_hat.shape = (self._nx,)
        start = time.time()
        _, u_mpc_traj, x_ss, u_ss = self.solve_optimization_problem(x_hat, ref)
        self._computational_times.append(time.time() - start)
        return self.encapsulate(u_mpc_traj, x_ss, u_ss, q_t)

    def _spec(self):
        return MPCSpec(self._A, self._B, self._Q, self._R, self._N, P_term=self._P, T_ss=self._Tout,
                       stage_x=None if self._X is None else _Ab(self._X),
                       stage_u=None if self._U is None else _Ab(self._U),
                       terminal=None if self._Xf is None else _Ab(self._Xf), terminal_eq=self._Xf is None)

    def generate_optimization_problem(self):
        self._prob = BatchedQP(self._spec(), Kss=self._K)

    def solve_optimization_problem(self, x_init, ref, verbose_MPC=False):
        ref.shape = (self._nx,)
        x_init.shape = (self._nx,)
        z, _ = self._solve_one(self._prob, x_init, ref)
        if self.status != "optimal":
            print(f"Status of tracking MPC is {self.status}")
        if not _ok(self._last_status):
            return None, None, None, None
        x, u, xb, ub = self._prob.split(z)
        return x[0].copy(), u[0].copy(), xb[0].copy(), ub[0].copy()

    def encapsulate(self, u_mpc, x_bar, u_bar, q_t):
        return self._packet(u_mpc, x_bar, u_bar, q_t)

    def determine_Xf(self):
        Hu, hu = _Ab(self._U)
        Hx, hx = _Ab(self._X)
        self._Xf = _augmented_terminal_set(self._A, self._B, self._K, self._Acl, Hx, hx, Hu, hu, self._lambda,
                                           self._nx, self._nu)

    def setup_optimization(self):
        self.determine_Xf()
        self.generate_optimization_problem()


class TubeTrackingMPC(_TrackingMixin, TubeRegulatorMPC):
    """``TubeTrackingMPC.py:20-246`` (Limon tube tracking MPC / remote tube MPC of Umsonst & Barbosa)."""

    def __init__(self, A, B, Q, R, N, lambda_param=0.99999):
        TubeRegulatorMPC.__init__(self, A, B, Q, R, N)
        self._lambda = lambda_param
        self._Tout = 10 * self._P
        self._K_ancillary = None
        self._Acl_plant = None
        self._computational_times = []

    def determine_Xf(self):
        Hu, hu = _Ab(self._Uc)
        Hx, hx = _Ab(self._Xc)
        self._Xf = _augmented_terminal_set(self._A, self._B, self._K, self._Acl, Hx, hx, Hu, hu, self._lambda,
                                           self._nx, self._nu)

    def determine_mRPI(self, W, epsilon=1e-4, Acl=None, rpi_method=0, skip_wasted_pass=False):
        if Acl is None:
            Acl = self._Acl if self._Acl_plant is None else self._Acl_plant
        K = self._K if self._K_ancillary is None else self._K_ancillary
        return TubeRegulatorMPC.determine_mRPI(self, W, epsilon, Acl=Acl, K=K, rpi_method=rpi_method,
                                               skip_wasted_pass=skip_wasted_pass)

    def tighten_constraints(self):
        K = self._K if self._K_ancillary is None else self._K_ancillary
        self._Uc = up.pont_diff(self._U, up.scale(self._Z, -K))
        self._Xc = up.pont_diff(self._X, self._Z)

    def _spec(self, fixed_initial_state=False):
        return MPCSpec(self._A, self._B, self._Q, self._R, self._N, P_term=self._P, T_ss=self._Tout,
                       stage_x=_Ab(self._Xc), stage_u=_Ab(self._Uc), terminal=_Ab(self._Xf),
                       tube_init=None if fixed_initial_state else _Ab(self._Z))

    def generate_optimization_problem(self, fixed_initial_state=False):
        self._prob = BatchedQP(self._spec(fixed_initial_state), Kss=self._K)

    def setup_optimization(self, W, fixed_initial_state=False, rpi_method=0, skip_wasted_pass=False):
        """``TubeTrackingMPC.py:158-168``; ``skip_wasted_pass``: see ``TubeRegulatorMPC.determine_mRPI``."""
        self.determine_mRPI(W, rpi_method=rpi_method, skip_wasted_pass=skip_wasted_pass)
        self.tighten_constraints()
        self.determine_Xf()
        self.generate_optimization_problem(fixed_initial_state)

    def load_sets(self, Z, Xc, Uc, Xf, fixed_initial_state=False):
        """Skip the (slow, host-side) set computations by supplying previously computed sets."""
        self._Z, self._Xc, self._Uc, self._Xf = Z, Xc, Uc, Xf
        self.generate_optimization_problem(fixed_initial_state)

    def _solve_named(self, prob, x_init, ref, label):
        z, _ = self._solve_one(prob, x_init, ref)
        if self.status != "optimal":
            print(f"Status of {label}: {self.status}")
        if not _ok(self._last_status):
            return None, None, None, None
        x, u, xb, ub = prob.split(z)
        return x[0].copy(), u[0].copy(), xb[0].copy(), ub[0].copy()

    def solve_optimization_problem(self, x_init, ref):
        ref.shape = (self._nx,)
        x_init.shape = (self._nx,)
        x_nom, u_nom, x_ss, u_ss = self._solve_named(self._prob, x_init, ref, "tube tracking MPC is")
        return x_nom, u_nom, x_ss, u_ss

    def determine_packet(self, x_hat, ref, q_t):
        x_hat.shape = (self._nx,)
        start = time.time()
        x_nom_traj, u_nom_traj, x_ss, u_ss = self.solve_optimization_problem(x_hat, ref)
        self._computational_times.append(time.time() - start)
        return self.encapsulate(u_nom_traj, u_ss, x_ss, q_t)

    def encapsulate(self, u_nom_traj, u_steady_state, x_steady_state, q_t):
        # NB argument order differs from TrackingMPC.encapsulate (TubeTrackingMPC.py:211 vs TrackingMPC.py:143)
        return self._packet(u_nom_traj, x_steady_state, u_steady_state, q_t)

    def set_ancillary_controller_gain(self, K_ancillary):
        self._K_ancillary = K_ancillary
        self._Acl_plant = self._A - self._B @ K_ancillary

    def get_ancillary_controller_gain(self):
        return self._K if self._K_ancillary is None else self._K_ancillary


class ExtendedTubeTrackingMPC(TubeTrackingMPC):
    """``TubeTrackingMPC.py:249-369`` (Section IV.F of the paper).  ``strict_terminal=False`` keeps the
    reference's behaviour where the "packet received" problem's terminal rows bind the *other*
    problem's x_N and u_bar (``:293``), i.e. only constrain x_bar to the projection of Xf."""

    def __init__(self, A, B, Q, R, N, lambda_param=0.99999, strict_terminal=False):
        super().__init__(A, B, Q, R, N, lambda_param)
        self._strict_terminal = strict_terminal
        self._prob_packet_received = None

    def generate_optimization_problem_when_packet_received(self, W):
        self._ZmW = up.pont_diff(self._Z, W)
        spec = MPCSpec(self._A, self._B, self._Q, self._R, self._N, P_term=self._P, T_ss=self._Tout,
                       stage_x=_Ab(self._Xc), stage_u=_Ab(self._Uc), terminal=_Ab(self._Xf),
                       tube_init=_Ab(self._ZmW), g2_free_terminal=not self._strict_terminal)
        self._prob_packet_received = BatchedQP(spec, Kss=self._K)

    def setup_optimization(self, W, fixed_initial_state=False, rpi_method=0, skip_wasted_pass=False):
        super().setup_optimization(W, fixed_initial_state=fixed_initial_state, rpi_method=rpi_method,
                                   skip_wasted_pass=skip_wasted_pass)
        self.generate_optimization_problem_when_packet_received(W)

    def load_sets(self, Z, Xc, Uc, Xf, W=None, ZmW=None, fixed_initial_state=False):
        super().load_sets(Z, Xc, Uc, Xf, fixed_initial_state)
        if ZmW is not None:
            self._ZmW = ZmW
            spec = MPCSpec(self._A, self._B, self._Q, self._R, self._N, P_term=self._P, T_ss=self._Tout,
                           stage_x=_Ab(Xc), stage_u=_Ab(Uc), terminal=_Ab(Xf), tube_init=_Ab(ZmW),
                           g2_free_terminal=not self._strict_terminal)
            self._prob_packet_received = BatchedQP(spec, Kss=self._K)
        else:
            self.generate_optimization_problem_when_packet_received(W)

    def solve_optimization_problem(self, x_init, ref, gamma_t=0):
        if gamma_t == 1:
            x_init.shape = (self._nx,)
            ref.shape = (self._nx,)
            return self._solve_named(self._prob_packet_received, x_init, ref,
                                     "extended tube MPC when packet has been received")
        return self._solve_named(self._prob, np.asarray(x_init).reshape(self._nx), np.asarray(ref).reshape(self._nx),
                                 "extended tube MPC when packet has not been received")

    def determine_packet(self, x_hat, ref, q_t, gamma_t=0):
        x_hat.shape = (self._nx,)
        start = time.time()
        x_nom_traj, u_nom_traj, x_ss, u_ss = self.solve_optimization_problem(x_hat, ref, gamma_t)
        self._computational_times.append(time.time() - start)
        packet = self.encapsulate(u_nom_traj, u_ss, x_ss, q_t)
        x_nom_0 = x_nom_traj[:, 0] if x_nom_traj is not None else None
        packet["x_nom_0"] = x_nom_0
        return packet, x_nom_0

    def determine_packet_batch(self, x_hat, ref, q_t, gamma_t):
        """Batched: each instance uses the problem selected by its own gamma_t."""
        gamma_t = np.ascontiguousarray(gamma_t, np.int32)
        start = time.time()
        a = self.solve_batch(x_hat, ref, prob=self._prob_packet_received, sel=gamma_t, sel_value=1)
        b = self.solve_batch(x_hat, ref, prob=self._prob, sel=gamma_t, sel_value=0)
        self._computational_times.append(time.time() - start)
        recv = gamma_t == 1
        U = np.where(recv[:, None, None], a["U_t"], b["U_t"])
        x0 = np.where(recv[:, None], a["x"][:, :, 0], b["x"][:, :, 0])
        st = np.where(recv, a["status"], b["status"])
        return {"U_t": U, "q_t": np.asarray(q_t), "status": st, "x_nom_0": x0}

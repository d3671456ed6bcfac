"""Oracle (test infrastructure): single-instance restatement of the local/remote state machines.

Literal semantics of ``SmartActuator.py`` and ``Estimator.py`` -- including the parts the GPU
kernel replaces by O(1) equivalents: the *growing* theta history with the product over
``theta[q_t+1:]`` (``SmartActuator.py:57-71``) and the *unbounded list* of sent control sequences
indexed by ``s_t`` (``Estimator.py:34-41,55``).  Column vectors are plain 1-D arrays here.
"""
import numpy as np


class SmartActuator:
    """``SmartActuator.py:11-123`` (Pezzutto et al. smart actuator)."""

    def __init__(self, K):
        self.K = np.atleast_2d(np.asarray(K, float))
        self.t = 0
        self.theta_hist = []
        self.q_t = 0
        self.s_t = 0
        self.Theta_t = 0
        self.u_traj = None

    # eq. (17): Theta_t = theta_t * prod_{k=q_t+1..t} theta_k          (``:57-71``)
    def update_Theta_t(self, theta_t, q_t):
        self.theta_hist.append(int(theta_t))
        if theta_t == 1:
            self.q_t = int(q_t)
            self.Theta_t = int(np.prod(self.theta_hist[self.q_t + 1:])) if self.theta_hist[self.q_t + 1:] else 1
        else:
            self.Theta_t = 0
        return self.Theta_t

    # eq. (18)                                                          (``:73-79``)
    def update_s_t(self):
        self.s_t = int(self.Theta_t * self.t + (1 - self.Theta_t) * self.s_t)
        return self.s_t

    def update_local_information(self, U_t, x_nom_0=None):              # (``:81-88``)
        if self.Theta_t == 1:
            self.u_traj = np.array(U_t, float)

    # eq. (19)                                                          (``:90-107``)
    def compute_u_t(self, x):
        N = self.u_traj.shape[1] - 1
        k = self.t - self.s_t
        if k < N:
            return self.u_traj[:, k].copy()
        return self.u_traj[:, -1] - self.K @ np.asarray(x, float).flatten()

    def encapsulate(self, x):                                           # (``:115-123``)
        return {"x_t": np.array(x, float).flatten(), "s_t": self.s_t}

    def process_packet(self, packet, x_t, theta_t):                     # (``:31-54``)
        self.update_Theta_t(theta_t, packet["q_t"])
        self.update_s_t()
        self.update_local_information(packet["U_t"])
        u = self.compute_u_t(x_t)
        pkt = self.encapsulate(x_t)
        self.t += 1
        return u, pkt


class ConsistentActuator(SmartActuator):
    """``SmartActuator.py:125-231``: adds the nominal model and the ancillary law."""

    def __init__(self, A, B, K, K_plant, x0, is_extended_MPC_used=False):
        super().__init__(K)
        self.A = np.asarray(A, float)
        self.B = np.asarray(B, float)
        self.K_plant = np.atleast_2d(np.asarray(K_plant, float))
        self.x_nom = np.array(x0, float).flatten()
        self.extended = bool(is_extended_MPC_used)

    def update_local_information(self, U_t, x_nom_0=None):              # (``:215-222``)
        if self.Theta_t == 1:
            self.u_traj = np.array(U_t, float)
            if x_nom_0 is not None:
                self.x_nom = np.array(x_nom_0, float).flatten()

    def process_packet(self, packet, x_t, theta_t):                     # (``:174-213``)
        x_t = np.asarray(x_t, float).flatten()
        self.update_Theta_t(theta_t, packet["q_t"])
        self.update_s_t()
        self.update_local_information(packet["U_t"], packet.get("x_nom_0"))
        x_nom = self.x_nom.copy()
        u_nom = self.compute_u_t(x_nom)
        u = u_nom - self.K_plant @ (x_t - x_nom)                        # eq. (5) (``:166-172``)
        if self.extended:
            pkt = self.encapsulate(x_t)
            pkt["x_nom_t"] = x_nom                                      # (``:224-231``)
        else:
            pkt = self.encapsulate(x_nom)                               # x_t field carries x_nom (``:207``)
        self.x_nom = self.A @ x_nom + self.B @ u_nom                    # eq. (4) (``:146-152``)
        self.t += 1
        return u, pkt

    def get_x_nom(self):
        return self.x_nom


class Estimator:
    """``Estimator.py:9-98``."""

    def __init__(self, A, B, K, x0, N):
        self.A = np.asarray(A, float)
        self.B = np.asarray(B, float)
        self.K = np.atleast_2d(np.asarray(K, float))
        self.x_hat = np.array(x0, float).flatten()
        self.t = 0
        self.q_t = 0
        self.N = int(N)
        self.sequences = []

    def store_sent_control_sequence(self, U_t):                         # (``:34-41``)
        self.sequences.append(None if U_t is None else np.array(U_t, float))

    def update_estimate(self, packet, gamma_t):                         # (``:43-78``)
        if gamma_t == 1:
            x_t, s_t = packet["x_t"], packet["s_t"]
            seq = self.sequences[s_t]
            k = self.t - s_t
            u = seq[:, k] if k < self.N else seq[:, -1] - self.K @ x_t
            self.x_hat = self.A @ x_t + self.B @ u
        else:
            self.x_hat = self.A @ self.x_hat + self.B @ self.sequences[-1][:, 0]
        self.q_t = gamma_t * self.t + (1 - gamma_t) * self.q_t           # (``:87-92``)
        self.t += 1

    def get_estimate(self):
        return self.x_hat

    def get_qt(self):
        return self.q_t


class RobustEstimator(Estimator):
    """``Estimator.py:101-161``."""

    def __init__(self, A, B, K, K_plant, x0, N):
        super().__init__(A, B, K, x0, N)
        self.K_plant = np.atleast_2d(np.asarray(K_plant, float))
        self.x_nom0_mpc = None

    def store_current_optimal_inital_nominal_plant_states(self, x_nom_0):   # (``:158-162``)
        self.x_nom0_mpc = np.array(x_nom_0, float).flatten()

    def update_estimate(self, packet, gamma_t):                         # (``:113-156``)
        if gamma_t == 1:
            x_t, s_t, x_nom = packet["x_t"], packet["s_t"], packet["x_nom_t"]
            seq = self.sequences[s_t]
            k = self.t - s_t
            u_nom = seq[:, k] if k < self.N else seq[:, -1] - self.K @ x_nom
            u = u_nom - self.K_plant @ (x_t - x_nom)
            self.x_hat = self.A @ x_t + self.B @ u
        else:
            self.x_hat = self.A @ self.x_nom0_mpc + self.B @ self.sequences[-1][:, 0]
        self.q_t = gamma_t * self.t + (1 - gamma_t) * self.q_t
        self.t += 1


def encapsulate_controller_packet(u_nom, x_bar, u_bar, K, q_t, x_nom_0=None):
    """``TubeTrackingMPC.encapsulate`` (``TubeTrackingMPC.py:211-227``) / ``TrackingMPC.encapsulate``
    (``TrackingMPC.py:143-158``): U_t = [u_0 .. u_{N-1}, u_bar + K x_bar]."""
    if x_bar is None:
        pkt = {"U_t": None, "q_t": q_t}
    else:
        u_ss = (np.asarray(u_bar, float).flatten() + np.atleast_2d(K) @ np.asarray(x_bar, float).flatten())
        pkt = {"U_t": np.hstack([u_nom, u_ss[:, None]]), "q_t": q_t}
    if x_nom_0 is not None:
        pkt["x_nom_0"] = np.array(x_nom_0, float).flatten()
    return pkt


BULLET_POLE_INERTIA = 0.1 * (0.05 ** 2 + 1.0 ** 2) / 12.0      # box 0.05 x 0.05 x 1.0, mass 0.1 (cartpole.urdf:61-72)
BULLET_LINK_DAMPING = 0.04                                     # pybullet default linearDamping = angularDamping


def cartpole_ode_step(x, F, dt=1.0 / 500.0, M=1.0, m=0.1, I=0.001, g=9.8, l=0.5, damping=0.0):
    """Analytic replacement of the PyBullet plant (parameters from
    ``Results/results_nonlinear_system.py:29-37`` and ``cartpole.urdf:37-38,61-63``): semi-implicit
    Euler at 1/500 s.  State (pos, vel, phi, phidot), phi = 0 upright; linearises to (Ac, Bc).
    ``damping`` k > 0 adds Bullet's link damping (force ``-m v (k + k|v|)`` on each link's linear velocity, torque
    ``-I w (k + k|w|)``); ``I = BULLET_POLE_INERTIA`` is what ``loadURDF`` (no ``URDF_USE_INERTIA_FROM_FILE``,
    ``Results/Cartpole/cartpole.py:14-16``) recomputes from the pole's collision box."""
    pos, vel, phi, om = x
    s, c = np.sin(phi), np.cos(phi)
    fe = F + m * l * om * om * s
    ge = m * g * l * s
    if damping > 0.0:
        vpx, vpz = vel + l * om * c, -l * om * s
        kp = damping + damping * np.hypot(vpx, vpz)
        fpx, fpz = -m * vpx * kp, -m * vpz * kp
        fe += -M * vel * (damping + damping * abs(vel)) + fpx
        ge += -I * om * (damping + damping * abs(om)) + fpx * l * c - fpz * l * s
    D = (M + m) * (I + m * l * l) - (m * l * c) ** 2
    acc = ((I + m * l * l) * fe - m * l * c * ge) / D
    alp = ((M + m) * ge - m * l * c * fe) / D
    vel2 = vel + dt * acc
    om2 = om + dt * alp
    return np.array([pos + dt * vel2, vel2, phi + dt * om2, om2])


def estimate_model_error(x0s, K, Acl, n_steps, substeps=10, dt=1.0 / 500.0, I=0.001, damping=0.0):
    """``Results/estimate_W_for_Cartpole.py:79-127`` with the analytic plant: for every initial condition hold
    ``u = -K x`` over ``substeps`` physics steps (the script's ``lim_zoh`` counter, ``:96-113``), record
    ``w(k) = x(k) - Acl x(k-1)`` at the control instants (``:104-110``).  Returns w [runs, n_steps, 4] and the final
    states."""
    x0s = np.asarray(x0s, float).reshape(-1, 4)
    K = np.asarray(K, float).reshape(1, 4)
    W = np.zeros((len(x0s), n_steps, 4))
    XF = np.zeros((len(x0s), 4))
    for r, x in enumerate(x0s):
        x = x.copy()
        for k in range(n_steps):
            u = float(-(K @ x)[0])
            prev = x.copy()
            for _ in range(substeps):
                x = cartpole_ode_step(x, u, dt=dt, I=I, damping=damping)
            W[r, k] = x - Acl @ prev
        XF[r] = x
    return W, XF

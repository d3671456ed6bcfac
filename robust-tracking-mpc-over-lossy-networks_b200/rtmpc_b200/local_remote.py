"""Local-side and remote-side objects with the reference's names and call order, state on the GPU.

Mirrors ``SmartActuator`` / ``ConsistentActuator`` (``SmartActuator.py``) and ``Estimator`` /
``RobustEstimator`` (``Estimator.py``).  A single object may hold one instance (reference shapes:
column vectors ``[nx,1]``, ``U_t`` as ``[nu, N+1]``) or a batch (``x0`` given as ``[B, nx]`` with
``batch=True``; ``U_t`` as ``[B, N+1, nu]``).  The arithmetic runs in ``rtmpc_actuator_process`` /
``rtmpc_estimator_update`` (one thread per instance); torch tensors are only the device buffers.
For whole closed-loop rollouts use :class:`rtmpc_b200.rollout.RemoteLoop`, which fuses both sides
and the plant step into one kernel per control step.
"""
import numpy as np
import torch

from . import _lib


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _t64(a, dev):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device=dev)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class SmartActuator:
    """``SmartActuator.py:11-123``."""
    _kind = _lib.ACT_SMART

    def __init__(self, K, batch_size=None):
        self._dev = _dev()
        self._K = np.atleast_2d(np.asarray(K, float))
        self._nu, self._nx = self._K.shape
        self._single = batch_size is None
        self._Bn = 1 if batch_size is None else int(batch_size)
        self._t = 0
        self._dK = _t64(self._K, self._dev)
        self._dA = self._dB = self._dKp = None
        self._x_nom_d = None
        self._N = None
        self._buf = None
        z = dict(device=self._dev, dtype=torch.int32)
        self._s_t_d = torch.zeros(self._Bn, **z)
        self._Theta_d = torch.zeros(self._Bn, **z)
        self._last_loss_d = torch.full((self._Bn,), -1, **z)
        self._u_d = torch.zeros(self._Bn, self._nu, device=self._dev, dtype=torch.float64)
        self._pkt_x_d = torch.zeros(self._Bn, self._nx, device=self._dev, dtype=torch.float64)
        self._pkt_xnom_d = torch.zeros(self._Bn, self._nx, device=self._dev, dtype=torch.float64)

    def update_time(self):
        self._t += 1

    # -- helpers ------------------------------------------------------------------------------
    def _U_to_device(self, U_t):
        if self._single:
            U = np.asarray(U_t, float)                     # [nu, N+1]
            N = U.shape[1] - 1
            U = U.T.reshape(1, N + 1, self._nu)
        else:
            U = U_t
            N = U.shape[1] - 1
        Ud = U.to(self._dev, torch.float64) if torch.is_tensor(U) else _t64(U, self._dev)
        if self._buf is None:
            self._N = N
            self._buf = torch.zeros(self._Bn, N + 1, self._nu, device=self._dev, dtype=torch.float64)
        return Ud.contiguous()

    def _vec(self, v, width):
        if torch.is_tensor(v):
            return v.to(self._dev, torch.float64).reshape(self._Bn, width).contiguous()
        return _t64(np.asarray(v, float).reshape(self._Bn, width), self._dev)

    def _ivec(self, v):
        if torch.is_tensor(v):
            return v.to(self._dev, torch.int32).reshape(self._Bn).contiguous()
        return torch.as_tensor(np.asarray(v).reshape(-1).astype(np.int32) * np.ones(self._Bn, np.int32), device=self._dev)

    def _run(self, packet, x_t, theta_t, x_nom0):
        U = self._U_to_device(packet["U_t"])
        xd = self._vec(x_t, self._nx)
        q = self._ivec(packet["q_t"])
        th = self._ivec(theta_t)
        x0d = None if x_nom0 is None else self._vec(x_nom0, self._nx)
        L = _lib.lib()
        p = _lib.ptr
        _lib.check(L.rtmpc_actuator_process(
            self._Bn, self._nx, self._nu, self._N, self._kind, self._t, p(self._dA), p(self._dB), p(self._dK),
            p(self._dKp), p(xd), p(U), p(x0d), p(q), p(th), p(self._buf), p(self._x_nom_d), p(self._s_t_d),
            p(self._Theta_d), p(self._last_loss_d), p(self._u_d), p(self._pkt_x_d), p(self._pkt_xnom_d), _stream()),
            "rtmpc_actuator_process")

    def _out_u(self):
        if self._single:
            return self._u_d.cpu().numpy().reshape(self._nu, 1)
        return self._u_d.clone()

    def _pkt(self, extended=False):
        if self._single:
            pkt = {"x_t": self._pkt_x_d.cpu().numpy().reshape(self._nx, 1), "s_t": int(self._s_t_d.item())}
            if extended:
                pkt["x_nom_t"] = self._pkt_xnom_d.cpu().numpy().reshape(self._nx, 1)
        else:
            pkt = {"x_t": self._pkt_x_d.clone(), "s_t": self._s_t_d.clone()}
            if extended:
                pkt["x_nom_t"] = self._pkt_xnom_d.clone()
        return pkt

    # -- reference surface ----------------------------------------------------------------------
    def process_packet(self, packet, x_t, theta_t):
        if packet["U_t"] is None:
            raise ValueError("controller packet carries U_t = None (infeasible MPC problem)")
        self._run(packet, x_t, theta_t, None)
        u = self._out_u()
        pkt = self._pkt()
        self.update_time()
        return u, pkt

    def get_s_t(self):
        return int(self._s_t_d.item()) if self._single else self._s_t_d.clone()

    def get_Theta_t(self):
        return int(self._Theta_d.item()) if self._single else self._Theta_d.clone()


class ConsistentActuator(SmartActuator):
    """``SmartActuator.py:125-231``."""
    _kind = _lib.ACT_CONSISTENT

    def __init__(self, A, B, K, K_plant, x0, is_extended_MPC_used=False, batch=False):
        x0 = np.asarray(x0, float)
        bs = x0.shape[0] if batch else None
        super().__init__(K, batch_size=bs)
        self._A, self._B = np.asarray(A, float), np.asarray(B, float)
        self._K_plant = np.atleast_2d(np.asarray(K_plant, float))
        self._dA, self._dB, self._dKp = (_t64(M, self._dev) for M in (self._A, self._B, self._K_plant))
        self._x_nom_d = _t64(x0.reshape(self._Bn, self._nx), self._dev)
        self._is_extended_MPC_used = bool(is_extended_MPC_used)
        if self._is_extended_MPC_used:
            self._kind = _lib.ACT_EXTENDED

    def get_x_nom(self):
        if self._single:
            return self._x_nom_d.cpu().numpy().reshape(self._nx, 1)
        return self._x_nom_d.clone()

    def reset_x_nom(self, x_nom_0):
        self._x_nom_d.copy_(self._vec(x_nom_0, self._nx))

    def process_packet(self, packet, x_t, theta_t):
        if packet["U_t"] is None:
            raise ValueError("controller packet carries U_t = None (infeasible MPC problem)")
        self._run(packet, x_t, theta_t, packet.get("x_nom_0"))
        u = self._out_u()
        pkt = self._pkt(extended=self._is_extended_MPC_used)
        self.update_time()
        return u, pkt


class Estimator:
    """``Estimator.py:9-98``."""
    _robust = 0

    def __init__(self, A, B, K, x0, N, batch=False):
        self._dev = _dev()
        self._A, self._B = np.asarray(A, float), np.asarray(B, float)
        self._K = np.atleast_2d(np.asarray(K, float))
        self._nx, self._nu = self._B.shape
        x0 = np.asarray(x0, float)
        self._single = not batch
        self._Bn = x0.shape[0] if batch else 1
        self._N = int(N)
        self._t = 0
        self._dA, self._dB, self._dK = (_t64(M, self._dev) for M in (self._A, self._B, self._K))
        self._dKp = None
        self._x_hat_d = _t64(x0.reshape(self._Bn, self._nx), self._dev)
        self._q_t_d = torch.zeros(self._Bn, device=self._dev, dtype=torch.int32)
        self._hist = torch.zeros(64, self._Bn, self._N + 1, self._nu, device=self._dev, dtype=torch.float64)
        self._n_hist = 0
        self._x_nom0_d = None

    def update_time(self):
        self._t += 1

    def store_sent_control_sequence(self, Ut):
        """``Estimator.py:34-41``: append to the (unbounded) list of sent sequences."""
        if Ut is None:
            raise ValueError("cannot store U_t = None (infeasible MPC problem)")
        if self._n_hist == self._hist.shape[0]:
            self._hist = torch.cat([self._hist, torch.zeros_like(self._hist)], dim=0)
        if self._single:
            U = _t64(np.asarray(Ut, float).T.reshape(1, self._N + 1, self._nu), self._dev)
        else:
            U = Ut.to(self._dev, torch.float64) if torch.is_tensor(Ut) else _t64(Ut, self._dev)
        self._hist[self._n_hist].copy_(U.reshape(self._Bn, self._N + 1, self._nu))
        self._n_hist += 1

    def _vec(self, v, width):
        if torch.is_tensor(v):
            return v.to(self._dev, torch.float64).reshape(self._Bn, width).contiguous()
        return _t64(np.asarray(v, float).reshape(self._Bn, width), self._dev)

    def _ivec(self, v):
        if torch.is_tensor(v):
            return v.to(self._dev, torch.int32).reshape(self._Bn).contiguous()
        return torch.as_tensor(np.asarray(v).reshape(-1).astype(np.int32) * np.ones(self._Bn, np.int32), device=self._dev)

    def update_estimate(self, packet, gamma_t):
        px = self._vec(packet["x_t"], self._nx)
        ps = self._ivec(packet["s_t"])
        pxn = self._vec(packet["x_nom_t"], self._nx) if self._robust else None
        g = self._ivec(gamma_t)
        L = _lib.lib()
        p = _lib.ptr
        _lib.check(L.rtmpc_estimator_update(
            self._Bn, self._nx, self._nu, self._N, self._robust, self._t, self._n_hist, p(self._dA), p(self._dB),
            p(self._dK), p(self._dKp), p(px), p(pxn), p(ps), p(g), p(self._hist), p(self._x_nom0_d),
            p(self._x_hat_d), p(self._q_t_d), _stream()), "rtmpc_estimator_update")
        self.update_time()

    def get_estimate(self):
        if self._single:
            return self._x_hat_d.cpu().numpy().reshape(self._nx, 1)
        return self._x_hat_d.clone()

    def get_qt(self):
        return int(self._q_t_d.item()) if self._single else self._q_t_d.clone()


class RobustEstimator(Estimator):
    """``Estimator.py:101-161``."""
    _robust = 1

    def __init__(self, A, B, K, K_plant, x0, N, batch=False):
        super().__init__(A, B, K, x0, N, batch=batch)
        self._K_plant = np.atleast_2d(np.asarray(K_plant, float))
        self._dKp = _t64(self._K_plant, self._dev)
        self._x_nom0_d = torch.zeros(self._Bn, self._nx, device=self._dev, dtype=torch.float64)

    def store_current_optimal_inital_nominal_plant_states(self, x_nom_0):
        self._x_nom0_d.copy_(self._vec(x_nom_0, self._nx))

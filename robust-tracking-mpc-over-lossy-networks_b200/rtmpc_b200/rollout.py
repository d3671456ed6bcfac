"""Batched closed-loop rollouts: thousands of Monte-Carlo instances of the reference's experiment
loop (``Results/results_linear_system.py:209-255``, ``..._with_extendedMPC.py:247-378``) at once.

Per control step, for all B instances: one QP launch (two for the extended variant, switched per
instance on gamma_{t-1}) followed by one fused loop-step launch (consistent actuator, nominal
model, ancillary law, plant, estimator).  Everything stays on the GPU between steps; the estimate
x_hat the next solve needs is read straight from the loop's device state.  torch only provides
the output buffers and the stream.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

CART_PARAMS = (1.0, 0.1, 0.001, 9.8, 0.5, 1.0 / 500.0, 10.0, 0.0)   # M, m, I, g, l, dt, sub-steps, link damping
# The plant the reference actually simulates (Results/Cartpole/cartpole.py:14-16: loadURDF without
# URDF_USE_INERTIA_FROM_FILE, default dynamics): Bullet recomputes the pole's inertia from its collision box
# (m (0.05^2 + 1^2) / 12, cartpole.urdf:61-72) and damps every link with linearDamping = angularDamping = 0.04.
BULLET_POLE_INERTIA = 0.1 * (0.05 ** 2 + 1.0 ** 2) / 12.0
BULLET_LINK_DAMPING = 0.04
CART_PARAMS_BULLET = (1.0, 0.1, BULLET_POLE_INERTIA, 9.8, 0.5, 1.0 / 500.0, 10.0, BULLET_LINK_DAMPING)
PLANTS = {"linear": (_lib.PLANT_LINEAR, CART_PARAMS), "cartpole": (_lib.PLANT_CARTPOLE, CART_PARAMS),
          "cartpole_bullet": (_lib.PLANT_CARTPOLE, CART_PARAMS_BULLET)}


class _DeviceArray:
    """Minimal ``__cuda_array_interface__`` carrier for library-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _view(ptr, shape, dtype, dev):
    """torch view (no copy) of a device array owned by the C library."""
    return torch.as_tensor(_DeviceArray(ptr, shape, "<f8" if dtype == torch.float64 else "<i4"), device=dev)


class RemoteLoop:
    """B closed-loop instances of one controller.

    mpc       TubeTrackingMPC / ExtendedTubeTrackingMPC / TrackingMPC object from :mod:`rtmpc_b200.mpc`
              with its optimisation problem(s) generated
    kind      'tube' (ConsistentActuator + Estimator), 'extended' (+ RobustEstimator, x_nom_0 in the
              packet, gamma-switched QP) or 'track' (SmartActuator + Estimator, Pezzutto remote MPC)
    plant     'linear', 'cartpole' (analytic ODE with the parameters of the reference's linear model, 10 sub-steps of
              1/500 s, no disturbance) or 'cartpole_bullet' (the same ODE with the pole inertia and link damping Bullet
              uses for the reference's URDF: the plant behind figures/TrackingErrorNonlinear.png)
    """

    def __init__(self, mpc, B, kind="tube", plant="linear", w_half=None, Z=None, K_plant=None):
        _lib.require_cuda()
        self.L = _lib.lib()
        self.mpc = mpc
        self.B = int(B)
        self.kind = kind
        self.dev = torch.device("cuda", torch.cuda.current_device())
        A, Bm, K = np.asarray(mpc._A, float), np.asarray(mpc._B, float), np.atleast_2d(mpc._K)
        self.nx, self.nu = Bm.shape
        self.N = mpc._N
        if K_plant is None:
            K_plant = mpc.get_ancillary_controller_gain() if hasattr(mpc, "get_ancillary_controller_gain") else K
        d = _lib.LoopDesc()
        d.nx, d.nu, d.N = self.nx, self.nu, self.N
        d.actuator = {"track": _lib.ACT_SMART, "tube": _lib.ACT_CONSISTENT, "extended": _lib.ACT_EXTENDED}[kind]
        d.plant, cart = PLANTS[plant]
        keep = [_lib.f64(A), _lib.f64(Bm), _lib.f64(K), _lib.f64(np.atleast_2d(K_plant)),
                _lib.f64(np.zeros(self.nx) if w_half is None else w_half)]
        dp = C.POINTER(C.c_double)
        d.A, d.B, d.K, d.K_plant, d.w_half = (k.ctypes.data_as(dp) for k in keep)
        if Z is not None:
            Hz, hz = _lib.f64(Z.A), _lib.f64(np.asarray(Z.b).flatten())
            d.nz_rows, d.Hz, d.hz = Hz.shape[0], Hz.ctypes.data_as(dp), hz.ctypes.data_as(dp)
            keep += [Hz, hz]
        for i, v in enumerate(cart):
            d.cart_params[i] = v
        h = C.c_void_p()
        _lib.check(self.L.rtmpc_loop_create(C.byref(d), self.B, C.byref(h)), "rtmpc_loop_create")
        self._h = h
        f64, i32 = torch.float64, torch.int32
        B_, nx, nu, N = self.B, self.nx, self.nu, self.N
        g = lambda name: getattr(self.L, "rtmpc_loop_" + name)(self._h)        # noqa: E731
        self.x = _view(g("x"), (B_, nx), f64, self.dev)
        self.x_nom = _view(g("x_nom"), (B_, nx), f64, self.dev)
        self.x_hat = _view(g("x_hat"), (B_, nx), f64, self.dev)
        self.q_t = _view(g("q_t"), (B_,), i32, self.dev)
        self.s_t = _view(g("s_t"), (B_,), i32, self.dev)
        self.Theta = _view(g("Theta"), (B_,), i32, self.dev)
        self.alive = _view(g("alive"), (B_,), i32, self.dev)
        self.err_acc = _view(g("err_acc"), (B_,), f64, self.dev)
        self.tube_max = _view(g("tube_max"), (B_,), f64, self.dev)
        self.u = _view(g("u"), (B_, nu), f64, self.dev)
        self.gamma_last = _view(g("gamma"), (B_,), i32, self.dev)
        nz = mpc._prob.nz
        self.z = torch.zeros(B_, nz, device=self.dev, dtype=f64)
        self.U = torch.zeros(B_, N + 1, nu, device=self.dev, dtype=f64)
        self.status = torch.zeros(B_, device=self.dev, dtype=i32)
        self.iters = torch.zeros(B_, device=self.dev, dtype=i32)
        self.iters_total = torch.zeros(3, device=self.dev, dtype=torch.int64)   # IPM iterations, active-set steps, rounds
        # per-instance warm-start state of each QP (last certified active set), -1 = none
        self.warm = torch.full((B_, mpc._prob.warm_stride), -1, device=self.dev, dtype=i32)
        pr = getattr(mpc, "_prob_packet_received", None)
        self.warm_recv = None if pr is None else torch.full((B_, pr.warm_stride), -1, device=self.dev, dtype=i32)
        self.status_count = torch.zeros(4, device=self.dev, dtype=torch.int64)
        # rtmpc_loop_rollout's counters: solves by status [4], IPM iterations, active-set steps, rounds, flops
        self.stats = torch.zeros(8, device=self.dev, dtype=torch.int64)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.L.rtmpc_loop_destroy(h)
            self._h = None

    def reset(self, x0=None):
        """x = x_nom = x_hat = x0 ([nx] or [B, nx]; None: zeros), t = 0.  A host array goes through ``rtmpc_loop_reset``
        (returns when the state is in place); a CUDA tensor or None through ``rtmpc_loop_reset_device`` (enqueued on the
        current stream, nothing synchronises)."""
        if x0 is None or torch.is_tensor(x0):
            if x0 is not None:
                x0 = x0.to(self.dev, torch.float64).reshape(-1, self.nx).expand(self.B, self.nx).contiguous()
            _lib.check(self.L.rtmpc_loop_reset_device(self._h, _lib.ptr(x0), torch.cuda.current_stream().cuda_stream),
                       "rtmpc_loop_reset_device")
        else:
            x0 = _lib.f64(np.broadcast_to(np.asarray(x0, float).reshape(-1, self.nx), (self.B, self.nx)))
            _lib.check(self.L.rtmpc_loop_reset(self._h, _lib.ptr(x0)), "rtmpc_loop_reset")
        self.iters_total.zero_()
        self.status_count.zero_()
        self.stats.zero_()
        self.warm.fill_(-1)
        if self.warm_recv is not None:
            self.warm_recv.fill_(-1)

    @property
    def t(self):
        return int(self.L.rtmpc_loop_time(self._h))

    def step(self, ref_d, theta=None, gamma=None, w=None, p_loss=None, seed=0, id_offset=0, traj=None, stats=True):
        """One control step for all instances.  ``ref_d`` [B,nx] device tensor.  Either explicit
        device arrays theta/gamma [B] int32 and w [B,nx], or p_loss [B] for on-device Philox draws."""
        stream = torch.cuda.current_stream().cuda_stream
        p = _lib.ptr
        if self.kind == "extended":
            self.mpc._prob_packet_received.solve_device(self.x_hat, ref_d, self.z, self.U, self.status, self.iters,
                                                        sel=self.gamma_last, sel_value=1, stream=stream,
                                                        warm=self.warm_recv)
            self.mpc._prob.solve_device(self.x_hat, ref_d, self.z, self.U, self.status, self.iters,
                                        sel=self.gamma_last, sel_value=0, stream=stream, warm=self.warm)
            x_nom0, stride = self.z, self.z.shape[1]
        else:
            self.mpc._prob.solve_device(self.x_hat, ref_d, self.z, self.U, self.status, self.iters, stream=stream,
                                        warm=self.warm)
            x_nom0, stride = None, 0
        if stats:
            # instances whose controller already returned None take no further steps (loop_step_begin) and are not
            # counted again - the persistent rollout stops solving them altogether
            live = self.alive != 0
            self.accumulate_iters(live)
            self.status_count += torch.bincount(self.status[live].clamp(min=0), minlength=4)[:4]
        _lib.check(self.L.rtmpc_loop_step(self._h, p(self.U), p(self.status), p(x_nom0), stride, p(ref_d), p(theta),
                                          p(gamma), p(w), p(p_loss), int(seed), int(id_offset), p(traj),
                                          0 if traj is None else traj.shape[1] * traj.shape[2], stream),
                   "rtmpc_loop_step")

    def accumulate_iters(self, live=None):
        it = self.iters if live is None else self.iters[live]
        self.iters_total += torch.stack(((it & 0xFFF).sum(), ((it >> 12) & 0xFFF).sum(), ((it >> 24) & 0xF).sum()))

    def run(self, T, ref, p_loss=None, theta=None, gamma=None, w=None, seed=0, id_offset=0, record=False, stats=True,
            fused=None, out=None):
        """T steps.  ``ref`` [nx], [T,nx] or [T,B,nx]; explicit arrays theta/gamma [T,B], w [T,B,nx] (host or
        device) or p_loss [B].  Returns the trajectory tensor [B,T+1,nx] when ``record``.

        ``fused`` (default) runs all T steps in one persistent launch
        (``rtmpc_loop_rollout``); otherwise one QP launch + one loop-step launch per control step.  Both
        give identical results."""
        f64 = torch.float64
        if record and self.t != 0:
            raise _lib.RtmpcError("record=True writes x_t at row t of a [B, T+1, nx] buffer: reset() the loop first")
        if fused is None:
            fused = True
        if fused:
            return self._run_fused(T, ref, p_loss, theta, gamma, w, seed, id_offset, record, out)
        ref = np.asarray(ref, float)
        if ref.ndim == 1:
            ref = np.broadcast_to(ref, (T, self.nx))
        if ref.ndim == 2:
            ref = np.broadcast_to(ref[:, None, :], (T, self.B, self.nx))
        ref_d = torch.as_tensor(np.ascontiguousarray(ref), device=self.dev, dtype=f64)
        traj = torch.zeros(self.B, T + 1, self.nx, device=self.dev, dtype=f64) if record else None
        if theta is not None:
            theta = torch.as_tensor(np.ascontiguousarray(theta), device=self.dev).to(torch.int32).contiguous()
            gamma = torch.as_tensor(np.ascontiguousarray(gamma), device=self.dev).to(torch.int32).contiguous()
            w = None if w is None else torch.as_tensor(np.ascontiguousarray(w), device=self.dev, dtype=f64).contiguous()
        else:
            p_loss = self._p_loss(p_loss)
        for k in range(T):
            if theta is not None:
                self.step(ref_d[k], theta[k], gamma[k], None if w is None else w[k], traj=traj, stats=stats)
            else:
                self.step(ref_d[k], p_loss=p_loss, seed=seed, id_offset=id_offset, traj=traj, stats=stats)
        return traj

    def _p_loss(self, p_loss):
        """Loss probability per instance as a contiguous FP64 device vector [B] (scalar / host array / tensor of any
        dtype or device; None: no losses)."""
        if p_loss is None:
            return torch.zeros(self.B, device=self.dev, dtype=torch.float64)
        if not torch.is_tensor(p_loss):
            p_loss = torch.as_tensor(np.asarray(p_loss, dtype=float))
        return p_loss.to(self.dev, torch.float64).reshape(-1).expand(self.B).contiguous()

    def _run_fused(self, T, ref, p_loss, theta, gamma, w, seed, id_offset, record, out=None):
        f64 = torch.float64
        recv = None
        if self.kind == "extended":
            # both problems padded to the same row count (one kernel instantiation handles either)
            rows = max(self.mpc._prob.rows, self.mpc._prob_packet_received.rows)
            self.mpc._prob = self.mpc._prob.with_rows(rows)
            self.mpc._prob_packet_received = self.mpc._prob_packet_received.with_rows(rows)
            recv = self.mpc._prob_packet_received._h
        if torch.is_tensor(ref):
            ref_d = ref.to(self.dev, f64).contiguous()
        else:
            ref_d = torch.as_tensor(np.ascontiguousarray(np.asarray(ref, float)), device=self.dev, dtype=f64)
        nx, B = self.nx, self.B
        if ref_d.ndim == 1:
            st, sb = 0, 0
        elif ref_d.ndim == 2:
            st, sb = nx, 0
        else:
            st, sb = B * nx, nx
        if record and out is not None:
            # caller's buffer [B, T+1, nx]: device memory, or PINNED host memory - the kernel then writes the trajectory
            # through the bus while it runs (no device buffer, no copy afterwards); rows after an instance stops are left as they are
            if out.dtype != f64 or tuple(out.shape) != (B, T + 1, nx) or not out.is_contiguous():
                raise _lib.RtmpcError("out: contiguous float64 tensor [B, T+1, nx] expected")
            if not (out.is_cuda or out.is_pinned()):
                raise _lib.RtmpcError("out: device tensor or pinned host tensor expected")
            traj = out
        else:
            traj = torch.zeros(B, T + 1, nx, device=self.dev, dtype=f64) if record else None
        if theta is not None:
            theta = torch.as_tensor(np.ascontiguousarray(theta), device=self.dev).to(torch.int32).contiguous()
            gamma = torch.as_tensor(np.ascontiguousarray(gamma), device=self.dev).to(torch.int32).contiguous()
            w = None if w is None else torch.as_tensor(np.ascontiguousarray(w), device=self.dev, dtype=f64).contiguous()
            p_loss = None
        else:
            p_loss = self._p_loss(p_loss)
        p = _lib.ptr
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(self.L.rtmpc_loop_rollout(self._h, self.mpc._prob._h, recv, int(T), p(ref_d), st, sb, p(theta), p(gamma), p(w),
                                             p(p_loss), int(seed), int(id_offset), p(traj),
                                             0 if traj is None else (T + 1) * nx, p(self.stats), stream),
                   "rtmpc_loop_rollout")
        self.status_count = self.stats[:4].clone()
        self.iters_total = self.stats[4:7].clone()
        return traj

    def tracking_error(self, T):
        """1/T sqrt(sum_t ||x_t - ref_t||^2) per instance (``Results/results_linear_system.py:291``)."""
        return torch.sqrt(self.err_acc) / T

"""Device-resident batched QP: the object behind every controller class' ``_prob``.

``BatchedQP(spec)`` condenses and equilibrates the problem on the host (once), uploads it through
``rtmpc_qp_create`` and then solves any number of (x_init, ref) instances per call on the GPU.
It plays the role of the cvxpy ``Problem`` + Clarabel pair in the reference
(``TubeTrackingMPC.py:153,183``) and has no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from .condense import MPCSpec, condense
from .ipm_data import prepare


def desc_arrays(spec: MPCSpec, Kss=None, max_iter=60, min_rows=0):
    """Everything ``rtmpc_qp_desc`` (include/rtmpc.h) points at, prepared on the host: returns (cq, d, ints, floats,
    arrays) with the arrays C-contiguous and named like the struct's fields.  Used by :class:`BatchedQP` and by
    ``examples/dump_qp_desc.py`` (which writes them to a flat file for a C host program)."""
    cq = condense(spec)
    d = prepare(cq)
    nz = cq.Phi.shape[0]
    has_ss = spec.T_ss is not None
    Phi = np.zeros((nz, d.npad))
    Phi[:, :cq.n] = cq.Phi
    Dp = np.zeros(d.npad)
    Dp[:cq.n] = d.D
    Hs = np.zeros((d.npad, d.npad))
    Hs[:cq.n, :cq.n] = d.Hs
    if has_ss and Kss is None:
        raise ValueError("Kss (steady-state gain) is required for the tracking variants")
    arrays = dict(Hs=_lib.f64(Hs), Hinv=_lib.f64(d.Hinv), G=_lib.f64(d.Gs), Y=_lib.f64(d.Y), Fx=_lib.f64(d.Fx),
                  Fr=_lib.f64(d.Fr), lo0=_lib.f64(d.lo0), up0=_lib.f64(d.up0), Lx=_lib.f64(d.Lx),
                  Ux=_lib.f64(d.Ux), parC=_lib.f64(d.par_C), parh=_lib.f64(d.par_h), Dscale=_lib.f64(Dp),
                  Phi=_lib.f64(Phi), Psi=_lib.f64(cq.Psi))
    if Kss is not None:
        arrays["Kss"] = _lib.f64(np.atleast_2d(Kss))
    arrays["has_lo"] = np.ascontiguousarray(d.has_lo, np.uint8)
    arrays["has_up"] = np.ascontiguousarray(d.has_up, np.uint8)
    arrays["shift"] = np.ascontiguousarray(d.shift, np.int32)
    ints = dict(nx=cq.nx, nu=cq.nu, N=cq.N, n=cq.n, npad=d.npad, m=cq.m, mpad=d.mpad, np=len(d.par_h), nz=nz,
                nss=(cq.nx + cq.nu) if has_ss else 0, max_iter=int(max_iter), min_rows=int(min_rows))
    floats = dict(s_floor=float(d.s_floor), sc_b=float(d.sc_b))
    return cq, d, ints, floats, arrays


class BatchedQP:
    def __init__(self, spec: MPCSpec, Kss=None, max_iter=60, min_rows=0):
        self.spec = spec
        self.cq, self.data, ints, floats, keep = desc_arrays(spec, Kss, max_iter, min_rows)
        cq = self.cq
        self.nx, self.nu, self.N = cq.nx, cq.nu, cq.N
        self.nz = ints["nz"]
        self.has_ss = spec.T_ss is not None
        self.n, self.m = cq.n, cq.m
        L = _lib.lib()
        _lib.require_cuda()
        desc = _lib.QPDesc()
        for k, v in ints.items():
            setattr(desc, k, v)
        for k, v in floats.items():
            setattr(desc, k, v)
        dp = C.POINTER(C.c_double)
        for k, v in keep.items():
            if v.dtype == np.float64:
                setattr(desc, k, v.ctypes.data_as(dp) if v.size else None)
        desc.has_lo = keep["has_lo"].ctypes.data_as(C.POINTER(C.c_uint8))
        desc.has_up = keep["has_up"].ctypes.data_as(C.POINTER(C.c_uint8))
        desc.shift = keep["shift"].ctypes.data_as(C.POINTER(C.c_int32))
        h = C.c_void_p()
        _lib.check(L.rtmpc_qp_create(C.byref(desc), C.byref(h)), "rtmpc_qp_create")
        self._h = h
        self._L = L
        self.warm_stride = int(L.rtmpc_qp_warm_stride(h))
        self.rows = int(L.rtmpc_qp_rows(h))          # rows the kernels work on (padded)
        self._Kss = Kss
        self._max_iter = max_iter
        self._method = "active_set"
        self._step_cap = 0

    def with_rows(self, min_rows):
        """The same problem padded to at least ``min_rows`` rows (two problems one rollout switches between)."""
        if self.rows >= min_rows:
            return self
        q = BatchedQP(self.spec, Kss=self._Kss, max_iter=self._max_iter, min_rows=min_rows)
        q.set_method(self._method)          # the handle's settings travel with it
        q.set_step_cap(self._step_cap)
        return q

    def set_method(self, method):
        """'active_set' (default: dual active-set kernel, interior point as fallback) or 'interior_point'."""
        m = {"active_set": _lib.METHOD_ACTIVE_SET, "interior_point": _lib.METHOD_INTERIOR_POINT}[method]
        _lib.check(self._L.rtmpc_qp_set_method(self._h, m), "rtmpc_qp_set_method")
        self._method = method

    def set_step_cap(self, max_steps):
        """Active-set steps after which an instance goes to the interior-point kernel (<= 0: default)."""
        _lib.check(self._L.rtmpc_qp_set_step_cap(self._h, int(max_steps)), "rtmpc_qp_set_step_cap")
        self._step_cap = int(max_steps)

    def set_work_counter(self, counter):
        """uint64 device tensor (1 element) accumulating the active-set kernel's algorithmic flops, or None."""
        _lib.check(self._L.rtmpc_qp_set_work_counter(self._h, _lib.ptr(counter)), "rtmpc_qp_set_work_counter")

    @property
    def rollout_kernel(self):
        """Name of the rollout-kernel instantiation ``rtmpc_loop_rollout`` launches for this problem."""
        return self._L.rtmpc_qp_rollout_kernel(self._h).decode()

    def warm_reset(self):
        _lib.check(self._L.rtmpc_qp_warm_reset(self._h), "rtmpc_qp_warm_reset")

    @staticmethod
    def decode_iters(iters):
        """packed d_iters -> (interior-point iterations, active-set steps, certification rounds)"""
        return iters & 0xFFF, (iters >> 12) & 0xFFF, (iters >> 24) & 0xF

    @staticmethod
    def decode_why(iters):
        """Diagnostics: last reason the active-set kernel gave up on a factorisation (0 = never; 1 step cap,
        2 dependent row with a tiny violation, 3 no free slot, 4 certification failed, 5 certification rounds,
        6 contradiction found late in a leg)."""
        return (iters >> 28) & 7

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.rtmpc_qp_destroy(h)
            self._h = None

    # -- host buffers (numpy in / numpy out): the reference-facing call ----------------------
    def solve_host(self, x_init, ref=None, sel=None, sel_value=1, want_z=True, warm=False):
        x_init = _lib.f64(np.atleast_2d(x_init))
        B = x_init.shape[0]
        ref = None if ref is None else _lib.f64(np.broadcast_to(np.atleast_2d(ref), x_init.shape))
        z = np.empty((B, self.nz)) if want_z else None
        U = np.empty((B, self.N + 1, self.nu))
        if not self.has_ss:
            U[:, self.N, :] = np.nan
        status = np.full(B, -1, np.int32)
        iters = np.zeros(B, np.int32)
        selp = None if sel is None else np.ascontiguousarray(sel, np.int32)
        _lib.check(self._L.rtmpc_qp_solve_host(self._h, B, _lib.ptr(x_init), _lib.ptr(ref), _lib.ptr(selp), sel_value,
                                               1 if warm else 0, _lib.ptr(z), _lib.ptr(U), _lib.ptr(status), _lib.ptr(iters)),
                   "rtmpc_qp_solve_host")
        return z, U, status, iters

    # -- device buffers (torch tensors used purely as memory) ---------------------------------
    def solve_device(self, x_init, ref, z, U, status, iters, sel=None, sel_value=1, stream=None, warm=None):
        """``warm``: int32 device tensor [B, warm_stride] initialised with -1 (per-instance warm-start state)."""
        B = x_init.shape[0]
        _lib.check(self._L.rtmpc_qp_solve(self._h, B, _lib.ptr(x_init), _lib.ptr(ref), _lib.ptr(sel), sel_value,
                                          _lib.ptr(warm), _lib.ptr(z), _lib.ptr(U), _lib.ptr(status), _lib.ptr(iters), stream),
                   "rtmpc_qp_solve")

    def split(self, z):
        """[B,nz] -> x[B,nx,N+1], u[B,nu,N], (x_bar[B,nx], u_bar[B,nu])"""
        nx, nu, N = self.nx, self.nu, self.N
        B = z.shape[0]
        x = z[:, :nx * (N + 1)].reshape(B, N + 1, nx).transpose(0, 2, 1)
        u = z[:, nx * (N + 1):nx * (N + 1) + nu * N].reshape(B, N, nu).transpose(0, 2, 1)
        if self.has_ss:
            o = nx * (N + 1) + nu * N
            return x, u, z[:, o:o + nx], z[:, o + nx:o + nx + nu]
        return x, u

"""ctypes binding of ``librtmpc_b200.so`` (declared in ``include/rtmpc.h``).

There is no CPU fallback: importing this module without the built library, or calling a compute
entry point without a CUDA device, raises.  Build with ``python __graft_entry__.py`` (or
``rtmpc_b200.build.build()``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librtmpc_b200.so")

# every symbol include/rtmpc.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "rtmpc_abi_version", "rtmpc_last_error", "rtmpc_device_count", "rtmpc_set_device", "rtmpc_set_tuning", "rtmpc_get_tuning",
    "rtmpc_qp_create", "rtmpc_qp_destroy", "rtmpc_qp_solve", "rtmpc_qp_solve_host", "rtmpc_launch_count",
    "rtmpc_qp_warm_stride", "rtmpc_qp_rows", "rtmpc_qp_rollout_kernel", "rtmpc_qp_warm_reset", "rtmpc_qp_set_method", "rtmpc_qp_set_work_counter", "rtmpc_qp_set_step_cap",
    "rtmpc_loop_create", "rtmpc_loop_destroy", "rtmpc_loop_reset", "rtmpc_loop_reset_device", "rtmpc_loop_x", "rtmpc_loop_x_nom",
    "rtmpc_loop_x_hat", "rtmpc_loop_q_t", "rtmpc_loop_s_t", "rtmpc_loop_Theta", "rtmpc_loop_alive",
    "rtmpc_loop_err_acc", "rtmpc_loop_tube_max", "rtmpc_loop_u", "rtmpc_loop_gamma", "rtmpc_loop_time",
    "rtmpc_loop_step", "rtmpc_loop_rollout",
    "rtmpc_actuator_process", "rtmpc_estimator_update", "rtmpc_support_sweep", "rtmpc_support_sweep_host",
    "rtmpc_model_error_sweep", "rtmpc_model_error_sweep_host", "rtmpc_lp_solve", "rtmpc_lp_solve_host",
]

OPTIMAL, MAX_ITER, INFEASIBLE, OPTIMAL_INACCURATE = 0, 1, 2, 3
METHOD_ACTIVE_SET, METHOD_INTERIOR_POINT = 0, 1
ABI_VERSION = 3
TUNE_ROLLOUT_QUANTUM, TUNE_ROLLOUT_WARPS, TUNE_AS_WARPS, TUNE_ROLLOUT_CARRY, TUNE_ROLLOUT_FIXED_DIMS, TUNE_CERT_FACTORED = 0, 1, 2, 3, 4, 5
ACT_SMART, ACT_CONSISTENT, ACT_EXTENDED = 0, 1, 2
PLANT_LINEAR, PLANT_CARTPOLE = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class QPDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("nx", "nu", "N", "n", "npad", "m", "mpad", "np", "nz", "nss")] + \
               [(k, _dp) for k in ("Hs", "Hinv", "G", "Y", "Fx", "Fr", "lo0", "up0", "Lx", "Ux")] + \
               [("has_lo", _bp), ("has_up", _bp)] + \
               [(k, _dp) for k in ("parC", "parh", "Dscale", "Phi", "Psi", "Kss")] + \
               [("s_floor", C.c_double), ("sc_b", C.c_double), ("max_iter", C.c_int32), ("min_rows", C.c_int32),
                ("shift", _ip)]


class LoopDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("nx", "nu", "N", "actuator", "plant", "nz_rows")] + \
               [(k, _dp) for k in ("A", "B", "K", "K_plant", "Hz", "hz", "w_half")] + \
               [("cart_params", C.c_double * 8)]


class RtmpcError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtmpcError(f"{LIB_PATH} is missing: build the CUDA library first (python __graft_entry__.py); "
                         "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.rtmpc_abi_version.restype = C.c_int
    if L.rtmpc_abi_version() != ABI_VERSION:
        raise RtmpcError(f"{LIB_PATH} has ABI version {L.rtmpc_abi_version()}, expected {ABI_VERSION}: rebuild it")
    L.rtmpc_last_error.restype = C.c_char_p
    L.rtmpc_device_count.restype = C.c_int
    L.rtmpc_set_device.argtypes = [C.c_int]
    L.rtmpc_set_tuning.argtypes = [C.c_int32, C.c_int32]
    L.rtmpc_get_tuning.argtypes = [C.c_int32]
    L.rtmpc_get_tuning.restype = C.c_int32
    L.rtmpc_qp_create.argtypes = [C.POINTER(QPDesc), C.POINTER(vp)]
    L.rtmpc_qp_destroy.argtypes = [vp]
    L.rtmpc_qp_destroy.restype = None
    L.rtmpc_qp_solve.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.rtmpc_qp_solve_host.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.rtmpc_qp_warm_stride.argtypes = [vp]
    L.rtmpc_qp_warm_stride.restype = C.c_int32
    L.rtmpc_qp_rows.argtypes = [vp]
    L.rtmpc_qp_rows.restype = C.c_int32
    L.rtmpc_qp_rollout_kernel.argtypes = [vp]
    L.rtmpc_qp_rollout_kernel.restype = C.c_char_p
    L.rtmpc_qp_warm_reset.argtypes = [vp]
    L.rtmpc_qp_set_method.argtypes = [vp, C.c_int32]
    L.rtmpc_qp_set_work_counter.argtypes = [vp, vp]
    L.rtmpc_qp_set_step_cap.argtypes = [vp, C.c_int32]
    L.rtmpc_launch_count.restype = C.c_int64
    L.rtmpc_loop_create.argtypes = [C.POINTER(LoopDesc), C.c_int32, C.POINTER(vp)]
    L.rtmpc_loop_destroy.argtypes = [vp]
    L.rtmpc_loop_destroy.restype = None
    L.rtmpc_loop_reset.argtypes = [vp, vp]
    L.rtmpc_loop_reset_device.argtypes = [vp, vp, vp]
    for name in ("x", "x_nom", "x_hat", "q_t", "s_t", "Theta", "alive", "err_acc", "tube_max", "u", "gamma"):
        f = getattr(L, "rtmpc_loop_" + name)
        f.argtypes = [vp]
        f.restype = vp
    L.rtmpc_loop_time.argtypes = [vp]
    L.rtmpc_loop_time.restype = C.c_int32
    L.rtmpc_loop_step.argtypes = [vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, C.c_uint64, C.c_int64, vp,
                                  C.c_int64, vp]
    L.rtmpc_loop_rollout.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int64, C.c_int64, vp, vp, vp, vp, C.c_uint64, C.c_int64,
                                     vp, C.c_int64, vp, vp]
    i32 = C.c_int32
    L.rtmpc_actuator_process.argtypes = [i32] * 6 + [vp] * 18
    L.rtmpc_estimator_update.argtypes = [i32] * 7 + [vp] * 13
    L.rtmpc_support_sweep.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int64, vp, vp]
    L.rtmpc_support_sweep_host.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int64, vp]
    L.rtmpc_model_error_sweep.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.rtmpc_model_error_sweep_host.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    L.rtmpc_lp_solve.argtypes = [vp, vp, i32, i32, i32, vp, vp, C.c_double, vp, i32, C.c_int64, C.c_double, vp, vp, vp, vp, vp]
    L.rtmpc_lp_solve_host.argtypes = [vp, vp, i32, i32, vp, vp, C.c_double, vp, i32, C.c_int64, C.c_double, vp, vp, vp, vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise RtmpcError(f"{what} failed: {lib().rtmpc_last_error().decode()}")


def set_tuning(knob, value):
    """Process-wide launch tuning (``rtmpc_set_tuning``): ``knob`` one of TUNE_ROLLOUT_QUANTUM / TUNE_ROLLOUT_WARPS /
    TUNE_AS_WARPS / TUNE_ROLLOUT_CARRY / TUNE_ROLLOUT_FIXED_DIMS / TUNE_CERT_FACTORED; a negative value restores the default.  Results never depend on it, except in
    the last bits for TUNE_ROLLOUT_CARRY (include/rtmpc.h)."""
    check(lib().rtmpc_set_tuning(int(knob), int(value)), "rtmpc_set_tuning")


def get_tuning(knob):
    return int(lib().rtmpc_get_tuning(int(knob)))


def require_cuda():
    n = lib().rtmpc_device_count()
    if n <= 0:
        raise RtmpcError("no CUDA device visible: rtmpc_b200 has no CPU fallback")
    return n


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """Pointer of a numpy array / torch tensor / None as a void pointer value."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))

"""Wire formats either side of the hot path: the reference's packet dicts <-> the struct-of-arrays
the batched kernels work on, and H-representation I/O compatible with ``polytope.Polytope``.

Reference formats
  controller -> plant   ``{'U_t': ndarray [nu, N+1] | None, 'q_t': int}``                    TubeTrackingMPC.py:223-227
                        ``+ 'x_nom_0': ndarray [nx] | None`` (extended MPC)                  TubeTrackingMPC.py:363-369
  plant -> controller   ``{'x_t': ndarray [nx,1], 's_t': int}``                              SmartActuator.py:119-123
                        ``+ 'x_nom_t': ndarray [nx,1]`` (extended)                           SmartActuator.py:228-229
Batched formats (what ``determine_packet_batch`` returns and the batch mode of the actuator / estimator classes takes)
  controller            ``{'U_t': [B, N+1, nu] (time-major; NaN where infeasible), 'q_t': int32 [B], 'status': int32 [B]
                          [, 'x_nom_0': [B, nx]]}``
  plant                 ``{'x_t': [B, nx], 's_t': int32 [B][, 'x_nom_t': [B, nx]]}``
Everything here is host-side numpy; device tensors are accepted where noted (they are copied to the host).
"""
import numpy as np

from . import _lib
from .polytope import Polytope


def _np(a):
    if hasattr(a, "detach"):                      # torch tensor (any device)
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def pack_controller_packets(packets):
    """List of B reference controller packets -> one batched packet.  ``U_t = None`` (infeasible problem,
    ``TubeTrackingMPC.py:220-221``) becomes a NaN payload with ``status = RTMPC_INFEASIBLE``."""
    B = len(packets)
    if B == 0:
        raise ValueError("no packets")
    shape = next((np.asarray(p["U_t"]).shape for p in packets if p["U_t"] is not None), None)
    if shape is None:
        raise ValueError("every packet is infeasible: payload shape unknown")
    nu, N1 = shape
    U = np.full((B, N1, nu), np.nan)
    status = np.zeros(B, np.int32)
    for b, p in enumerate(packets):
        if p["U_t"] is None:
            status[b] = _lib.INFEASIBLE
        else:
            Ub = np.asarray(p["U_t"], float)
            if Ub.shape != (nu, N1):
                raise ValueError(f"packet {b}: U_t has shape {Ub.shape}, expected {(nu, N1)}")
            U[b] = Ub.T
    out = {"U_t": U, "q_t": np.array([int(p["q_t"]) for p in packets], np.int32), "status": status}
    if any("x_nom_0" in p for p in packets):
        nx = next(np.asarray(p["x_nom_0"]).size for p in packets if p.get("x_nom_0") is not None)
        X = np.full((B, nx), np.nan)
        for b, p in enumerate(packets):
            if p.get("x_nom_0") is not None:
                X[b] = np.asarray(p["x_nom_0"], float).reshape(nx)
        out["x_nom_0"] = X
    return out


def unpack_controller_packets(batched):
    """Batched controller packet -> list of B packets in the reference's shapes (``U_t`` [nu, N+1] or None)."""
    U = _np(batched["U_t"])
    q = np.broadcast_to(_np(batched["q_t"]).reshape(-1), (U.shape[0],))
    st = _np(batched["status"]).reshape(-1) if "status" in batched else np.zeros(U.shape[0], np.int32)
    X = _np(batched["x_nom_0"]) if batched.get("x_nom_0") is not None else None
    out = []
    for b in range(U.shape[0]):
        dead = st[b] == _lib.INFEASIBLE or not np.all(np.isfinite(U[b]))
        p = {"U_t": None if dead else np.array(U[b].T, float), "q_t": int(q[b])}
        if X is not None:
            p["x_nom_0"] = None if dead else np.array(X[b], float)
        out.append(p)
    return out


def pack_plant_packets(packets):
    """List of B reference plant packets -> ``{'x_t': [B,nx], 's_t': [B][, 'x_nom_t': [B,nx]]}``."""
    if not packets:
        raise ValueError("no packets")
    out = {"x_t": np.stack([np.asarray(p["x_t"], float).reshape(-1) for p in packets]),
           "s_t": np.array([int(p["s_t"]) for p in packets], np.int32)}
    if "x_nom_t" in packets[0]:
        out["x_nom_t"] = np.stack([np.asarray(p["x_nom_t"], float).reshape(-1) for p in packets])
    return out


def unpack_plant_packets(batched):
    """Batched plant packet (host arrays or device tensors) -> list of reference plant packets (column vectors)."""
    X = _np(batched["x_t"])
    s = _np(batched["s_t"]).reshape(-1)
    Xn = _np(batched["x_nom_t"]) if "x_nom_t" in batched else None
    out = []
    for b in range(X.shape[0]):
        p = {"x_t": np.array(X[b], float).reshape(-1, 1), "s_t": int(s[b])}
        if Xn is not None:
            p["x_nom_t"] = np.array(Xn[b], float).reshape(-1, 1)
        out.append(p)
    return out


# ---- polytope I/O ---------------------------------------------------------------------------------------------------
def as_polytope(obj):
    """Anything with ``.A`` [m, n] and ``.b`` ([m], [m,1] or [1,m]) -> this package's ``Polytope`` with unit-length rows,
    i.e. what ``polytope.Polytope(A, b)`` holds after its own normalisation (zero rows dropped).  Objects of this
    package are returned unchanged."""
    if isinstance(obj, Polytope):
        return obj
    if not (hasattr(obj, "A") and hasattr(obj, "b")):
        raise TypeError(f"{type(obj).__name__} has no .A / .b: not an H-representation")
    A = np.array(obj.A, dtype=float)
    b = np.array(obj.b, dtype=float).reshape(-1)
    if A.ndim != 2 or A.shape[0] != b.size:
        raise ValueError(f"inconsistent H-representation: A {A.shape}, b {b.shape}")
    V = getattr(obj, "vertices", None)
    return Polytope(A, b, vertices=None if V is None else np.array(V, float))


def to_hrep(poly):
    """(A, b) copies of any H-representation object, ``b`` flat."""
    return np.array(poly.A, float), np.array(poly.b, float).reshape(-1)


def to_polytope_package(poly):
    """``polytope.Polytope`` with the same rows, when that third-party package is importable (it is not in this image;
    the reference's scripts use it)."""
    try:
        import polytope as pc
    except ImportError as e:
        raise ImportError("the third-party `polytope` package is not installed") from e
    A, b = to_hrep(poly)
    return pc.Polytope(A, b)

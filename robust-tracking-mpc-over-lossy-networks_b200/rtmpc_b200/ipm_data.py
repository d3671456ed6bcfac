"""Host-side, once-per-problem preparation of everything the interior-point kernel keeps resident.

Given a :class:`~rtmpc_b200.condense.CondensedQP` this module
  1. equilibrates it (modified Ruiz on the KKT matrix [[H, G'],[G, 0]]): zeta = D zeta_hat,
     row i of G scaled by e_i, cost scaled by c,
  2. prepares ``Hinv = H^{-1}`` (start point = unconstrained minimiser, and the active-set
     polish operator ``Y = G Hinv``),
  3. pads columns to a multiple of 4 doubles and rows to a multiple of 32 (one row per lane and
     slot), with padding rows carrying no finite bound.

All arrays are FP64 and row-major.
"""
from dataclasses import dataclass

import numpy as np

from .condense import CondensedQP, INF


def _pad(k, mult):
    return ((k + mult - 1) // mult) * mult


@dataclass
class IPMData:
    n: int
    m: int
    nx: int
    npad: int               # columns padded (multiple of 4)
    mpad: int               # rows padded (multiple of 32)
    D: np.ndarray           # [n]   variable scaling
    Ev: np.ndarray          # [m]   row scaling
    c: float                # cost scaling
    Hs: np.ndarray          # [n,n]  scaled Hessian
    Gs: np.ndarray          # [mpad,npad] scaled constraint matrix (zero padded)
    Hinv: np.ndarray        # [npad,npad]
    Y: np.ndarray           # [mpad,npad]   Gs @ Hinv
    Fx: np.ndarray          # [npad,nx]   q_hat = Fx x_init + Fr ref
    Fr: np.ndarray          # [npad,nx]
    lo0: np.ndarray         # [mpad]   scaled bounds: lo = lo0 + Lx x_init
    up0: np.ndarray         # [mpad]
    Lx: np.ndarray          # [mpad,nx]
    Ux: np.ndarray          # [mpad,nx]
    has_lo: np.ndarray      # [mpad] uint8
    has_up: np.ndarray      # [mpad] uint8
    par_C: np.ndarray       # [mp,nx]  parameter rows  par_C x_init <= par_h (unscaled)
    par_h: np.ndarray       # [mp]
    shift: np.ndarray       # [mpad] int32 warm-start map (row holding the same constraint one stage earlier)
    s_floor: float          # smallest initial slack (scaled units)
    sc_b: float             # 1 + typical |bound| (scaled units), for relative primal residuals


def ruiz_equilibrate(H, G, iters=15):
    n = H.shape[0]
    m = G.shape[0]
    D = np.ones(n)
    E = np.ones(m)
    Hs = H.copy()
    Gs = G.copy()
    for _ in range(iters):
        cn = np.maximum(np.abs(Hs).max(axis=0), np.abs(Gs).max(axis=0) if m else 0.0)
        rn = np.abs(Gs).max(axis=1) if m else np.zeros(0)
        dD = 1.0 / np.sqrt(np.where(cn > 1e-12, cn, 1.0))
        dE = 1.0 / np.sqrt(np.where(rn > 1e-12, rn, 1.0))
        Hs = dD[:, None] * Hs * dD[None, :]
        Gs = dE[:, None] * Gs * dD[None, :]
        D *= dD
        E *= dE
    c = 1.0 / max(np.mean(np.abs(Hs).max(axis=0)), 1e-12)
    return D, E, c, c * Hs, Gs


def prepare(cq: CondensedQP) -> IPMData:
    n, m, nx = cq.n, cq.m, cq.nx
    npad = _pad(n, 4)
    mpad = _pad(max(m, 1), 32)
    D, Ev, c, Hs, Gs = ruiz_equilibrate(cq.H, cq.G)
    Hs = 0.5 * (Hs + Hs.T)
    Gp = np.zeros((mpad, npad))
    Gp[:m, :n] = Gs
    Hinv = np.zeros((npad, npad))
    Hinv[:n, :n] = np.linalg.inv(Hs)
    Hinv = 0.5 * (Hinv + Hinv.T)
    Y = Gp @ Hinv

    Fx = np.zeros((npad, nx))
    Fr = np.zeros((npad, nx))
    Fx[:n] = c * D[:, None] * cq.Fx
    Fr[:n] = c * D[:, None] * cq.Fr

    fin_lo = cq.lo0 > -INF / 2
    fin_up = cq.up0 < INF / 2
    lo0 = np.zeros(mpad)
    up0 = np.zeros(mpad)
    Lx = np.zeros((mpad, nx))
    Ux = np.zeros((mpad, nx))
    lo0[:m] = np.where(fin_lo, Ev * cq.lo0, 0.0)
    up0[:m] = np.where(fin_up, Ev * cq.up0, 0.0)
    Lx[:m] = np.where(fin_lo[:, None], Ev[:, None] * cq.Lx, 0.0)
    Ux[:m] = np.where(fin_up[:, None], Ev[:, None] * cq.Ux, 0.0)
    has_lo = np.zeros(mpad, np.uint8)
    has_up = np.zeros(mpad, np.uint8)
    has_lo[:m] = fin_lo
    has_up[:m] = fin_up
    shift = np.full(mpad, -1, np.int32)
    if cq.shift is not None:
        shift[:m] = cq.shift
    bmag = np.r_[np.abs(lo0[:m][fin_lo]), np.abs(up0[:m][fin_up])]
    sc_b = 1.0 + (float(np.median(bmag)) if bmag.size else 0.0)
    return IPMData(n=n, m=m, nx=nx, npad=npad, mpad=mpad, D=D, Ev=Ev, c=c, Hs=Hs, Gs=Gp, Hinv=Hinv, Y=Y,
                   Fx=Fx, Fr=Fr, lo0=lo0, up0=up0, Lx=Lx, Ux=Ux, has_lo=has_lo, has_up=has_up,
                   par_C=cq.par_C.copy(), par_h=cq.par_h.copy(), shift=shift, s_floor=1e-3 * sc_b, sc_b=sc_b)

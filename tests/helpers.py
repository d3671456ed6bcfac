"""Shared test helpers: load golden fixtures and turn them into product-side problem specs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def spec_tube_tracking(s, fixed_initial_state=True):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["Xc_A"], s["Xc_b"]), stage_u=(s["Uc_A"], s["Uc_b"]), terminal=(s["Xf_A"], s["Xf_b"]),
                   tube_init=None if fixed_initial_state else (s["Z_A"], s["Z_b"]))


def spec_ext_received(s, strict=False):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["Xc_A"], s["Xc_b"]), stage_u=(s["Uc_A"], s["Uc_b"]), terminal=(s["Xf_A"], s["Xf_b"]),
                   tube_init=(s["ZmW_A"], s["ZmW_b"]), g2_free_terminal=not strict)


def spec_tracking(s):
    from rtmpc_b200.condense import MPCSpec
    return MPCSpec(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), P_term=s["P"], T_ss=10 * s["P"],
                   stage_x=(s["X_A"], s["X_b"]), stage_u=(s["U_A"], s["U_b"]), terminal=(s["Xf_track_A"], s["Xf_track_b"]))

#!/usr/bin/env python
"""Write the host-prepared description of the cartpole remote tube MPC problem (BASELINE configs[1]) to a flat binary
file that a C host program can read straight into ``rtmpc_qp_desc`` (include/rtmpc.h) - see examples/c_abi_demo.c.

    python examples/dump_qp_desc.py /tmp/qp_cp.bin

Layout (little endian): magic "RTMPCQP1"; int32[12] nx nu N n npad m mpad np nz nss max_iter min_rows; float64[2]
s_floor sc_b; then for each array, in the order of ARRAYS below: int64 element count followed by the elements
(float64, except has_lo / has_up uint8 and shift int32; count 0 = NULL)."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

ARRAYS = ("Hs", "Hinv", "G", "Y", "Fx", "Fr", "lo0", "up0", "Lx", "Ux", "has_lo", "has_up", "parC", "parh", "Dscale", "Phi",
          "Psi", "Kss", "shift")
INTS = ("nx", "nu", "N", "n", "npad", "m", "mpad", "np", "nz", "nss", "max_iter", "min_rows")


def main(path):
    import helpers as H
    from rtmpc_b200.qp import desc_arrays
    s = H.load("sets_cp.npz")
    _, _, ints, floats, arrays = desc_arrays(H.spec_tube_tracking(s), Kss=s["K"])
    with open(path, "wb") as f:
        f.write(b"RTMPCQP1")
        f.write(struct.pack("<12i", *[ints[k] for k in INTS]))
        f.write(struct.pack("<2d", floats["s_floor"], floats["sc_b"]))
        for k in ARRAYS:
            a = arrays.get(k)
            n = 0 if a is None else a.size
            f.write(struct.pack("<q", n))
            if n:
                f.write(np.ascontiguousarray(a).tobytes())
    print(path, os.path.getsize(path), "bytes; n =", ints["n"], "rows =", ints["m"])


if __name__ == "__main__":
    main(sys.argv[1])

"""Latency of ONE solve through the reference-facing call (determine_packet of the drop-in class), the quantity the
reference's Fig. 3d histograms (2.5-20 ms with cvxpy + Clarabel): wall clock per call incl. ctypes, H2D, kernel, D2H."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench
import helpers as H
g = H.load("loop_cp_tube.npz")
mpc, Z = bench.build_controller()
xs = g["tube_xhat_in"][1]            # a closed loop of the golden run (loss 0.3), 250 states
refs = g["refs"]
for label, reset in (("warm-started along the closed loop", False), ("cold", True)):
    ts = []
    for rep in range(3):
        mpc._prob.warm_reset()
        for x, r in zip(xs, refs):
            if reset:
                mpc._prob.warm_reset()
            t0 = time.perf_counter()
            pkt = mpc.determine_packet(x.copy(), r.copy(), 0)
            ts.append(time.perf_counter() - t0)
            assert pkt["U_t"] is not None
    ts = np.array(ts[len(xs):]) * 1e3      # drop the first pass (page-in)
    print(f"determine_packet, {label}: median {np.median(ts):.3f} ms, 90 % {np.quantile(ts, 0.9):.3f} ms, max {ts.max():.3f} ms over {len(ts)} calls")

#!/usr/bin/env python
"""Batched counterpart of the reference's Results/results_linear_system.py: remote tube MPC (RT-MPC) against Pezzutto's
remote MPC (R-MPC) on the linearised cartpole, 10 packet-loss probabilities x N_MC Monte-Carlo runs x 250 steps, every
run a closed-loop instance on the GPU.  Prints the statistics the script prints (tracking errors, failed executions of
R-MPC, amortised solve time); plotting is out of scope.

    python examples/results_linear_system.py --n-mc 20            # the paper's experiment: 200 instances per controller
    python examples/results_linear_system.py --n-mc 2000          # 20 000 instances per controller
    python examples/results_linear_system.py --compute-sets       # mRPI set, tightening and terminal sets from scratch
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mc", type=int, default=20)
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--seed", type=int, default=679)
    ap.add_argument("--compute-sets", action="store_true", help="run the set pipeline instead of loading tests/golden/sets_cp.npz")
    args = ap.parse_args()
    import helpers as H
    from rtmpc_b200 import mpc
    from rtmpc_b200.experiments import linear_system_experiment
    s = H.load("sets_cp.npz")
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])                      # Results/results_linear_system.py:76-91
    if args.compute_sets:
        from rtmpc_b200.polytope import box
        tube = mpc.TubeTrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
        tube.set_input_constraints(H.poly(s, "U"))
        tube.set_state_constraints(H.poly(s, "X"))
        tube.setup_optimization(box(hw), fixed_initial_state=True, rpi_method=1)
        track = mpc.TrackingMPC(s["A"], s["B"], s["Q"], s["R"], int(s["N"]))
        track.set_input_constraints(H.poly(s, "U"))
        track.set_state_constraints(H.poly(s, "X"))
        track.setup_optimization()
        Z = tube._Z
    else:
        tube, track, Z = H.make_tube_mpc(s), H.make_track_mpc(s), H.poly(s, "Z")
    res = linear_system_experiment(tube, track, Z, hw, n_mc=args.n_mc, T=args.steps, seed=args.seed)
    print(res.summary())


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: batched tube-MPC QP solves per second (= closed-loop steps per second).

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): the reference's linearised-cartpole remote tube
MPC experiment (Results/results_linear_system.py: nx=4, nu=1, N=20, T=250 control steps, x0=0,
ref=(0.5,0,0,0), loss probability {0,...,0.9}[i mod 10], theta/gamma ~ Bernoulli, w ~ U(-hw,hw)) with
4096 independent closed-loop instances per GPU.  One bench "step" = one full 250-step rollout of the
batch = 1 024 000 QP solves per GPU.  Weak scaling: every rank runs its own 4096 instances (global
instance id = rank*4096 + i seeds the Philox draws), no collective on the hot path; one NCCL
all-gather of the per-instance tracking errors after the timed region.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W     # CPU arm: the oracle port on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
sys.path.insert(0, ROOT)

METRIC = "batched tube-MPC QP solves/sec (closed-loop steps/s)"
B_PER_GPU = 4096
T_STEPS = 250
HW = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
REF = np.array([0.5, 0.0, 0.0, 0.0])
SEED = 679


def load_sets():
    return np.load(os.path.join(ROOT, "tests", "golden", "sets_cp.npz"))


def ipm_flops_per_iteration(n, m):
    """Algorithmic FP64 flops of one interior-point iteration of the condensed QP:
    Schur assembly 2*m*n(n+1)/2, five G mat-vecs 2*m*n each (t, ta, tz, two G' products; the third
    G' product shares its pass), Cholesky n^3/3, two triangular solve pairs 2*2*n^2."""
    return 2 * m * n * (n + 1) // 2 + 5 * 2 * m * n + n ** 3 // 3 + 4 * n * n


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference's own cvxpy+Clarabel path cannot be installed here)
# --------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _W["tp"] = threadpool_limits(1)
    except Exception:
        pass
    from oracle import ref_qp as rq
    from oracle.ref_polytope import Polytope
    s = load_sets()
    P = lambda k: Polytope(s[k + "_A"], s[k + "_b"], normalize=False)      # noqa: E731
    _W["qp"] = rq.build_tube_tracking(s["A"], s["B"], s["Q"], s["R"], int(s["N"]), s["P"], P("Xc"), P("Uc"), P("Xf"),
                                      None, True)
    _W["rq"] = rq
    _W["sets"] = s


def _cpu_worker_loops(args):
    """Closed loops of the benchmark workload on one host core, the way Results/results_linear_system.py:209-255 runs them
    (estimator -> QP -> packet -> actuator -> plant -> estimator, oracle restatements of the reference's classes): instance
    id has loss probability 0.1 * (id % 10) on both links, disturbances uniform in the box HW, reference REF, x0 = 0.
    Returns the number of QP solves."""
    ids, T = args
    rq, qp, s = _W["rq"], _W["qp"], _W["sets"]
    from oracle import ref_loop as rl
    A, B, K, N = s["A"], s["B"], s["K"], int(s["N"])
    n = 0
    for i in ids:
        rng = np.random.default_rng(SEED * 1000003 + int(i))
        p = 0.1 * (int(i) % 10)
        x0 = np.zeros(A.shape[0])
        est = rl.Estimator(A, B, K, x0, N)
        act = rl.ConsistentActuator(A, B, K, K, x0)
        x = x0.copy()
        xhat = est.get_estimate()
        for t in range(T):
            theta = 1 if t == 0 or rng.random() >= p else 0
            gamma = 1 if t == 0 or rng.random() >= p else 0
            w = HW * (2.0 * rng.random(A.shape[0]) - 1.0)
            q_t = est.get_qt()
            (xn, un, xb, ub), res = rq.solve_param(qp, xhat.copy(), REF.copy())
            n += 1
            if res.status != "optimal":
                break
            pkt = rl.encapsulate_controller_packet(un, xb, ub, K, q_t)
            est.store_sent_control_sequence(pkt["U_t"])
            u, ppkt = act.process_packet(pkt, x, theta)
            x = A @ x + B @ u + w
            est.update_estimate(ppkt, gamma)
            xhat = est.get_estimate()
    return n


def run_cpu_arm(solves_per_step, steps, warmup, cores=None):
    """The reference's CPU path on a bounded sample of the SAME workload: whole closed loops (T_STEPS control steps each,
    instance ids 0, 1, ..: all ten loss rates), one process per host core; `solves_per_step` sets how many loops
    (at least one per core)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    n_loops = max(cores, int(round(solves_per_step / T_STEPS)))
    chunks = [(list(range(c, n_loops, cores)), T_STEPS) for c in range(cores)]
    ctx = mp.get_context("spawn")      # the parent may hold a CUDA context
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        for _ in range(warmup):
            pool.map(_cpu_worker_loops, [([c], 2) for c in range(cores)])
        t0 = time.perf_counter()
        total = 0
        for _ in range(steps):
            total += sum(pool.map(_cpu_worker_loops, chunks))
        dt = time.perf_counter() - t0
    return total / dt, dt / steps, cores, total, n_loops


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms (the timed region is tens of
    milliseconds), nvidia-smi every 200 ms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.sm, self.sm_max, self.reasons = [], None, set()
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        if nv is not None:
            bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                    (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                    (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                    (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
            while not self._stop.is_set():
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    for bit, name in bits:
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop.wait(0.005)
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.sm.append(float(f[0]))
                    self.sm_max = float(f[1])
                    for i, n in enumerate(self.NAMES):
                        if f[2 + i].lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=3)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no clock samples"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
_SETUP = {}


def build_controller(extended=False):
    """The controller of Results/results_linear_system.py:26-125, built by the PRODUCT's own set-up call on the GPU
    (``setup_optimization(W, fixed_initial_state=True, rpi_method=1)``: Darup RPI by support sweeps, reduce, tightening,
    the 9-D terminal-set iteration on batched LPs, QP generation) - not loaded from a fixture.  The sets it produces are
    compared with tests/golden/sets_cp.npz (same row counts, bounds after sorting) and the result goes into ``checks``."""
    from rtmpc_b200 import mpc, numerics
    from rtmpc_b200.polytope import box
    A, B = numerics.cartpole_linear(0.02)
    Q, R, N = np.diag([100.0, 10.0, 100.0, 10.0]), 0.1 * np.eye(1), 20
    cls = mpc.ExtendedTubeTrackingMPC if extended else mpc.TubeTrackingMPC
    c = cls(A, B, Q, R, N)
    c.set_input_constraints(box([10.0]))
    c.set_state_constraints(box([5.0, 5.0, 0.3, 2.0]))
    t0 = time.perf_counter()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):           # the set-up prints k_star / convergence lines like the reference
        c.setup_optimization(box(HW), fixed_initial_state=True, rpi_method=1, skip_wasted_pass=True)
    dt = time.perf_counter() - t0
    s = load_sets()
    same = True
    worst = 0.0
    for k in ("Z", "Xc", "Uc", "Xf"):
        P = getattr(c, "_" + k)
        if P.A.shape != s[k + "_A"].shape:
            same = False
            continue
        worst = max(worst, float(np.abs(np.sort(P.b) - np.sort(s[k + "_b"])).max()))
    _SETUP["extended" if extended else "tube"] = {"setup_seconds": dt, "rows_Z_Xf": [int(c._Z.A.shape[0]), int(c._Xf.A.shape[0])],
                                                   "same_row_counts_as_fixture": same, "max_sorted_bound_diff_vs_fixture": worst}
    return c, c._Z


# BASELINE.json configs[1..3]: the same closed loop with another controller variant / plant
WORKLOADS = {
    "c2": dict(kind="tube", plant="linear", extended=False, seed=679,
               text="results_linear_system.py remote tube MPC (linearised cartpole nx=4 nu=1 N=20)"),
    "c3": dict(kind="extended", plant="linear", extended=True, seed=347,
               text="results_linear_system_with_extendedMPC.py extended remote tube MPC (two QPs switched by gamma_{t-1}, "
                    "robust estimator, x_nom_0 in the packet; linearised cartpole)"),
    "c4": dict(kind="tube", plant="cartpole", extended=False, seed=124,
               text="results_nonlinear_system.py remote tube MPC on the analytic cartpole ODE (10 sub-steps of 1/500 s per "
                    "control step, no added disturbance)"),
}


def measure_fp64_peak(torch, dev):
    """FP64 denominator: cuBLAS DGEMM 6144^3 (best of 5), measured in this run -- MEASURED_PEAKS.json
    only holds HBM and bf16 figures."""
    n = 6144
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from rtmpc_b200 import _lib
    from rtmpc_b200 import distributed as D
    from rtmpc_b200.rollout import RemoteLoop

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if args.rollout_warps:
        _lib.set_tuning(_lib.TUNE_ROLLOUT_WARPS, args.rollout_warps)
    if args.rollout_quantum >= 0:
        _lib.set_tuning(_lib.TUNE_ROLLOUT_QUANTUM, args.rollout_quantum)
    if args.rollout_carry >= 0:
        _lib.set_tuning(_lib.TUNE_ROLLOUT_CARRY, args.rollout_carry)
    if args.rollout_fixed_dims >= 0:
        _lib.set_tuning(_lib.TUNE_ROLLOUT_FIXED_DIMS, args.rollout_fixed_dims)
    wl = WORKLOADS[args.workload]
    SEED = wl["seed"]
    mpc, Z = build_controller(extended=wl["extended"])
    B, T = args.instances, T_STEPS
    loop = RemoteLoop(mpc, B, kind=wl["kind"], plant=wl["plant"], w_half=HW if wl["plant"] == "linear" else None, Z=Z)
    ids0, _ = D.shard(B * world, rank, world)          # weak scaling: B instances per rank, global ids
    p_loss = torch.as_tensor(np.array([0.1 * ((ids0 + i) % 10) for i in range(B)]), device=dev)
    ref_d = torch.as_tensor(np.tile(REF, (B, 1)), device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)    # > 126 MB L2
    n, m = mpc._prob.n, mpc._prob.m
    f_it = ipm_flops_per_iteration(n, m)
    stream = torch.cuda.current_stream()

    def rollout(seed):
        # one persistent launch: every warp takes its instance through all T control steps
        loop.reset()
        loop.run(T, ref_d[0], p_loss=p_loss, seed=seed, id_offset=ids0, fused=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for wi in range(args.warmup):
        rollout(SEED + 1000 + wi)
    barrier()
    # the instantiation rtmpc_loop_rollout launches (from the library; read after the first rollout: the extended controller's
    # two problems are padded to a common row count when the loop first runs)
    kernel_name = mpc._prob.rollout_kernel.replace("*", "true" if wl["extended"] else "false")
    fp64_peak = measure_fp64_peak(torch, dev) if rank == 0 else None
    # ---- timed region: K rollouts, device-timed, L2 flushed between them -------------------------
    launches_timed = 0          # our kernels launched between the timing events (the loop re-initialisation is outside them)
    total_ms = 0.0
    solve_ms = 0.0
    iters_sum = np.zeros(3, np.int64)
    as_flops = 0
    status_sum = np.zeros(4, np.int64)
    with ClockSampler(local) as clk:
        barrier()
        for k in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            loop.reset()
            l0 = L.rtmpc_launch_count()
            e0.record(stream)
            loop.run(T, ref_d[0], p_loss=p_loss, seed=SEED + k, id_offset=ids0, fused=True)
            e1.record(stream)
            launches_timed += L.rtmpc_launch_count() - l0
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            solve_ms = total_ms                 # the rollout IS the kernel: one launch per bench step
            st = loop.stats.cpu().numpy()
            iters_sum += st[4:7]
            status_sum += st[:4]
            as_flops += int(st[7])
        barrier()
    launches = launches_timed
    err = loop.tracking_error(T)
    tube_max = float(loop.tube_max.max().item())
    t_ms = D.all_reduce_max(torch.tensor([total_ms], device=dev, dtype=torch.float64))
    g0 = time.perf_counter()
    err_all = D.all_gather_instances(err, B * world)     # the only collectives: statistics after the timed region
    status_all = D.all_reduce_sum(torch.as_tensor(status_sum, device=dev)).cpu().numpy()
    torch.cuda.synchronize()
    gather_ms = (time.perf_counter() - g0) * 1e3 if world > 1 else 0.0
    total_ms_max = float(t_ms.item())
    solves = B * T * args.steps
    value = solves * world / (total_ms_max * 1e-3)

    # ---- e2e: the same rollouts through the host-facing call: host x0 / loss rates / reference in,
    # full state trajectories + tracking errors out (what the reference's script collects), copies timed ----
    e2e, e2e_same = None, None
    if not args.no_e2e:
        x0_h = torch.zeros(B, 4, dtype=torch.float64).pin_memory()
        p_h = torch.as_tensor(np.array([0.1 * ((ids0 + i) % 10) for i in range(B)])).pin_memory()
        ref_h = torch.as_tensor(REF.copy()).pin_memory()
        traj_h = torch.zeros(B, T + 1, 4, dtype=torch.float64).pin_memory()
        err_h = torch.zeros(B, dtype=torch.float64).pin_memory()
        h2d = x0_h.numel() * 8 + p_h.numel() * 8 + ref_h.numel() * 8
        d2h = traj_h.numel() * 8 + err_h.numel() * 8

        def rollout_host(seed):
            loop.reset(x0_h.numpy())                                            # H2D of the initial states
            p_d = p_h.to(dev, non_blocking=True)
            r_d = ref_h.to(dev, non_blocking=True)
            if args.e2e_staged:
                tr = loop.run(T, r_d, p_loss=p_d, seed=seed, id_offset=ids0, record=True, fused=True)
                traj_h.copy_(tr, non_blocking=True)
            else:
                # the rollout kernel writes every x_t straight into the pinned host buffer while it runs
                loop.run(T, r_d, p_loss=p_d, seed=seed, id_offset=ids0, record=True, fused=True, out=traj_h)
            err_h.copy_(loop.tracking_error(T), non_blocking=True)
            torch.cuda.synchronize()
        rollout_host(SEED + 500)
        barrier()
        t0 = time.perf_counter()
        ksteps = args.steps
        for k in range(ksteps):
            rollout_host(SEED + k)
        barrier()
        dt = D.all_reduce_max(torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64))
        loop.reset(x0_h.numpy())
        tr_dev = loop.run(T, ref_h.to(dev), p_loss=p_h.to(dev), seed=SEED + ksteps - 1, id_offset=ids0, record=True, fused=True)
        e2e_same = bool(torch.equal(tr_dev.cpu(), traj_h))
        del tr_dev
        e2e = {"value": B * T * ksteps * world / float(dt.item()), "unit": "solves/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "RemoteLoop.reset(host x0) + RemoteLoop.run(T, host ref, host loss rates, record=True, out=pinned "
                      "host buffer) -> rtmpc_loop_reset + rtmpc_loop_rollout; " +
                      ("trajectories [B,T+1,nx] recorded on the device and copied back" if args.e2e_staged else
                       "the kernel writes the trajectories [B,T+1,nx] into the pinned host buffer while it runs (they cross "
                       "the bus once, inside the timed region; checked against a device-recorded run: "
                       "checks.e2e_host_trajectory_matches_device)") +
                      ", tracking errors copied back to pinned host memory, wall clock"}

    extra, gather = {}, {}
    if not args.no_extra:
        del loop
        extra = extra_workloads(args, torch, dist, D, RemoteLoop, dev, rank, world, fp64_peak)
        gather = trajectory_all_gather(torch, dist, D, RemoteLoop, dev, rank, world, mpc, Z, ids0, p_loss, ref_d)

    if rank == 0:
        peaks, traffic = {}, None
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch of this workload, from the
            # committed `ncu --set full` capture (profiles/)
            if args.workload == "c2" and B == B_PER_GPU:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "rollout_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        flops = as_flops + f_it * int(iters_sum[0])      # rank 0's timed region: active-set kernel + IPM fallback
        achieved = flops / (solve_ms * 1e-3) / 1e12 if solve_ms > 0 else 0.0
        cpu_value, cpu_step_s, cores, cpu_n, cpu_loops = run_cpu_arm(args.cpu_solves, 1, 1)
        out = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"] + f", {B} closed-loop instances per GPU x {T} control steps per bench step",
                       "baseline_config": args.workload,
                       "instances_per_gpu": B, "control_steps": T, "qp_n": n, "qp_rows_two_sided": m,
                       "l2": "256 MB buffer written between timed rollouts (L2 flush)",
                       "rng": f"Philox4x32-10 on device, seed {SEED}, counter = global instance id"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "pipe": "fp64 FMA pipe (the QP kernel issues DFMA only - no tensor-core instruction; tcgen05 has no f64 "
                                                     "kind; 'tensor' = the compute roof of the bench schema, its denominator is a DMMA DGEMM)",
                         "kernel": kernel_name + " (dual active-set QP + closed-loop step, one warp per instance for all T steps; ipm_solve_kernel on handed-over instances)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "limiter": "no single roof, and neither arithmetic nor DRAM: in the committed ncu --set full capture of this launch "
                                    "(profiles/r2_rollout_ncu.md section 5) the GPC-level instruction cache serves requests at 85 % of its peak "
                                    "rate (gcc__cache_requests_type_instruction; SM instruction cache hit rate 74 %), the L1 data pipe runs at "
                                    "62 % of its wavefront rate (102 KB of shared-table reads per solve + step, L1 hit 99 %), issue slots are "
                                    "42 % busy with four warps per scheduler (128 registers), the FP64 pipe is 16 % busy; the method needs "
                                    "24 kflop per solve (round 1: 36), so frac moves less than the throughput",
                         "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json has no fp64 figure; "
                                        f"its hbm_gbs={peaks.get('hbm_gbs')}, bf16_tflops={peaks.get('bf16_tflops')})",
                         "algorithmic_flops_active_set": as_flops, "flops_per_ipm_iteration": f_it,
                         "ipm_iterations_in_timed_region": int(iters_sum[0]),
                         "active_set_steps_in_timed_region": int(iters_sum[1]),
                         "certification_rounds_in_timed_region": int(iters_sum[2]),
                         "kernel_ms_in_timed_region": solve_ms, "kernel_share_of_step": solve_ms / total_ms,
                         "mean_active_set_steps_per_solve": float(iters_sum[1]) / solves},
            "cpu_baseline": {"value": cpu_value, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_loops} closed loops of the same workload x {T} control steps = {cpu_n} QP solves "
                                       "(instance ids 0.., all ten loss rates; estimator -> QP -> packet -> actuator -> plant "
                                       "per step as in Results/results_linear_system.py:209-255), one process per core",
                             "same_config": True,
                             "caveat": "the reference's cvxpy + Clarabel stack is not installable offline: the QP is solved by the "
                                       "oracle's numpy Mehrotra interior-point port with a certified polish (about 15 ms per solve "
                                       "and core, inside the 2.5-20 ms the reference's own histogram of this call spans), the loop "
                                       "objects are the oracle's restatements of SmartActuator.py / Estimator.py (pinned to the "
                                       "reference's own code by tests/test_reference_pin.py).  A reported baseline, not a "
                                       "like-for-like solver comparison; it does not grow with --gpus"},
            "checks": {"status_counts[optimal,max_iter,infeasible,inaccurate]": status_all.tolist(),
                       "max_tube_violation": tube_max, "mean_tracking_error": float(err_all.mean().item()),
                       "stats_all_gather_ms": gather_ms, **gather,
                       "e2e_host_trajectory_matches_device": e2e_same,
                       "controller_setup_on_gpu": _SETUP},
            "extra_workloads": extra,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def timed_rollouts(torch, D, loop, T, ref_vec, p_loss, seed, ids0, steps, warmup, flush, barrier):
    """`steps` device-timed rollouts of an already built loop (L2 flushed between them); returns (ms max over ranks, stats)."""
    stream = torch.cuda.current_stream()
    for wi in range(warmup):
        loop.reset()
        loop.run(T, ref_vec, p_loss=p_loss, seed=seed + 1000 + wi, id_offset=ids0, fused=True)
    barrier()
    total_ms, stats = 0.0, np.zeros(8, np.int64)
    for k in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loop.reset()
        e0.record(stream)
        loop.run(T, ref_vec, p_loss=p_loss, seed=seed + k, id_offset=ids0, fused=True)
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        stats += loop.stats.cpu().numpy().astype(np.int64)
    barrier()
    t_ms = float(D.all_reduce_max(torch.tensor([total_ms], device=loop.dev, dtype=torch.float64)).item())
    stats = D.all_reduce_sum(torch.as_tensor(stats, device=loop.dev)).cpu().numpy()
    return t_ms, stats


def extra_workloads(args, torch, dist, D, RemoteLoop, dev, rank, world, fp64_peak):
    """BASELINE.json configs[2..4] in the same process, after the headline's timed region (so that the driver's BENCH / SCALE
    records carry them): c3 = extended variant, 65 536 instances in total split over the ranks (STRONG scaling: total work
    fixed as N grows); c4 = analytic cartpole plant, 32 768 instances per GPU (configs[3]'s 262 144 on 8 GPUs, weak);
    c5 = support sweep over 10^6 directions split over the ranks (strong).  Same timing rules as the headline (warm-up,
    L2 flush, CUDA events, max over ranks), 2 timed steps each."""
    from rtmpc_b200 import _lib
    out = {}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    steps, warmup = 2, 1
    for name, total, per_gpu, scaling in (("c3", 65536, None, "strong"), ("c4", None, 32768, "weak")):
        wl = WORKLOADS[name]
        mpc, Z = build_controller(extended=wl["extended"])
        if total is not None:
            ids0, B = D.shard(total, rank, world)
        else:
            B = per_gpu
            ids0, _ = D.shard(B * world, rank, world)
        loop = RemoteLoop(mpc, B, kind=wl["kind"], plant=wl["plant"], w_half=HW if wl["plant"] == "linear" else None, Z=Z)
        p_loss = torch.as_tensor(np.array([0.1 * ((ids0 + i) % 10) for i in range(B)]), device=dev)
        ref_vec = torch.as_tensor(REF.copy(), device=dev)
        t_ms, st = timed_rollouts(torch, D, loop, T_STEPS, ref_vec, p_loss, wl["seed"], ids0, steps, warmup, flush, barrier)
        n_inst = total if total is not None else B * world
        solves = n_inst * T_STEPS * steps
        flops = int(st[7]) + ipm_flops_per_iteration(mpc._prob.n, mpc._prob.m) * int(st[4])
        ach = flops / (t_ms * 1e-3) / 1e12 / world            # per GPU
        out[name] = {"workload": wl["text"], "instances_total": int(n_inst), "instances_per_gpu": int(B), "scaling": scaling,
                     "value": solves / (t_ms * 1e-3), "unit": "solves/s", "ms_per_step": t_ms / steps, "steps": steps,
                     "warmup": warmup, "kernel": mpc._prob.rollout_kernel.replace("*", "true" if wl["extended"] else "false"),
                     "roofline": {"achieved_tflops_per_gpu": ach, "peak": fp64_peak, "frac": (ach / fp64_peak) if fp64_peak else None},
                     "status_counts[optimal,max_iter,infeasible,inaccurate]": [int(v) for v in st[:4]],
                     "mean_active_set_steps_per_solve": float(st[5]) / max(solves, 1),
                     "max_tube_violation": float(D.all_reduce_max(loop.tube_max.max().reshape(1)).item())}
        del loop, mpc
    # c5: support sweep, 10^6 directions in total
    from rtmpc_b200 import polytope as pc
    s = load_sets()
    Mtot = 1_000_000
    off, M = D.shard(Mtot, rank, world)
    dirs = torch.as_tensor(support_directions(s, Mtot)[off:off + M], device=dev)
    V_h = np.ascontiguousarray(pc.extreme(pc.Polytope(s["Z_A"], s["Z_b"], normalize=False)))
    V = torch.as_tensor(V_h, device=dev)
    res = torch.empty(M, device=dev, dtype=torch.float64)
    L = _lib.lib()
    stream = torch.cuda.current_stream()
    sweep = lambda: _lib.check(L.rtmpc_support_sweep(_lib.ptr(V), V.shape[0], V.shape[1], _lib.ptr(dirs), M, _lib.ptr(res),   # noqa: E731
                                                    stream.cuda_stream), "rtmpc_support_sweep")
    sweep()
    barrier()
    total_ms = 0.0
    for k in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); sweep(); e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier()
    t_ms = float(D.all_reduce_max(torch.tensor([total_ms], device=dev, dtype=torch.float64)).item())
    fl = 2.0 * V.shape[1] * V.shape[0] * M * steps / (total_ms * 1e-3) / 1e12      # this rank's kernel
    out["c5"] = {"workload": "support-function sweep h_Z(a), 10^6 directions in total split over the ranks, Z = cartpole tube "
                             f"({V.shape[0]} vertices in {V.shape[1]}-D)", "directions_total": Mtot, "scaling": "strong",
                 "value": Mtot * steps / (t_ms * 1e-3), "unit": "directions/s", "ms_per_step": t_ms / steps, "steps": steps,
                 "warmup": 1, "kernel": "support_sweep_kernel",
                 "roofline": {"achieved_tflops_per_gpu": fl, "peak": fp64_peak, "frac": (fl / fp64_peak) if fp64_peak else None}}
    return out


def trajectory_all_gather(torch, dist, D, RemoteLoop, dev, rank, world, mpc, Z, ids0, p_loss, ref_d):
    """north_star / SURVEY 8(e): the one collective of the path - an all-gather of the state trajectories [B/G, T+1, nx]
    (and of the per-instance statistics) AFTER the rollouts.  One recorded rollout of the headline workload, then the
    gather timed with CUDA events (second call: the first one pays NCCL's connection set-up)."""
    B = p_loss.shape[0]
    loop = RemoteLoop(mpc, B, kind="tube", plant="linear", w_half=HW, Z=Z)
    loop.reset()
    traj = loop.run(T_STEPS, ref_d[0], p_loss=p_loss, seed=SEED, id_offset=ids0, record=True, fused=True)
    out = {"trajectory_all_gather_bytes_per_rank": int(traj.numel() * 8)}
    if world == 1:
        out.update(trajectory_all_gather_ms=0.0, trajectory_all_gather_gbs=None, trajectory_checksum=float(traj.sum().item()))
        return out
    D.all_gather_instances(traj, B * world)                   # connection set-up
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    allt = D.all_gather_instances(traj, B * world)
    e1.record()
    torch.cuda.synchronize()
    ms = float(D.all_reduce_max(torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)).item())
    recv = traj.numel() * 8 * (world - 1)                     # bytes every rank receives
    out.update(trajectory_all_gather_ms=ms, trajectory_all_gather_gbs=recv / (ms * 1e-3) / 1e9,
               trajectory_all_gather_shape=list(allt.shape), trajectory_checksum=float(allt.sum().item()))
    return out


def support_directions(s, M, seed=1):
    """BASELINE configs[4]: rows of Hd Acl^j and Hw Acl^j of the cartpole (the directions Darup's RPI
    and the constraint tightening evaluate), filled up with unit-normal Gaussian directions to M rows."""
    Acl = s["A"] - s["B"] @ np.atleast_2d(s["K"])
    Hd = np.vstack([s["X_A"], -s["U_A"] @ np.atleast_2d(s["K"])])
    rows, P = [], np.eye(4)
    for _ in range(308):                      # k* of the cartpole's Darup RPI
        rows.append(Hd @ P)
        rows.append(s["W_A"] @ P)
        P = P @ Acl
    D = np.vstack(rows)
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(np.vstack([D, rng.normal(size=(M - len(D), 4))])[:M])


def run_support_arm(args):
    """--workload c5: h_Z(a) = max over the vertices of the tube Z for 10^6 directions, sharded over the ranks."""
    import torch
    import torch.distributed as dist
    from rtmpc_b200 import _lib, polytope as pc, sets as up
    from rtmpc_b200 import distributed as D
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    s = load_sets()
    Mtot = 1_000_000
    off, M = D.shard(Mtot, rank, world)
    dirs_h = support_directions(s, Mtot)[off:off + M]
    V_h = np.ascontiguousarray(pc.extreme(pc.Polytope(s["Z_A"], s["Z_b"], normalize=False)))
    nv, dim = V_h.shape
    V = torch.as_tensor(V_h, device=dev); dirs = torch.as_tensor(dirs_h, device=dev); out = torch.empty(M, device=dev, dtype=torch.float64)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream()

    def sweep():
        _lib.check(L.rtmpc_support_sweep(_lib.ptr(V), nv, dim, _lib.ptr(dirs), M, _lib.ptr(out), stream.cuda_stream), "rtmpc_support_sweep")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        sweep()
    barrier()
    total_ms = 0.0
    launches0 = L.rtmpc_launch_count()
    with ClockSampler(local) as clk:
        barrier()
        for k in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); sweep(); e1.record(stream)
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        barrier()
    launches = L.rtmpc_launch_count() - launches0
    t_ms = float(D.all_reduce_max(torch.tensor([total_ms], device=dev, dtype=torch.float64)).item())
    value = Mtot * args.steps / (t_ms * 1e-3)
    # e2e: host directions in, host values out through the reference-facing call (sets.support_sweep)
    hv = up.support_sweep(V_h, dirs_h[:1000])
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        hv = up.support_sweep(V_h, dirs_h)
    barrier()
    dt = float(D.all_reduce_max(torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)).item())
    same = bool(np.array_equal(hv, out.cpu().numpy()))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = measure_fp64_peak(torch, dev)
        flops = 2.0 * dim * nv * M
        byts = (dim + 1) * 8.0 * M
        ms_launch = total_ms / args.steps
        # CPU arm: the oracle's support function (one HiGHS LP per direction) on a bounded sample
        from oracle import ref_sets as rs
        from oracle.ref_polytope import Polytope as OP
        Zo = OP(s["Z_A"], s["Z_b"], normalize=False)
        n_cpu = 300
        c0 = time.perf_counter()
        ref_vals = np.array([rs.support(Zo, d) for d in dirs_h[:n_cpu]])
        cpu_dt = time.perf_counter() - c0
        err = float(np.abs(ref_vals - hv[:n_cpu]).max())
        print(json.dumps({
            "metric": "support-function sweep directions/sec (BASELINE configs[4])", "value": value, "unit": "directions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"h_Z(a) over 10^6 directions (rows of Hd Acl^j, Hw Acl^j of the cartpole + Gaussian directions, seed 1) "
                                   f"sharded over the ranks; Z = the cartpole tube, {nv} vertices in {dim}-D staged in shared memory",
                       "baseline_config": "c5", "l2": "256 MB buffer written between timed sweeps (L2 flush)"},
            "e2e": {"value": Mtot * args.steps / dt, "unit": "directions/s", "h2d_bytes_per_step": int(dirs_h.nbytes + V_h.nbytes),
                    "d2h_bytes_per_step": int(M * 8), "api": "rtmpc_b200.sets.support_sweep(V, dirs) on host arrays -> rtmpc_support_sweep_host",
                    "identical_to_device_call": same},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA)", "kernel": "support_sweep_kernel",
                         "achieved": flops / (ms_launch * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": flops / (ms_launch * 1e-3) / 1e12 / fp64_peak, "traffic": None,
                         "algorithmic_hbm_gbs": byts / (ms_launch * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                         "peak_source": "cuBLAS DGEMM 6144^3 measured in this run"},
            "cpu_baseline": {"value": n_cpu / cpu_dt, "unit": "directions/s", "cores": 1, "kind": "port",
                             "sample": f"{n_cpu} directions, oracle support() = one HiGHS LP each (utils_polytope.py:12-23)"},
            "checks": {"max_abs_diff_vs_oracle_lp": err}}))
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    value, step_s, cores, total, n_loops = run_cpu_arm(args.cpu_solves, args.steps, args.warmup)
    s = load_sets()
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOADS["c2"]["text"] + f", {n_loops} closed-loop instances x {T_STEPS} control steps per bench step "
                                  "(a bounded sample of the GPU arm's 4096 per GPU: same controller, loss rates, disturbance box, reference)",
                      "baseline_config": "c2", "instances": n_loops, "control_steps": T_STEPS,
                      "note": "the reference's cvxpy+Clarabel stack cannot be installed offline; this is the oracle port of the same "
                              "closed loop (restated SmartActuator / Estimator objects, dense Mehrotra IPM + certified polish for the QP) "
                              "on all host cores, one process per core"},
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                            "sample": f"{n_loops} closed loops x {T_STEPS} steps per bench step, {args.steps} steps"},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "terminal_rows": int(s["Xf_A"].shape[0])}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-solves", type=int, default=8192, dest="cpu_solves",
                    help="QP solves of the CPU arm per step: round(N / 250) whole closed loops of the workload, at least one per core (~10 s on 16 cores)")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e")
    ap.add_argument("--e2e-staged", action="store_true", dest="e2e_staged",
                    help="e2e leg: record the trajectories in device memory and copy them back afterwards (default: the "
                         "kernel writes them into the pinned host buffer itself; 100 vs 107 M solves/s)")
    ap.add_argument("--no-extra", action="store_true", dest="no_extra",
                    help="skip the extra_workloads object (BASELINE configs[2..4] measured after the headline) and the "
                         "trajectory all-gather")
    ap.add_argument("--rollout-warps", type=int, default=0, dest="rollout_warps",
                    help="development: rtmpc_set_tuning(RTMPC_TUNE_ROLLOUT_WARPS) (0 = the library's choice)")
    ap.add_argument("--rollout-quantum", type=int, default=-1, dest="rollout_quantum",
                    help="development: rtmpc_set_tuning(RTMPC_TUNE_ROLLOUT_QUANTUM) (-1 = the library's default)")
    ap.add_argument("--rollout-carry", type=int, default=-1, dest="rollout_carry",
                    help="development: rtmpc_set_tuning(RTMPC_TUNE_ROLLOUT_CARRY) (-1 = the library's default, 1)")
    ap.add_argument("--rollout-fixed-dims", type=int, default=-1, dest="rollout_fixed_dims",
                    help="development: rtmpc_set_tuning(RTMPC_TUNE_ROLLOUT_FIXED_DIMS) (-1 = the library's default, 1)")
    ap.add_argument("--instances", type=int, default=B_PER_GPU, help="closed-loop instances per GPU (BASELINE configs[1]: 4096; configs[2] = --workload c3 --instances 8192 on 8 GPUs, "
                         "configs[3] = --workload c4 --instances 32768 on 8 GPUs)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"],
                    help="c2 (default, the headline): BASELINE configs[1]; c3: extended variant; c4: analytic cartpole plant; "
                         "c5: support-function sweep over 10^6 directions (its own metric)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "c5":
        run_support_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

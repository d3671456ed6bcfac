/*
 * rtmpc.h -- C ABI of the B200-native batched tube-MPC hot path.
 *
 * The reference (EricssonResearch/Robust-Tracking-MPC-over-Lossy-Networks) has no FFI of its own:
 * its "boundary" for this path is a Python call into cvxpy -> Clarabel (a Rust extension module)
 * plus a handful of numpy statements.  Each entry point below names the reference interface it
 * replaces (file:line under the reference root).  All functions
 *   - take plain pointers and sizes (no torch / numpy types),
 *   - return 0 on success and a negative code on failure (rtmpc_last_error() has the text),
 *   - never throw and never fall back to a CPU implementation,
 *   - enqueue their kernels on the CUDA stream passed as `stream` (a cudaStream_t cast to void*;
 *     NULL = default stream) and do not synchronise unless stated.
 * Pointers named d_* are DEVICE pointers, h_* are HOST pointers.  All floating point is FP64,
 * all matrices are row-major.
 */
#ifndef RTMPC_H
#define RTMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTMPC_ABI_VERSION 3

/* per-instance solver status (replaces cvxpy's `prob.status` string, TubeTrackingMPC.py:185) */
#define RTMPC_OPTIMAL            0   /* KKT-certified active-set point                          */
#define RTMPC_MAX_ITER           1   /* no convergence (reference: status printed, values kept) */
#define RTMPC_INFEASIBLE         2   /* reference: `.value is None` -> U_t = None (:215-221)    */
#define RTMPC_OPTIMAL_INACCURATE 3   /* usable solution without the 1e-11 certificate: interior-point tolerance reached, or
                                      rows that contradict each other by less than 1e-8 (relative) */
#define RTMPC_FALLBACK_STATUS  (-2)  /* transient, never returned: active-set kernel handed the instance
                                        to the interior-point kernel inside rtmpc_qp_solve           */

/* solution methods of rtmpc_qp_solve (rtmpc_qp_set_method) */
#define RTMPC_METHOD_ACTIVE_SET      0  /* dual active-set kernel, interior point only as fallback (default) */
#define RTMPC_METHOD_INTERIOR_POINT  1  /* interior-point kernel for every instance                          */

typedef struct rtmpc_qp rtmpc_qp;       /* one condensed QP resident on one GPU                  */
typedef struct rtmpc_loop rtmpc_loop;   /* closed-loop model + per-instance state on one GPU     */

/*
 * Host-prepared description of one condensed, equilibrated QP (see DESIGN.md "data layout").
 *    min 1/2 z'Hs z + (Fx x_init + Fr ref)'z    s.t.  lo0 + Lx x_init <= G z <= up0 + Ux x_init
 * plus parameter rows  parC x_init <= parh.  Replaces the cvxpy problem objects built in
 * TubeTrackingMPC.py:104-156 / :253-299, TrackingMPC.py:62-114, TubeRegulatorMPC.py:109-143,
 * RegulatorMPC.py:45-76.
 */
typedef struct rtmpc_qp_desc {
    int32_t nx, nu, N;      /* system sizes and horizon                                          */
    int32_t n;              /* decision variables                                                */
    int32_t npad;           /* padded columns (multiple of 4, >= n)                              */
    int32_t m;              /* two-sided rows                                                    */
    int32_t mpad;           /* padded rows (multiple of 32, >= m)                                */
    int32_t np;             /* parameter rows                                                    */
    int32_t nz;             /* length of the un-condensed vector [x_0..x_N | u | x_bar | u_bar]  */
    int32_t nss;            /* nx+nu if the variant has (x_bar,u_bar), else 0                    */
    const double* Hs;       /* [npad*npad]                                                       */
    const double* Hinv;     /* [npad*npad]                                                       */
    const double* G;        /* [mpad*npad]                                                       */
    const double* Y;        /* [mpad*npad]   G Hinv                                              */
    const double* Fx;       /* [npad*nx]                                                         */
    const double* Fr;       /* [npad*nx]                                                         */
    const double* lo0;      /* [mpad]                                                            */
    const double* up0;      /* [mpad]                                                            */
    const double* Lx;       /* [mpad*nx]                                                         */
    const double* Ux;       /* [mpad*nx]                                                         */
    const uint8_t* has_lo;  /* [mpad]                                                            */
    const uint8_t* has_up;  /* [mpad]                                                            */
    const double* parC;     /* [np*nx]                                                           */
    const double* parh;     /* [np]                                                              */
    const double* Dscale;   /* [npad]  z_unscaled = Dscale .* z                                  */
    const double* Phi;      /* [nz*npad]  un-condensed = Phi z_unscaled + Psi x_init             */
    const double* Psi;      /* [nz*nx]                                                           */
    const double* Kss;      /* [nu*nx]  steady-state gain K used in U_t's last column, or NULL   */
    double s_floor;         /* smallest initial slack (scaled units)                             */
    double sc_b;            /* 1 + typical |bound| (scaled units)                                */
    int32_t max_iter;       /* interior-point iteration cap (Clarabel default 200; we use 60)    */
    int32_t min_rows;       /* pad the rows at least to this count (0: as few as possible); two problems
                               that one rollout switches between need the same count, see rtmpc_qp_rows */
    const int32_t* shift;   /* [mpad] row holding the same constraint one stage earlier (-1: none),
                               used to move a warm-start active set one control step on; or NULL    */
} rtmpc_qp_desc;

/* Library / device ------------------------------------------------------------------------- */
int         rtmpc_abi_version(void);
const char* rtmpc_last_error(void);                 /* thread-local text of the last failure   */
int         rtmpc_device_count(void);
int         rtmpc_set_device(int device);

/*
 * Launch tuning, process-wide (nothing is read from the environment; results never depend on these, only speed - except the last
 * bits with RTMPC_TUNE_ROLLOUT_CARRY, see there):
 *   RTMPC_TUNE_ROLLOUT_QUANTUM  control steps a warp runs one closed loop for before it hands the loop's state on and
 *                               draws the next (instance, chunk) ticket (rtmpc_loop_rollout; default 25).  0: every warp
 *                               keeps its instance for the whole rollout.  Time slicing needs all thread blocks of the
 *                               launch resident at once: the library launches it cooperatively and falls back to 0 by
 *                               itself when the device / context cannot guarantee that (SM-limited MPS, green contexts).
 *   RTMPC_TUNE_ROLLOUT_WARPS    warps per thread block of the rollout kernel (0: automatic)
 *   RTMPC_TUNE_AS_WARPS         upper bound on the warps per thread block of the active-set kernel of rtmpc_qp_solve
 *                               (0: automatic)
 *   RTMPC_TUNE_ROLLOUT_CARRY    1 (default): rtmpc_loop_rollout warm-starts every solve of a closed loop from the previous
 *                               control step's certified working set AS IT IS and keeps that set's inverse on chip from
 *                               one step to the next (no factorisation per solve; the inverse is rebuilt from scratch at
 *                               control steps that are multiples of RTMPC_TUNE_ROLLOUT_QUANTUM, so the results do not depend
 *                               on how the batch is cut, sliced or spread over GPUs).  0: the working set is moved one stage
 *                               earlier (desc.shift) and inverted from scratch in every solve, exactly as rtmpc_qp_solve
 *                               does.  The one knob that touches results, in the last bits only: every solution carries the
 *                               same KKT certificate either way, but the certification refines the multipliers with the
 *                               inverse as approximate inverse, so with 1 a rollout agrees with the step-by-step path
 *                               (rtmpc_qp_solve + rtmpc_loop_step) to rounding, with 0 bit for bit.
 *   RTMPC_TUNE_ROLLOUT_FIXED_DIMS 1 (default): a single-problem rollout whose controller has the dimensions of the
 *                               reference's cartpole study (nx = 4, nu = 1, N = 20, fixed initial state: 21 unknowns) runs the
 *                               kernel instantiation that has those dimensions as compile-time constants.  0: the general
 *                               instantiation.  Same arithmetic in the same order: identical bits.
 *   RTMPC_TUNE_CERT_FACTORED    1 (default): the certification of a solve accepts an evaluation of the row values G z - up other than
 *                               G' z where every row outside the working set clears the tolerance by a per-row bound on that
 *                               evaluation's distance from G' z (rounding of the evaluations and of the tables, computed by
 *                               rtmpc_qp_create in long double): first the values the active-set steps arrived at (nothing is
 *                               read), then Ex x + Tr r - up0 - W[:,A] (s lam) from scratch (2 nx columns, |A| rows); otherwise,
 *                               and with 0 always, the rows are recomputed from G' z.  The solution z never depends on it; 92 % /
 *                               4 % / 4 % of the benchmark's certifications end in the three tiers.
 * A negative value restores the default.  rtmpc_get_tuning returns the value in force (-1: unknown knob).
 */
#define RTMPC_TUNE_ROLLOUT_QUANTUM 0
#define RTMPC_TUNE_ROLLOUT_WARPS   1
#define RTMPC_TUNE_AS_WARPS        2
#define RTMPC_TUNE_ROLLOUT_CARRY   3
#define RTMPC_TUNE_ROLLOUT_FIXED_DIMS 4
#define RTMPC_TUNE_CERT_FACTORED   5
int     rtmpc_set_tuning(int32_t knob, int32_t value);
int32_t rtmpc_get_tuning(int32_t knob);

/* QP object: replaces `generate_optimization_problem` (TubeTrackingMPC.py:104-156 and siblings) */
int  rtmpc_qp_create(const rtmpc_qp_desc* desc, rtmpc_qp** out);
void rtmpc_qp_destroy(rtmpc_qp* qp);

/*
 * Solve B independent instances.  Replaces `self._prob.solve(solver=CLARABEL, tol_gap_abs=1e-7,
 * tol_gap_rel=1e-7)` + reading the variables' `.value` (TubeTrackingMPC.py:170-194, :307-349;
 * TrackingMPC.py:116-135; TubeRegulatorMPC.py:156-166; RegulatorMPC.py:78-91) and `encapsulate`
 * (TubeTrackingMPC.py:211-227, :363-369; TrackingMPC.py:143-158).
 *   d_x_init [B*nx], d_ref [B*nx] (may be NULL for the regulator variants)
 *   d_sel    [B] or NULL: instance b is solved only if d_sel[b] == sel_value (the gamma_t switch of
 *                          ExtendedTubeTrackingMPC.solve_optimization_problem, :307-349)
 *   d_warm   [B*rtmpc_qp_warm_stride()] int32 or NULL: per-instance warm-start state, read and
 *            rewritten by every solve.  Entry 0 = number of active rows of the last certified
 *            solution (-1 = none, initialise the array with -1), then the signed rows.  The next
 *            solve moves that set one stage earlier (desc.shift) and starts from it; the result
 *            does not depend on it (every solution is KKT-certified), only the work does.
 *   d_z      [B*nz]        un-condensed solution [x_0..x_N | u_0..u_{N-1} | x_bar | u_bar], or NULL
 *   d_U_t    [B*(N+1)*nu]  packet payload, time-major: U_t[b][k][:] ; last column u_bar + K x_bar
 *                          (only u_0..u_{N-1} are written when the variant has no steady state)
 *   d_status [B] (may be NULL).  With d_sel, entries of unselected instances are left as they are, except that a
 *            stale value <= RTMPC_FALLBACK_STATUS (uninitialised memory) is replaced by -1: the hand-over to the
 *            interior-point kernel goes through this array
 *   d_iters  [B] (may be NULL): bits 0-11 interior-point iterations, bits 12-23 active-set steps
 *            (rows added + rows dropped), bits 24-27 certification / endgame rounds (saturating), bits 28-30
 *            diagnostics: why the active-set kernel last refactorised or gave up (0 = it never did)
 * Infeasible instances get NaN payloads (the reference returns None).
 */
int rtmpc_qp_solve(rtmpc_qp* qp, int32_t B, const double* d_x_init, const double* d_ref,
                   const int32_t* d_sel, int32_t sel_value, int32_t* d_warm, double* d_z, double* d_U_t,
                   int32_t* d_status, int32_t* d_iters, void* stream);
int32_t rtmpc_qp_warm_stride(rtmpc_qp* qp);     /* int32 entries per instance of d_warm (npad + 1) */
int32_t rtmpc_qp_rows(rtmpc_qp* qp);            /* rows the kernels work on (m padded to a multiple of 64) */
const char* rtmpc_qp_rollout_kernel(rtmpc_qp* qp);  /* name of the kernel instantiation rtmpc_loop_rollout launches for this
                                                   problem, e.g. "rollout_kernel<5,16,*>" (profiles, bench.py); * = false /
                                                   true: one problem / the extended controller's two              */

/* Same call with HOST buffers: copies in, solves, copies out, synchronises.  This is the
 * reference-facing plugin call (numpy arrays in, numpy arrays out).  warm != 0 keeps the
 * warm-start state inside the handle between calls (instance b of consecutive calls with the same
 * B is taken to be the same closed loop one control step later; rtmpc_qp_warm_reset forgets it). */
int rtmpc_qp_solve_host(rtmpc_qp* qp, int32_t B, const double* h_x_init, const double* h_ref,
                        const int32_t* h_sel, int32_t sel_value, int32_t warm, double* h_z, double* h_U_t,
                        int32_t* h_status, int32_t* h_iters);
int rtmpc_qp_warm_reset(rtmpc_qp* qp);

/* RTMPC_METHOD_*: which kernel rtmpc_qp_solve runs (default: active set + interior-point fallback) */
int rtmpc_qp_set_method(rtmpc_qp* qp, int32_t method);
/* active-set steps (rows added + dropped) after which an instance is handed to the interior-point kernel;
 * <= 0 restores the default 16*npad + 128.  The result does not depend on it, only which kernel produces it. */
int rtmpc_qp_set_step_cap(rtmpc_qp* qp, int32_t max_steps);
/* device counter (or NULL) to which the active-set kernel adds the algorithmic FP64 flops it executes
 * (bench.py's roofline numerator) */
int rtmpc_qp_set_work_counter(rtmpc_qp* qp, uint64_t* d_counter);

/* number of kernels this library has launched so far in this process (bench's gpu_launches) */
int64_t rtmpc_launch_count(void);

/* Closed loop ------------------------------------------------------------------------------- */

#define RTMPC_ACT_SMART       0   /* SmartActuator            (SmartActuator.py:11-123)        */
#define RTMPC_ACT_CONSISTENT  1   /* ConsistentActuator       (SmartActuator.py:125-231)       */
#define RTMPC_ACT_EXTENDED    2   /* ConsistentActuator(is_extended_MPC_used=True) + RobustEstimator */

#define RTMPC_PLANT_LINEAR    0   /* x+ = A x + B u + w   (Results/results_linear_system.py:248) */
#define RTMPC_PLANT_CARTPOLE  1   /* analytic cartpole ODE, 10 sub-steps of 1/500 s, no w
                                     (replaces PyBullet, Results/results_nonlinear_system.py:255-345) */

typedef struct rtmpc_loop_desc {
    int32_t nx, nu, N;
    int32_t actuator;       /* RTMPC_ACT_*                                                      */
    int32_t plant;          /* RTMPC_PLANT_*                                                    */
    int32_t nz_rows;        /* rows of the tube polytope Z for the containment check, 0 = off   */
    const double* A;        /* [nx*nx]                                                          */
    const double* B;        /* [nx*nu]                                                          */
    const double* K;        /* [nu*nx] steady-state gain (u = -K x)                             */
    const double* K_plant;  /* [nu*nx] ancillary gain                                           */
    const double* Hz;       /* [nz_rows*nx] or NULL                                             */
    const double* hz;       /* [nz_rows]    or NULL                                             */
    const double* w_half;   /* [nx] half-widths of the disturbance box (RNG mode)               */
    double cart_params[8];  /* M, m, I, g, l, dt, substeps, unused                              */
} rtmpc_loop_desc;

int  rtmpc_loop_create(const rtmpc_loop_desc* desc, int32_t B, rtmpc_loop** out);
void rtmpc_loop_destroy(rtmpc_loop* loop);

/* (re)initialise all B instances: x = x_nom = x_hat = x0[b], t = 0, q_t = s_t = 0
 * (constructors of SmartActuator.py:13-24,129-144 and Estimator.py:11-26).
 * rtmpc_loop_reset:        h_x0 is [B*nx] on the HOST; returns when the state is in place (waits for the default stream
 *                          only, never for the whole device).
 * rtmpc_loop_reset_device: d_x0 is [B*nx] on the DEVICE (NULL: all zeros); enqueued on `stream`, no synchronisation. */
int rtmpc_loop_reset(rtmpc_loop* loop, const double* h_x0);
int rtmpc_loop_reset_device(rtmpc_loop* loop, const double* d_x0, void* stream);

/* device views of the per-instance state, for the host classes and the tests */
double*  rtmpc_loop_x(rtmpc_loop* loop);        /* [B*nx] plant state                            */
double*  rtmpc_loop_x_nom(rtmpc_loop* loop);    /* [B*nx] nominal state (ConsistentActuator)     */
double*  rtmpc_loop_x_hat(rtmpc_loop* loop);    /* [B*nx] remote estimate                        */
int32_t* rtmpc_loop_q_t(rtmpc_loop* loop);      /* [B]    Estimator.get_qt()                     */
int32_t* rtmpc_loop_s_t(rtmpc_loop* loop);      /* [B]    SmartActuator.get_s_t()                */
int32_t* rtmpc_loop_Theta(rtmpc_loop* loop);    /* [B]    SmartActuator.get_Theta_t()            */
int32_t* rtmpc_loop_alive(rtmpc_loop* loop);    /* [B]    0 once a controller returned U_t=None  */
double*  rtmpc_loop_err_acc(rtmpc_loop* loop);  /* [B]    running sum of ||x_t - ref_t||^2       */
double*  rtmpc_loop_tube_max(rtmpc_loop* loop); /* [B]    max_t max_i (Hz (x - x_nom) - hz)_i    */
double*  rtmpc_loop_u(rtmpc_loop* loop);        /* [B*nu] last applied input                     */
int32_t* rtmpc_loop_gamma(rtmpc_loop* loop);    /* [B]    gamma of the last step (1 before step 0): the
                                                   switch ExtendedTubeTrackingMPC sees, TubeTrackingMPC.py:312 */
int32_t  rtmpc_loop_time(rtmpc_loop* loop);     /* internal timer t (SmartActuator._t == Estimator._t) */

/*
 * One closed-loop step for every instance, given this step's controller packets
 * (ConsistentActuator.process_packet SmartActuator.py:174-213 -> plant step
 * Results/results_linear_system.py:248 -> Estimator.update_estimate Estimator.py:43-78 /
 * RobustEstimator.update_estimate Estimator.py:113-156).
 *   d_U_t [B*(N+1)*nu], d_status [B] from rtmpc_qp_solve
 *   d_x_nom0: x_nom[:,0] of this step's solve, element k of instance b at d_x_nom0[b*x_nom0_stride+k]
 *             (pass rtmpc_qp_solve's d_z with stride nz), or NULL; used by the extended variant only
 *   d_ref [B*nx] (only for the tracking-error accumulator)
 *   d_theta, d_gamma [B] (0/1) and d_w [B*nx]: explicit arrays; pass NULL for all three to draw
 *   them on the device (Philox4x32-10, key = seed, counter = (instance id + id_offset, t)) with loss
 *   probability d_p_loss[b]; step 0 is forced lossless (Results/results_linear_system.py:211-214).
 */
int rtmpc_loop_step(rtmpc_loop* loop, const double* d_U_t, const int32_t* d_status,
                    const double* d_x_nom0, int64_t x_nom0_stride, const double* d_ref,
                    const int32_t* d_theta, const int32_t* d_gamma, const double* d_w,
                    const double* d_p_loss, uint64_t seed, int64_t id_offset,
                    double* d_traj_x, int64_t traj_stride, void* stream);

/*
 * T closed-loop steps of every instance in ONE launch (the whole experiment loop of
 * Results/results_linear_system.py:209-255 for B Monte-Carlo instances): per step and instance the
 * controller's QP (TubeTrackingMPC.determine_packet, TubeTrackingMPC.py:196-209; TrackingMPC for the
 * smart actuator) followed by the step of rtmpc_loop_step.  One warp owns one instance for all T
 * steps; x_hat feeds the next solve on chip and each solve is warm-started from the previous one.
 *   qp       the controller's problem (default method RTMPC_METHOD_ACTIVE_SET); sizes must match
 *   qp_received  NULL, or for RTMPC_ACT_EXTENDED the "packet received" problem of ExtendedTubeTrackingMPC
 *            (TubeTrackingMPC.py:253-299): each step solves it instead of qp where gamma_{t-1} = 1 (:307-349) and
 *            the packet's x_nom_0 resets the local nominal state (SmartActuator.py:219-222); both problems must be
 *            padded to the same row count (rtmpc_qp_desc.min_rows)
 *   d_ref    reference of instance b at step k (0-based within this call) at
 *            d_ref[k*ref_stride_t + b*ref_stride_b + 0..nx)  (strides in doubles; 0 = shared)
 *   d_theta, d_gamma [T*B], d_w [T*B*nx]: explicit realisations, or NULL for the device RNG of
 *            rtmpc_loop_step (same counters, so a rollout equals T single steps bit for bit)
 *   d_traj_x [B*traj_stride] or NULL: x_t at d_traj_x[b*traj_stride + t*nx + :], t = 0..T.  Device memory, or pinned
 *            (mapped) host memory: the kernel then writes the trajectory through the bus while it runs and no copy
 *            follows (RemoteLoop.run(..., out=pinned tensor); bench.py's e2e leg: 107 against 100 M solves/s staged)
 *   d_stats  [8] uint64 or NULL, accumulated: solves by status [4], interior-point iterations,
 *            active-set steps, certification rounds, algorithmic flops of the active-set method
 * Instances the active-set method hands over are solved by the interior-point kernel between
 * relaunches; the call synchronises `stream` before it returns (it has to read how many instances were handed over;
 * nothing else is staged through the host).  With more instances than resident warps the loops are time-sliced
 * (RTMPC_TUNE_ROLLOUT_QUANTUM above); results are identical either way.
 */
int rtmpc_loop_rollout(rtmpc_loop* loop, rtmpc_qp* qp, rtmpc_qp* qp_received, int32_t T, const double* d_ref, int64_t ref_stride_t,
                       int64_t ref_stride_b, const int32_t* d_theta, const int32_t* d_gamma, const double* d_w,
                       const double* d_p_loss, uint64_t seed, int64_t id_offset, double* d_traj_x,
                       int64_t traj_stride, uint64_t* d_stats, void* stream);

/*
 * The reference's per-object call order, on caller-owned device arrays (all [B*...], FP64/int32):
 *   u, plant_packet = actuator.process_packet(packet, x_t, theta_t)   SmartActuator.py:31-54 / :174-213
 * kind = RTMPC_ACT_SMART (d_A, d_B, d_K_plant, d_x_nom may be NULL) or RTMPC_ACT_CONSISTENT /
 * RTMPC_ACT_EXTENDED.  t is the actuator's internal timer (the caller increments it afterwards).
 * State in/out: d_buf [B*(N+1)*nu], d_x_nom [B*nx], d_s_t, d_Theta, d_last_loss [B] (last_loss starts
 * at -1: the O(1) form of the growing theta vector, SmartActuator.py:57-71).
 * Outputs: d_u_out [B*nu]; plant packet fields d_pkt_x [B*nx] ('x_t': x_nom for the consistent
 * actuator, x otherwise, SmartActuator.py:204-207) and d_pkt_xnom [B*nx] ('x_nom_t', may be NULL).
 */
int rtmpc_actuator_process(int32_t B, int32_t nx, int32_t nu, int32_t N, int32_t kind, int32_t t,
                           const double* d_A, const double* d_B, const double* d_K, const double* d_K_plant,
                           const double* d_x_t, const double* d_U_t, const double* d_x_nom0,
                           const int32_t* d_q_pkt, const int32_t* d_theta, double* d_buf, double* d_x_nom,
                           int32_t* d_s_t, int32_t* d_Theta, int32_t* d_last_loss, double* d_u_out,
                           double* d_pkt_x, double* d_pkt_xnom, void* stream);

/*
 *   estimator.update_estimate(plant_packet, gamma_t)   Estimator.py:43-78 / RobustEstimator :113-156
 * d_hist holds every sent control sequence, [n_hist][B][(N+1)*nu] (the reference's unbounded
 * `_controlSequences` list, Estimator.py:34-41); d_pkt_s [B] is the packet's s_t.  t is the
 * estimator's internal timer.  robust != 0 selects RobustEstimator (needs d_pkt_xnom, d_K_plant and
 * d_x_nom0_mpc = the value passed to store_current_optimal_inital_nominal_plant_states).
 */
int rtmpc_estimator_update(int32_t B, int32_t nx, int32_t nu, int32_t N, int32_t robust, int32_t t,
                           int32_t n_hist, const double* d_A, const double* d_B, const double* d_K,
                           const double* d_K_plant, const double* d_pkt_x, const double* d_pkt_xnom,
                           const int32_t* d_pkt_s, const int32_t* d_gamma, const double* d_hist,
                           const double* d_x_nom0_mpc, double* d_x_hat, int32_t* d_q_t, void* stream);

/* Support-function sweep: out[j] = max_v <dirs[j,:], V[v,:]>  (utils_polytope.support,
 * utils_polytope.py:12-23, evaluated on the vertex representation instead of one LP per call). */
int rtmpc_support_sweep(const double* d_V, int32_t nv, int32_t dim, const double* d_dirs, int64_t M,
                        double* d_out, void* stream);
int rtmpc_support_sweep_host(const double* h_V, int32_t nv, int32_t dim, const double* h_dirs, int64_t M,
                             double* h_out);

/*
 * Batched low-dimensional LPs over one shared H-representation (the LPs behind the offline set
 * pipeline once a set has no tractable vertex representation - utils_polytope.support utils_polytope.py:12-23 inside the
 * maximal-output-admissible-set iteration :247-268, polytope's subset test, polytope.reduce TubeRegulatorMPC.py:74):
 *     val[b] = max  obj[b]' x   s.t.  H x <= h  (bound of row relax_row[b] relaxed by relax_by),
 *                                     extra[b][e][0..dim) x <= extra[b][e][dim]  for e < n_extra,   |x_i| <= box
 * One warp per LP, dual simplex from the box vertex (csrc/rtmpc_lp.cu).  dim <= 16.
 *   d_HT [dim*mpad]  H TRANSPOSED (row r, coordinate k at d_HT[k*mpad + r]); d_h [mpad]; rows >= m are ignored
 *   d_obj [B*dim]; d_relax_row [B] or NULL (-1: no row relaxed); d_extra [B*n_extra*(dim+1)] or NULL
 *   d_val [B]; d_x [B*dim] or NULL (the optimal vertex); d_iters [B] or NULL (simplex iterations)
 *   d_status [B]: 0 optimal, 1 iteration limit / singular basis, 2 infeasible (val = -1e300),
 *                 4 unbounded inside the box (val = +1e300)
 * rtmpc_lp_solve_host: the same with HOST buffers and H row-major [m*dim]; copies, solves, copies back, synchronises.
 */
int rtmpc_lp_solve(const double* d_HT, const double* d_h, int32_t m, int32_t mpad, int32_t dim, const double* d_obj,
                   const int32_t* d_relax_row, double relax_by, const double* d_extra, int32_t n_extra, int64_t B,
                   double box, double* d_val, double* d_x, int32_t* d_status, int32_t* d_iters, void* stream);
int rtmpc_lp_solve_host(const double* h_H, const double* h_h, int32_t m, int32_t dim, const double* h_obj,
                        const int32_t* h_relax_row, double relax_by, const double* h_extra, int32_t n_extra, int64_t B,
                        double box, double* h_val, double* h_x, int32_t* h_status, int32_t* h_iters);

/* Model-error sweep for the disturbance set W (Results/estimate_W_for_Cartpole.py:79-127, analytic cartpole ODE instead
 * of PyBullet): B independent runs; run b starts at x0[b], the plant is driven by the zero-order-hold LQR law
 * u_k = -K x_k for T control periods (cart_params as in rtmpc_loop_desc: M, m, I, g, l, dt, substeps), and
 * w[b][k][:] = x_{k+1} - Acl x_k  is recorded.  x_final may be NULL.  nx = 4, nu = 1. */
int rtmpc_model_error_sweep(const double cart_params[8], int32_t B, int32_t T, const double* d_x0, const double* d_K,
                            const double* d_Acl, double* d_w, double* d_x_final, void* stream);
int rtmpc_model_error_sweep_host(const double cart_params[8], int32_t B, int32_t T, const double* h_x0,
                                 const double* h_K, const double* h_Acl, double* h_w, double* h_x_final);

#ifdef __cplusplus
}
#endif
#endif /* RTMPC_H */

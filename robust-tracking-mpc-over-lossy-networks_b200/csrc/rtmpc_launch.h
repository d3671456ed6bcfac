// Host-side launchers of the QP kernels (one translation unit per kernel family so they compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "rtmpc_ipm.cuh"
#include "rtmpc_loop.cuh"

namespace rtmpc {

struct QPLaunch {
    int B;
    const double *x_init, *ref;
    const int* sel;
    int sel_value;
    double *z, *U;
    int *status, *iters, *warm;
    unsigned long long* work;
    cudaStream_t stream;
};

// Resident warps per SM for B instances on num_sms SMs (one warp per instance, at most wpb per SM): the count that
// minimises  rounds(w) * (1 + 0.02 w)  -- a round with w warps per SM takes about (1 + 0.02 w) times a lone warp's
// time (measured: 10.2 ms at 1 warp, 13.4 ms at 16), so 4096 instances run as 2 x 14 rather than 16 + 12 per SM.
inline int balanced_warps(int B, int num_sms, int wpb) {
    int best = 1;
    double best_cost = 1e300;
    for (int w = 1; w <= wpb; ++w) {
        const long long slots = (long long)num_sms * w;
        const double rounds = (double)((B + slots - 1) / slots);
        const double cost = rounds * (1.0 + 0.02 * w);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = w; }
    }
    return best;
}

// Process-wide launch tuning (rtmpc_set_tuning, include/rtmpc.h); never read from the environment.
struct Tuning {
    int rollout_quantum = 25;   // control steps per ticket of the time-sliced rollout; 0: whole chains per warp
    int rollout_warps = 0;      // warps per CTA of the rollout kernel; 0: automatic
    int as_warps = 0;           // cap on the warps per CTA of the active-set solve kernel; 0: automatic
    int rollout_carry = 1;      // 1: the rollout carries each instance's working set and its inverse from one control step to the next
    int rollout_fixed_dims = 1; // 1: problems with the cartpole controller's dimensions run the instantiation that has them as constants
    int cert_factored = 1;      // 1: the certification accepts row values through the factored tables where they clear the tolerance by the
                                //    rounding bound (QPDev::kap); 0: every certification forms the rows from G' z
};
Tuning& tuning();

// interior-point kernel: picks the instantiation for (n, mpad); returns false if none fits
bool ipm_configure(const QPDev& P, int max_smem, int* wpb, size_t* smem, cudaError_t* err);
cudaError_t ipm_launch(const QPDev& P, int wpb, size_t smem, int num_sms, const QPLaunch& a);

// dual active-set kernel
int as_padded_rows(int mpad);    // rows the active-set kernel's instantiation for mpad works on (multiple of 64), -1: none
bool as_configure(const QPDev& P, int max_smem, int* wpb, size_t* smem, cudaError_t* err);
cudaError_t as_launch(const QPDev& P, int wpb, size_t smem, int num_sms, const QPLaunch& a);

// persistent closed-loop rollout (rtmpc_rollout.cu); uses the active-set kernel's launch shape
struct RolloutArgs {
    int B, T, t0;                 // instances; steps t0 .. T-1 are taken
    const double* ref;            // reference of instance b at step t: ref[(t - t0) * ref_stride_t + b * ref_stride_b + :]
    long long ref_stride_t, ref_stride_b;
    const int *theta, *gamma;     // explicit realisations [T - t0][B] (and w [T - t0][B][nx]) or NULL: device RNG
    const double* w;
    const double* p_loss;
    unsigned long long seed;
    long long id_offset;
    double* traj;
    long long traj_stride;
    int *warm, *warm1;            // [B * (npad + 1)] warm-start records of the problem(s)
    int two;                      // 1: two problems, switched per step on gamma_{t-1}
    long long z_stride;           // entries per instance of z
    double *U, *z;                // [B * (N+1) * nu] packet payloads; [B * nz] or NULL
    int *status, *iters;          // [B] of the last solve
    int *inst_t, *pending;        // [B] per-instance time; 1 = current step solved by the interior-point kernel
    double* ref_pending;          // [B * nx] reference of the step a parked instance waits at
    int* n_pending;               // instances parked by this launch
    // time slicing: a ticket is (instance, chunk of `quantum` control steps); warps draw tickets from `next`, a ticket
    // waits for its instance's previous one through `done` (both zeroed before every launch).  quantum <= 0: whole rollouts
    int* next;
    int* done;
    int quantum;
    int carry;                    // 1: warm starts on the working set and inverse carried from the previous control step
                                  //    (as_solve_instance); 0: moved one stage and inverted from scratch every step
    int refresh;                  // the carried inverse is dropped at control steps t with t % refresh == 0 (0: never)
    unsigned long long* stats;    // [8] status counts[4], IPM iterations, active-set steps, rounds, flops (or NULL)
};
bool rollout_configure(const QPDev& P, int max_smem, cudaError_t* err);
const char* rollout_kernel_name(const QPDev& P);     // instantiation rollout_launch picks for this problem
cudaError_t rollout_launch(const QPDev& P, const QPDev& P1, const LoopDev& L, int wpb, int num_sms, int max_smem,
                           const RolloutArgs& a, cudaStream_t stream);

}  // namespace rtmpc

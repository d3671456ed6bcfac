// Persistent closed-loop rollout: one warp owns one instance for all T control steps.
//
// Per control step the warp solves the instance's tracking-MPC QP with the dual active-set method
// (rtmpc_as.cuh, warm-started from its own previous step) and then takes the closed-loop step of
// rtmpc_loop.cuh (consistent actuator, nominal model, ancillary law, plant, estimator; lane 0, with the
// tube-containment statistic spread over the lanes).  Instances never exchange data, so a rollout is a
// single launch: no per-step launch latency and no waiting for the slowest instance of every step.
// An instance whose solve has to be handed to the interior-point kernel parks itself (inst_t, pending,
// ref_pending); the host runs that kernel on the parked instances and relaunches, which resumes them.
#include "rtmpc_as.cuh"

#include "rtmpc_launch.h"
#include "rtmpc_loop.cuh"

namespace rtmpc {

#ifdef RTMPC_AS_DEBUG
extern "C" int rtmpc_debug_counters(unsigned long long* out, int reset) {
    if (cudaMemcpyFromSymbol(out, g_as_dbg, sizeof(g_as_dbg)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_as_dbg, z, sizeof(z)); }
    return 0;
}
#endif

// per-warp shared memory of the rollout, in doubles:
//   closed-loop state | packet payload | x_nom_0 of this step's solve | warm-start record of each QP | solver scratch
__host__ __device__ inline int rollout_fixed_doubles(const QPDev& P0, const QPDev& P1, bool two) {
    return loop_smem_doubles(P0.N, P0.nu) + loop_even((P0.N + 1) * P0.nu) + loop_even(P0.nx) +
           ((((P0.npad + 2) >> 1) + 1) & ~1) + (two ? ((((P1.npad + 2) >> 1) + 1) & ~1) : 0);   // every part even (16-byte alignment)
}
__host__ __device__ inline int rollout_warp_doubles(const QPDev& P0, const QPDev& P1, bool two) {
    const int s0 = as_warp_doubles(P0), s1 = as_warp_doubles(P1);
    return rollout_fixed_doubles(P0, P1, two) + (s0 > s1 ? s0 : s1);
}

// the same sizes inside the kernel, from the instantiation's dimensions (ASDimsFix: compile-time constants)
template <class D, bool TWO>
__device__ __forceinline__ int rollout_fixed_doubles_d(const QPDev& P0, const QPDev& P1) {
    return loop_smem_doubles(D::N(P0), D::nu(P0)) + loop_even((D::N(P0) + 1) * D::nu(P0)) + loop_even(D::nx(P0)) +
           ((((D::npad(P0) + 2) >> 1) + 1) & ~1) + (TWO ? ((((P1.npad + 2) >> 1) + 1) & ~1) : 0);
}
template <class D, bool TWO>
__device__ __forceinline__ int rollout_warp_doubles_d(const QPDev& P0, const QPDev& P1) {
    const int s0 = D::n(P0) * (D::npad(P0) + 2) + 5 * D::npad(P0) + 22;       // (= as_warp_doubles)
    if (!TWO) return rollout_fixed_doubles_d<D, TWO>(P0, P1) + s0;            // (P1 == P0)
    const int s1 = as_warp_doubles(P1);
    return rollout_fixed_doubles_d<D, TWO>(P0, P1) + (s0 > s1 ? s0 : s1);
}

// closed-loop step and tube statistic: straight to the compile-time-sized versions when the instantiation fixes the sizes
template <class D, class State>
__device__ __forceinline__ double rollout_tube_rows(const LoopDev& L, const State S, int lane) {
    if constexpr (D::kFixed) return loop_tube_rows_warp_t<D::kNx>(L, S, lane);
    else return loop_tube_rows_warp(L, S, lane);
}
template <class D, class State>
__device__ __forceinline__ void rollout_step_warp(const LoopDev& L, const State S, int lane, int t, const double* Ub,
                                                  const double* x_nom0_b, const double* ref_b, int theta_in, int gamma_in,
                                                  const double* w_in_b, double p, unsigned long long seed,
                                                  unsigned long long id, double* traj_b, double tube_worst) {
    if constexpr (D::kFixed)
        loop_step_body_warp_t<D::kNx, D::kNu>(L, S, lane, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
    else
        loop_step_warp(L, S, lane, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
}

// P0: the controller's problem; P1: the "packet received" problem of ExtendedTubeTrackingMPC, chosen per step and
// instance on gamma_{t-1} (TubeTrackingMPC.py:307-349); P1 == P0 for the single-problem controllers.
// Hand-over of an instance between warps of different SMs.  Everything the previous owner wrote is read by the next one with
// ld.global.cg (served by L2, the point of coherence, and never allocated in L1), so no stale L1 line can exist and the
// reader needs no acquire: ld.acquire.gpu / fence.*.gpu make ptxas emit CCTL.IVALL, which throws away the SM's whole L1 -
// the shared tables of all 16 warps - at every ticket (twice per ticket with __threadfence on both sides: L1 hit rate
// 85.8 % -> 81.9 % when time slicing came in).  The writer releases (MEMBAR.ALL.GPU + strong store, no invalidation): the
// other lanes' stores are ordered before lane 0's release by the __syncwarp in front of it (cumulativity).
#ifndef RTMPC_RO_FENCE
#define RTMPC_RO_FENCE 0      // 1: the round-1 hand-over (two __threadfence per ticket), kept for A/B measurements
#endif
__device__ __forceinline__ int rollout_ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rollout_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int rollout_take_next(int* next, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(next, 1);
    return __shfl_sync(RTMPC_FULL_MASK, v, 0);
}

#ifndef RTMPC_RO_MAXW5
#define RTMPC_RO_MAXW5 16      // warps per CTA of the cartpole-sized instantiation (development knob)
#endif
#ifdef RTMPC_RO_MAXNREG
#define RTMPC_RO_BOUNDS __maxnreg__(RTMPC_RO_MAXNREG)
#else
#define RTMPC_RO_BOUNDS __launch_bounds__(MAXW * 32, 1)
#endif
// TWO = false: one problem (every controller but ExtendedTubeTrackingMPC).  The solver then reads the problem description
// straight from the kernel's parameter bank (constant operands inside the instructions); with two problems every field
// access is an indexed constant load into a register first - 1.2 % of the executed instructions plus their latency.
// D: ASDimsDyn (any problem) or ASDimsFix (the problem's dimensions as compile-time constants; one problem only).
template <int R2, int MAXW, bool TWO, class D>
__global__ void RTMPC_RO_BOUNDS
rollout_kernel(QPDev P0, QPDev P1, LoopDev L, RolloutArgs a) {
    static_assert(!(D::kFixed && TWO), "fixed dimensions: single-problem controllers only");
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nx = D::nx(P0), nu = D::nu(P0);
    const int usz = (D::N(P0) + 1) * nu;
    const int npad0 = D::npad(P0);
    constexpr bool two = TWO;
    double* wbase = smem + warp * rollout_warp_doubles_d<D, TWO>(P0, P1);
    LoopSmemState S;
    S.base = wbase;
    double* U_s = S.base + loop_smem_doubles(D::N(P0), nu);              // this step's packet payload
    double* x0_s = U_s + loop_even(usz);                                  // x_nom[:,0] of this step's solve
    int* warm0_s = reinterpret_cast<int*>(x0_s + loop_even(nx));          // warm-start records
    int* warm1_s = warm0_s + 2 * ((((npad0 + 2) >> 1) + 1) & ~1);
    double* scratch = wbase + rollout_fixed_doubles_d<D, TWO>(P0, P1);
    const bool ext = L.actuator == RTMPC_ACT_EXTENDED;

    // Time slicing.  An instance is a chain of T dependent steps, so B chains on S warp slots take ceil(B / S) chain
    // lengths when every warp keeps its chain to the end (4096 chains on 2368 slots: 2, with the slots of the second round
    // 73 % used).  Instead a warp runs a chain for `quantum` steps, puts its state back and draws the next ticket
    // (instance, chunk) from a global counter; all slots stay busy until the work runs out (1.73 chain lengths).  A ticket
    // waits for the previous ticket of its instance (`done`), which an earlier draw - a warp that is running - holds.
    // State written by one SM is read by another: these loads go to L2 (__ldcg), the hand-over is fence + flag.
    // Tickets end at multiples of the quantum in ABSOLUTE time, and the carried working-set inverse of as_solve_instance is
    // dropped at multiples of a.refresh (= the configured quantum, also when the chains are not sliced): where an inverse
    // is built from scratch does then not depend on the launch shape, so sliced and whole chains, any cut of the batch and
    // any number of GPUs give the same bits.
    const int Q = a.quantum;
    const int c0 = Q > 0 ? a.t0 / Q : 0;
    const int nchunks = Q > 0 ? (a.T + Q - 1) / Q - c0 : 1;
    const long long ntickets = (long long)a.B * nchunks;
    const int first_wave = gridDim.x * wpb;       // the first tickets are dealt statically: consecutive instances on different SMs
#pragma unroll 1
    for (long long ticket = warp * gridDim.x + blockIdx.x; ticket < ntickets;
         ticket = (long long)first_wave + rollout_take_next(a.next, lane)) {
        const int inst = (int)(ticket % a.B), chunk = (int)(ticket / a.B);
        if (chunk > 0) {
            if (lane == 0) while (rollout_ld_relaxed(a.done + inst) < chunk) __nanosleep(128);
            __syncwarp();
#if RTMPC_RO_FENCE
            __threadfence();
#endif
        }
        int t = __ldcg(a.inst_t + inst);
        const int t_stop = Q > 0 ? min(a.T, (c0 + chunk + 1) * Q) : a.T;
        bool have = __ldcg(a.pending + inst) != 0;       // this step was solved by the interior-point kernel ...
        const bool parked = have && __ldcg(a.status + inst) <= RTMPC_FALLBACK;     // ... or is still waiting for it
        if (t < t_stop && !parked) {
        // per-ticket statistics: status counts packed 16 bits each (flushed before they can overflow), solver counters summed
        // over the ticket's solves (the flop count is linear in them: evaluated once per ticket, as_flops)
        unsigned long long n_status = 0;
        unsigned n_ipm = 0, n_rounds = 0, n_solves0 = 0, n_solves1 = 0;
        ASCounters tot0, tot1;
        tot0.steps = tot0.rounds = tot0.rows = tot0.sq = 0;
        tot1 = tot0;
        int last = -1;            // packed iteration word | status of the last solve, stored when the ticket ends (-1: nothing to store)
        int next_refresh = a.refresh > 0 ? ((t + a.refresh - 1) / a.refresh) * a.refresh : -1;
        double* traj_b = a.traj ? a.traj + (size_t)inst * a.traj_stride : nullptr;
        const double p = a.p_loss ? a.p_loss[inst] : 0.0;
        // ---- state, payload and warm-start records move on chip for the whole rollout -------------------
        if (lane < nx) {
            S.x(lane) = __ldcg(L.x + (size_t)inst * nx + lane);
            S.x_nom(lane) = __ldcg(L.x_nom + (size_t)inst * nx + lane);
            S.x_hat(lane) = __ldcg(L.x_hat + (size_t)inst * nx + lane);
            // x_nom_0 of a step the interior-point kernel solved: that kernel wrote z with its problem's own stride
            const int zs = (two && __ldcg(L.gamma_last + inst) == 1) ? P1.nz : P0.nz;
            x0_s[lane] = (have && ext) ? __ldcg(a.z + (size_t)inst * zs + lane) : 0.0;
        }
        if (lane < nu) S.u_last(lane) = __ldcg(L.u_last + (size_t)inst * nu + lane);
        for (int i = lane; i < usz; i += 32) {
            S.buf()[i] = __ldcg(L.buf + (size_t)inst * usz + i);
            U_s[i] = __ldcg(a.U + (size_t)inst * usz + i);
        }
        for (int i = lane; i < npad0 + 1; i += 32) warm0_s[i] = __ldcg(a.warm + (size_t)inst * (npad0 + 1) + i);
        if (two) for (int i = lane; i < P1.npad + 1; i += 32) warm1_s[i] = __ldcg(a.warm1 + (size_t)inst * (P1.npad + 1) + i);
        if (lane == 0) {
            S.err_acc() = __ldcg(L.err_acc + inst); S.tube_max() = __ldcg(L.tube_max + inst);
            S.q_t() = __ldcg(L.q_t + inst); S.s_t() = __ldcg(L.s_t + inst); S.Theta() = __ldcg(L.Theta + inst);
            S.alive() = __ldcg(L.alive + inst);
            S.last_loss() = __ldcg(L.last_loss + inst); S.gamma_last() = __ldcg(L.gamma_last + inst);
            S.ints()[6] = 0;          // carried working set / inverse of as_solve_instance: none at the start of a ticket
            S.ints()[7] = 0;
        }
        __syncwarp();
#pragma unroll 1
        for (; t < t_stop; ++t) {
            if (!S.alive()) { t = a.T; break; }
            if (((t + 1) & 0x7fff) == 0 && a.stats && lane == 0) {      // (whole chains of more than 65535 steps)
                for (int i = 0; i < 4; ++i) if ((n_status >> (16 * i)) & 0xffffull) atomicAdd(a.stats + i, (n_status >> (16 * i)) & 0xffffull);
                n_status = 0;
            }
            if (t == next_refresh) {
                next_refresh += a.refresh;
                if (lane == 0) S.ints()[6] = 0;
                __syncwarp();
            }
            const int k = t - a.t0;
            const double* ref_t = a.ref ? a.ref + (size_t)k * a.ref_stride_t + (size_t)inst * a.ref_stride_b : nullptr;
            int status;
            if (have) {
                have = false;
                status = __ldcg(a.status + inst);
                const int it = __ldcg(a.iters + inst);
                n_ipm += it & 0xFFF;
                n_rounds += (it >> 24) & 0xF;
                if (lane == 0) a.pending[inst] = 0;
            } else {
                ASCounters cnt;
                cnt.steps = 0; cnt.rounds = 0; cnt.rows = 0; cnt.sq = 0;
                const bool recv = two && S.gamma_last() == 1;           // gamma of the previous step
                const QPDev& P = (two && recv) ? P1 : P0;
                ASWarp w = as_carve<D>(scratch, P);
// rows of W / columns of G' per streaming pass.  One: the smaller loop bodies are worth more to the rollout kernel (instruction
// fetch) than the extra loads in flight of two (116.4 -> 120.4 M solves/s on the benchmark; same FMA order, same bits)
#ifndef RTMPC_RO_ILP
#define RTMPC_RO_ILP 1
#endif
                status = as_solve_instance<R2, RTMPC_RO_ILP, D>(P, w, lane, &S.x_hat(0), ref_t,
                                                                         recv ? warm1_s : warm0_s, ext ? x0_s : nullptr,
                                                                         nx, U_s, cnt, a.carry ? S.ints() + 6 : nullptr,
                                                                         recv ? 2 : 1);
                if (two && recv) {
                    tot1.steps += cnt.steps; tot1.rounds += cnt.rounds; tot1.rows += cnt.rows; tot1.sq += cnt.sq;
                    n_solves1 += 1;
                } else {
                    tot0.steps += cnt.steps; tot0.rounds += cnt.rounds; tot0.rows += cnt.rows; tot0.sq += cnt.sq;
                    n_solves0 += 1;
                }
                if (status == RTMPC_FALLBACK) {
                    if (lane == 0) {
                        a.status[inst] = recv ? RTMPC_FALLBACK - 1 : RTMPC_FALLBACK;     // which problem the hand-over is for
                        a.iters[inst] = as_pack_iters(cnt, w);
                        a.pending[inst] = 1;
                        for (int j = 0; j < nx; ++j) a.ref_pending[(size_t)inst * nx + j] = ref_t ? ref_t[j] : 0.0;
                        atomicAdd(a.n_pending, 1);
                    }
                    last = -1;
                    break;
                }
                last = as_pack_iters(cnt, w) | (status & 0xf);     // (the low 12 bits of the word count interior-point iterations: 0 here)
            }
            if (status >= 0 && status < 4) n_status += 1ull << (16 * status);
            __syncwarp();
            // ---- closed-loop step ------------------------------------------------------------
            int go = 0;
            if (lane == 0) go = loop_step_begin(L, S, t, status, traj_b) ? 1 : 0;
            go = __shfl_sync(RTMPC_FULL_MASK, go, 0);
            if (go) {
                double worst = 0.0;
                if (L.nz_rows > 0) worst = as_wmax(rollout_tube_rows<D>(L, S, lane));
                const bool expl = a.theta != nullptr;
                rollout_step_warp<D>(L, S, lane, t, U_s, ext ? x0_s : nullptr, ref_t,
                                    expl ? a.theta[(size_t)k * a.B + inst] : -1, expl ? a.gamma[(size_t)k * a.B + inst] : -1,
                                    (expl && a.w) ? a.w + ((size_t)k * a.B + inst) * nx : nullptr, p, a.seed,
                                    (unsigned long long)(a.id_offset + inst), traj_b, worst);
            }
            __syncwarp();
        }
        // ---- back to global memory (also when the instance parks for the interior-point kernel) -----------
        if (lane < nx) {
            L.x[(size_t)inst * nx + lane] = S.x(lane);
            L.x_nom[(size_t)inst * nx + lane] = S.x_nom(lane);
            L.x_hat[(size_t)inst * nx + lane] = S.x_hat(lane);
        }
        if (lane < nu) L.u_last[(size_t)inst * nu + lane] = S.u_last(lane);
        for (int i = lane; i < usz; i += 32) {
            L.buf[(size_t)inst * usz + i] = S.buf()[i];
            a.U[(size_t)inst * usz + i] = U_s[i];
        }
        for (int i = lane; i < npad0 + 1; i += 32) a.warm[(size_t)inst * (npad0 + 1) + i] = warm0_s[i];
        if (two) for (int i = lane; i < P1.npad + 1; i += 32) a.warm1[(size_t)inst * (P1.npad + 1) + i] = warm1_s[i];
        if (lane == 0) {
            L.err_acc[inst] = S.err_acc(); L.tube_max[inst] = S.tube_max();
            L.q_t[inst] = S.q_t(); L.s_t[inst] = S.s_t(); L.Theta[inst] = S.Theta(); L.alive[inst] = S.alive();
            L.last_loss[inst] = S.last_loss(); L.gamma_last[inst] = S.gamma_last();
            a.inst_t[inst] = t;
            if (last != -1) { a.status[inst] = last & 0xf; a.iters[inst] = last & ~0xf; }
            if (a.stats) {
                for (int i = 0; i < 4; ++i) if ((n_status >> (16 * i)) & 0xffffull) atomicAdd(a.stats + i, (n_status >> (16 * i)) & 0xffffull);
                if (n_ipm) atomicAdd(a.stats + 4, (unsigned long long)n_ipm);
                const unsigned n_steps = tot0.steps + tot1.steps, n_rounds_as = tot0.rounds + tot1.rounds;
                if (n_steps) atomicAdd(a.stats + 5, (unsigned long long)n_steps);
                if (n_rounds + n_rounds_as) atomicAdd(a.stats + 6, (unsigned long long)(n_rounds + n_rounds_as));
                unsigned long long n_flops = as_flops_total(P0, tot0, n_solves0);
                if (two) n_flops += as_flops_total(P1, tot1, n_solves1);
                if (n_flops) atomicAdd(a.stats + 7, n_flops);
            }
        }
        }   // (ticket had work)
        // hand the instance on: everything this warp wrote is visible before the flag moves
#if RTMPC_RO_FENCE
        __threadfence();
#endif
        __syncwarp();
        if (Q > 0 && lane == 0) rollout_st_release(a.done + inst, chunk + 1);
    }
}

typedef void (*ro_fn)(QPDev, QPDev, LoopDev, RolloutArgs);
struct RoChoice { int r2, maxw; ro_fn fn[3]; };      // fn[1]: two problems; fn[2]: one problem with the cartpole controller's dimensions (or NULL)
// the reference's cartpole controllers (results_linear_system.py / results_nonlinear_system.py): nx = 4, nu = 1, N = 20,
// condensed problem with 21 unknowns (fixed initial state), rows padded to 320
typedef ASDimsFix<21, 24, 4, 1, 20> RoCartpoleDims;
#define RO_ENTRY(R2, W, FIX) {R2, W, {rollout_kernel<R2, W, false, ASDimsDyn>, rollout_kernel<R2, W, true, ASDimsDyn>, FIX}}
static const RoChoice kRo[] = {RO_ENTRY(2, 24, nullptr), RO_ENTRY(5, RTMPC_RO_MAXW5, (rollout_kernel<5, RTMPC_RO_MAXW5, false, RoCartpoleDims>)),
                               // (nine row pairs per lane: 12 warps at 168 registers and 98 KB of shared memory - the 100 KB carve-out -
                               //  beat 16 warps at 128 registers with 168 B of spills and 131 KB: c3 136 -> 154 M solves/s; 10 / 14 warps: 136 / 127)
                               RO_ENTRY(9, 12, nullptr), RO_ENTRY(12, 16, nullptr), RO_ENTRY(16, 16, nullptr)};
// the instantiation a launch uses: index into RoChoice::fn
static int pick_fn(const RoChoice* kc, const QPDev& P, const LoopDev* L, bool two) {
    if (two) return 1;
    if (kc->fn[2] && tuning().rollout_fixed_dims && P.n == 21 && P.npad == 24 && P.nx == 4 && P.nu == 1 && P.N == 20 &&
        (!L || (L->nx == 4 && L->nu == 1 && L->N == 20)))
        return 2;
    return 0;
}
static const RoChoice* pick(int mpad) {
    const int r_need = (mpad + 63) / 64;
    for (const auto& c : kRo)
        if (c.r2 >= r_need) return &c;
    return nullptr;
}

const char* rollout_kernel_name(const QPDev& P) {
    // (the third template argument is true for ExtendedTubeTrackingMPC's pair of problems, false otherwise)
    static const char* names[] = {"rollout_kernel<2,24,*>", "rollout_kernel<5,16,*>", "rollout_kernel<9,12,*>", "rollout_kernel<12,16,*>",
                                  "rollout_kernel<16,16,*>"};
    const RoChoice* kc = pick(P.mpad);
    if (kc && pick_fn(kc, P, nullptr, false) == 2) return "rollout_kernel<5,16,*,cartpole dims>";
    return kc ? names[kc - kRo] : "none";
}

bool rollout_configure(const QPDev& P, int max_smem, cudaError_t* err) {
    const RoChoice* kc = pick(P.mpad);
    if (!kc) { *err = cudaSuccess; return false; }
    for (int k = 0; k < 3; ++k) {
        if (!kc->fn[k]) continue;
        *err = cudaFuncSetAttribute((const void*)kc->fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (*err != cudaSuccess) return false;
    }
    return true;
}

// Time slicing makes warps wait for tickets held by warps of OTHER CTAs, which is only safe when every CTA of the grid
// is resident at the same time.  The sliced launch is therefore a cooperative launch (the runtime refuses it instead of
// letting it hang when the grid cannot be co-resident: SM-limited MPS / green contexts) with the grid clamped to what
// the occupancy calculator says fits; when the device or context cannot give that guarantee the rollout runs with
// whole chains per warp (quantum 0: no cross-CTA waits, same results bit for bit).
cudaError_t rollout_launch(const QPDev& P_in, const QPDev& P1_in, const LoopDev& L, int wpb, int num_sms, int max_smem,
                           const RolloutArgs& a, cudaStream_t stream) {
    QPDev P = P_in, P1 = P1_in;
    if (!tuning().cert_factored) { P.kap = nullptr; P1.kap = nullptr; }      // RTMPC_TUNE_CERT_FACTORED
    const RoChoice* kc = pick(P.mpad);
    if (!kc || P.mpad != 64 * kc->r2 || (a.two && P1.mpad != P.mpad)) return cudaErrorInvalidValue;    // (rows padded by rtmpc_qp_create)
    if (wpb > kc->maxw) wpb = kc->maxw;
    const size_t per_warp = (size_t)rollout_warp_doubles(P, P1, a.two != 0) * sizeof(double);
    while (wpb > 1 && per_warp * wpb > (size_t)max_smem) --wpb;
    int warps = balanced_warps(a.B, num_sms, wpb);
    RolloutArgs b = a;
    const Tuning& tn = tuning();
    b.carry = tn.rollout_carry;
    b.refresh = tn.rollout_carry ? tn.rollout_quantum : 0;
    // more chains than warp slots: slice them (see the kernel) and use every slot
    const int q = tn.rollout_quantum;
    bool sliced = q > 0 && (long long)a.B > (long long)num_sms * wpb && a.T - a.t0 >= 2 * q;
    if (sliced) warps = wpb;
    if (tn.rollout_warps >= 1 && tn.rollout_warps <= wpb) warps = tn.rollout_warps;      // RTMPC_TUNE_ROLLOUT_WARPS
    int blocks = (a.B + warps - 1) / warps;
    if (blocks > num_sms) blocks = num_sms;
    const size_t smem = per_warp * warps;
    const ro_fn fn = kc->fn[pick_fn(kc, P, &L, a.two != 0)];
    if (sliced) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, warps * 32, smem);
        if (e != cudaSuccess) return e;
        const long long fit = (long long)per_sm * num_sms;
        if (coop && fit >= 1) {
            if (blocks > fit) blocks = (int)fit;
            b.quantum = q;
            QPDev p0 = P, p1 = P1;
            LoopDev ld = L;
            void* args[] = {&p0, &p1, &ld, &b};
            e = cudaLaunchCooperativeKernel((const void*)fn, dim3(blocks), dim3(warps * 32), args, smem, stream);
            if (e == cudaSuccess) return cudaGetLastError();
            if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorNotSupported) return e;
            (void)cudaGetLastError();            // co-residency refused: fall through to whole chains
        }
        b.quantum = 0;
        warps = balanced_warps(a.B, num_sms, wpb);
        blocks = (a.B + warps - 1) / warps;
        if (blocks > num_sms) blocks = num_sms;
    }
    fn<<<blocks, warps * 32, per_warp * warps, stream>>>(P, P1, L, b);
    return cudaGetLastError();
}

}  // namespace rtmpc

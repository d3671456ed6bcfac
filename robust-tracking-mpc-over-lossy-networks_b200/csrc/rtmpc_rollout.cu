// Persistent closed-loop rollout: one warp owns one instance for all T control steps.
//
// Per control step the warp solves the instance's tracking-MPC QP with the dual active-set method
// (rtmpc_as.cuh, warm-started from its own previous step) and then takes the closed-loop step of
// rtmpc_loop.cuh (consistent actuator, nominal model, ancillary law, plant, estimator; lane 0, with the
// tube-containment statistic spread over the lanes).  Instances never exchange data, so a rollout is a
// single launch: no per-step launch latency and no waiting for the slowest instance of every step.
// An instance whose solve has to be handed to the interior-point kernel parks itself (inst_t, pending,
// ref_pending); the host runs that kernel on the parked instances and relaunches, which resumes them.
#include <cstdlib>

#include "rtmpc_as.cuh"
#include "rtmpc_launch.h"
#include "rtmpc_loop.cuh"

namespace rtmpc {

// per-warp shared memory of the rollout: solver scratch | closed-loop state | packet payload | warm-start record
__host__ __device__ inline int rollout_warp_doubles(const QPDev& P) {
    return as_warp_doubles(P) + loop_smem_doubles(P.N, P.nu) + (((P.N + 1) * P.nu + 1) & ~1) +
           ((((P.npad + 2) >> 1) + 1) & ~1);     // every part even: 16-byte alignment of the next warp's block
}

template <int R2, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
rollout_kernel(QPDev P, LoopDev L, RolloutArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nx = P.nx, nu = P.nu;
    const int usz = (P.N + 1) * nu;
    double* wbase = smem + (size_t)warp * rollout_warp_doubles(P);
    ASWarp w = as_carve(wbase, P);
    LoopSmemState S;
    S.base = wbase + as_warp_doubles(P);
    double* U_s = S.base + loop_smem_doubles(P.N, nu);                  // this step's packet payload
    int* warm_s = reinterpret_cast<int*>(U_s + ((usz + 1) & ~1));        // warm-start record

#pragma unroll 1
    // consecutive instances go to different SMs, so a last partial round is spread over all of them
    for (int inst = warp * gridDim.x + blockIdx.x; inst < a.B; inst += gridDim.x * wpb) {
        int t = a.inst_t[inst];
        if (t >= a.T) continue;
        bool have = a.pending[inst] != 0;       // this step was solved by the interior-point kernel
        unsigned n_status[4] = {0, 0, 0, 0}, n_ipm = 0, n_steps = 0, n_rounds = 0;
        unsigned long long n_flops = 0;
        double* z_inst = a.z ? a.z + (size_t)inst * P.nz : nullptr;
        double* traj_b = a.traj ? a.traj + (size_t)inst * a.traj_stride : nullptr;
        const double p = a.p_loss ? a.p_loss[inst] : 0.0;
        // ---- state, payload and warm-start record move on chip for the whole rollout -------------------
        if (lane < nx) {
            S.x(lane) = L.x[(size_t)inst * nx + lane];
            S.x_nom(lane) = L.x_nom[(size_t)inst * nx + lane];
            S.x_hat(lane) = L.x_hat[(size_t)inst * nx + lane];
        }
        if (lane < nu) S.u_last(lane) = L.u_last[(size_t)inst * nu + lane];
        for (int i = lane; i < usz; i += 32) {
            S.buf()[i] = L.buf[(size_t)inst * usz + i];
            U_s[i] = a.U[(size_t)inst * usz + i];
        }
        for (int i = lane; i < P.npad + 1; i += 32) warm_s[i] = a.warm[(size_t)inst * (P.npad + 1) + i];
        if (lane == 0) {
            S.err_acc() = L.err_acc[inst]; S.tube_max() = L.tube_max[inst];
            S.q_t() = L.q_t[inst]; S.s_t() = L.s_t[inst]; S.Theta() = L.Theta[inst]; S.alive() = L.alive[inst];
            S.last_loss() = L.last_loss[inst]; S.gamma_last() = L.gamma_last[inst];
        }
        __syncwarp();
#pragma unroll 1
        for (; t < a.T; ++t) {
            if (!S.alive()) { t = a.T; break; }
            const int k = t - a.t0;
            const double* ref_t = a.ref ? a.ref + (size_t)k * a.ref_stride_t + (size_t)inst * a.ref_stride_b : nullptr;
            int status;
            if (have) {
                have = false;
                status = a.status[inst];
                const int it = a.iters[inst];
                n_ipm += it & 0xFFF;
                n_rounds += (it >> 24) & 0xFF;
                if (lane == 0) a.pending[inst] = 0;
            } else {
                ASCounters cnt;
                cnt.steps = 0; cnt.rounds = 0; cnt.rows = 0; cnt.sq = 0;
                status = as_solve_instance<R2, (MAXW * R2 <= 100) ? 2 : 1>(P, w, lane, &S.x_hat(0), ref_t, warm_s, z_inst,
                                                                         U_s, cnt);
                n_steps += cnt.steps; n_rounds += cnt.rounds; n_flops += as_flops(P, cnt, z_inst != nullptr);
                if (status == RTMPC_FALLBACK) {
                    if (lane == 0) {
                        a.status[inst] = RTMPC_FALLBACK;
                        a.pending[inst] = 1;
                        for (int j = 0; j < nx; ++j) a.ref_pending[(size_t)inst * nx + j] = ref_t ? ref_t[j] : 0.0;
                        atomicAdd(a.n_pending, 1);
                    }
                    break;
                }
                if (lane == 0) { a.status[inst] = status; a.iters[inst] = as_pack_iters(cnt); }
            }
            if (status >= 0 && status < 4) n_status[status] += 1;
            __syncwarp();
            // ---- closed-loop step ------------------------------------------------------------
            int go = 0;
            if (lane == 0) go = loop_step_begin(L, S, t, status, traj_b) ? 1 : 0;
            go = __shfl_sync(RTMPC_FULL_MASK, go, 0);
            if (go) {
                double worst = 0.0;
                if (L.nz_rows > 0) worst = as_wmax(loop_tube_rows(L, S, lane, 32));
                if (lane == 0) {
                    const bool expl = a.theta != nullptr;
                    loop_step_body(L, S, t, U_s, (L.actuator == RTMPC_ACT_EXTENDED) ? z_inst : nullptr, ref_t,
                                   expl ? a.theta[(size_t)k * a.B + inst] : -1, expl ? a.gamma[(size_t)k * a.B + inst] : -1,
                                   (expl && a.w) ? a.w + ((size_t)k * a.B + inst) * nx : nullptr, p, a.seed,
                                   (unsigned long long)(a.id_offset + inst), traj_b, worst);
                }
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
        for (int i = lane; i < P.npad + 1; i += 32) a.warm[(size_t)inst * (P.npad + 1) + i] = warm_s[i];
        if (lane == 0) {
            L.err_acc[inst] = S.err_acc(); L.tube_max[inst] = S.tube_max();
            L.q_t[inst] = S.q_t(); L.s_t[inst] = S.s_t(); L.Theta[inst] = S.Theta(); L.alive[inst] = S.alive();
            L.last_loss[inst] = S.last_loss(); L.gamma_last[inst] = S.gamma_last();
            a.inst_t[inst] = t;
            if (a.stats) {
                for (int i = 0; i < 4; ++i) if (n_status[i]) atomicAdd(a.stats + i, (unsigned long long)n_status[i]);
                if (n_ipm) atomicAdd(a.stats + 4, (unsigned long long)n_ipm);
                if (n_steps) atomicAdd(a.stats + 5, (unsigned long long)n_steps);
                if (n_rounds) atomicAdd(a.stats + 6, (unsigned long long)n_rounds);
                if (n_flops) atomicAdd(a.stats + 7, n_flops);
            }
        }
        __syncwarp();
    }
}

typedef void (*ro_fn)(QPDev, LoopDev, RolloutArgs);
struct RoChoice { int r2, maxw; ro_fn fn; };
static const RoChoice kRo[] = {
    {2, 24, rollout_kernel<2, 24>},  {5, 16, rollout_kernel<5, 16>},  {9, 16, rollout_kernel<9, 16>},
    {12, 16, rollout_kernel<12, 16>}, {16, 16, rollout_kernel<16, 16>},
};
static const RoChoice kRoExp[] = {{5, 20, rollout_kernel<5, 20>}, {5, 28, rollout_kernel<5, 28>}};
static const RoChoice* pick(int mpad) {
    const int r_need = (mpad + 63) / 64;
    const char* v = getenv("RTMPC_RO_MAXW");       // experiment knob
    const int want = v ? atoi(v) : 0;
    if (want && r_need == 5)
        for (const auto& c : kRoExp)
            if (c.maxw == want) return &c;
    for (const auto& c : kRo)
        if (c.r2 >= r_need) return &c;
    return nullptr;
}

bool rollout_configure(const QPDev& P, int max_smem, cudaError_t* err) {
    const RoChoice* kc = pick(P.mpad);
    if (!kc) { *err = cudaSuccess; return false; }
    *err = cudaFuncSetAttribute((const void*)kc->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    return *err == cudaSuccess;
}

cudaError_t rollout_launch(const QPDev& P, const LoopDev& L, int wpb, int num_sms, const RolloutArgs& a,
                           cudaStream_t stream) {
    const RoChoice* kc = pick(P.mpad);
    if (wpb > kc->maxw) wpb = kc->maxw;
    int per_cta = (a.B + num_sms - 1) / num_sms;
    int warps = per_cta < wpb ? per_cta : wpb;
    if (warps < 1) warps = 1;
    int blocks = (a.B + warps - 1) / warps;
    if (blocks > num_sms) blocks = num_sms;
    const size_t per_warp = (size_t)rollout_warp_doubles(P) * sizeof(double);
    kc->fn<<<blocks, warps * 32, per_warp * warps, stream>>>(P, L, a);
    return cudaGetLastError();
}

}  // namespace rtmpc

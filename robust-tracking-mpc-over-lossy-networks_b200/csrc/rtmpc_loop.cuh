// Fused closed-loop step: consistent actuator (buffer + consistency flag), nominal model,
// ancillary law, plant step and remote estimator for every instance, one thread per instance.
//
// Reference semantics (one Python object per instance there):
//   SmartActuator.update_Theta_t / update_s_t / update_local_information / compute_u_t   SmartActuator.py:57-107
//   ConsistentActuator.process_packet / ancillary_controller / update_x_nom             SmartActuator.py:146-222
//   Estimator.update_estimate / update_q_t                                               Estimator.py:43-92
//   RobustEstimator.update_estimate                                                      Estimator.py:113-156
//   plant step x+ = A x + B u + w                                                        Results/results_linear_system.py:248
// O(1) equivalents of the reference's O(t) bookkeeping (checked against the literal restatement
// in oracle/ref_loop.py by tests/test_oracle_loop.py and, on the reference's own sequences, tests/test_reference_pin.py):
//   * Theta_t = theta_t * prod(theta[q_t+1 .. t])  ==  theta_t && (last lost step <= q_t)
//   * the estimator's `controlSequences[s_t]` is by construction the actuator's current buffer.
#pragma once
#include "rtmpc_common.cuh"
#include "../../include/rtmpc.h"

namespace rtmpc {

constexpr int LOOP_MAX_NX = 8;
constexpr int LOOP_MAX_NU = 4;

struct LoopDev {
    int nx, nu, N, actuator, plant, nz_rows;
    int tube_sym;      // 1: the tube's facets come in pairs with opposite normals; Hz holds one normal per pair and hz the two
                       //    bounds (h+, h-) of the pair: value max(Hz d - h+, -Hz d - h-), one dot product for both facets
    const double *A, *Bm, *K, *Kp, *Hz, *hz, *w_half;
    const double* tube_cut;   // [2 * (nz_rows / 32 + 1)] per block of 32 tube rows: smallest facet distance and smallest |normal| from that block on
    double cart[8];
    double *x, *x_nom, *x_hat, *buf, *u_last, *err_acc, *tube_max;
    int *q_t, *s_t, *Theta, *alive, *last_loss, *gamma_last;
};

// Where the per-instance state lives: the step-by-step kernels keep it in global memory (LoopDev's arrays,
// instance b); the rollout kernel keeps it in a per-warp block of shared memory for all T steps.
struct LoopGlobalState {
    const LoopDev* L;
    int b;
    __device__ __forceinline__ double& x(int k) const { return L->x[(size_t)b * L->nx + k]; }
    __device__ __forceinline__ double& x_nom(int k) const { return L->x_nom[(size_t)b * L->nx + k]; }
    __device__ __forceinline__ double& x_hat(int k) const { return L->x_hat[(size_t)b * L->nx + k]; }
    __device__ __forceinline__ double& u_last(int j) const { return L->u_last[(size_t)b * L->nu + j]; }
    __device__ __forceinline__ double* buf() const { return L->buf + (size_t)b * (L->N + 1) * L->nu; }
    __device__ __forceinline__ double& err_acc() const { return L->err_acc[b]; }
    __device__ __forceinline__ double& tube_max() const { return L->tube_max[b]; }
    __device__ __forceinline__ int& q_t() const { return L->q_t[b]; }
    __device__ __forceinline__ int& s_t() const { return L->s_t[b]; }
    __device__ __forceinline__ int& Theta() const { return L->Theta[b]; }
    __device__ __forceinline__ int& alive() const { return L->alive[b]; }
    __device__ __forceinline__ int& last_loss() const { return L->last_loss[b]; }
    __device__ __forceinline__ int& gamma_last() const { return L->gamma_last[b]; }
};
// shared-memory block: x[8] x_nom[8] x_hat[8] u_last[4] err_acc tube_max | 6 ints + 2 ints of the rollout's carried working set (tag, slot mask) | buf[(N+1) nu]
// (fixed offsets on purpose: offsets that depend on nx cost registers the rollout kernel does not have)
__host__ __device__ inline int loop_even(int v) { return (v + 1) & ~1; }
constexpr int LOOP_SMEM_FIXED = 3 * LOOP_MAX_NX + LOOP_MAX_NU + 2 + 4;
__host__ __device__ inline int loop_smem_doubles(int N, int nu) { return LOOP_SMEM_FIXED + loop_even((N + 1) * nu); }
struct LoopSmemState {
    double* base;
    __device__ __forceinline__ double& x(int k) const { return base[k]; }
    __device__ __forceinline__ double& x_nom(int k) const { return base[LOOP_MAX_NX + k]; }
    __device__ __forceinline__ double& x_hat(int k) const { return base[2 * LOOP_MAX_NX + k]; }
    __device__ __forceinline__ double& u_last(int j) const { return base[3 * LOOP_MAX_NX + j]; }
    __device__ __forceinline__ double& err_acc() const { return base[3 * LOOP_MAX_NX + LOOP_MAX_NU]; }
    __device__ __forceinline__ double& tube_max() const { return base[3 * LOOP_MAX_NX + LOOP_MAX_NU + 1]; }
    __device__ __forceinline__ int* ints() const { return reinterpret_cast<int*>(base + 3 * LOOP_MAX_NX + LOOP_MAX_NU + 2); }
    __device__ __forceinline__ int& q_t() const { return ints()[0]; }
    __device__ __forceinline__ int& s_t() const { return ints()[1]; }
    __device__ __forceinline__ int& Theta() const { return ints()[2]; }
    __device__ __forceinline__ int& alive() const { return ints()[3]; }
    __device__ __forceinline__ int& last_loss() const { return ints()[4]; }
    __device__ __forceinline__ int& gamma_last() const { return ints()[5]; }
    __device__ __forceinline__ double* buf() const { return base + LOOP_SMEM_FIXED; }
};

// Analytic cartpole (replaces the reference's PyBullet plant, Results/Cartpole/cartpole.py:32-41): cart mass M on a
// prismatic joint, pole of mass m with its centre of mass l from the pivot and inertia I about it, semi-implicit Euler
// at dt (Bullet's multibody integrator), `nsub` physics steps per control period with the input held.
// c[7] = k > 0 adds Bullet's default LINK DAMPING (btMultiBody: force -m v (k + k |v|) on every link's linear
// velocity, torque -I w (k + k |w|) on its angular velocity; pybullet default linearDamping = angularDamping = 0.04).
// With k = 0.04 and I = the inertia Bullet recomputes from the pole's collision box (loadURDF without
// URDF_USE_INERTIA_FROM_FILE: m (0.05^2 + 1.0^2) / 12, cartpole.urdf:61-72) the model-error quantiles of
// Results/estimate_W_for_Cartpole.py:79-127 land on the constants the reference hard-codes
// (Results/results_linear_system.py:76-91) to within 0.5 % - see DESIGN.md section 7.
static __device__ __noinline__ void cartpole_substeps(double* x, double F, const double* c);
// The plant's parameters are read out of the kernel's parameter block element by element: taking the ADDRESS of a member of
// a by-value kernel parameter (L.cart) makes the compiler keep a copy of the whole struct in local memory, and every later
// L.field access of the kernel then reads that copy (12 local loads per control step of the rollout kernel).
template <class LoopDevT>
__device__ __forceinline__ void cartpole_substeps_of(const LoopDevT& L, double* x, double F) {
    double cc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cc[k] = L.cart[k];
    cartpole_substeps(x, F, cc);
}
static __device__ __noinline__ void cartpole_substeps(double* x, double F, const double* c) {
    const double M = c[0], m = c[1], I = c[2], g = c[3], l = c[4], dt = c[5], kd = c[7];
    const int nsub = (int)c[6];
    double pos = x[0], vel = x[1], phi = x[2], om = x[3];
    for (int s = 0; s < nsub; ++s) {
        double sn, cs;
        sincos(phi, &sn, &cs);
        double fe = F + m * l * om * om * sn;          // generalised force on the cart coordinate
        double ge = m * g * l * sn;                    // ... on the pole angle
        if (kd > 0.0) {
            const double vpx = vel + l * om * cs, vpz = -l * om * sn;          // velocity of the pole's centre of mass
            const double kp = kd + kd * sqrt(vpx * vpx + vpz * vpz);
            const double fpx = -m * vpx * kp, fpz = -m * vpz * kp;
            fe += -M * vel * (kd + kd * fabs(vel)) + fpx;
            ge += -I * om * (kd + kd * fabs(om)) + fpx * l * cs - fpz * l * sn;
        }
        const double D = (M + m) * (I + m * l * l) - (m * l * cs) * (m * l * cs);
        const double iD = 1.0 / D;                      // (one division per sub-step; the oracle's two differ in the last bit)
        const double acc = ((I + m * l * l) * fe - m * l * cs * ge) * iD;
        const double alp = ((M + m) * ge - m * l * cs * fe) * iD;
        vel += dt * acc;
        om += dt * alp;
        pos += dt * vel;
        phi += dt * om;
    }
    x[0] = pos; x[1] = vel; x[2] = phi; x[3] = om;
}

// tube containment statistic  max_i (Hz (x - x_nom) - hz)_i  of instance b, rows i = first, first+stride, ...
template <int NX, class State>
__device__ __forceinline__ double loop_tube_rows_t(const LoopDev& L, const State S, int first, int stride) {
    const int nx = NX ? NX : L.nx;
    constexpr int AX = NX ? NX : LOOP_MAX_NX;
    double d[AX];
#pragma unroll
    for (int k = 0; k < AX; ++k) d[k] = (k < nx) ? S.x(k) - S.x_nom(k) : 0.0;
    double worst = -1e300;
    const bool sym = L.tube_sym != 0;
#ifndef RTMPC_TUBE_UNROLL
#define RTMPC_TUBE_UNROLL 1     // rolled: the rollout kernel is bound by instruction fetch (4: 84.6, 2: 86.7, 1: 87.2 M solves/s)
#endif
    constexpr int kUnroll = RTMPC_TUBE_UNROLL;
#pragma unroll kUnroll
    for (int i = first; i < L.nz_rows; i += stride) {
        double acc = sym ? 0.0 : -__ldg(L.hz + i);
        if (NX > 0 && (NX & 1) == 0) {
            const double2* __restrict__ h2 = reinterpret_cast<const double2*>(L.Hz + i * NX);      // rows are 16-byte aligned
#pragma unroll
            for (int k = 0; k < AX / 2; ++k) {
                const double2 hh = __ldg(h2 + k);
                acc = fma(hh.x, d[2 * k], acc);
                acc = fma(hh.y, d[2 * k + 1], acc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < AX; ++k) if (k < nx) acc = fma(__ldg(L.Hz + i * nx + k), d[k], acc);
        }
        if (sym) {                                          // both facets of the pair at once
            const double2 hh = __ldg(reinterpret_cast<const double2*>(L.hz) + i);
            const double a1 = acc - hh.x, a2 = -acc - hh.y;
            acc = (a1 > a2) ? a1 : a2;
        }
        if (acc > worst) worst = acc;          // (a compare and a select: FP64 fmax is a seven-instruction sequence)
    }
    return worst;
}
// The same statistic for one instance per warp (row 32 j + lane in trip j; the rows are sorted by the facet's distance
// from the origin and padded to whole blocks, rtmpc_loop_create).  Every lane returns the maximum over its own rows; the
// scan stops as soon as one lane holds a value no remaining row can reach (value of a row <= |a| (|d| - h / |a|)), so
// the maximum over the lanes is the maximum over all rows, bit for bit.
template <int NX, class State>
__device__ __forceinline__ double loop_tube_rows_warp_t(const LoopDev& L, const State S, int lane) {
    const int nx = NX ? NX : L.nx;
    constexpr int AX = NX ? NX : LOOP_MAX_NX;
    double d[AX];
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < AX; ++k) { d[k] = (k < nx) ? S.x(k) - S.x_nom(k) : 0.0; s2 = fma(d[k], d[k], s2); }
    // |d| rounded up (single-precision root with a margin far above its error)
    const double nd = (double)sqrtf((float)s2) * (1.0 + 1e-6) + 1e-30;
    double worst = -1e300;
    const bool sym = L.tube_sym != 0;
    const double2* __restrict__ cut = reinterpret_cast<const double2*>(L.tube_cut);
#pragma unroll 1
    for (int i = lane, j = 1; i < L.nz_rows; i += 32, ++j) {
        double acc = sym ? 0.0 : -__ldg(L.hz + i);
        if (NX > 0 && (NX & 1) == 0) {
            const double2* __restrict__ h2 = reinterpret_cast<const double2*>(L.Hz + i * NX);
#pragma unroll
            for (int k = 0; k < AX / 2; ++k) {
                const double2 hh = __ldg(h2 + k);
                acc = fma(hh.x, d[2 * k], acc);
                acc = fma(hh.y, d[2 * k + 1], acc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < AX; ++k) if (k < nx) acc = fma(__ldg(L.Hz + i * nx + k), d[k], acc);
        }
        if (sym) {
            const double2 hh = __ldg(reinterpret_cast<const double2*>(L.hz) + i);
            const double a1 = acc - hh.x, a2 = -acc - hh.y;
            acc = (a1 > a2) ? a1 : a2;
        }
        if (acc > worst) worst = acc;
        // rows from block j on: at most c.y * (|d| - c.x) when that is negative (c.x their smallest distance, c.y their
        // smallest |a|); the margins are orders of magnitude above the rounding of a row value
        const double2 c = __ldg(cut + j);
        const double g = nd - c.x;
        if (__any_sync(0xffffffffu, g < 0.0 && worst > fma(c.y * g, 1.0 - 1e-9, 1e-12))) break;
    }
    return worst;
}
template <class State>
__device__ __forceinline__ double loop_tube_rows_warp(const LoopDev& L, const State S, int lane) {
    if (L.nx == 4) return loop_tube_rows_warp_t<4>(L, S, lane);
    if (L.nx == 2) return loop_tube_rows_warp_t<2>(L, S, lane);
    return loop_tube_rows_warp_t<0>(L, S, lane);
}
template <class State>
__device__ __forceinline__ double loop_tube_rows(const LoopDev& L, const State S, int first, int stride) {
    if (L.nx == 4) return loop_tube_rows_t<4>(L, S, first, stride);
    if (L.nx == 2) return loop_tube_rows_t<2>(L, S, first, stride);
    return loop_tube_rows_t<0>(L, S, first, stride);
}

// Start of a control step for instance b: records x_0, retires the instance when the controller
// returned None.  Returns false when the instance takes no step.
template <class State>
__device__ __forceinline__ bool loop_step_begin(const LoopDev& L, const State S, int t, int status_b, double* traj_b) {
    if (!S.alive()) return false;
    const int nx = L.nx;
    if (traj_b && t == 0) for (int k = 0; k < nx; ++k) traj_b[k] = S.x(k);
    // controller returned None (infeasible): the reference stops this controller's run
    if (status_b == RTMPC_INFEASIBLE) { S.alive() = 0; return false; }
    return true;
}

// One closed-loop step of instance b (one thread): statistics, network / disturbance realisation,
// local side, plant, remote side.  Ub: this step's packet payload [(N+1)*nu]; x_nom0_b: x_nom[:,0]
// of this step's solve or NULL; theta_in < 0 selects the device RNG.
// (no __restrict__: inside the rollout kernel these buffers are written by the same warp)
// NX, NU > 0: sizes known at compile time (state in registers, loops unrolled); 0: taken from L.
#ifdef RTMPC_AS_CALLS
static __device__ __noinline__
#else
__device__ __forceinline__
#endif
Philox4 loop_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32_10(c0, c1, c2, c3, k0, k1);
}

template <int NX, int NU, class State>
static __device__ __noinline__ void loop_step_body_t(const LoopDev& L, const State S, int t, const double* Ub,
                                                     const double* x_nom0_b, const double* ref_b, int theta_in,
                                                     int gamma_in, const double* w_in_b, double p,
                                                     unsigned long long seed, unsigned long long id, double* traj_b,
                                                     double tube_worst) {
    const int nx = NX ? NX : L.nx, nu = NU ? NU : L.nu, N = L.N;
    constexpr int AX = NX ? NX : LOOP_MAX_NX, AU = NU ? NU : LOOP_MAX_NU;
    double x[AX], xn[AX], xh[AX], w[AX];
    for (int k = 0; k < nx; ++k) {
        x[k] = S.x(k);
        xn[k] = S.x_nom(k);
        xh[k] = S.x_hat(k);
    }
    // statistics on the pre-step state (x_traj[:, t] in the reference's scripts)
    if (ref_b) {
        double e = 0.0;
        for (int k = 0; k < nx; ++k) { double d = x[k] - ref_b[k]; e = fma(d, d, e); }
        S.err_acc() += e;
    }
    if (L.nz_rows > 0) S.tube_max() = fmax(S.tube_max(), tube_worst);

    // network and disturbance realisation
    int theta, gamma;
    if (theta_in >= 0) {
        theta = theta_in;
        gamma = gamma_in;
        for (int k = 0; k < nx; ++k) w[k] = w_in_b ? w_in_b[k] : 0.0;
    } else {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        Philox4 r = loop_philox((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, 0u, k0, k1);
        theta = (t == 0) ? 1 : (u01_from_bits(r.x, r.y) < p ? 0 : 1);
        gamma = (t == 0) ? 1 : (u01_from_bits(r.z, r.w) < p ? 0 : 1);
        for (int k = 0; k < nx; k += 2) {
            Philox4 q = loop_philox((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, 1u + (uint32_t)(k >> 1), k0, k1);
            w[k] = L.w_half[k] * (2.0 * u01_from_bits(q.x, q.y) - 1.0);
            if (k + 1 < nx) w[k + 1] = L.w_half[k + 1] * (2.0 * u01_from_bits(q.z, q.w) - 1.0);
        }
    }

    // ---- local side ---------------------------------------------------------------------
    const int q_pkt = S.q_t();                 // q_t carried by the controller packet
    int last_loss = S.last_loss();
    int Theta = 0;
    if (theta == 1) Theta = (last_loss <= q_pkt) ? 1 : 0;
    else last_loss = t;
    int s_t = S.s_t();
    double* buf = S.buf();
    if (Theta) {
        s_t = t;
        for (int i = 0; i < (N + 1) * nu; ++i) buf[i] = Ub[i];
        if (L.actuator != RTMPC_ACT_SMART && x_nom0_b)    // packet carries x_nom_0 (SmartActuator.py:183-187,219-222)
            for (int k = 0; k < nx; ++k) xn[k] = x_nom0_b[k];
    }
    const int kk = t - s_t;
    const double* xfb = (L.actuator == RTMPC_ACT_SMART) ? x : xn;   // state fed to compute_u_t
    double u_nom[AU], u[AU];
    for (int j = 0; j < nu; ++j) {
        if (kk < N) u_nom[j] = buf[kk * nu + j];
        else {
            double acc = buf[N * nu + j];
            for (int k = 0; k < nx; ++k) acc = fma(-L.K[j * nx + k], xfb[k], acc);
            u_nom[j] = acc;
        }
        if (L.actuator == RTMPC_ACT_SMART) u[j] = u_nom[j];
        else {
            double acc = u_nom[j];
            for (int k = 0; k < nx; ++k) acc = fma(-L.Kp[j * nx + k], x[k] - xn[k], acc);
            u[j] = acc;
        }
    }
    // plant packet content (pre-update values)
    double xp[AX], xnp[AX];
    for (int k = 0; k < nx; ++k) {
        xnp[k] = xn[k];
        xp[k] = (L.actuator == RTMPC_ACT_CONSISTENT) ? xn[k] : x[k];
    }
    // nominal model
    if (L.actuator != RTMPC_ACT_SMART) {
        double nxt[AX];
        for (int i = 0; i < nx; ++i) {
            double acc = 0.0;
            for (int k = 0; k < nx; ++k) acc = fma(L.A[i * nx + k], xnp[k], acc);
            for (int j = 0; j < nu; ++j) acc = fma(L.Bm[i * nu + j], u_nom[j], acc);
            nxt[i] = acc;
        }
        for (int i = 0; i < nx; ++i) xn[i] = nxt[i];
    }
    // ---- plant ----------------------------------------------------------------------------
    double xnew[AX];
    if (L.plant == RTMPC_PLANT_CARTPOLE) {
        for (int k = 0; k < nx; ++k) xnew[k] = x[k];
        cartpole_substeps_of(L, xnew, u[0]);
    } else {
        for (int i = 0; i < nx; ++i) {
            double acc = 0.0;
            for (int k = 0; k < nx; ++k) acc = fma(L.A[i * nx + k], x[k], acc);
            for (int j = 0; j < nu; ++j) acc = fma(L.Bm[i * nu + j], u[j], acc);
            xnew[i] = acc + w[i];
        }
    }
    // ---- remote side ----------------------------------------------------------------------
    double uh[AU];
    const double* xbase;
    double xn0[AX];
    if (gamma == 1) {
        // \hat u(k|k) from the sequence the plant is using (== buf) and the packet's state
        for (int j = 0; j < nu; ++j) {
            double un;
            if (kk < N) un = buf[kk * nu + j];
            else {
                double acc = buf[N * nu + j];
                const double* xs = (L.actuator == RTMPC_ACT_EXTENDED) ? xnp : xp;
                for (int k = 0; k < nx; ++k) acc = fma(-L.K[j * nx + k], xs[k], acc);
                un = acc;
            }
            if (L.actuator == RTMPC_ACT_EXTENDED) {
                double acc = un;
                for (int k = 0; k < nx; ++k) acc = fma(-L.Kp[j * nx + k], xp[k] - xnp[k], acc);
                un = acc;
            }
            uh[j] = un;
        }
        xbase = xp;
    } else {
        for (int j = 0; j < nu; ++j) uh[j] = Ub[j];           // first input of the latest sent sequence
        if (L.actuator == RTMPC_ACT_EXTENDED && x_nom0_b) {
            for (int k = 0; k < nx; ++k) xn0[k] = x_nom0_b[k];
            xbase = xn0;
        } else xbase = xh;
    }
    double xhn[AX];
    for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int k = 0; k < nx; ++k) acc = fma(L.A[i * nx + k], xbase[k], acc);
        for (int j = 0; j < nu; ++j) acc = fma(L.Bm[i * nu + j], uh[j], acc);
        xhn[i] = acc;
    }
    // ---- write back -----------------------------------------------------------------------
    for (int k = 0; k < nx; ++k) {
        S.x(k) = xnew[k];
        S.x_nom(k) = xn[k];
        S.x_hat(k) = xhn[k];
    }
    for (int j = 0; j < nu; ++j) S.u_last(j) = u[j];
    if (gamma == 1) S.q_t() = t;
    S.s_t() = s_t;
    S.Theta() = Theta;
    S.last_loss() = last_loss;
    S.gamma_last() = gamma;
    if (traj_b) for (int k = 0; k < nx; ++k) traj_b[(size_t)(t + 1) * nx + k] = xnew[k];
}

// sizes known at compile time for the reference's two systems (double integrator, cartpole), generic otherwise
template <class State>
__device__ __forceinline__ void loop_step_body(const LoopDev& L, const State S, int t, const double* Ub, const double* x_nom0_b,
                                               const double* ref_b, int theta_in, int gamma_in, const double* w_in_b,
                                               double p, unsigned long long seed, unsigned long long id, double* traj_b,
                                               double tube_worst) {
    if (L.nx == 4 && L.nu == 1)
        loop_step_body_t<4, 1>(L, S, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
    else if (L.nx == 2 && L.nu == 1)
        loop_step_body_t<2, 1>(L, S, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
    else
        loop_step_body_t<0, 0>(L, S, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
}

// The same closed-loop step taken by a whole warp (rollout kernel; state in shared memory): lane i owns row i of
// the three mat-vecs, the Philox draws run in parallel lanes, scalars are computed redundantly.  Every output
// element is produced by the same FMA sequence as in loop_step_body_t, so both give identical bits.
template <class State>
__device__ __forceinline__ void loop_step_body_warp(const LoopDev& L, const State S, int lane, int t, const double* Ub,
                                                    const double* x_nom0_b, const double* ref_b, int theta_in,
                                                    int gamma_in, const double* w_in_b, double p,
                                                    unsigned long long seed, unsigned long long id, double* traj_b,
                                                    double tube_worst) {
    const int nx = L.nx, nu = L.nu, N = L.N;
    // ---- statistics on the pre-step state ---------------------------------------------------------------
    if (lane == 0) {
        if (ref_b) {
            double e = 0.0;
#pragma unroll 1
            for (int k = 0; k < nx; ++k) { const double d = S.x(k) - ref_b[k]; e = fma(d, d, e); }
            S.err_acc() += e;
        }
        if (L.nz_rows > 0) S.tube_max() = fmax(S.tube_max(), tube_worst);
    }
    // ---- network and disturbance realisation ------------------------------------------------------------
    int theta, gamma;
    double wi = 0.0;                                     // lane i < nx: disturbance on state i
    if (theta_in >= 0) {
        theta = theta_in;
        gamma = gamma_in;
        if (lane < nx && w_in_b) wi = w_in_b[lane];
    } else {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        // lane 0 draws (theta, gamma), lane 1 + j draws (w_2j, w_2j+1): counters as in loop_step_body_t
        Philox4 r;
        r.x = r.y = r.z = r.w = 0u;
        if (lane <= ((nx + 1) >> 1)) r = loop_philox((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (uint32_t)lane, k0, k1);
        const int th0 = (t == 0) ? 1 : (u01_from_bits(r.x, r.y) < p ? 0 : 1);
        const int ga0 = (t == 0) ? 1 : (u01_from_bits(r.z, r.w) < p ? 0 : 1);
        theta = __shfl_sync(0xffffffffu, th0, 0);
        gamma = __shfl_sync(0xffffffffu, ga0, 0);
        double wa = 0.0, wb = 0.0;
        if (lane >= 1 && lane <= ((nx + 1) >> 1)) {
            const int k = 2 * (lane - 1);
            wa = __ldg(L.w_half + k) * (2.0 * u01_from_bits(r.x, r.y) - 1.0);
            if (k + 1 < nx) wb = __ldg(L.w_half + k + 1) * (2.0 * u01_from_bits(r.z, r.w) - 1.0);
        }
        const int src = 1 + ((lane < nx ? lane : 0) >> 1);
        const double ga = __shfl_sync(0xffffffffu, wa, src), gb = __shfl_sync(0xffffffffu, wb, src);
        if (lane < nx) wi = (lane & 1) ? gb : ga;
    }
    // ---- local side (scalars redundantly in every lane) ---------------------------------------------------
    const int q_pkt = S.q_t();
    int last_loss = S.last_loss();
    int Theta = 0;
    if (theta == 1) Theta = (last_loss <= q_pkt) ? 1 : 0;
    else last_loss = t;
    int s_t = S.s_t();
    double* buf = S.buf();
    __syncwarp();
    if (Theta) {
        s_t = t;
        for (int i = lane; i < (N + 1) * nu; i += 32) buf[i] = Ub[i];
        if (L.actuator != RTMPC_ACT_SMART && x_nom0_b && lane < nx) S.x_nom(lane) = x_nom0_b[lane];
    }
    __syncwarp();
    const int kk = t - s_t;
    const bool smart = L.actuator == RTMPC_ACT_SMART;
    double u_nom[LOOP_MAX_NU], u[LOOP_MAX_NU], uh[LOOP_MAX_NU];
#pragma unroll
    for (int j = 0; j < LOOP_MAX_NU; ++j) {
        u_nom[j] = 0.0; u[j] = 0.0; uh[j] = 0.0;
        if (j < nu) {
            if (kk < N) u_nom[j] = buf[kk * nu + j];
            else {
                double acc = buf[N * nu + j];
#pragma unroll 1
                for (int k = 0; k < nx; ++k) acc = fma(-__ldg(L.K + j * nx + k), smart ? S.x(k) : S.x_nom(k), acc);
                u_nom[j] = acc;
            }
            if (smart) u[j] = u_nom[j];
            else {
                double acc = u_nom[j];
#pragma unroll 1
                for (int k = 0; k < nx; ++k) acc = fma(-__ldg(L.Kp + j * nx + k), S.x(k) - S.x_nom(k), acc);
                u[j] = acc;
            }
        }
    }
    // remote side's input estimate (plant packet content = pre-update values: x_t field carries x_nom for the
    // consistent actuator, x otherwise; the extended packet carries both)
    const bool cons = L.actuator == RTMPC_ACT_CONSISTENT, ext = L.actuator == RTMPC_ACT_EXTENDED;
    if (gamma == 1) {
#pragma unroll
        for (int j = 0; j < LOOP_MAX_NU; ++j) {
            if (j < nu) {
                double un;
                if (kk < N) un = buf[kk * nu + j];
                else {
                    double acc = buf[N * nu + j];
#pragma unroll 1
                    for (int k = 0; k < nx; ++k) acc = fma(-__ldg(L.K + j * nx + k), (ext || cons) ? S.x_nom(k) : S.x(k), acc);
                    un = acc;
                }
                if (ext) {
                    double acc = un;
#pragma unroll 1
                    for (int k = 0; k < nx; ++k) acc = fma(-__ldg(L.Kp + j * nx + k), S.x(k) - S.x_nom(k), acc);
                    un = acc;
                }
                uh[j] = un;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < LOOP_MAX_NU; ++j) if (j < nu) uh[j] = Ub[j];      // first input of the latest sent sequence
    }
    // ---- lane i: row i of nominal model, plant and estimator ----------------------------------------------
    double xn_new = 0.0, x_new = 0.0, xh_new = 0.0;
    if (lane < nx) {
        const int i = lane;
        if (!smart) {
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < nx; ++k) acc = fma(__ldg(L.A + i * nx + k), S.x_nom(k), acc);
#pragma unroll
            for (int j = 0; j < LOOP_MAX_NU; ++j) if (j < nu) acc = fma(__ldg(L.Bm + i * nu + j), u_nom[j], acc);
            xn_new = acc;
        } else xn_new = S.x_nom(i);
        if (L.plant != RTMPC_PLANT_CARTPOLE) {
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < nx; ++k) acc = fma(__ldg(L.A + i * nx + k), S.x(k), acc);
#pragma unroll
            for (int j = 0; j < LOOP_MAX_NU; ++j) if (j < nu) acc = fma(__ldg(L.Bm + i * nu + j), u[j], acc);
            x_new = acc + wi;
        }
        {
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < nx; ++k) {
                double xb;
                if (gamma == 1) xb = cons ? S.x_nom(k) : S.x(k);
                else xb = (ext && x_nom0_b) ? x_nom0_b[k] : S.x_hat(k);
                acc = fma(__ldg(L.A + i * nx + k), xb, acc);
            }
#pragma unroll
            for (int j = 0; j < LOOP_MAX_NU; ++j) if (j < nu) acc = fma(__ldg(L.Bm + i * nu + j), uh[j], acc);
            xh_new = acc;
        }
    }
    double xc[4] = {0.0, 0.0, 0.0, 0.0};
    if (L.plant == RTMPC_PLANT_CARTPOLE && lane == 0) {
        for (int k = 0; k < 4; ++k) xc[k] = S.x(k);
        cartpole_substeps_of(L, xc, u[0]);
    }
    __syncwarp();
    // ---- write back -----------------------------------------------------------------------------------------
    if (lane < nx) {
        if (L.plant != RTMPC_PLANT_CARTPOLE) S.x(lane) = x_new;
        S.x_nom(lane) = xn_new;
        S.x_hat(lane) = xh_new;
    }
    if (lane == 0) {
        if (L.plant == RTMPC_PLANT_CARTPOLE) for (int k = 0; k < 4; ++k) S.x(k) = xc[k];
        for (int j = 0; j < nu; ++j) S.u_last(j) = u[j];
        if (gamma == 1) S.q_t() = t;
        S.s_t() = s_t;
        S.Theta() = Theta;
        S.last_loss() = last_loss;
        S.gamma_last() = gamma;
    }
    __syncwarp();
    if (traj_b && lane < nx) traj_b[(size_t)(t + 1) * nx + lane] = S.x(lane);
}

// The warp step with the sizes known at compile time (the reference's two systems: nx = 4 / 2, nu = 1).  Lane group
// g = lane / NX takes one of the three mat-vecs (0: nominal model, 1: plant, 2: estimator), lane i = lane % NX its row i;
// the scalars (inputs, flags) are computed redundantly in every lane.  Straight-line code, a third of the generic
// version's instructions; every output element goes through the same FMA sequence as loop_step_body_t (same bits).
template <int NX, int NU, class State>
__device__ __forceinline__ void loop_step_body_warp_t(const LoopDev& L, const State S, int lane, int t, const double* Ub,
                                                      const double* x_nom0_b, const double* ref_b, int theta_in,
                                                      int gamma_in, const double* w_in_b, double p,
                                                      unsigned long long seed, unsigned long long id, double* traj_b,
                                                      double tube_worst) {
    static_assert(3 * NX <= 32 && ((NX + 1) >> 1) + 1 <= 32, "lane groups");
    const int N = L.N;
    const int g = lane / NX, i = lane - g * NX;          // (NX is a compile-time constant)
    // ---- statistics on the pre-step state ---------------------------------------------------------------
    if (lane == 0) {
        if (ref_b) {
            double e = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) { const double d = S.x(k) - ref_b[k]; e = fma(d, d, e); }
            S.err_acc() += e;
        }
        if (L.nz_rows > 0) S.tube_max() = fmax(S.tube_max(), tube_worst);
    }
    // ---- network and disturbance realisation ------------------------------------------------------------
    int theta, gamma;
    double wi = 0.0;                                     // lane NX + i (plant group): disturbance on state i
    if (theta_in >= 0) {
        theta = theta_in;
        gamma = gamma_in;
        if (g == 1 && w_in_b) wi = w_in_b[i];
    } else {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        Philox4 r;
        r.x = r.y = r.z = r.w = 0u;
        if (lane <= ((NX + 1) >> 1)) r = loop_philox((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (uint32_t)lane, k0, k1);
        const int th0 = (t == 0) ? 1 : (u01_from_bits(r.x, r.y) < p ? 0 : 1);
        const int ga0 = (t == 0) ? 1 : (u01_from_bits(r.z, r.w) < p ? 0 : 1);
        theta = __shfl_sync(0xffffffffu, th0, 0);
        gamma = __shfl_sync(0xffffffffu, ga0, 0);
        double wa = 0.0, wb = 0.0;
        if (lane >= 1 && lane <= ((NX + 1) >> 1)) {
            const int k = 2 * (lane - 1);
            wa = __ldg(L.w_half + k) * (2.0 * u01_from_bits(r.x, r.y) - 1.0);
            if (k + 1 < NX) wb = __ldg(L.w_half + k + 1) * (2.0 * u01_from_bits(r.z, r.w) - 1.0);
        }
        const int src = 1 + (i >> 1);
        const double ga = __shfl_sync(0xffffffffu, wa, src), gb = __shfl_sync(0xffffffffu, wb, src);
        if (g == 1) wi = (i & 1) ? gb : ga;
    }
    // ---- local side ---------------------------------------------------------------------------------------
    const int q_pkt = S.q_t();
    int last_loss = S.last_loss();
    int Theta = 0;
    if (theta == 1) Theta = (last_loss <= q_pkt) ? 1 : 0;
    else last_loss = t;
    int s_t = S.s_t();
    double* buf = S.buf();
    const bool smart = L.actuator == RTMPC_ACT_SMART, cons = L.actuator == RTMPC_ACT_CONSISTENT,
               ext = L.actuator == RTMPC_ACT_EXTENDED;
    __syncwarp();
    if (Theta) {
        s_t = t;
        for (int k = lane; k < (N + 1) * NU; k += 32) buf[k] = Ub[k];
        if (!smart && x_nom0_b && lane < NX) S.x_nom(lane) = x_nom0_b[lane];
    }
    __syncwarp();
    const int kk = t - s_t;
    double x[NX], xn[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) { x[k] = S.x(k); xn[k] = S.x_nom(k); }
    double u_nom[NU], u[NU], uh[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        if (kk < N) u_nom[j] = buf[kk * NU + j];
        else {
            double acc = buf[N * NU + j];
#pragma unroll
            for (int k = 0; k < NX; ++k) acc = fma(-__ldg(L.K + j * NX + k), smart ? x[k] : xn[k], acc);
            u_nom[j] = acc;
        }
        u[j] = u_nom[j];
        if (!smart) {
            double acc = u_nom[j];
#pragma unroll
            for (int k = 0; k < NX; ++k) acc = fma(-__ldg(L.Kp + j * NX + k), x[k] - xn[k], acc);
            u[j] = acc;
        }
        // remote side's input estimate (plant packet content = pre-update values)
        if (gamma == 1) {
            double un;
            if (kk < N) un = buf[kk * NU + j];
            else {
                double acc = buf[N * NU + j];
#pragma unroll
                for (int k = 0; k < NX; ++k) acc = fma(-__ldg(L.K + j * NX + k), (ext || cons) ? xn[k] : x[k], acc);
                un = acc;
            }
            if (ext) {
                double acc = un;
#pragma unroll
                for (int k = 0; k < NX; ++k) acc = fma(-__ldg(L.Kp + j * NX + k), x[k] - xn[k], acc);
                un = acc;
            }
            uh[j] = un;
        } else uh[j] = Ub[j];                                 // first input of the latest sent sequence
    }
    // ---- group g: row i of nominal model (0), plant (1), estimator (2) --------------------------------------
    double out = 0.0;
    if (g < 3) {
        // the vector this group multiplies: one pointer select instead of a select per element
        const double* src = &S.x_nom(0);
        if (g == 1) src = &S.x(0);
        else if (g == 2) {
            if (gamma == 1) src = cons ? &S.x_nom(0) : &S.x(0);
            else src = (ext && x_nom0_b) ? x_nom0_b : &S.x_hat(0);
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) acc = fma(__ldg(L.A + i * NX + k), src[k], acc);
#pragma unroll
        for (int j = 0; j < NU; ++j) acc = fma(__ldg(L.Bm + i * NU + j), (g == 0) ? u_nom[j] : (g == 1) ? u[j] : uh[j], acc);
        out = (g == 1) ? acc + wi : acc;
        if (g == 0 && smart) out = S.x_nom(i);                // (no nominal model in Pezzutto's scheme)
    }
    // (the array the plant routine works on lives in local memory: only inside the branch that needs it)
    double xc0 = 0.0, xc1 = 0.0, xc2 = 0.0, xc3 = 0.0;
    const bool cart = L.plant == RTMPC_PLANT_CARTPOLE;
    if (NX == 4 && cart && lane == 0) {
        double xc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xc[k] = x[k < NX ? k : 0];
        cartpole_substeps_of(L, xc, u[0]);
        xc0 = xc[0]; xc1 = xc[1]; xc2 = xc[2]; xc3 = xc[3];
    }
    __syncwarp();
    // ---- write back -----------------------------------------------------------------------------------------
    if (g == 0) S.x_nom(i) = out;
    else if (g == 1) { if (!cart) S.x(i) = out; }
    else if (g == 2) S.x_hat(i) = out;
    if (lane == 0) {
        if (NX == 4 && cart) { S.x(0) = xc0; S.x(NX > 1 ? 1 : 0) = xc1; S.x(NX > 2 ? 2 : 0) = xc2; S.x(NX > 3 ? 3 : 0) = xc3; }
#pragma unroll
        for (int j = 0; j < NU; ++j) S.u_last(j) = u[j];
        if (gamma == 1) S.q_t() = t;
        S.s_t() = s_t;
        S.Theta() = Theta;
        S.last_loss() = last_loss;
        S.gamma_last() = gamma;
    }
    __syncwarp();
    if (traj_b && lane < NX) traj_b[(size_t)(t + 1) * NX + lane] = S.x(lane);
}

template <class State>
__device__ __forceinline__ void loop_step_warp(const LoopDev& L, const State S, int lane, int t, const double* Ub,
                                               const double* x_nom0_b, const double* ref_b, int theta_in, int gamma_in,
                                               const double* w_in_b, double p, unsigned long long seed,
                                               unsigned long long id, double* traj_b, double tube_worst) {
    if (L.nx == 4 && L.nu == 1)
        loop_step_body_warp_t<4, 1>(L, S, lane, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
    else if (L.nx == 2 && L.nu == 1)
        loop_step_body_warp_t<2, 1>(L, S, lane, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
    else
        loop_step_body_warp(L, S, lane, t, Ub, x_nom0_b, ref_b, theta_in, gamma_in, w_in_b, p, seed, id, traj_b, tube_worst);
}

#ifdef RTMPC_LOOP_KERNELS   // the non-template kernels are compiled in one translation unit (rtmpc_capi.cu)
__global__ void loop_step_kernel(LoopDev L, int B, int t, const double* __restrict__ U_t,
                                 const int* __restrict__ status, const double* __restrict__ x_nom0,
                                 long long x_nom0_stride, const double* __restrict__ ref,
                                 const int* __restrict__ theta_in, const int* __restrict__ gamma_in,
                                 const double* __restrict__ w_in, const double* __restrict__ p_loss,
                                 unsigned long long seed, long long id_offset, double* __restrict__ traj,
                                 long long traj_stride) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nx = L.nx, nu = L.nu, N = L.N;
    double* traj_b = traj ? traj + (size_t)b * traj_stride : nullptr;
    LoopGlobalState S;
    S.L = &L; S.b = b;
    if (!loop_step_begin(L, S, t, status ? status[b] : RTMPC_OPTIMAL, traj_b)) return;
    const double worst = (L.nz_rows > 0) ? loop_tube_rows(L, S, 0, 1) : 0.0;
    loop_step_body(L, S, t, U_t + (size_t)b * (N + 1) * nu, x_nom0 ? x_nom0 + (size_t)b * x_nom0_stride : nullptr,
                   ref ? ref + (size_t)b * nx : nullptr, theta_in ? theta_in[b] : -1, theta_in ? gamma_in[b] : -1,
                   w_in ? w_in + (size_t)b * nx : nullptr, p_loss ? p_loss[b] : 0.0, seed,
                   (unsigned long long)(id_offset + b), traj_b, worst);
}

// out[j] = max_v <dirs[j,:], V[v,:]> as a product  dirs [M x dim] * V' [dim x nv]  reduced by a row maximum, on the FP64
// tensor pipe: mma.sync m8n8k4 (the instruction cuBLAS DGEMM runs on here; tcgen05 has no f64 kind).  One warp owns
// RT tiles of 8 directions (A fragments stay in registers for the whole sweep); the vertices are staged once per
// block in shared memory as [KC][nv8][4] (KC chunks of four coordinates, zero padded; nv rounded up to a multiple
// of 8 by repeating vertex 0), so the B fragment of a tile of 8 vertices is 32 consecutive doubles: one conflict-free
// 8-byte load per lane, shared by the RT products.  Each product is followed by two compare-selects per lane; the
// maxima of a row's four lanes are combined at the very end.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int RT, int KC>
__global__ void __launch_bounds__(512, 1)
support_sweep_kernel(const double* __restrict__ V, int nv, int dim, const double* __restrict__ dirs, long long M,
                     double* __restrict__ out) {
    extern __shared__ __align__(16) double sv[];
    const int nv8 = (nv + 7) & ~7;
    for (int i = threadIdx.x; i < KC * nv8 * 4; i += blockDim.x) {
        const int kc = i / (nv8 * 4), r = i - kc * (nv8 * 4), v = r >> 2, k = kc * 4 + (r & 3);
        sv[i] = (k < dim) ? V[(size_t)(v < nv ? v : 0) * dim + k] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = lane >> 2, col = lane & 3;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp0 * (8 * RT); base < M; base += nwarps * (8 * RT)) {
        double a[RT][KC];
#pragma unroll
        for (int t = 0; t < RT; ++t) {
            const long long j = base + 8 * t + row;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                const int k = kc * 4 + col;
                a[t][kc] = (j < M && k < dim) ? dirs[j * dim + k] : 0.0;
            }
        }
        // Running maxima as SIGNED 64-BIT INTEGERS on the bit patterns: for doubles >= 0 the integer order is the order of
        // the values, and every negative double is a negative integer, i.e. below every non-negative one - so the integer
        // maximum is the maximum whenever it is >= 0, which it is for a set that contains the origin (every set the tube
        // pipeline sweeps).  The compares then run on the integer pipe instead of sharing the FP64 pipe with the DMMAs
        // (DSETP: the kernel was limited by that pipe at 66 % DMMA utilisation).  A direction whose maximum comes out
        // negative is redone with FP64 compares below.
        long long k0[RT], k1[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) { k0[t] = (long long)0x8000000000000000ull; k1[t] = k0[t]; }
#pragma unroll 2
        for (int vt = 0; vt < nv8; vt += 8) {
            double b[KC];
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) b[kc] = sv[(kc * nv8 + vt) * 4 + lane];
#pragma unroll
            for (int t = 0; t < RT; ++t) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) dmma884(c0, c1, a[t][kc], b[kc], c0, c1);
                const long long i0 = __double_as_longlong(c0), i1 = __double_as_longlong(c1);
                k0[t] = (i0 > k0[t]) ? i0 : k0[t];
                k1[t] = (i1 > k1[t]) ? i1 : k1[t];
            }
        }
        bool redo = false;
#pragma unroll
        for (int t = 0; t < RT; ++t) {
            long long m = (k0[t] > k1[t]) ? k0[t] : k1[t];
            long long o = __shfl_xor_sync(RTMPC_FULL_MASK, m, 1);
            m = (o > m) ? o : m;
            o = __shfl_xor_sync(RTMPC_FULL_MASK, m, 2);
            m = (o > m) ? o : m;
            const long long j = base + 8 * t + row;
            if (j < M && m < 0) redo = true;
            if (col == 0 && j < M) out[j] = __longlong_as_double(m);
        }
        if (!__any_sync(RTMPC_FULL_MASK, redo)) continue;
        // (rare) some direction of this trip has a negative maximum: FP64 compares for the whole trip
        double best0[RT], best1[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) { best0[t] = -RTMPC_INF; best1[t] = -RTMPC_INF; }
#pragma unroll 1
        for (int vt = 0; vt < nv8; vt += 8) {
            double b[KC];
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) b[kc] = sv[(kc * nv8 + vt) * 4 + lane];
#pragma unroll
            for (int t = 0; t < RT; ++t) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) dmma884(c0, c1, a[t][kc], b[kc], c0, c1);
                best0[t] = (c0 > best0[t]) ? c0 : best0[t];       // (a compare and a select; FP64 fmax is a longer sequence)
                best1[t] = (c1 > best1[t]) ? c1 : best1[t];
            }
        }
#pragma unroll
        for (int t = 0; t < RT; ++t) {
            double m = (best0[t] > best1[t]) ? best0[t] : best1[t];
            double o = __shfl_xor_sync(RTMPC_FULL_MASK, m, 1);
            m = (o > m) ? o : m;
            o = __shfl_xor_sync(RTMPC_FULL_MASK, m, 2);
            m = (o > m) ? o : m;
            const long long j = base + 8 * t + row;
            if (col == 0 && j < M) out[j] = m;
        }
    }
}

}  // namespace rtmpc

// ------------------------------------------------------------------------------------------------
// Split entry points for the reference's per-object call order
//   u, plant_packet = actuator.process_packet(packet, x, theta)   (SmartActuator.py:31-54, :174-213)
//   ... caller steps the plant ...
//   estimator.update_estimate(plant_packet, gamma)                 (Estimator.py:43-78, :113-156)
// They operate on caller-owned device arrays (the Python classes keep them in torch tensors).
// ------------------------------------------------------------------------------------------------
namespace rtmpc {

struct ActArgs {
    int nx, nu, N, kind, t, has_xnom0;
    const double *A, *Bm, *K, *Kp;
    const double *x_t, *U_t, *x_nom0;
    const int *q_pkt, *theta;
    double *buf, *x_nom, *u_out, *pkt_x, *pkt_xnom;
    int *s_t, *Theta, *last_loss;
};

__global__ void actuator_process_kernel(ActArgs a, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nx = a.nx, nu = a.nu, N = a.N;
    double x[LOOP_MAX_NX], xn[LOOP_MAX_NX];
    for (int k = 0; k < nx; ++k) {
        x[k] = a.x_t[(size_t)b * nx + k];
        xn[k] = (a.kind == RTMPC_ACT_SMART) ? 0.0 : a.x_nom[(size_t)b * nx + k];
    }
    int last_loss = a.last_loss[b], Theta = 0, s_t = a.s_t[b];
    if (a.theta[b] == 1) Theta = (last_loss <= a.q_pkt[b]) ? 1 : 0;
    else last_loss = a.t;
    double* buf = a.buf + (size_t)b * (N + 1) * nu;
    if (Theta) {
        s_t = a.t;
        for (int i = 0; i < (N + 1) * nu; ++i) buf[i] = a.U_t[(size_t)b * (N + 1) * nu + i];
        if (a.kind != RTMPC_ACT_SMART && a.has_xnom0)
            for (int k = 0; k < nx; ++k) xn[k] = a.x_nom0[(size_t)b * nx + k];
    }
    const int kk = a.t - s_t;
    const double* xfb = (a.kind == RTMPC_ACT_SMART) ? x : xn;
    double u_nom[LOOP_MAX_NU];
    for (int j = 0; j < nu; ++j) {
        if (kk < N) u_nom[j] = buf[kk * nu + j];
        else {
            double acc = buf[N * nu + j];
            for (int k = 0; k < nx; ++k) acc = fma(-a.K[j * nx + k], xfb[k], acc);
            u_nom[j] = acc;
        }
        double u = u_nom[j];
        if (a.kind != RTMPC_ACT_SMART)
            for (int k = 0; k < nx; ++k) u = fma(-a.Kp[j * nx + k], x[k] - xn[k], u);
        a.u_out[(size_t)b * nu + j] = u;
    }
    for (int k = 0; k < nx; ++k) {
        a.pkt_x[(size_t)b * nx + k] = (a.kind == RTMPC_ACT_CONSISTENT) ? xn[k] : x[k];
        if (a.pkt_xnom) a.pkt_xnom[(size_t)b * nx + k] = xn[k];
    }
    if (a.kind != RTMPC_ACT_SMART) {
        for (int i = 0; i < nx; ++i) {
            double acc = 0.0;
            for (int k = 0; k < nx; ++k) acc = fma(a.A[i * nx + k], xn[k], acc);
            for (int j = 0; j < nu; ++j) acc = fma(a.Bm[i * nu + j], u_nom[j], acc);
            a.x_nom[(size_t)b * nx + i] = acc;
        }
    }
    a.s_t[b] = s_t;
    a.Theta[b] = Theta;
    a.last_loss[b] = last_loss;
}

struct EstArgs {
    int nx, nu, N, robust, t, n_hist;
    const double *A, *Bm, *K, *Kp;
    const double *pkt_x, *pkt_xnom, *x_nom0_mpc, *hist;   // hist[(time)*B*(N+1)*nu + b*(N+1)*nu + ...]
    const int *pkt_s, *gamma;
    double* x_hat;
    int* q_t;
};

__global__ void estimator_update_kernel(EstArgs e, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nx = e.nx, nu = e.nu, N = e.N;
    const size_t seq = (size_t)(N + 1) * nu;
    double base[LOOP_MAX_NX], uh[LOOP_MAX_NU];
    if (e.gamma[b] == 1) {
        const int s_t = e.pkt_s[b];
        if (s_t < 0 || s_t >= e.n_hist) {
            // the packet names a sequence that was never stored (Estimator.py:55 would raise IndexError): poison the
            // estimate instead of reading out of bounds
            for (int i = 0; i < nx; ++i) e.x_hat[(size_t)b * nx + i] = __longlong_as_double(0x7ff8000000000000LL);
            return;
        }
        const double* sq = e.hist + ((size_t)s_t * B + b) * seq;
        const int kk = e.t - s_t;
        for (int k = 0; k < nx; ++k) base[k] = e.pkt_x[(size_t)b * nx + k];
        for (int j = 0; j < nu; ++j) {
            double un;
            if (kk < N) un = sq[kk * nu + j];
            else {
                double acc = sq[N * nu + j];
                const double* xs = e.robust ? e.pkt_xnom : e.pkt_x;
                for (int k = 0; k < nx; ++k) acc = fma(-e.K[j * nx + k], xs[(size_t)b * nx + k], acc);
                un = acc;
            }
            if (e.robust)
                for (int k = 0; k < nx; ++k)
                    un = fma(-e.Kp[j * nx + k], e.pkt_x[(size_t)b * nx + k] - e.pkt_xnom[(size_t)b * nx + k], un);
            uh[j] = un;
        }
    } else {
        const double* sq = e.hist + ((size_t)(e.n_hist - 1) * B + b) * seq;
        for (int j = 0; j < nu; ++j) uh[j] = sq[j];
        for (int k = 0; k < nx; ++k)
            base[k] = e.robust ? e.x_nom0_mpc[(size_t)b * nx + k] : e.x_hat[(size_t)b * nx + k];
    }
    double out[LOOP_MAX_NX];
    for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int k = 0; k < nx; ++k) acc = fma(e.A[i * nx + k], base[k], acc);
        for (int j = 0; j < nu; ++j) acc = fma(e.Bm[i * nu + j], uh[j], acc);
        out[i] = acc;
    }
    for (int i = 0; i < nx; ++i) e.x_hat[(size_t)b * nx + i] = out[i];
    if (e.gamma[b] == 1) e.q_t[b] = e.t;
}

// (re)initialisation of every instance in one launch: x_nom = x_hat = x (already copied in), empty buffers, t = 0 state
// of the actuator / estimator objects (SmartActuator.py:13-24,129-144, Estimator.py:11-26), warm-start records cleared
__global__ void loop_reset_kernel(LoopDev L, int B, int* __restrict__ warm0, int stride0, int* __restrict__ warm1, int stride1) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nx = L.nx, nu = L.nu, usz = (L.N + 1) * nu;
    for (int k = 0; k < nx; ++k) {
        const double v = L.x[(size_t)b * nx + k];
        L.x_nom[(size_t)b * nx + k] = v;
        L.x_hat[(size_t)b * nx + k] = v;
    }
    for (int i = 0; i < usz; ++i) L.buf[(size_t)b * usz + i] = 0.0;
    for (int j = 0; j < nu; ++j) L.u_last[(size_t)b * nu + j] = 0.0;
    L.err_acc[b] = 0.0;
    L.tube_max[b] = -1e300;
    L.q_t[b] = 0; L.s_t[b] = 0; L.Theta[b] = 0;
    L.alive[b] = 1; L.last_loss[b] = -1; L.gamma_last[b] = 1;
    if (warm0) for (int i = 0; i < stride0; ++i) warm0[(size_t)b * stride0 + i] = -1;
    if (warm1) for (int i = 0; i < stride1; ++i) warm1[(size_t)b * stride1 + i] = -1;
}

// Model-error sweep (estimate_W_for_Cartpole.py:79-127): one thread per run, T control periods of the nonlinear plant
// under the zero-order-hold LQR law, w_k = x_{k+1} - Acl x_k.
__global__ void model_error_kernel(const double* __restrict__ cart, int B, int T, const double* __restrict__ x0,
                                   const double* __restrict__ K, const double* __restrict__ Acl,
                                   double* __restrict__ w, double* __restrict__ x_final) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double c[8], x[4], k[4], a[16];
    for (int i = 0; i < 8; ++i) c[i] = cart[i];
    for (int i = 0; i < 4; ++i) { x[i] = x0[(size_t)b * 4 + i]; k[i] = K[i]; }
    for (int i = 0; i < 16; ++i) a[i] = Acl[i];
    for (int t = 0; t < T; ++t) {
        double u = 0.0, lin[4];
        for (int i = 0; i < 4; ++i) u = fma(-k[i], x[i], u);
        for (int i = 0; i < 4; ++i) {
            double acc = 0.0;
            for (int j = 0; j < 4; ++j) acc = fma(a[i * 4 + j], x[j], acc);
            lin[i] = acc;
        }
        cartpole_substeps(x, u, c);
        double* wo = w + ((size_t)b * T + t) * 4;
        for (int i = 0; i < 4; ++i) wo[i] = x[i] - lin[i];
    }
    if (x_final) for (int i = 0; i < 4; ++i) x_final[(size_t)b * 4 + i] = x[i];
}

#endif  // RTMPC_LOOP_KERNELS

}  // namespace rtmpc

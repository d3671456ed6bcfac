// Batched dual active-set QP kernel (Goldfarb-Idnani on operators shared by the whole batch), FP64, sm_100a.
//
// One warp owns one instance of the condensed, equilibrated problem of rtmpc_ipm.cuh
//     min 1/2 z'Hs z + q'z   s.t.  lo <= G z <= up,   q = Fx x_init + Fr ref,  lo/up affine in x_init.
// Everything that depends only on (A, B, Q, R, sets) is prepared once per problem and shared by all
// instances:  Hinv,  Y = G Hinv,  W = G Hinv G'  and the parameter maps
//     z_u = Zx x_init + Zr ref  (unconstrained minimiser),   G z_u = Tx x_init + Tr ref.
// With those, a dual active-set step never touches H or G: for the working set A (signed rows n_a)
//     multipliers  lam_A = M (N_A' z_u - b_A),  M = (N_A' Hinv N_A)^-1 = inverse of a signed sub-matrix of W,
//     adding the violated row p moves every row value by  -step * (s_p W[:,p] - W[:,A] (s_A r)),  r = M W[A,p].
// M is kept explicitly and updated by bordering (row added) / a rank-one downdate (row dropped): a step
// is one small mat-vec, a few warp reductions and one pass over |A|+1 rows of W.  No factorisation.
//
// Data placement.  Per lane, in registers: e_i = (G z)_i - up_i of its 2*R2 rows (rows r2*64 + 2*lane + {0,1}:
// every table is read with 16-byte loads; the lower side of a row is -e_i - wid_i with the constant width
// wid = up - lo read from a table), plus the multiplier and row of working-set slot `lane`.  Per warp, in shared
// memory: M (one row per slot), z_u, z and two scratch vectors.  Shared by all warps, read through L1: W, G', Y
// and the parameter tables.  The kernel is kept small on purpose (real calls for the dense helpers, rolled
// loops): the warps of an SM are all at different points of their solves and share one instruction cache.
//
// Per instance (mirrored by tools/as_model.py::solve_as_inv):
//   0. parameter rows; z_u; row values at z_u; no row violated -> z_u is optimal.
//   1. warm start: last control step's certified active set moved one stage earlier (QPDev.shift): its signed
//      sub-matrix of W is inverted in place (Gauss-Jordan, rows that depend on earlier ones skipped -- the same
//      pivots as bordering row by row), multipliers = M rhs, rows with negative multipliers are dropped one at
//      a time until the set is dual feasible.
//   2. Goldfarb-Idnani: add the most violated row with primal/dual ratio test (partial steps drop
//      the blocking row) until no row is violated by more than 1e-11 (scaled).  Strictly increasing
//      dual objective: no cycling.  Linear dependence without a blocking row = primal infeasible.
//   3. certification: multipliers refined against the true rows of G with M as approximate inverse
//      (three passes), z = z_u - Y_A' (s lam), every row recomputed from G' z and re-checked, every
//      multiplier re-checked (KKT certificate, same tolerances as the interior-point kernel's
//      endgame).  A violated row found here sends the instance back to 2 with exact row values.
// Anything unexpected (step cap, refinement not converging, negative multiplier at certification)
// returns RTMPC_FALLBACK; the C ABI then runs the interior-point kernel on those instances only.
#pragma once
#include "rtmpc_ipm.cuh"

namespace rtmpc {

#ifdef RTMPC_AS_DEBUG
// development counters (scratch builds only): [0] solves past the unconstrained test, [1] carried, [2] moved, [3] changes,
// [4] old set sizes, [5] from-scratch inversions, [6] their candidates, [7] GI adds, [8] GI drops, [9] warm multiplier drops,
// [16 + k] histogram of the certified working-set size
static __device__ unsigned long long g_as_dbg[64];      // one copy per translation unit; the rollout's is read out
#define AS_DBG(i, v) do { if (lane == 0) atomicAdd(&g_as_dbg[i], (unsigned long long)(v)); } while (0)
#else
#define AS_DBG(i, v) do { } while (0)
#endif

constexpr int RTMPC_FALLBACK = RTMPC_FALLBACK_STATUS;
// as_gi: the entering row is a combination of the working set, no multiplier blocks, and its violation is tiny
constexpr int AS_STALL = 7;

// per-warp shared memory, in doubles: M (one row per working-set slot), zu, z, v, coef, 16 parameters, slot lists
// (2*npad ints).  Kept as small as possible: what the warps do not take is L1 for the shared tables, and
// the L1 / shared split moves in steps (.., 100, 132, .. KB): 16 warps of the cartpole problem fit the 100 KB step.
// Problem dimensions as the solver sees them: read from the problem description (ASDimsDyn, any problem) or compile-time
// constants (ASDimsFix: the rollout kernel's instantiation for the reference's cartpole controller - offsets into the
// per-warp shared memory become immediates, the short loops over nx / npad get constant trip counts; same arithmetic in
// the same order, so the same bits).
struct ASDimsDyn {
    static constexpr bool kFixed = false;
    static constexpr int kNx = 0, kNu = 0;
    static __device__ __forceinline__ int n(const QPDev& P) { return P.n; }
    static __device__ __forceinline__ int npad(const QPDev& P) { return P.npad; }
    static __device__ __forceinline__ int nx(const QPDev& P) { return P.nx; }
    static __device__ __forceinline__ int nu(const QPDev& P) { return P.nu; }
    static __device__ __forceinline__ int N(const QPDev& P) { return P.N; }
};
template <int N_VARS, int NPAD, int NX, int NU, int HORIZON>
struct ASDimsFix {
    static constexpr bool kFixed = true;
    static constexpr int kNx = NX, kNu = NU;
    static __device__ __forceinline__ constexpr int n(const QPDev&) { return N_VARS; }
    static __device__ __forceinline__ constexpr int npad(const QPDev&) { return NPAD; }
    static __device__ __forceinline__ constexpr int nx(const QPDev&) { return NX; }
    static __device__ __forceinline__ constexpr int nu(const QPDev&) { return NU; }
    static __device__ __forceinline__ constexpr int N(const QPDev&) { return HORIZON; }
};
__host__ __device__ inline int as_ms(const QPDev& P) { return P.npad + 2; }
__host__ __device__ inline int as_mrows(const QPDev& P) { return P.n; }      // one row per slot (<= n rows in the working set; the row stride is even)
__host__ __device__ inline int as_warp_doubles(const QPDev& P) {
    return as_mrows(P) * as_ms(P) + 4 * P.npad + 16 + P.npad + 6;
}

// (offsets are added to one base pointer where they are used: nine live pointers would not stay in registers)
struct ASWarp {
    double* base;
    int off;          // base as an offset (in doubles) into the kernel's dynamic shared memory, for the real-call helpers
    int npad, mm;     // mm = rows of M * ms
    __device__ __forceinline__ double* M() const { return base; }
    __device__ __forceinline__ int Mo() const { return off; }
    __device__ __forceinline__ int vo() const { return off + mm + 2 * npad; }
    __device__ __forceinline__ int rvo() const { return off + mm + 3 * npad; }
    __device__ __forceinline__ double* zu() const { return base + mm; }
    __device__ __forceinline__ double* z() const { return base + mm + npad; }
    __device__ __forceinline__ double* v() const { return base + mm + 2 * npad; }
    __device__ __forceinline__ double* coef() const { return base + mm + 3 * npad; }
    __device__ __forceinline__ double* rv() const { return base + mm + 3 * npad; }      // shares coef's storage
    // list of the rows a streaming pass adds to the row values, 16 bytes each: coefficient, offset of the row in W (in doubles).
    // Lives in v's and coef's storage (2 npad doubles = npad entries), which nothing else uses between the step's mat-vec and
    // its bordering / down-date.
    __device__ __forceinline__ double* lst() const { return base + mm + 2 * npad; }
    __device__ __forceinline__ double* xr() const { return base + mm + 4 * npad; }
    __device__ __forceinline__ int* act_row() const { return reinterpret_cast<int*>(base + mm + 4 * npad + 16); }
    __device__ __forceinline__ int* act_sgn() const { return act_row() + npad; }
    // rarely touched control state lives here rather than in registers: ctl()[0] the row tolerance in force,
    // ctl()[1] violation of the row a stall happened on, ctl()[2] sum of the magnitudes of every coefficient applied to the row
    // values since they were last evaluated from scratch (1e300: they no longer follow the multipliers; as_certify);
    // ictl()[4] value of ASCounters::rows at that evaluation; ictl()[0] steps at the last factorisation, [1] steps at
    // the last exact row values, [2] why the solve last refactorised or gave up (ASCounters), [3] refactorisations
    __device__ __forceinline__ double* ctl() const { return base + mm + 5 * npad + 16; }
    __device__ __forceinline__ int* ictl() const { return reinterpret_cast<int*>(base + mm + 5 * npad + 19); }
};

template <class D = ASDimsDyn>
__device__ __forceinline__ ASWarp as_carve(double* base, const QPDev& P) {
    extern __shared__ __align__(16) double as_smem[];
    ASWarp w;
    w.base = base; w.off = (int)(base - as_smem); w.npad = D::npad(P); w.mm = D::n(P) * (D::npad(P) + 2);
    return w;
}

// The small dense helpers are inlined at their call sites.  (Round 1 made them real calls, which paid when the kernel was
// 16 k instructions; with the kernel at half that size the call / return jumps cost more instruction fetches than the
// copies: all three inlined 88.5 -> 92.6 M solves/s, reductions as calls 84.6 -> 80.1; RTMPC_AS_CALLS restores the calls.)
#ifdef RTMPC_AS_CALLS
#define AS_FN_MATVEC static __device__ __noinline__
#define AS_FN_BORDER static __device__ __noinline__
#define AS_FN_DOWNDATE static __device__ __noinline__
#define AS_FN_INVERT static __device__ __noinline__
#else
#define AS_FN_MATVEC __device__ __forceinline__
#define AS_FN_BORDER __device__ __forceinline__
#define AS_FN_DOWNDATE __device__ __forceinline__
#define AS_FN_INVERT __device__ __forceinline__
#endif
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
// 16-byte read-only loads with an L1 policy.  The shared tables (1.3 MB for the cartpole) go through a ~118 KB L1 that
// 16 warps at different points of their solves compete for; what is streamed once per use should not evict what every
// warp re-reads every control step.  0: default, 1: no_allocate, 2: evict_first, 3: evict_last.
template <int HINT>
__device__ __forceinline__ double2 ld2_hint(const double* p) {
    double2 v;
    if (HINT == 1) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (HINT == 2) asm volatile("ld.global.nc.L1::evict_first.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (HINT == 3) asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else v = *reinterpret_cast<const double2*>(p);
    return v;
}
#ifndef RTMPC_HINT_GT
#define RTMPC_HINT_GT 1        // columns of G' (certification): streamed
#endif
#ifndef RTMPC_HINT_W
#define RTMPC_HINT_W 0         // rows of W (active-set steps)
#endif
#ifndef RTMPC_HINT_SETUP
#define RTMPC_HINT_SETUP 0     // ExT / TrT (row values at z_u, every solve)
#endif
#ifndef RTMPC_HINT_UX
#define RTMPC_HINT_UX 0        // UxT (certification)
#endif
__device__ __forceinline__ double2 ld2_stream(const double* p) { return ld2_hint<RTMPC_HINT_GT>(p); }

// steps: rows added + dropped; rounds: certifications; rows: passes over all m rows (rows of W streamed, columns of the
// certification's tables); sq: sum of na^2 over the small dense operations (mat-vec, bordering, downdate) and 6 n^2 per
// certification (refinement).  Algorithmic flops are derived from these: 2 m per row pass, 2 per unit of sq.
// why (kept in ASWarp::ictl()[2]): last reason the Goldfarb-Idnani / certification pair gave up on a factorisation (diagnostics; 0 = never)
//   1 step cap, 2 dependent row without a blocking multiplier and a tiny violation, 3 no free slot,
//   4 certification failed (negative multiplier / refinement), 5 certification rounds, 6 contradiction late in a leg
struct ASCounters { int steps, rounds, rows, sq; };

__device__ __forceinline__ unsigned long long as_flops(const QPDev& P, const ASCounters& c, bool with_z) {
    const unsigned long long n = P.n, m = P.m, nx = P.nx;
    unsigned long long f = 2ull * n * 2 * nx + 2ull * m * 4 * nx;                 // z_u, row values and bounds
    f += 2ull * m * (unsigned long long)c.rows + 2ull * (unsigned long long)c.sq; // steps and certifications (as_certify adds its passes to rows / sq)
    f += 2ull * (with_z ? (unsigned long long)P.nz : (unsigned long long)((P.N + 1) * P.nu)) * (n + nx);
    return f;
}

// the same count for `solves` solves whose counters were summed into c (as_flops is linear in them)
__device__ __forceinline__ unsigned long long as_flops_total(const QPDev& P, const ASCounters& c, unsigned solves) {
    const unsigned long long n = P.n, m = P.m, nx = P.nx;
    unsigned long long f = (unsigned long long)solves * (2ull * n * 2 * nx + 2ull * m * 4 * nx + 2ull * (unsigned long long)((P.N + 1) * P.nu) * (n + nx));
    f += 2ull * m * (unsigned long long)c.rows + 2ull * (unsigned long long)c.sq;
    return f;
}

// Warp reductions.  Maxima / arg-maxima go through the integer reduction unit (REDUX) on an
// order-preserving 64-bit key, two 32-bit halves: a handful of instructions instead of five shuffle stages.
__device__ __forceinline__ unsigned long long as_key(double v) {
    const long long b = __double_as_longlong(v);
    return (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double as_unkey(unsigned long long k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}
// (inlined at every call site: as real calls they cost 5 % - 80.1 against 84.6 M solves/s - although the code shrinks)
#ifdef RTMPC_AS_RED_CALLS
#define AS_RED static __device__ __noinline__
#else
#define AS_RED __device__ __forceinline__
#endif
AS_RED double as_wmax(double v) {
    const unsigned long long k = as_key(v);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(RTMPC_FULL_MASK, hi);
    const unsigned ml = __reduce_max_sync(RTMPC_FULL_MASK, hi == mh ? lo : 0u);
    return as_unkey(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ double as_wmin(double v) { return -as_wmax(-v); }
// lane holding the maximum (lowest lane among equals)
__device__ __forceinline__ int as_lane_max(double v) {
    const unsigned long long k = as_key(v);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(RTMPC_FULL_MASK, hi);
    const unsigned ml = __reduce_max_sync(RTMPC_FULL_MASK, hi == mh ? lo : 0u);
    return __ffs(__ballot_sync(RTMPC_FULL_MASK, hi == mh && lo == ml)) - 1;
}
struct ASArg { double v; int idx; };
AS_RED ASArg as_wargmax(double v, int idx) {
    const int src = as_lane_max(v);
    ASArg r;
    r.v = __shfl_sync(RTMPC_FULL_MASK, v, src);
    r.idx = __shfl_sync(RTMPC_FULL_MASK, idx, src);
    return r;
}
__device__ __forceinline__ ASArg as_wargmin(double v, int idx) {
    ASArg r = as_wargmax(-v, idx);
    r.v = -r.v;
    return r;
}
AS_RED double as_wsum(double v) { return warp_sum(v); }

// slot state of one lane: working-set slot `lane` holds signed row (ra, sa) with multiplier lam
struct ASSlot { int ra; double sa, lam; };

// out_a = sum_b M[a][b] vec[b] over the slots below hi (free slots hold zeros)
// (the helpers below are real calls; they address the per-warp block through offsets into the dynamic shared memory
//  array so that their loads and stores are shared-memory instructions, not generic ones)
AS_FN_MATVEC double as_matvec(int Mo, int ms, int hi, int lane, int vo) {
    extern __shared__ __align__(16) double as_smem[];
    double s0 = 0.0, s1 = 0.0;
    if (lane < hi) {
        const double* row = as_smem + Mo + lane * ms;
        const double* vec = as_smem + vo;
        // four columns per trip (npad is a multiple of 4; columns and entries beyond hi hold zeros)
#pragma unroll 1
        for (int b = 0; b < hi; b += 4) {
            const double2 m0 = ld2(row + b), v0 = ld2(vec + b), m1 = ld2(row + b + 2), v1 = ld2(vec + b + 2);
            s0 = fma(m0.x, v0.x, s0);
            s1 = fma(m0.y, v0.y, s1);
            s0 = fma(m1.x, v1.x, s0);
            s1 = fma(m1.y, v1.y, s1);
        }
    }
    return s0 + s1;
}

// M grows by slot s:  [[M + r r'/kappa, -r/kappa], [-r'/kappa, 1/kappa]]   (r in rv, zero on free slots;
// hi = even number of slots covering every occupied one and s)
// (nr = rows of M: hi is a multiple of 4 and may reach past them; those slots are never occupied)
AS_FN_BORDER void as_border(int Mo, int rvo, int ms, int hi, int nr, int lane, int s, double r_own, double kappa) {
    extern __shared__ __align__(16) double as_smem[];
    double* M = as_smem + Mo;
    const double* rv = as_smem + rvo;
    const double ik = __drcp_rn(kappa);
    if (lane < hi && lane < nr) {
        double* row = M + lane * ms;
        if (lane == s) {
#pragma unroll 1
            for (int b = 0; b < hi; b += 4) {
                const double2 r0 = ld2(rv + b), r1 = ld2(rv + b + 2);
                *reinterpret_cast<double2*>(row + b) = make_double2(-r0.x * ik, -r0.y * ik);
                *reinterpret_cast<double2*>(row + b + 2) = make_double2(-r1.x * ik, -r1.y * ik);
            }
            row[s] = ik;
        } else {
            const double c = r_own * ik;
            if (c != 0.0) {
#pragma unroll 1
                for (int b = 0; b < hi; b += 4) {
                    const double2 r0 = ld2(rv + b), r1 = ld2(rv + b + 2);
                    double2 m0 = ld2(row + b), m1 = ld2(row + b + 2);
                    m0.x = fma(c, r0.x, m0.x); m0.y = fma(c, r0.y, m0.y);
                    m1.x = fma(c, r1.x, m1.x); m1.y = fma(c, r1.y, m1.y);
                    *reinterpret_cast<double2*>(row + b) = m0;
                    *reinterpret_cast<double2*>(row + b + 2) = m1;
                }
            }
            row[s] = -c;
        }
    }
    __syncwarp();
}

// slot j leaves:  M <- M - m_j m_j' / M_jj, row and column j cleared (tmp: npad doubles of scratch)
AS_FN_DOWNDATE void as_downdate(int Mo, int tmpo, int ms, int hi, int nr, int lane, int j) {
    extern __shared__ __align__(16) double as_smem[];
    double* M = as_smem + Mo;
    double* tmp = as_smem + tmpo;
    if (lane < hi) tmp[lane] = M[j * ms + lane];      // (columns between the highest slot and hi hold zeros)
    __syncwarp();
    const double ij = __drcp_rn(tmp[j]);
    if (lane < hi && lane < nr) {
        double* row = M + lane * ms;
        const double c = (lane == j) ? -1.0 : -row[j] * ij;      // row j: m_j - m_j = 0
        if (c != 0.0) {
#pragma unroll 1
            for (int b = 0; b < hi; b += 4) {
                const double2 v0 = ld2(tmp + b), v1 = ld2(tmp + b + 2);
                double2 m0 = ld2(row + b), m1 = ld2(row + b + 2);
                m0.x = (lane == j) ? 0.0 : fma(c, v0.x, m0.x); m0.y = (lane == j) ? 0.0 : fma(c, v0.y, m0.y);
                m1.x = (lane == j) ? 0.0 : fma(c, v1.x, m1.x); m1.y = (lane == j) ? 0.0 : fma(c, v1.y, m1.y);
                *reinterpret_cast<double2*>(row + b) = m0;
                *reinterpret_cast<double2*>(row + b + 2) = m1;
            }
        }
        row[j] = 0.0;
    }
    __syncwarp();
}

// In-place Gauss-Jordan inverse of the nc x nc matrix in M (lane i owns row i; hi = even cover of nc).  A pivot that
// has lost its size relative to the original diagonal (diag0) marks a row that depends on the rows before it: it is
// skipped, and its row and column are cleared at the end, which leaves the inverse over the remaining rows.
// Returns the mask of skipped rows.
AS_FN_INVERT unsigned as_invert(int Mo, int ms, int nc, int hi, int lane, double diag0) {
    extern __shared__ __align__(16) double as_smem[];
    double* M = as_smem + Mo;
    unsigned dead = 0;
#pragma unroll 1
    for (int k = 0; k < nc; ++k) {
        const double pk = M[k * ms + k];
        const double dk = __shfl_sync(RTMPC_FULL_MASK, diag0, k);
        if (!(pk > 1e-11 * dk)) { dead |= 1u << k; continue; }
        const double ip = __drcp_rn(pk);
        if (lane == k) {
            double* row = M + k * ms;
            row[k] = 1.0;
#pragma unroll 1
            for (int b = 0; b < hi; b += 4) {
                double2 m0 = ld2(row + b), m1 = ld2(row + b + 2);
                m0.x *= ip; m0.y *= ip; m1.x *= ip; m1.y *= ip;
                *reinterpret_cast<double2*>(row + b) = m0;
                *reinterpret_cast<double2*>(row + b + 2) = m1;
            }
        }
        __syncwarp();
        if (lane < nc && lane != k) {
            double* row = M + lane * ms;
            const double* rk = M + k * ms;
            const double f = row[k];
            row[k] = 0.0;
            if (f != 0.0) {
#pragma unroll 1
                for (int b = 0; b < hi; b += 4) {
                    const double2 k0 = ld2(rk + b), k1 = ld2(rk + b + 2);
                    double2 m0 = ld2(row + b), m1 = ld2(row + b + 2);
                    m0.x = fma(-f, k0.x, m0.x); m0.y = fma(-f, k0.y, m0.y);
                    m1.x = fma(-f, k1.x, m1.x); m1.y = fma(-f, k1.y, m1.y);
                    *reinterpret_cast<double2*>(row + b) = m0;
                    *reinterpret_cast<double2*>(row + b + 2) = m1;
                }
            }
        }
        __syncwarp();
    }
    if (dead && lane < nc) {
        double* row = M + lane * ms;
#pragma unroll 1
        for (int b = 0; b < nc; ++b)
            if (((dead >> b) & 1u) || ((dead >> lane) & 1u)) row[b] = 0.0;
    }
    __syncwarp();
    return dead;
}

// flag / unflag row `row` as a member of the working set in its owner lane's bit masks
__device__ __forceinline__ void as_mark(int lane, int row, int sgn, bool on, unsigned& actu, unsigned& actl) {
    if (lane == ((row & 63) >> 1)) {
        const unsigned bit = 1u << (((row >> 6) << 1) | (row & 1));
        if (sgn > 0) actu = on ? (actu | bit) : (actu & ~bit);
        else actl = on ? (actl | bit) : (actl & ~bit);
    }
}

__device__ __forceinline__ int as_hi(unsigned amask) { return (35 - __clz(amask | 1u)) & ~3; }   // multiple of 4, covers amask

// M and the slot lists of one warp back to "empty working set"
template <class D>
__device__ __forceinline__ void as_clear(ASWarp& w, const QPDev& P, int lane) {
    const int npad = D::npad(P), ms = D::npad(P) + 2;
    if (lane < npad) {
        if (lane < D::n(P)) {
            double* row = w.M() + lane * ms;
#pragma unroll 1
            for (int b = 0; b < npad; b += 2) *reinterpret_cast<double2*>(row + b) = make_double2(0.0, 0.0);
        }
        w.act_row()[lane] = 0;
        w.act_sgn()[lane] = 0;
    }
    __syncwarp();
}

__device__ __forceinline__ void as_list_put(const ASWarp& w, int k, double coef, int off) {
    double* q = w.lst() + 2 * k;
    q[0] = coef;
    reinterpret_cast<int*>(q + 1)[0] = off;
}

// Goldfarb-Idnani iteration.  `apply_only`: the warm start left its multipliers in w.coef(); the first pass
// only moves the row values (the row streaming code exists once).  Returns 0 when no row is violated by
// more than tolp.
// ILP = 2: two rows of W / three columns of G' per pass (more loads in flight, more registers)
template <int R2, int ILP, class D>
__device__ __forceinline__ int as_gi(const QPDev& P, ASWarp& w, unsigned& amask, ASSlot& sl, int lane,
                                     double (&e)[2 * R2], unsigned& actu, unsigned& actl, int max_steps,
                                     bool apply_only, ASCounters& cnt) {
    const int npad = D::npad(P), n = D::n(P), ms = D::npad(P) + 2;
    constexpr int mpad = 64 * R2;          // (= P.mpad: rtmpc_qp_create pads the rows to the instantiation's 64 * R2)
    const unsigned slots = (1u << n) - 1u;          // n <= 30
    while (true) {
        int p = 0;
        double sp = 1.0, cp = 0.0, lam_p = 0.0, wpp = 1.0;
        if (!apply_only) {
            double best = -RTMPC_INF;
            int code = 0;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const int row = (r2 << 6) + 2 * lane;
                const double2 wd = ld2(P.wid + row);
                const double vu0 = e[2 * r2], vl0 = -e[2 * r2] - wd.x, vu1 = e[2 * r2 + 1], vl1 = -e[2 * r2 + 1] - wd.y;
                if (!((actu >> (2 * r2)) & 1u) && vu0 > best) { best = vu0; code = 2 * row; }
                if (!((actl >> (2 * r2)) & 1u) && vl0 > best) { best = vl0; code = 2 * row + 1; }
                if (!((actu >> (2 * r2 + 1)) & 1u) && vu1 > best) { best = vu1; code = 2 * row + 2; }
                if (!((actl >> (2 * r2 + 1)) & 1u) && vl1 > best) { best = vl1; code = 2 * row + 3; }
            }
            const ASArg bm = as_wargmax(best, code);
            if (bm.v <= w.ctl()[0]) return 0;
            p = bm.idx >> 1;
            sp = (bm.idx & 1) ? -1.0 : 1.0;
            cp = bm.v;
            wpp = P.W[(size_t)p * mpad + p];
        }
        const double* __restrict__ Wp = P.W + (size_t)p * mpad;
        while (true) {
            const bool occ = (amask >> lane) & 1u;
            const int na = __popc(amask);
            const int hi = as_hi(amask);
            double c = 0.0, rr = 0.0, kappa = 1.0, step = 0.0;
            bool full = false, dependent = false;
            int j1 = 0;
            if (!apply_only) {
                if (cnt.steps >= max_steps) { w.ictl()[2] = 1; return RTMPC_FALLBACK; }
                cnt.steps += 1;
                const double v = occ ? sl.sa * sp * Wp[sl.ra] : 0.0;
                if (lane < npad) w.v()[lane] = v;
                __syncwarp();
                rr = occ ? as_matvec(w.Mo(), ms, hi, lane, w.vo()) : 0.0;
                kappa = wpp - as_wsum(v * rr);
                dependent = !(kappa > 1e-11 * wpp) || na >= n;      // n independent rows already span everything
                const double rmax = as_wmax(fabs(rr));
                const double ratio = (occ && rr > 1e-13 * (1.0 + rmax)) ? sl.lam * __drcp_rn(rr) : RTMPC_INF;
                const ASArg rm = as_wargmin(ratio, lane);
                const double t1 = rm.v;
                j1 = rm.idx;
                const bool has_j = t1 < 0.5 * RTMPC_INF;
                if (dependent) {
                    // n_p is a combination of the active rows: without a blocking multiplier the
                    // constraints contradict each other (Farkas: y = (-r, 1) >= 0, N y = 0, b'y = -c_p < 0)
                    if (!has_j) {
                        w.ictl()[2] = (cp > 1e-6 * P.sc_b) ? 6 : 2;
                        w.ctl()[1] = cp;
                        return (cp > 1e-6 * P.sc_b) ? RTMPC_INFEASIBLE : AS_STALL;
                    }
                    step = t1;
                } else {
                    const double t2 = cp * __drcp_rn(kappa);
                    full = !(has_j && t1 < t2);
                    step = full ? t2 : t1;
                }
                if (occ) sl.lam = (!full && lane == j1) ? 0.0 : fma(-step, rr, sl.lam);
                lam_p += step;
                // the row values move by  step (|1| W_p + sum_a |r_a| W_a)  at most; a step along a dependent row moves the
                // multipliers but not the row values, which then no longer follow the multipliers exactly
                if (lane == 0) w.ctl()[2] = dependent ? 1e300 : fma(step, 1.0 + (double)na * rmax, w.ctl()[2]);
                cnt.sq += na * na;
                c = -step * sp;
                // rows of this pass: the entering row, then the working set's in slot order (na < n here whenever the pass
                // runs: at most n <= npad entries)
                // (every lane is past the mat-vec that read v: the reductions behind it are warp-wide)
                if (occ) as_list_put(w, 1 + __popc(amask & ((1u << lane) - 1u)), step * sl.sa * rr, sl.ra * mpad);
                if (lane == 0) as_list_put(w, 0, c, p * mpad);
                __syncwarp();
            }
            if (!dependent) {
                // row values move by  c W[p][:] + sum_a coef_a W[row_a][:]  (the list; the warm start's pass has no entering row)
                const int nl = apply_only ? na : na + 1;
                const double* __restrict__ lq = w.lst();
                if (ILP >= 2) {
#pragma unroll 1
                    for (int k = 0; k < nl; k += 2) {
                        const double ca = lq[2 * k];
                        const double* __restrict__ Wa = P.W + (unsigned)(reinterpret_cast<const int*>(lq + 2 * k + 1)[0] + 2 * lane);
                        const bool two = k + 1 < nl;
                        const double cb = two ? lq[2 * k + 2] : 0.0;
                        const double* __restrict__ Wb = two ? P.W + (unsigned)(reinterpret_cast<const int*>(lq + 2 * k + 3)[0] + 2 * lane) : Wa;
#pragma unroll
                        for (int r2 = 0; r2 < R2; ++r2) {
                            const double2 g = ld2_hint<RTMPC_HINT_W>(Wa + r2 * 64), h = ld2_hint<RTMPC_HINT_W>(Wb + r2 * 64);
                            e[2 * r2] = fma(cb, h.x, fma(ca, g.x, e[2 * r2]));
                            e[2 * r2 + 1] = fma(cb, h.y, fma(ca, g.y, e[2 * r2 + 1]));
                        }
                    }
                } else {
#pragma unroll 1
                    for (int k = 0; k < nl; ++k) {
                        const double ca = lq[2 * k];
                        const double* __restrict__ Wa = P.W + (unsigned)(reinterpret_cast<const int*>(lq + 2 * k + 1)[0] + 2 * lane);
#pragma unroll
                        for (int r2 = 0; r2 < R2; ++r2) {
                            const double2 g = ld2_hint<RTMPC_HINT_W>(Wa + r2 * 64);
                            e[2 * r2] = fma(ca, g.x, e[2 * r2]);
                            e[2 * r2 + 1] = fma(ca, g.y, e[2 * r2 + 1]);
                        }
                    }
                }
                cp = fma(-step, kappa, cp);
                cnt.rows += nl;
            }
            if (apply_only) { apply_only = false; break; }
            if (full) {
                AS_DBG(7, 1);
                const unsigned freem = ~amask & slots;
                if (na >= n || !freem) { w.ictl()[2] = 3; return RTMPC_FALLBACK; }
                const int s = __ffs(freem) - 1;
                __syncwarp();                 // rv shares coef's storage: every lane is done streaming
                if (lane < npad) w.rv()[lane] = rr;
                __syncwarp();
                as_border(w.Mo(), w.rvo(), ms, as_hi(amask | (1u << s)), n, lane, s, rr, kappa);
                if (lane == s) { sl.ra = p; sl.sa = sp; sl.lam = lam_p; w.act_row()[s] = p; w.act_sgn()[s] = (int)sp; }
                as_mark(lane, p, (int)sp, true, actu, actl);
                amask |= 1u << s;
                cnt.sq += na * na;
                __syncwarp();
                break;
            }
            // partial step: the blocking row leaves the working set
            AS_DBG(8, 1);
            as_mark(lane, w.act_row()[j1], w.act_sgn()[j1], false, actu, actl);
            as_downdate(w.Mo(), w.vo(), ms, hi, n, lane, j1);
            amask &= ~(1u << j1);
            cnt.sq += na * na;
        }
    }
}

// Certification on the working set.  Returns 0 when the KKT conditions hold, 1 when a row is still
// violated (t holds exact values: go back to as_gi), 2 on a negative multiplier / no convergence.
template <int R2, int ILP, class D>
__device__ __forceinline__ int as_certify(const QPDev& P, ASWarp& w, unsigned amask, ASSlot& sl, int lane,
                                          double (&e)[2 * R2], unsigned actu, unsigned actl, ASCounters& cnt) {
    const int n = D::n(P), npad = D::npad(P), nx = D::nx(P), ms = D::npad(P) + 2;
    constexpr int mpad = 64 * R2;          // (= P.mpad)
    const bool occ = (amask >> lane) & 1u;
    const int hi = as_hi(amask);
    double ba = 0.0;
    if (occ) {
        // bound of the signed row:  s > 0: up,  s < 0: -lo
        const double* tab = (sl.sa > 0) ? P.UxT : P.LxT;
        double b = (sl.sa > 0) ? P.upI[sl.ra] : P.loI[sl.ra];
#pragma unroll 1
        for (int k = 0; k < nx; ++k) b = fma(tab[(size_t)k * mpad + sl.ra], w.xr()[k], b);
        ba = (sl.sa > 0) ? b : -b;
    }
    const double* __restrict__ Ga = P.G + (size_t)(occ ? sl.ra : 0) * npad;
    double lam = 0.0, resid = 0.0;
    double zj = (lane < n) ? w.zu()[lane] : 0.0;
    // pass 0 starts from the multipliers the Goldfarb-Idnani steps arrived at (usually already within the certificate's
    // tolerance of the true rows: one pass instead of two), passes 1..3 refine with M as approximate inverse
    bool refined = false;                            // multipliers moved away from the ones the row values were built with
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        double dl;
        refined = pass > 0;
        if (pass == 0) dl = occ ? sl.lam : 0.0;
        else {
            if (lane < npad) w.v()[lane] = resid;
            __syncwarp();
            dl = occ ? as_matvec(w.Mo(), ms, hi, lane, w.vo()) : 0.0;
        }
        lam += dl;
        // z -= Y_A' (s dl): the rows of Y as a packed list (coefficient, offset), two per trip
        __syncwarp();                                 // (the list lives in v's storage: every lane is past the mat-vec)
        if (occ) as_list_put(w, __popc(amask & ((1u << lane) - 1u)), dl * sl.sa, sl.ra * npad);
        __syncwarp();
        if (lane < n) {
            double z2 = 0.0;
            const int nl = __popc(amask);
            const double* __restrict__ lq = w.lst();
            const double* __restrict__ Yl = P.Y + lane;
#pragma unroll 1
            for (int k = 0; k < nl; k += 2) {
                const bool two = k + 1 < nl;
                const double ca = lq[2 * k], cb = two ? lq[2 * k + 2] : 0.0;
                const int oa = reinterpret_cast<const int*>(lq + 2 * k + 1)[0];
                const int ob = two ? reinterpret_cast<const int*>(lq + 2 * k + 3)[0] : oa;
                zj = fma(-ca, Yl[oa], zj);
                z2 = fma(-cb, Yl[ob], z2);
            }
            zj += z2;
        }
        if (lane < npad) w.z()[lane] = zj;
        __syncwarp();
        resid = 0.0;
        if (occ) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 2
            for (int k = 0; k < npad; k += 2) {
                const double2 g = ld2(Ga + k), zz = ld2(w.z() + k);
                a0 = fma(g.x, zz.x, a0);
                a1 = fma(g.y, zz.y, a1);
            }
            resid = sl.sa * (a0 + a1) - ba;
        }
        // converged (every active row on its bound to well below the certificate's tolerance): stop refining
        if (!__any_sync(RTMPC_FULL_MASK, fabs(resid) > 1e-3 * w.ctl()[0])) break;      // (a vote: only the threshold matters)
    }
    AS_DBG(11, 1);
    cnt.rounds += 1;
    cnt.sq += 6 * n * n;                              // refinement
    const double tolp = w.ctl()[0];
    const double lmaxabs = as_wmax(occ ? fabs(lam) : 0.0);
    const int na = __popc(amask);
    // Row values at z, in three tiers.  Every evaluation of a row differs from (G z - up)_i by rounding only, and the
    // rounding is bounded:  kap_i S f,  kap from rtmpc_qp_create (rounding of one evaluation and of the tables, in long
    // double), S = 1 + |x|_1 + |ref|_1 + the magnitudes of the coefficients that went into the value, f = 1 + (passes over
    // the rows since the value was started from scratch) / 32.  A row outside the working set that clears the tolerance by
    // its bound is certified without G z being formed.
    //   tier 0: the values the Goldfarb-Idnani steps arrived at, e = e_u - sum of every step's W rows (nothing is read).
    //           Only when the refinement left the multipliers where they were (one pass) and no step moved multipliers
    //           without moving rows (ctl()[2] < 1e300).
    //   tier 1: through the factored tables from scratch, e = Ex x + Tr r - up0 - W[:,A] (s lam): 2 nx columns and |A|
    //           rows the solve has just read.
    //   tier 2 (below): from G' z, n + nx columns streamed from L2 - when a row is within its bound of the tolerance, or
    //           really violated; the decision is then taken on those values.
    bool exact = true, violated = false;
    if (P.kap) {                                      // (NULL: RTMPC_TUNE_CERT_FACTORED 0)
        double S0 = 1.0;
#pragma unroll 1
        for (int k = 0; k < nx; ++k) S0 += fabs(w.xr()[k]) + fabs(w.xr()[8 + k]);
        const double hist = w.ctl()[2];
        const int since = cnt.rows - w.ictl()[4];
#pragma unroll 1
        for (int tier = (refined || !(hist < 1e299)) ? 1 : 0; tier < 2; ++tier) {
            double S;
            if (tier == 0) {
                AS_DBG(12, 1);
                S = (S0 + hist) * (1.0 + (double)since * 0.03125) * (1.0 + 1e-9);
            } else {
                AS_DBG(13, 1);
                cnt.rows += 2 * nx + na;
                S = (S0 + (double)na * lmaxabs) * (1.0 + 1e-9);
#pragma unroll
                for (int r2 = 0; r2 < R2; ++r2) {
                    const double2 uu = ld2(P.upI + r2 * 64 + 2 * lane);
                    e[2 * r2] = -uu.x;
                    e[2 * r2 + 1] = -uu.y;
                }
#pragma unroll 1
                for (int k = 0; k < nx; ++k) {
                    const double xk = w.xr()[k], rk = w.xr()[8 + k];
                    const size_t o = (size_t)k * mpad + 2 * lane;
#pragma unroll
                    for (int r2 = 0; r2 < R2; ++r2) {
                        const double2 a = ld2_hint<RTMPC_HINT_SETUP>(P.ExT + o + r2 * 64), b = ld2_hint<RTMPC_HINT_SETUP>(P.TrT + o + r2 * 64);
                        e[2 * r2] = fma(b.x, rk, fma(a.x, xk, e[2 * r2]));
                        e[2 * r2 + 1] = fma(b.y, rk, fma(a.y, xk, e[2 * r2 + 1]));
                    }
                }
                __syncwarp();
                if (lane < npad) w.coef()[lane] = occ ? -sl.sa * lam : 0.0;
                __syncwarp();
#pragma unroll 1
                for (unsigned mk = amask; mk; mk &= mk - 1) {
                    const int a = __ffs(mk) - 1;
                    const double ca = w.coef()[a];
                    const double* __restrict__ Wa = P.W + (unsigned)(w.act_row()[a] * mpad + 2 * lane);
#pragma unroll
                    for (int r2 = 0; r2 < R2; ++r2) {
                        const double2 g = ld2_hint<RTMPC_HINT_W>(Wa + r2 * 64);
                        e[2 * r2] = fma(ca, g.x, e[2 * r2]);
                        e[2 * r2 + 1] = fma(ca, g.y, e[2 * r2 + 1]);
                    }
                }
            }
            // (negated comparisons: a NaN counts as not cleared)
            bool close = false;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const double2 wd = ld2(P.wid + r2 * 64 + 2 * lane), kp = ld2(P.kap + r2 * 64 + 2 * lane);
                const double m0 = fma(kp.x, S, -tolp), m1 = fma(kp.y, S, -tolp);      // e + kap S <= tolp  <=>  e + m <= 0
                close = close || (!((actu >> (2 * r2)) & 1u) && !(e[2 * r2] + m0 <= 0.0)) ||
                        (!((actl >> (2 * r2)) & 1u) && !(-e[2 * r2] - wd.x + m0 <= 0.0)) ||
                        (!((actu >> (2 * r2 + 1)) & 1u) && !(e[2 * r2 + 1] + m1 <= 0.0)) ||
                        (!((actl >> (2 * r2 + 1)) & 1u) && !(-e[2 * r2 + 1] - wd.y + m1 <= 0.0));
            }
            exact = __any_sync(RTMPC_FULL_MASK, close);
            if (!exact) break;
        }
    }
    if (exact) {
    AS_DBG(10, 1);
    cnt.rows += n + nx;
    // the rows are about to hold fresh values again, built with `lam`; multipliers clamped below (tiny negative ones) leave them
    __syncwarp();
    if (lane == 0) {
        w.ctl()[2] = (double)na * lmaxabs;
        w.ictl()[4] = cnt.rows;
    }
    if (__any_sync(RTMPC_FULL_MASK, occ && lam < 0.0) && lane == 0) w.ctl()[2] = 1e300;
    // exact row values at z:  e = G z - up through the transposed copies (coalesced 16-byte loads)
#pragma unroll
    for (int r2 = 0; r2 < R2; ++r2) {
        const double2 uu = ld2(P.upI + r2 * 64 + 2 * lane);
        e[2 * r2] = -uu.x;
        e[2 * r2 + 1] = -uu.y;
    }
#pragma unroll 1
    for (int k = 0; k < nx; ++k) {
        const double xk = w.xr()[k];
        const double* __restrict__ u = P.UxT + (size_t)k * mpad + 2 * lane;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) {
            const double2 c = ld2_hint<RTMPC_HINT_UX>(u + r2 * 64);
            e[2 * r2] = fma(-c.x, xk, e[2 * r2]);
            e[2 * r2 + 1] = fma(-c.y, xk, e[2 * r2 + 1]);
        }
    }
    // (z is zero beyond n, G' is zero-padded)
    if (ILP >= 2) {
        // three columns per pass keep 3*R2 loads in flight
#pragma unroll 1
        for (int k = 0; k < n; k += 3) {
            const int k1 = (k + 1 < n) ? k + 1 : k, k2 = (k + 2 < n) ? k + 2 : k;
            const double z0 = w.z()[k], z1 = (k + 1 < n) ? w.z()[k1] : 0.0, z2 = (k + 2 < n) ? w.z()[k2] : 0.0;
            const double* __restrict__ g0 = P.GT + (size_t)k * mpad + 2 * lane;
            const double* __restrict__ g1 = P.GT + (size_t)k1 * mpad + 2 * lane;
            const double* __restrict__ g2 = P.GT + (size_t)k2 * mpad + 2 * lane;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const double2 a = ld2_stream(g0 + r2 * 64), b = ld2_stream(g1 + r2 * 64), c = ld2_stream(g2 + r2 * 64);
                e[2 * r2] = fma(c.x, z2, fma(b.x, z1, fma(a.x, z0, e[2 * r2])));
                e[2 * r2 + 1] = fma(c.y, z2, fma(b.y, z1, fma(a.y, z0, e[2 * r2 + 1])));
            }
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const double z0 = w.z()[k];
            const double* __restrict__ g0 = P.GT + (size_t)k * mpad + 2 * lane;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const double2 a = ld2(g0 + r2 * 64);
                e[2 * r2] = fma(a.x, z0, e[2 * r2]);
                e[2 * r2 + 1] = fma(a.y, z0, e[2 * r2 + 1]);
            }
        }
    }
    // any row outside the working set violated by more than the tolerance?  (not after the factored values cleared it)
    bool viol = false;
#pragma unroll
    for (int r2 = 0; r2 < R2; ++r2) {
        const double2 wd = ld2(P.wid + r2 * 64 + 2 * lane);
        viol = viol || (!((actu >> (2 * r2)) & 1u) && e[2 * r2] > tolp) || (!((actl >> (2 * r2)) & 1u) && e[2 * r2] + wd.x < -tolp) ||
               (!((actu >> (2 * r2 + 1)) & 1u) && e[2 * r2 + 1] > tolp) ||
               (!((actl >> (2 * r2 + 1)) & 1u) && e[2 * r2 + 1] + wd.y < -tolp);
    }
    violated = __any_sync(RTMPC_FULL_MASK, viol);
    }   // (exact rows)
    const bool bad = (fabs(resid) > tolp) ||                      // refinement did not converge
                     (occ && lam < -1e-9 * (1.0 + lmaxabs));     // a negative multiplier
    if (occ) sl.lam = fmax(lam, 0.0);
    if (__any_sync(RTMPC_FULL_MASK, bad)) return 2;
    return violated ? 1 : 0;
}

// One instance, one warp.  x_init / ref: this instance's parameters (ref may be NULL); warm_inst: this
// instance's warm-start record (npad + 1 ints) or NULL; z_out_inst (its first z_rows entries are written) /
// U_out_inst: this instance's outputs or NULL.  Returns the status; on RTMPC_FALLBACK nothing has been written.
// (no __restrict__ on the instance pointers: the rollout kernel writes them from the same warp)
// carry (rollout kernel only; NULL elsewhere): two ints that outlive the call, [0] = tag of the problem whose solve left
// its certified working set and that set's inverse in w (0: nothing usable), [1] = the set's slot mask; carry_tag (> 0)
// names P.  With a carry the warm start is the previous control step's working set AS IT IS, not moved one stage
// earlier: M depends on the rows of the working set only, not on x_init, so the carried inverse is exact for it and the
// solve starts with one mat-vec (multipliers at the new x_init) instead of |A| pivots.  Measured on the closed loops of
// the benchmarks (profiles/r2_rollout_ncu.md) the unmoved set is also the better guess: 2.0 rank-one changes per
// constrained solve against 4.1 after moving the set (plus the 3.8 changes the move itself takes on a carried inverse).
// The warm-start record is then used unmoved as well (first step of a ticket).  Results agree with the step-by-step
// path to rounding, not bit for bit: the certificate is the same, but the certification refines the multipliers with M
// as an approximate inverse, so the last bits follow M's history.
template <int R2, int ILP, class D = ASDimsDyn>
__device__ __forceinline__ int as_solve_instance(const QPDev& P, ASWarp& w, int lane, const double* x_init,
                                                 const double* ref, int* warm_inst, double* z_out_inst, int z_rows,
                                                 double* U_out_inst, ASCounters& cnt, int* carry = nullptr,
                                                 int carry_tag = 0) {
    const int n = D::n(P), npad = D::npad(P), nx = D::nx(P), ms = D::npad(P) + 2;
    constexpr int mpad = 64 * R2;          // (= P.mpad)
    const double tolp = 1e-11 * P.sc_b;
    if (lane < nx) {
        w.xr()[lane] = x_init[lane];
        w.xr()[8 + lane] = ref ? ref[lane] : 0.0;
    }
    // the previous solve of this instance was for this problem and left (M, slots) behind
    bool carried = carry && carry[0] == carry_tag;
    const unsigned carried_mask = carried ? (unsigned)carry[1] : 0u;
    bool m_clean = carried && carried_mask == 0u;       // M is known to hold zeros only
    __syncwarp();
    if (carry && lane == 0) carry[0] = 0;                // (set again when this solve ends with a certified working set)
    bool par_bad = false;
#pragma unroll 1
    for (int i = lane; i < P.np; i += 32) {
        double acc = -P.parh[i];
#pragma unroll 1
        for (int k = 0; k < nx; ++k) acc = fma(P.parC[i * nx + k], w.xr()[k], acc);
        if (acc > 1e-9 * (1.0 + fabs(P.parh[i]))) par_bad = true;
    }
    par_bad = __any_sync(RTMPC_FULL_MASK, par_bad);
    double zuj = 0.0;
    if (lane < n) {
#pragma unroll 1
        for (int k = 0; k < nx; ++k) {
            zuj = fma(P.Zx[lane * nx + k], w.xr()[k], zuj);
            zuj = fma(P.Zr[lane * nx + k], w.xr()[8 + k], zuj);
        }
    }
    if (lane < npad) { w.zu()[lane] = zuj; w.z()[lane] = zuj; }
    // e = G z_u - up for every row (rows are measured from their upper bound; the lower side is e + wid >= 0)
    double e[2 * R2];
    int status = RTMPC_OPTIMAL;
    unsigned amask = 0, actu = 0, actl = 0;
    ASSlot sl;
    sl.ra = 0; sl.sa = 0.0; sl.lam = 0.0;
    const int max_steps = P.as_max_steps;
    w.ctl()[0] = tolp;
    w.ictl()[2] = 0;
    w.ictl()[3] = 0;           // refactorisations so far
    bool carry_started = false, cold = false;
    // A long sequence of bordering / downdating steps on a nearly dependent working set lets the explicit inverse
    // drift (seen as a wrong sign of kappa, a failed refinement, a contradiction that is none).  The cure is the warm
    // start's own procedure: start over from z_u with the CURRENT working set as the candidate list, inverted from
    // scratch.  `restart` counts those refactorisations.
#pragma unroll 1
    for (;;) {
        const bool first = w.ictl()[3] == 0;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) {
            const double2 uu = ld2(P.upI + r2 * 64 + 2 * lane);
            e[2 * r2] = -uu.x;
            e[2 * r2 + 1] = -uu.y;
        }
#pragma unroll 1
        for (int k = 0; k < nx; ++k) {
            const double xk = w.xr()[k], rk = w.xr()[8 + k];
            const size_t o = (size_t)k * mpad + 2 * lane;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const double2 a = ld2_hint<RTMPC_HINT_SETUP>(P.ExT + o + r2 * 64), b = ld2_hint<RTMPC_HINT_SETUP>(P.TrT + o + r2 * 64);
                e[2 * r2] = fma(b.x, rk, fma(a.x, xk, e[2 * r2]));
                e[2 * r2 + 1] = fma(b.y, rk, fma(a.y, xk, e[2 * r2 + 1]));
            }
        }
        // (the row values start from scratch here: history of applied coefficients for as_certify)
        if (lane == 0) { w.ctl()[2] = 0.0; w.ictl()[4] = cnt.rows; }
        int nc = 0;
        if (first) {
            // is any row on or beyond a bound?  (only the sign matters here: compares and a vote, no FP64 max chain)
            bool touched = false;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const double2 wd = ld2(P.wid + r2 * 64 + 2 * lane);
                touched = touched || e[2 * r2] >= 0.0 || e[2 * r2] + wd.x <= 0.0 || e[2 * r2 + 1] >= 0.0 || e[2 * r2 + 1] + wd.y <= 0.0;
            }
            const bool feasible_u = !__any_sync(RTMPC_FULL_MASK, touched);
            __syncwarp();
            if (par_bad) { status = RTMPC_INFEASIBLE; break; }
            if (feasible_u) {                             // the unconstrained minimiser is feasible
                if (carry && !m_clean) as_clear<D>(w, P, lane);      // the empty working set is carried with an empty M
                break;
            }
        }
        // ---- 1a. carried working set and inverse: nothing to build ---------------------------------------------
        bool moved = false;
        AS_DBG(0, 1);
        if (first && carried) {
            AS_DBG(1, 1);
            moved = true;
            amask = carried_mask;
            if ((amask >> lane) & 1u) { sl.ra = w.act_row()[lane]; sl.sa = (double)w.act_sgn()[lane]; }
        }
        const bool on_carried = moved;
        carry_started = carry_started || moved;
        carried = false;
        m_clean = false;
        // M starts empty
        if (!moved) as_clear<D>(w, P, lane);
        // ---- 1. candidates: the previous step's working set moved one stage (warm start), or the current one ----
        if (!moved) {
            int prow = -1;
            double psg = 1.0;
            if (first) {
                const int wn = (warm_inst && (P.shift || carry) && !cold) ? warm_inst[0] : 0;
                if (lane < wn && lane < n) {
                    const int code = warm_inst[1 + lane];
                    const int row0 = code >> 1;
                    psg = (code & 1) ? -1.0 : 1.0;
                    if (row0 >= 0 && row0 < mpad) prow = carry ? row0 : P.shift[row0];
                    // kept if the row has that bound
                    if (prow >= 0 && ((psg > 0) ? !(P.upI[prow] < 0.5 * RTMPC_INF) : !(P.loI[prow] > -0.5 * RTMPC_INF))) prow = -1;
                }
            } else if ((amask >> lane) & 1u) {
                prow = sl.ra;
                psg = sl.sa;
            }
            const unsigned okm = __ballot_sync(RTMPC_FULL_MASK, prow >= 0);
            nc = __popc(okm);
            if (prow >= 0) {
                const int pos = __popc(okm & ((1u << lane) - 1u));
                w.act_row()[pos] = prow;
                w.act_sgn()[pos] = (int)psg;
            }
            amask = 0; actu = 0; actl = 0;
            sl.ra = 0; sl.sa = 0.0; sl.lam = 0.0;
            __syncwarp();
        }
        actu = 0; actl = 0;
        bool apply = false;
        if (nc > 0 || (moved && amask)) {
            if (!moved) {
                // candidate c sits in slot c: S = signed sub-matrix of W, inverted in place
                const int hi = (nc + 3) & ~3;
                double diag0 = 1.0;
                if (lane < nc) {
                    sl.ra = w.act_row()[lane];
                    sl.sa = (double)w.act_sgn()[lane];
                    const double* __restrict__ Wa = P.W + (size_t)sl.ra * mpad;
                    double* row = w.M() + lane * ms;
#pragma unroll 1
                    for (int b = 0; b < nc; ++b) row[b] = Wa[w.act_row()[b]] * sl.sa * (double)w.act_sgn()[b];
                    diag0 = Wa[sl.ra];
                }
                __syncwarp();
                AS_DBG(5, 1); AS_DBG(6, nc);
                const unsigned dead = as_invert(w.Mo(), ms, nc, hi, lane, diag0);
                amask = (((nc >= 32) ? 0xffffffffu : ((1u << nc) - 1u))) & ~dead;
                cnt.sq += nc * nc * nc / 2;
            }
            // multipliers of the equality-constrained problem; drop negative ones, most negative first
            double rhs = 0.0;
            if ((amask >> lane) & 1u) {
                double ev = -P.upI[sl.ra];
#pragma unroll 1
                for (int k = 0; k < nx; ++k) {
                    ev = fma(P.ExT[(size_t)k * mpad + sl.ra], w.xr()[k], ev);
                    ev = fma(P.TrT[(size_t)k * mpad + sl.ra], w.xr()[8 + k], ev);
                }
                rhs = (sl.sa > 0) ? ev : -(ev + P.wid[sl.ra]);   // s (t_u - up)  or  -(t_u - lo)
            }
#pragma unroll 1
            while (amask) {
                const bool occ = (amask >> lane) & 1u;
                const int na = __popc(amask);
                const int hi = as_hi(amask);
                if (lane < npad) w.v()[lane] = occ ? rhs : 0.0;
                __syncwarp();
                const double lamv = occ ? as_matvec(w.Mo(), ms, hi, lane, w.vo()) : 0.0;
                const ASArg lm = as_wargmin(occ ? lamv : RTMPC_INF, lane);
                cnt.sq += na * na;
                cnt.steps += 1;
                // (the size of the multipliers only matters when the smallest one is negative)
                if (!(lm.v < 0.0) || !(lm.v < -1e-9 * (1.0 + as_wmax(fabs(lamv))))) {
                    sl.lam = occ ? fmax(lamv, 0.0) : 0.0;
                    break;
                }
                AS_DBG(9, 1);
                as_downdate(w.Mo(), w.vo(), ms, hi, n, lane, lm.idx);
                amask &= ~(1u << lm.idx);
                cnt.sq += na * na;
            }
            if (amask) {
                const bool occ = (amask >> lane) & 1u;
                if (occ) as_list_put(w, __popc(amask & ((1u << lane) - 1u)), -sl.sa * sl.lam, sl.ra * mpad);      // as_gi's first pass
                const double lw = as_wmax(occ ? sl.lam : 0.0);
                if (lane == 0) w.ctl()[2] = (double)__popc(amask) * lw;
                __syncwarp();
#pragma unroll 1
                for (unsigned mk = amask; mk; mk &= mk - 1) {
                    const int a = __ffs(mk) - 1;
                    as_mark(lane, w.act_row()[a], w.act_sgn()[a], true, actu, actl);
                }
                apply = true;         // as_gi moves the row values first
            }
        }
        // ---- 2./3. Goldfarb-Idnani, then certification -----------------------------------------
        // The row values e move with every step by (multiplier change) x (row of W): with the huge multipliers of a
        // nearly empty feasible set their rounding error can exceed tolp.  A stall (AS_STALL) is therefore first
        // answered by certification, which recomputes e exactly; a stall on exact values means the rows contradict
        // each other by `stall_cp`: up to 1e-8 (relative, the reference solver's own feasibility tolerance) the
        // solve carries on with that much slack and reports RTMPC_OPTIMAL_INACCURATE.
        w.ictl()[0] = on_carried ? -(1 << 20) : cnt.steps;      // (a contradiction found on a carried inverse is never believed)
        w.ictl()[1] = -1;
#pragma unroll 1
        for (int refresh = 0;; ++refresh) {
            status = as_gi<R2, ILP, D>(P, w, amask, sl, lane, e, actu, actl, max_steps, apply, cnt);
            apply = false;
            if (status == AS_STALL) {
                status = RTMPC_FALLBACK;
                if (cnt.steps == w.ictl()[1] + 1) {
                    const double stall_cp = w.ctl()[1];
                    if (stall_cp > 1e-8 * P.sc_b || refresh >= 12) break;
                    w.ctl()[0] = 2.0 * stall_cp;
                    w.ictl()[1] = cnt.steps;
                    continue;
                }
            } else if (status != 0) break;
            const int c = as_certify<R2, ILP, D>(P, w, amask, sl, lane, e, actu, actl, cnt);
            if (c == 0) { status = (w.ctl()[0] > tolp) ? RTMPC_OPTIMAL_INACCURATE : RTMPC_OPTIMAL; break; }
            if (c == 2 || refresh >= 12) { status = RTMPC_FALLBACK; w.ictl()[2] = (c == 2) ? 4 : 5; break; }
            w.ictl()[1] = cnt.steps;
        }
        if (status == RTMPC_OPTIMAL || status == RTMPC_OPTIMAL_INACCURATE) break;
        // a contradiction found a few steps after a fresh factorisation is believed; anything else starts over
        if (status == RTMPC_INFEASIBLE && cnt.steps - w.ictl()[0] <= 8) break;
        if (cnt.steps >= max_steps || w.ictl()[3] >= 4) {
            // out of steps or refactorisations.  A solve that started on a carried working set gets one more attempt, from
            // the empty set with a fresh budget, before it is handed over: a carried start is then never worse than a cold one
            if (!carry_started || cold) { if (status != RTMPC_INFEASIBLE || cnt.steps < max_steps) status = RTMPC_FALLBACK; break; }
            cold = true;
            cnt.steps = 0;
            __syncwarp();
            if (lane == 0) { w.ictl()[3] = 0; w.ctl()[0] = tolp; }
            __syncwarp();
            continue;
        }
        if (lane == 0) w.ictl()[3] += 1;
        __syncwarp();
    }

    // ---- outputs ------------------------------------------------------------------------------
    if (status != RTMPC_FALLBACK) {
        const bool has_sol = (status == RTMPC_OPTIMAL || status == RTMPC_OPTIMAL_INACCURATE);
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        const int nu = D::nu(P), N = D::N(P);
        const int nrow = (N + 1) * nu;
        if (U_out_inst) {
            // packet payload straight from the scaled decision: U = UPhi z + UPsi x_init, last column u_bar + K x_bar
            if (P.Uidx) {
                // every payload row is one scaled decision variable
#pragma unroll 1
                for (int i = lane; i < nrow; i += 32) {
                    if (i >= N * nu && P.nss == 0) continue;
                    U_out_inst[i] = has_sol ? P.Ucoef[i] * w.z()[P.Uidx[i]] : nanv;
                }
            } else {
#pragma unroll 1
                for (int i = lane; i < nrow; i += 32) {
                    if (i >= N * nu && P.nss == 0) continue;
                    double acc = 0.0, acc2 = 0.0;
                    const double* __restrict__ up = P.UPhiT + i;          // zero-padded to npad columns
                    const double* zz = w.z();
    #pragma unroll 2
                    for (int k = 0; k < npad; k += 2) {
                        acc = fma(up[0], zz[k], acc);
                        acc2 = fma(up[nrow], zz[k + 1], acc2);
                        up += 2 * nrow;
                    }
                    const double* __restrict__ ps = P.UPsiT + i;
    #pragma unroll 1
                    for (int k = 0; k < nx; ++k) { acc = fma(ps[0], w.xr()[k], acc); ps += nrow; }
                    U_out_inst[i] = has_sol ? acc + acc2 : nanv;
                }
            }
        }
        if (z_out_inst) {
            if (lane < npad) w.coef()[lane] = (lane < n) ? w.z()[lane] * P.D[lane] : 0.0;   // unscaled decision
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < z_rows; i += 32) {      // leading rows of [x_0..x_N | u | x_bar | u_bar]
                double acc = 0.0;
#pragma unroll 1
                for (int k = 0; k < n; ++k) acc = fma(P.Phi[(size_t)i * npad + k], w.coef()[k], acc);
#pragma unroll 1
                for (int k = 0; k < nx; ++k) acc = fma(P.Psi[(size_t)i * nx + k], w.xr()[k], acc);
                z_out_inst[i] = has_sol ? acc : nanv;
            }
        }
    }
    AS_DBG(16 + __popc(amask), 1);
    if (carry && lane == 0) {
        // (M, slots) of a certified working set stay in w for the next control step of this instance
        carry[1] = (int)amask;
        carry[0] = (status == RTMPC_OPTIMAL || status == RTMPC_OPTIMAL_INACCURATE) ? carry_tag : 0;
    }
    if (warm_inst) {
        // certified working set, compacted (the slots are sparse)
        const bool solved = (status == RTMPC_OPTIMAL || status == RTMPC_OPTIMAL_INACCURATE);
        const bool occ = solved && ((amask >> lane) & 1u);
        const unsigned om = __ballot_sync(RTMPC_FULL_MASK, occ);
        const int pos = __popc(om & ((1u << lane) - 1u));
        if (lane == 0) warm_inst[0] = solved ? __popc(om) : -1;
        if (occ) warm_inst[1 + pos] = 2 * sl.ra + (sl.sa < 0 ? 1 : 0);
    }
    __syncwarp();
    return status;
}

__device__ __forceinline__ int as_pack_iters(const ASCounters& cnt, const ASWarp& w) {
    return ((cnt.steps > 4095 ? 4095 : cnt.steps) << 12) | ((cnt.rounds > 15 ? 15 : cnt.rounds) << 24) | ((w.ictl()[2] & 7) << 28);
}

template <int R2, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
as_solve_kernel(QPDev P, int B, const double* __restrict__ x_init, const double* __restrict__ ref,
                const int* __restrict__ sel, int sel_value, double* __restrict__ z_out, double* __restrict__ U_out,
                int* __restrict__ status_out, int* __restrict__ iters_out, int* __restrict__ warm,
                unsigned long long* __restrict__ work) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nx = P.nx;
    ASWarp w = as_carve(smem + (size_t)warp * as_warp_doubles(P), P);
    const size_t usz = (size_t)(P.N + 1) * P.nu;
#pragma unroll 1
    // consecutive instances go to different SMs, so a last partial round is spread over all of them
    for (int inst = warp * gridDim.x + blockIdx.x; inst < B; inst += gridDim.x * wpb) {
        if (sel && sel[inst] != sel_value) continue;
        ASCounters cnt;
        cnt.steps = 0; cnt.rounds = 0; cnt.rows = 0; cnt.sq = 0;
        const int status = as_solve_instance<R2, (MAXW * R2 <= 100) ? 2 : 1>(P, w, lane, x_init + (size_t)inst * nx,
                                                 ref ? ref + (size_t)inst * nx : nullptr,
                                                 warm ? warm + (size_t)inst * (P.npad + 1) : nullptr,
                                                 z_out ? z_out + (size_t)inst * P.nz : nullptr, P.nz,
                                                 U_out ? U_out + inst * usz : nullptr, cnt);
        if (lane == 0) {
            if (status_out) status_out[inst] = status;
            if (iters_out) iters_out[inst] = as_pack_iters(cnt, w);
            if (work) atomicAdd(work, as_flops(P, cnt, z_out != nullptr));
        }
        __syncwarp();
    }
}

}  // namespace rtmpc

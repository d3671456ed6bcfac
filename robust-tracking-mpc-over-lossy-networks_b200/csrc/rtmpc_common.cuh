// Shared device helpers for the rtmpc_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RTMPC_FULL_MASK 0xffffffffu
#define RTMPC_INF 1e30

namespace rtmpc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RTMPC_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(RTMPC_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(RTMPC_FULL_MASK, v, o));
    return v;
}
// arg-max over the warp: returns the max value, idx receives the index belonging to it
__device__ __forceinline__ double warp_argmax(double v, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(RTMPC_FULL_MASK, v, o);
        int oi = __shfl_xor_sync(RTMPC_FULL_MASK, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return v;
}
__device__ __forceinline__ double warp_argmin(double v, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(RTMPC_FULL_MASK, v, o);
        int oi = __shfl_xor_sync(RTMPC_FULL_MASK, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return v;
}

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, so draws depend only on
// (seed, instance id, time step) and are invariant to the number of GPUs / the launch shape. ----
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifndef RTMPC_PHILOX_UNROLL
#define RTMPC_PHILOX_UNROLL 1      // rolled rounds: instruction footprint of the rollout kernel
#endif
    constexpr int kPhiloxUnroll = RTMPC_PHILOX_UNROLL;
#pragma unroll kPhiloxUnroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 53-bit uniform in [0,1) from two 32-bit words
__host__ __device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo) {
    uint64_t v = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)v * (1.0 / 9007199254740992.0);
}

}  // namespace rtmpc

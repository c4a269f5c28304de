// Shared helpers for the mgnns_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/mgnns_b200.h"

namespace mgnns {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MG_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            ::mgnns::set_error(__VA_ARGS__);          \
            return 1;                                 \
        }                                             \
    } while (0)

#define MG_LAUNCH_CHECK(name)                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::mgnns::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return 2;                                                                \
        }                                                                            \
        ::mgnns::count_launch();                                                     \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline bool aligned16(const T* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Counter-based uniform in [0,1): splitmix64 finaliser of (seed, index).  Used
// for the in-kernel dropout masks so that backward can regenerate them.
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == MGNNS_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == MGNNS_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace mgnns

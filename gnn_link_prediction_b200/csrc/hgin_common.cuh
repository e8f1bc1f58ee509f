// Shared host/device helpers for libhgin (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "hgin.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhgin targets sm_100a (B200) only"
#endif

namespace hgin {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kWarp = 32;

// ---- error reporting across the C ABI (thread-local message, no exceptions) ------------------
char *error_buffer();  // defined in hgin_api.cu
inline int32_t fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define HGIN_CHECK_ARG(cond, ...)                                              \
    do {                                                                       \
        if (!(cond)) return ::hgin::fail(HGIN_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
    } while (0)

#define HGIN_CHECK_LAUNCH(name)                                                                  \
    do {                                                                                         \
        cudaError_t err__ = cudaGetLastError();                                                  \
        if (err__ != cudaSuccess)                                                                \
            return ::hgin::fail(HGIN_ERR_CUDA, "%s: %s", name, cudaGetErrorString(err__));       \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid for a grid-stride kernel: enough CTAs for `work_items` at `per_cta`, capped at a multiple
// of the SM count so the launch is whole waves.
inline int grid_for(int64_t work_items, int64_t per_cta, int ctas_per_sm) {
    int64_t want = ceil_div(work_items, per_cta);
    int64_t cap = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
    if (want < 1) want = 1;
    return static_cast<int>(want < cap ? want : cap);
}

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float *smem32) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem32[warp] = v;
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nwarps) ? smem32[threadIdx.x] : 0.f;
    if (warp == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

__device__ __forceinline__ float act_forward(float z, int act, float alpha) {
    if (act == HGIN_ACT_PRELU) return z > 0.f ? z : alpha * z;   // at::prelu: x > 0 ? x : w*x
    if (act == HGIN_ACT_RELU) return z > 0.f ? z : 0.f;
    return z;
}
__device__ __forceinline__ float act_backward(float g, float z, int act, float alpha) {
    if (act == HGIN_ACT_PRELU) return z > 0.f ? g : alpha * g;
    if (act == HGIN_ACT_RELU) return z > 0.f ? g : 0.f;
    return g;
}

}  // namespace hgin

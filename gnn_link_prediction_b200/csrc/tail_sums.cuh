// Sums-only pass over dz for the layers whose backward consumes dz in place but still needs column sums that the
// weight-gradient MMA does not produce: the rank-k2 tail dW[:, k1:k1+k2] = dz^T x2 of the readout's first layer
// (K = 128 + 3 raw input columns, models.py:366) and db when it cannot ride on the MMA.
//   part[cta][nn][k2 + 1]: columns [0, k2) = sum_m dz[m][nn] * x2[m][t], column k2 = sum_m dz[m][nn]
// A pure streaming read of dz: every thread owns 16 bytes of a row (4 floats / 8 bf16) and requests RIF rows before
// it consumes the first (two CTAs of 256 threads x 8 rows x 16 B = 64 KB of reads in flight per SM; the earlier
// one-size-fits-all dz_prepare pass kept 16-32 KB in flight and ran at 1-2 TB/s).
#pragma once

#include <cuda_bf16.h>

#include "hgin_common.cuh"

namespace hgin {
namespace tcgemm {

template <typename T>
struct TailCols;
template <>
struct TailCols<float> {
    static constexpr int value = 4;
    using Raw = float4;
    static __device__ __forceinline__ Raw load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void unpack(const Raw &t, float (&o)[4]) { o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
};
template <>
struct TailCols<__nv_bfloat16> {
    static constexpr int value = 8;
    using Raw = uint4;
    static __device__ __forceinline__ Raw load(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void unpack(const Raw &q, float (&o)[8]) {
        o[0] = __uint_as_float(q.x << 16); o[1] = __uint_as_float(q.x & 0xffff0000u);
        o[2] = __uint_as_float(q.y << 16); o[3] = __uint_as_float(q.y & 0xffff0000u);
        o[4] = __uint_as_float(q.z << 16); o[5] = __uint_as_float(q.z & 0xffff0000u);
        o[6] = __uint_as_float(q.w << 16); o[7] = __uint_as_float(q.w & 0xffff0000u);
    }
};

constexpr int TAIL_THREADS = 256;
inline int tail_ctas() { return kNumSMs * 2; }
template <typename T>
inline size_t tail_smem(int n) { return static_cast<size_t>(TAIL_THREADS / (n / TailCols<T>::value)) * n * 5 * sizeof(float); }

template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS, 2)
tail_sums_kernel(int64_t rows, int n, const T *__restrict__ g, int64_t ldg, const float *__restrict__ x2, int64_t ld2,
                 int k2, float *__restrict__ part) {
    constexpr int C = TailCols<T>::value;
    constexpr int RIF = C == 8 ? 5 : 8;   // (bf16 lanes carry 40 accumulators: fewer rows fit 128 registers)
    extern __shared__ float sm[];   // [slots][n][5]
    const int tpr = n / C;
    const int slots = TAIL_THREADS / tpr;
    const int slot = threadIdx.x / tpr;
    const int cg = threadIdx.x % tpr;
    float tail[C][4], db[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        db[i] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) tail[i][t] = 0.f;
    }
    if (slot < slots) {
        const int64_t stride = static_cast<int64_t>(gridDim.x) * slots;
        for (int64_t m0 = static_cast<int64_t>(blockIdx.x) * slots + slot; m0 < rows; m0 += stride * RIF) {
            typename TailCols<T>::Raw raw[RIF];      // kept packed (4 registers) until consumed
            float xq[RIF][4];
#pragma unroll
            for (int u = 0; u < RIF; ++u) {
                const int64_t m = m0 + u * stride;
                raw[u] = typename TailCols<T>::Raw{};   // all-zero bits == 0.0 in both types
#pragma unroll
                for (int t = 0; t < 4; ++t) xq[u][t] = 0.f;
                if (m < rows) {
                    raw[u] = TailCols<T>::load(g + m * ldg + cg * C);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t < k2) xq[u][t] = __ldg(x2 + m * ld2 + t);
                }
            }
#pragma unroll
            for (int u = 0; u < RIF; ++u) {   // rows past the end were zero-filled: they add nothing
                float d[C];
                TailCols<T>::unpack(raw[u], d);
#pragma unroll
                for (int i = 0; i < C; ++i) {
                    db[i] += d[i];
#pragma unroll
                    for (int t = 0; t < 4; ++t) tail[i][t] = fmaf(d[i], xq[u][t], tail[i][t]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
            float *dst = sm + (static_cast<int64_t>(slot) * n + cg * C + i) * 5;
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = tail[i][t];
            dst[4] = db[i];
        }
    }
    __syncthreads();
    const int kp = k2 + 1;
    for (int i = threadIdx.x; i < n * kp; i += TAIL_THREADS) {   // row slots combined in a fixed order
        const int nn = i / kp, t = i % kp;
        const int src_t = (t == k2) ? 4 : t;
        float s = 0.f;
        for (int sl = 0; sl < slots; ++sl) s += sm[(static_cast<int64_t>(sl) * n + nn) * 5 + src_t];
        part[(static_cast<int64_t>(blockIdx.x) * n + nn) * kp + t] = s;
    }
}

}  // namespace tcgemm
}  // namespace hgin

// Register-tiled fp32 SIMT GEMM engine shared by the dense-layer kernels (HGIN_MATH_FP32).
//
// C[BM x BN] += A[BM x k] * B[BN x k]^T over a contraction range, 256 threads per CTA, each thread
// a TM x TN micro-tile kept in registers, operands staged through shared memory in BK = 16 slices
// with register prefetch of the next slice (global loads in flight during the FMAs).
// Operand values come from accessor functors so that one engine serves
//   forward      z  = [x1|x2] W^T            (A = activations,          B = W rows)
//   input grad   dx = dz W                   (A = g * act'(z) on load,  B = W columns)
//   weight grad  dW = dz^T [x1|x2|1]         (A = dz columns,           B = x columns, split over rows)
// This is the parity path (fp32 FMA, rel 1e-5 vs the CPU reference); the tensor-core path for
// the 128-wide layers lives in linear_tc.cu.
#pragma once

#include "hgin_common.cuh"

namespace hgin {
namespace simt {

constexpr int BK = 16;
constexpr int THREADS = 256;

template <int BM_, int BN_, int TM_, int TN_>
struct Tile {
    static constexpr int BM = BM_, BN = BN_, TM = TM_, TN = TN_;
    static constexpr int TX = BN / TN;  // threads along the N side
    static constexpr int TY = BM / TM;  // threads along the M side
    static_assert(TX * TY == THREADS, "tile must use 256 threads");
    static constexpr int VM = TM < 4 ? TM : 4;  // contiguous run per thread along M
    static constexpr int VN = TN < 4 ? TN : 4;
    static constexpr int GM = TM / VM;          // runs per thread
    static constexpr int GN = TN / VN;
    static constexpr int LDA = BM + 4;          // smem row pitch (keeps 16B alignment of runs)
    static constexpr int LDB = BN + 4;
    static constexpr int A_ITERS = (BM * BK + THREADS - 1) / THREADS;
    static constexpr int B_ITERS = (BN * BK + THREADS - 1) / THREADS;
    static constexpr int SMEM_FLOATS = BK * LDA + BK * LDB;

    // local row / column index of micro-tile element i / j for thread (ty, tx)
    __device__ static __forceinline__ int row_of(int ty, int i) { return (i / VM) * (BM / GM) + ty * VM + (i % VM); }
    __device__ static __forceinline__ int col_of(int tx, int j) { return (j / VN) * (BN / GN) + tx * VN + (j % VN); }
};

// One operand slice (EXT x BK) global -> registers.  K_CONTIG: consecutive threads walk the
// contraction index (operand stored [ext][k]); otherwise they walk the tile extent (stored [k][ext]).
template <int EXT, int ITERS, bool K_CONTIG, class F>
__device__ __forceinline__ void fetch_slice(float (&r)[ITERS], const F &f, int64_t ext0, int64_t k0) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int idx = threadIdx.x + it * THREADS;
        int e, kk;
        if (K_CONTIG) { kk = idx % BK; e = idx / BK; } else { e = idx % EXT; kk = idx / EXT; }
        r[it] = (idx < EXT * BK) ? f(ext0 + e, k0 + kk) : 0.0f;
    }
}

template <int EXT, int LD, int ITERS, bool K_CONTIG>
__device__ __forceinline__ void stash_slice(float *smem, const float (&r)[ITERS]) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int idx = threadIdx.x + it * THREADS;
        int e, kk;
        if (K_CONTIG) { kk = idx % BK; e = idx / BK; } else { e = idx % EXT; kk = idx / EXT; }
        if (idx < EXT * BK) smem[kk * LD + e] = r[it];
    }
}

// acc += A(m0.., k) * B(n0.., k) for k in [kbeg, kend).  fa(m, k) / fb(n, k) return 0 out of range.
// `smem` holds Tile::SMEM_FLOATS floats.  All 256 threads must call (contains __syncthreads).
template <class T, bool A_K_CONTIG, bool B_K_CONTIG, class FA, class FB>
__device__ __forceinline__ void mainloop(float (&acc)[T::TM][T::TN], const FA &fa, const FB &fb, int64_t m0,
                                         int64_t n0, int64_t kbeg, int64_t kend, float *smem) {
    float *As = smem;
    float *Bs = smem + BK * T::LDA;
    const int tx = threadIdx.x % T::TX;
    const int ty = threadIdx.x / T::TX;
    float ra[T::A_ITERS], rb[T::B_ITERS];

    fetch_slice<T::BM, T::A_ITERS, A_K_CONTIG>(ra, fa, m0, kbeg);
    fetch_slice<T::BN, T::B_ITERS, B_K_CONTIG>(rb, fb, n0, kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        __syncthreads();  // previous slice fully consumed
        stash_slice<T::BM, T::LDA, T::A_ITERS, A_K_CONTIG>(As, ra);
        stash_slice<T::BN, T::LDB, T::B_ITERS, B_K_CONTIG>(Bs, rb);
        __syncthreads();
        if (k0 + BK < kend) {  // prefetch the next slice while this one is multiplied
            fetch_slice<T::BM, T::A_ITERS, A_K_CONTIG>(ra, fa, m0, k0 + BK);
            fetch_slice<T::BN, T::B_ITERS, B_K_CONTIG>(rb, fb, n0, k0 + BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[T::TM], b[T::TN];
#pragma unroll
            for (int g = 0; g < T::GM; ++g) {
                const float *p = As + kk * T::LDA + g * (T::BM / T::GM) + ty * T::VM;
                if constexpr (T::VM == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(p);
                    a[g * 4 + 0] = t.x; a[g * 4 + 1] = t.y; a[g * 4 + 2] = t.z; a[g * 4 + 3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < T::VM; ++i) a[g * T::VM + i] = p[i];
                }
            }
#pragma unroll
            for (int g = 0; g < T::GN; ++g) {
                const float *p = Bs + kk * T::LDB + g * (T::BN / T::GN) + tx * T::VN;
                if constexpr (T::VN == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(p);
                    b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
                } else {
#pragma unroll
                    for (int j = 0; j < T::VN; ++j) b[g * T::VN + j] = p[j];
                }
            }
#pragma unroll
            for (int i = 0; i < T::TM; ++i)
#pragma unroll
                for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
}

}  // namespace simt
}  // namespace hgin

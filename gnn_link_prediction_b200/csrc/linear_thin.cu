// Dense layers with a tiny contraction (K <= 8): layer 0 of the GIN stack, where the MLP input
// is [agg(3) | (1+eps) x_dst(3)] (models.py:287-290), and every GIN layer at config.json's
// NODE_EMBEDDING_SIZE = 8.  These are pure streaming kernels (24-32 B in, up to 1 KB out per row),
// so the tiled GEMM engine only wastes issue slots on them: here each lane owns 4 output columns,
// keeps its W slice in registers and walks rows, reading g/z and writing z/out as coalesced
// 128-bit vectors.  fp32 FMA throughout (this path serves both math modes).
//
// Backward computes, in ONE pass over g and z:
//   dW[n][k] = sum_m dz[m][n] h[m][k],  db[n] = sum_m dz[m][n],  dalpha = sum g * min(z, 0),
//   T[n][c]  = sum_m dz[m][n] dot_x[m][c]  ->  ddot = sum_{n,c} W[n][c0 + c] T[n][c]
// (the last equals sum_m sum_c (dz W)[m][c0+c] * dot_x[m][c] = d(eps) without materialising dz W).
// Reductions are deterministic: lanes -> CTA (shared memory, fixed order) -> partials -> final pass.
#include <cuda_bf16.h>

#include "hgin_common.cuh"
#include "linear_thin_api.h"

namespace hgin {
namespace thin {
namespace {

using bf16 = __nv_bfloat16;

// The WIDE side of these layers (z / out / g of the K <= 8 layers, x / dx of the head) is stored as float or bf16
// (HGIN_DTYPE_*); the narrow side (K <= 8 inputs, the single head column) and all arithmetic stay fp32.
// ldg4 / ld4 / st4: four consecutive elements as one 16-byte (float) or 8-byte (bf16) access.
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ldg4(const bf16 *p) {
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld4(const bf16 *p) {
    const uint2 q = *reinterpret_cast<const uint2 *>(p);
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
}
__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(bf16 *p, float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
}

constexpr int THREADS = 256;
constexpr int KMAX = 8;
constexpr int DMAX = 4;                      // max dot_x columns
constexpr int NACC = 4 * KMAX + 4 * DMAX + 4 + 1;  // per-thread accumulators: dW, T, db, dalpha = 53
constexpr int FWD_BLK = 256;                       // rows per CTA iteration of thin_fwd (>= the most row slots a CTA has)

template <typename TW>
__global__ void __launch_bounds__(THREADS)
thin_fwd_kernel(int64_t rows, const float *__restrict__ x, int64_t ldx, int k, const float *__restrict__ W,
                const float *__restrict__ bias, int n, int act, const float *__restrict__ alpha_ptr,
                TW *__restrict__ z, int64_t ldz, TW *__restrict__ out, int64_t ldo, int accumulate_out) {
    const int lpr = n >> 2;                        // lanes per row
    const int cg = threadIdx.x % lpr;              // column group: columns 4cg .. 4cg+3
    const int grp = threadIdx.x / lpr;             // row slot inside the CTA
    const int slots_per_cta = THREADS / lpr;
    float w[4][KMAX], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        b[i] = bias ? __ldg(bias + cg * 4 + i) : 0.0f;
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) w[i][kk] = kk < k ? __ldg(W + static_cast<int64_t>(cg * 4 + i) * k + kk) : 0.0f;
    }
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    // A write-mostly kernel whose only reads are the narrow input rows (24-32 bytes each, streamed from
    // DRAM): with the rows loaded by the threads that use them, every row costs a DRAM latency and the
    // stores trickle (3.2 TB/s measured).  Each CTA therefore walks contiguous blocks of FWD_BLK rows whose
    // inputs are staged into shared memory with cp.async one block AHEAD (double buffer, no registers),
    // so the compute loop only reads shared memory and issues stores.
    constexpr int XP = KMAX + 1;                       // padded pitch: row slots fall on distinct banks
    __shared__ float xs[2][FWD_BLK * XP];
    const int rpt = FWD_BLK / slots_per_cta;           // rows per thread and block (slots_per_cta <= 256 divides FWD_BLK)
    auto stage = [&](int buf, int64_t base) {
        for (int idx = threadIdx.x; idx < FWD_BLK * k; idx += THREADS) {
            const int r = idx / k, kk = idx % k;
            float *dst = &xs[buf][r * XP + kk];
            if (base + r < rows) {
                const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(x + (base + r) * ldx + kk) : "memory");
            } else {
                *dst = 0.0f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t step = static_cast<int64_t>(gridDim.x) * FWD_BLK;
    int64_t base = static_cast<int64_t>(blockIdx.x) * FWD_BLK;
    int buf = 0;
    if (base < rows) stage(0, base);
    for (; base < rows; base += step, buf ^= 1) {
        if (base + step < rows) {
            stage(buf ^ 1, base + step);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        for (int u = 0; u < rpt; ++u) {
            const int r = u * slots_per_cta + grp;
            const int64_t m = base + r;
            if (m >= rows) break;
            float xv[KMAX];
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? xs[buf][r * XP + kk] : 0.0f;
            float zz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.0f;
#pragma unroll
                for (int kk = 0; kk < KMAX; ++kk) s = fmaf(xv[kk], w[i][kk], s);
                zz[i] = s + b[i];
            }
            if (z) st4(z + m * ldz + cg * 4, zz[0], zz[1], zz[2], zz[3]);
            if (out) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = act_forward(zz[i], act, alpha);
                TW *po = out + m * ldo + cg * 4;
                if (accumulate_out) {
                    const float4 old = ld4(po);
                    o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
                }
                st4(po, o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();     // everyone is done with xs[buf] before the next iteration stages into it
    }
}

// CTA combine of the per-thread accumulators, 8 at a time: sm[thread][8] -> sum over row slots in slot
// order -> part[cta][n][kp]; dalpha -> alpha_part[cta].
__device__ __forceinline__ void thin_bwd_combine(const float (&acc)[NACC], float *sm, float *red, int lpr,
                                                 int slots_per_cta, int k, int d, int n, float *__restrict__ part,
                                                 float *__restrict__ alpha_part) {
    const int kp = k + d + 1;
    float *dst = part + static_cast<int64_t>(blockIdx.x) * n * kp;
#pragma unroll
    for (int round = 0; round < 7; ++round) {   // rounds 0-3: dW of column i; 4-5: T; 6: db (+ dalpha)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = round * 8 + j;
            sm[threadIdx.x * 8 + j] = a < NACC - 1 ? acc[a] : 0.0f;
        }
        __syncthreads();
        // one thread per (column group, j): sums the row slots
        for (int idx = threadIdx.x; idx < lpr * 8; idx += THREADS) {
            const int c_g = idx / 8, j = idx % 8;
            float s = 0.0f;
            for (int sl = 0; sl < slots_per_cta; ++sl) s += sm[(sl * lpr + c_g) * 8 + j];
            const int a = round * 8 + j;
            if (a < 32) {                       // dW: i = a / 8, kk = a % 8
                const int i = a >> 3, kk = a & 7;
                if (kk < k) dst[static_cast<int64_t>(c_g * 4 + i) * kp + kk] = s;
            } else if (a < 48) {                // T: i = (a-32)/4, c = (a-32)%4
                const int i = (a - 32) >> 2, c = (a - 32) & 3;
                if (c < d) dst[static_cast<int64_t>(c_g * 4 + i) * kp + k + c] = s;
            } else if (a < 52) {                // db
                dst[static_cast<int64_t>(c_g * 4 + (a - 48)) * kp + k + d] = s;
            }
        }
    }
    __syncthreads();
    const float da = block_sum(acc[52], red);
    if (threadIdx.x == 0 && alpha_part) alpha_part[blockIdx.x] = da;
}

// part[cta][n][kp], kp = k + d + 1: columns [0,k) = dW, [k, k+d) = T, k+d = db;  alpha_part[cta].
template <bool HAS_ACT, typename TW>
__global__ void __launch_bounds__(THREADS, 2)
thin_bwd_kernel(int64_t rows, const TW *__restrict__ g, int64_t ldg, const TW *__restrict__ z, int64_t ldz,
                int act, const float *__restrict__ alpha_ptr, const float *__restrict__ x, int64_t ldx, int k,
                const float *__restrict__ dot_x, int64_t ld_dot, int d, int n, float *__restrict__ part,
                float *__restrict__ alpha_part) {
    __shared__ float sm[THREADS * 8];
    __shared__ float red[32];
    const int lpr = n >> 2;
    const int cg = threadIdx.x % lpr;
    const int grp = threadIdx.x / lpr;             // row slot inside the CTA
    const int slots_per_cta = THREADS / lpr;
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
    // acc layout: [i*KMAX + kk] dW, [32 + i*DMAX + c] T, [48 + i] db, [52] dalpha
    // Each CTA iteration takes a CONTIGUOUS block of slots_per_cta * RIF rows: every thread requests the
    // g / z vectors of its RIF rows up front (with 128 registers only two CTAs fit an SM, and one row per
    // thread left ~16 KB of reads outstanding per SM), and the block's narrow x / dot_x rows (24 + 12
    // bytes per row, streamed from DRAM) are staged through shared memory with one coalesced load
    // instead of 12 latency-exposed scalar loads per row and thread.
    // (with ACT_NONE — g already is dz — only g is read, so twice as many rows fit the same registers)
    constexpr int RIF = HAS_ACT ? 4 : 8;
    constexpr int XP = KMAX + 1, DP = DMAX + 1;          // padded pitches: row slots fall on distinct banks
    __shared__ float xs[64 * 8 * XP];                    // up to 64 slots (n = 16) x 8 rows
    __shared__ float ds[64 * 8 * DP];
    const int rif = RIF < 512 / slots_per_cta ? RIF : 512 / slots_per_cta;   // tiny n: many slots, fewer rows each
    const int blk = slots_per_cta * rif;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * blk; base < rows; base += static_cast<int64_t>(gridDim.x) * blk) {
        float4 gq[RIF], zq[HAS_ACT ? RIF : 1];
#pragma unroll
        for (int u = 0; u < RIF; ++u) {
            const int64_t m = base + u * slots_per_cta + grp;
            gq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_ACT) zq[u] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (u < rif && m < rows) {
                gq[u] = ldg4(g + m * ldg + cg * 4);
                if (HAS_ACT) zq[u] = ldg4(z + m * ldz + cg * 4);
            }
        }
        __syncthreads();                                 // the previous block's tiles have been consumed
        for (int idx = threadIdx.x; idx < blk * k; idx += THREADS) {
            const int r = idx / k, kk = idx % k;
            xs[r * XP + kk] = (base + r < rows) ? __ldg(x + (base + r) * ldx + kk) : 0.0f;
        }
        for (int idx = threadIdx.x; idx < blk * d; idx += THREADS) {
            const int r = idx / d, c = idx % d;
            ds[r * DP + c] = (base + r < rows) ? __ldg(dot_x + (base + r) * ld_dot + c) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < RIF; ++u) {
            const int r = u * slots_per_cta + grp;
            if (u >= rif || base + r >= rows) break;
            float dz[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
            if (HAS_ACT) {
                const float zv[4] = {zq[u].x, zq[u].y, zq[u].z, zq[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (act == HGIN_ACT_PRELU && !(zv[i] > 0.0f)) acc[52] += dz[i] * zv[i];
                    dz[i] = act_backward(dz[i], zv[i], act, alpha);
                }
            }
            float xv[KMAX], dv[DMAX];
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? xs[r * XP + kk] : 0.0f;
#pragma unroll
            for (int c = 0; c < DMAX; ++c) dv[c] = c < d ? ds[r * DP + c] : 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int kk = 0; kk < KMAX; ++kk) acc[i * KMAX + kk] = fmaf(dz[i], xv[kk], acc[i * KMAX + kk]);
#pragma unroll
                for (int c = 0; c < DMAX; ++c) acc[32 + i * DMAX + c] = fmaf(dz[i], dv[c], acc[32 + i * DMAX + c]);
                acc[48 + i] += dz[i];
            }
        }
    }
    thin_bwd_combine(acc, sm, red, lpr, slots_per_cta, k, d, n, part, alpha_part);
}

// thin_bwd for ACT_NONE (g already is dz): every warp runs its own cp.async pipeline — each thread
// copies exactly the pieces of g it will consume itself and the lanes of a row group copy that
// row's x / dot_x entries — TWO blocks ahead into a three-slot ring, so no load latency sits between the
// row blocks and no CTA barrier is needed (a row group never spans warps: n <= 128).  Two blocks in flight
// per CTA, two CTAs per SM: ~128 KB (fp32) / 64 KB (bf16) of g outstanding per SM; the double-buffered version
// held half of that and ran at 2.7 TB/s.  Dynamic shared memory: NST x blk x (n x sizeof(TW) + (XP + DP) x 4).
constexpr int BWD_NST = 3;
inline int thin_bwd_rif(int n, int elem_bytes) {
    const int slots = THREADS / (n >> 2);
    const int max_rows = elem_bytes == 2 ? 96 : 56;           // rows per block: three slots x two CTAs fit one SM
    const int rif = max_rows / slots;
    return rif < 1 ? 1 : (rif > 16 ? 16 : rif);
}
inline size_t thin_bwd_async_smem(int n, int elem_bytes) {
    const int blk = (THREADS / (n >> 2)) * thin_bwd_rif(n, elem_bytes);
    return static_cast<size_t>(BWD_NST) * blk * (static_cast<size_t>(n) * elem_bytes + ((KMAX + 1) + (DMAX + 1)) * sizeof(float));
}

// the 4 columns a thread consumes: 16 bytes of fp32, 8 bytes of bf16
__device__ __forceinline__ void cp_async_cols4(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_cols4(bf16 *dst, const bf16 *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(float *dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src)
                 : "memory");
}

template <typename TW>
__global__ void __launch_bounds__(THREADS, 2)
thin_bwd_async_kernel(int64_t rows, const TW *__restrict__ g, int64_t ldg, const float *__restrict__ x, int64_t ldx,
                      int k, const float *__restrict__ dot_x, int64_t ld_dot, int d, int n, int rif,
                      float *__restrict__ part) {
    extern __shared__ __align__(16) uint8_t dyn_raw[];
    __shared__ float sm[THREADS * 8];
    __shared__ float red[32];
    constexpr int XP = KMAX + 1, DP = DMAX + 1;
    const int lpr = n >> 2;
    const int cg = threadIdx.x % lpr;
    const int grp = threadIdx.x / lpr;
    const int slots_per_cta = THREADS / lpr;
    const int blk = slots_per_cta * rif;
    TW *gs = reinterpret_cast<TW *>(dyn_raw);          // [NST][blk][n]
    float *xs = reinterpret_cast<float *>(gs + static_cast<size_t>(BWD_NST) * blk * n);   // [NST][blk][XP]
    float *ds = xs + static_cast<size_t>(BWD_NST) * blk * XP;   // [NST][blk][DP]
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;

    auto stage = [&](int buf, int64_t base) {     // always commits a group (possibly empty): uniform group counting
        if (base < rows) {
            for (int u = 0; u < rif; ++u) {
                const int r = u * slots_per_cta + grp;
                const int64_t m = base + r;
                if (m < rows) {
                    cp_async_cols4(gs + (static_cast<size_t>(buf) * blk + r) * n + cg * 4, g + m * ldg + cg * 4);
                    for (int kk = cg; kk < k; kk += lpr) cp_async4(xs + (static_cast<size_t>(buf) * blk + r) * XP + kk, x + m * ldx + kk);
                    for (int c = cg; c < d; c += lpr) cp_async4(ds + (static_cast<size_t>(buf) * blk + r) * DP + c, dot_x + m * ld_dot + c);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t step = static_cast<int64_t>(gridDim.x) * blk;
    int64_t base = static_cast<int64_t>(blockIdx.x) * blk;
    int buf = 0;
    stage(0, base);
    stage(1, base + step);
    for (; base < rows; base += step) {
        stage((buf + 2) % BWD_NST, base + 2 * step);
        asm volatile("cp.async.wait_group 2;" ::: "memory");   // the group of `buf` has landed
        __syncwarp();                                  // the x / dot_x entries were copied by other lanes of the row group
        for (int u = 0; u < rif; ++u) {
            const int r = u * slots_per_cta + grp;
            if (base + r >= rows) break;
            const float4 gv = ld4(gs + (static_cast<size_t>(buf) * blk + r) * n + cg * 4);
            const float dz[4] = {gv.x, gv.y, gv.z, gv.w};
            const float *xr = xs + (static_cast<size_t>(buf) * blk + r) * XP;
            const float *dr = ds + (static_cast<size_t>(buf) * blk + r) * DP;
            float xv[KMAX], dv[DMAX];
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? xr[kk] : 0.0f;
#pragma unroll
            for (int c = 0; c < DMAX; ++c) dv[c] = c < d ? dr[c] : 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int kk = 0; kk < KMAX; ++kk) acc[i * KMAX + kk] = fmaf(dz[i], xv[kk], acc[i * KMAX + kk]);
#pragma unroll
                for (int c = 0; c < DMAX; ++c) acc[32 + i * DMAX + c] = fmaf(dz[i], dv[c], acc[32 + i * DMAX + c]);
                acc[48 + i] += dz[i];
            }
        }
        __syncwarp();                                  // done with buf before a later iteration stages into it
        buf = (buf + 1) % BWD_NST;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    thin_bwd_combine(acc, sm, red, lpr, slots_per_cta, k, d, n, part, nullptr);
}

// Stage 2a (grid): out[i] = sum over CTAs of part[cta][i]; columns [0,k) -> dW, [k,k+d) -> tbuf
// (T = dz^T dot_x), k+d -> db.  One WARP per output element: lane l adds partials l, l+32, ... in
// order, then a fixed xor tree — deterministic, and 32 loads in flight instead of one chain.
__global__ void __launch_bounds__(256)
thin_finalize_kernel(const float *__restrict__ part, int num_part, int n, int k, int d, float *__restrict__ dW,
                     float *__restrict__ db, float *__restrict__ tbuf) {
    const int kp = k + d + 1;
    const int total = n * kp;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= total) return;
    float s = 0.0f;
    for (int p = lane; p < num_part; p += 32) s += part[static_cast<int64_t>(p) * total + i];
    s = warp_sum(s);
    if (lane != 0) return;
    const int nn = i / kp, c = i % kp;
    if (c < k) {
        if (dW) dW[nn * k + c] = s;
    } else if (c < k + d) {
        tbuf[nn * d + (c - k)] = s;
    } else if (db) {
        db[nn] = s;
    }
}

// Stage 2b (one CTA): ddot = sum_{n,c} W[n][c0+c] * T[n][c];  dalpha = sum of the CTA partials.
__global__ void __launch_bounds__(256)
thin_scalars_kernel(const float *__restrict__ tbuf, int n, int k, int d, const float *__restrict__ W, int c0,
                    float *__restrict__ ddot, const float *__restrict__ alpha_part, int num_part,
                    float *__restrict__ dalpha) {
    __shared__ float red[32];
    if (ddot) {
        float dot = 0.0f;
        for (int i = threadIdx.x; i < n * d; i += blockDim.x)
            dot = fmaf(__ldg(W + (i / d) * k + c0 + i % d), tbuf[i], dot);
        dot = block_sum(dot, red);
        if (threadIdx.x == 0) ddot[0] = dot;
    }
    if (dalpha) {
        float s = 0.0f;
        if (alpha_part)
            for (int i = threadIdx.x; i < num_part; i += blockDim.x) s += alpha_part[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) dalpha[0] = s;
    }
}

// ---- readout head: n = 1 output column (models.py:328), k <= 128 ---------------------------------
// k/4 lanes per row, each owning 4 input columns: forward is a segmented dot product, backward
// streams g, z, x once and produces dx = dz * w, dW = sum_m dz x, db, dalpha.
template <typename TX>
__global__ void __launch_bounds__(THREADS)
head_fwd_kernel(int64_t rows, const TX *__restrict__ x, int64_t ldx, int k, const float *__restrict__ W,
                const float *__restrict__ bias, int act, const float *__restrict__ alpha_ptr, float *__restrict__ z,
                int64_t ldz, float *__restrict__ out, int64_t ldo, int accumulate_out) {
    const int lpr = k >> 2;
    const int cg = threadIdx.x % lpr;
    const int64_t slots = (static_cast<int64_t>(gridDim.x) * THREADS) / lpr;
    const int64_t slot = (static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x) / lpr;
    const float4 w = __ldg(reinterpret_cast<const float4 *>(W) + cg);
    const float b = bias ? __ldg(bias) : 0.0f;
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    const int64_t iters = (rows + slots - 1) / slots;   // warp-uniform trip count for the shuffles
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t m = it * slots + slot;
        float s = 0.0f;
        if (m < rows) {
            const float4 xv = ldg4(x + m * ldx + cg * 4);
            s = fmaf(xv.x, w.x, fmaf(xv.y, w.y, fmaf(xv.z, w.z, xv.w * w.w)));
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (m < rows && cg == 0) {
            const float zz = s + b;
            if (z) z[m * ldz] = zz;
            if (out) {
                const float o = act_forward(zz, act, alpha);
                out[m * ldo] = accumulate_out ? out[m * ldo] + o : o;
            }
        }
    }
}

template <typename TX>
__global__ void __launch_bounds__(THREADS)
head_bwd_kernel(int64_t rows, const float *__restrict__ g, int64_t ldg, const float *__restrict__ z, int64_t ldz,
                int act, const float *__restrict__ alpha_ptr, const TX *__restrict__ x, int64_t ldx, int k,
                const float *__restrict__ W, TX *__restrict__ dx, int64_t lddx, float *__restrict__ part,
                float *__restrict__ alpha_part, const TX *__restrict__ pz, int64_t ldpz, int pact,
                const float *__restrict__ palpha_ptr, float *__restrict__ palpha_part) {
    // pz != NULL: dx leaves as dx * act'(pz) (the dz of the layer that produced x) and
    // palpha_part[cta] = sum dx * min(pz, 0) — see hgin_linear_bwd_post.
    __shared__ float sm[THREADS * 5];
    __shared__ float red[32];
    const int lpr = k >> 2;
    const int cg = threadIdx.x % lpr;
    const int grp = threadIdx.x / lpr;
    const int slots_per_cta = THREADS / lpr;
    const int64_t slots = static_cast<int64_t>(gridDim.x) * slots_per_cta;
    const float4 w = __ldg(reinterpret_cast<const float4 *>(W) + cg);
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    const float palpha = (pz && pact == HGIN_ACT_PRELU) ? __ldg(palpha_ptr) : 0.0f;
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, adb = 0.0f, adal = 0.0f, apal = 0.0f;
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * slots_per_cta + grp; m < rows; m += slots) {
        float dz = __ldg(g + m * ldg);
        if (act != HGIN_ACT_NONE) {
            const float zv = __ldg(z + m * ldz);
            if (cg == 0 && act == HGIN_ACT_PRELU && !(zv > 0.0f)) adal += dz * zv;
            dz = act_backward(dz, zv, act, alpha);
        }
        const float4 xv = ldg4(x + m * ldx + cg * 4);
        aw[0] = fmaf(dz, xv.x, aw[0]); aw[1] = fmaf(dz, xv.y, aw[1]);
        aw[2] = fmaf(dz, xv.z, aw[2]); aw[3] = fmaf(dz, xv.w, aw[3]);
        if (cg == 0) adb += dz;
        if (dx) {
            float o[4] = {dz * w.x, dz * w.y, dz * w.z, dz * w.w};
            if (pz) {
                const float4 p4 = ldg4(pz + m * ldpz + cg * 4);
                const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (pact == HGIN_ACT_PRELU && !(pv[i] > 0.0f)) apal = fmaf(o[i], pv[i], apal);
                    o[i] = act_backward(o[i], pv[i], pact, palpha);
                }
            }
            st4(dx + m * lddx + cg * 4, o[0], o[1], o[2], o[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sm[threadIdx.x * 5 + j] = aw[j];
    sm[threadIdx.x * 5 + 4] = adb;
    __syncthreads();
    float *dst = part + static_cast<int64_t>(blockIdx.x) * (k + 1);   // [k | db]
    for (int idx = threadIdx.x; idx < lpr * 5; idx += THREADS) {
        const int c_g = idx / 5, j = idx % 5;
        float s = 0.0f;
        for (int sl = 0; sl < slots_per_cta; ++sl) s += sm[(sl * lpr + c_g) * 5 + j];
        if (j < 4) dst[c_g * 4 + j] = s;
        else if (c_g == 0) dst[k] = s;
    }
    __syncthreads();
    const float da = block_sum(adal, red);
    if (threadIdx.x == 0 && alpha_part) alpha_part[blockIdx.x] = da;
    if (palpha_part) {
        const float dp = block_sum(apal, red);
        if (threadIdx.x == 0) palpha_part[blockIdx.x] = dp;
    }
}

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int thin_ctas(int64_t rows, int n, int per_sm = 4) {
    const int slots = THREADS / (n >> 2);
    const int64_t want = ceil_div(rows, static_cast<int64_t>(slots) * 4);
    const int64_t cap = static_cast<int64_t>(kNumSMs) * per_sm;
    return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

// rows of 4 wide-side columns must be vector-addressable: 16 B of fp32, 8 B of bf16
inline bool wide_ok(const void *p, int64_t ld, int dtype) {
    if (!p) return true;
    return dtype == HGIN_DTYPE_BF16 ? (ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 7u) == 0) : (ld % 4 == 0 && aligned16(p));
}

bool fwd_eligible(const float *x1, int k1, int k2, int n, const void *z, int64_t ldz, const void *out, int64_t ldo, int dtype) {
    return x1 && k2 == 0 && k1 <= KMAX && n >= 4 && n <= 128 && pow2(n) && wide_ok(z, ldz, dtype) && wide_ok(out, ldo, dtype);
}

bool bwd_eligible(const void *g, int64_t ldg, const void *z, int64_t ldz, int act, int k1, int k2, int n, int c0,
                  int c1, const void *dx, const float *dot_x, int dtype) {
    const int d = dot_x ? c1 - c0 : 0;
    return k2 == 0 && k1 <= KMAX && n >= 4 && n <= 128 && pow2(n) && dx == nullptr && d <= DMAX &&
           (c1 == c0 || dot_x != nullptr) && g && wide_ok(g, ldg, dtype) && (act == HGIN_ACT_NONE || (z && wide_ok(z, ldz, dtype)));
}

int64_t bwd_workspace_bytes(int n, int k) {
    const int64_t ctas = static_cast<int64_t>(kNumSMs) * 4;
    const int64_t thin = ctas * (static_cast<int64_t>(n) * ((k < KMAX ? k : KMAX) + DMAX + 1) + 1) + static_cast<int64_t>(n) * DMAX;
    const int64_t head = ctas * (k + 3) + 8;
    return align_up((thin > head ? thin : head) * 4, 256) + 256;
}

int32_t linear_fwd(int64_t rows, const float *x, int64_t ldx, int k, const float *W, const float *bias, int n, int act,
                   const float *alpha, void *z, int64_t ldz, void *out, int64_t ldo, int accumulate_out, int dtype,
                   cudaStream_t s) {
    if (dtype == HGIN_DTYPE_BF16)
        thin_fwd_kernel<bf16><<<thin_ctas(rows, n), THREADS, 0, s>>>(rows, x, ldx, k, W, bias, n, act, alpha, static_cast<bf16 *>(z),
                                                                    ldz, static_cast<bf16 *>(out), ldo, accumulate_out);
    else
        thin_fwd_kernel<float><<<thin_ctas(rows, n), THREADS, 0, s>>>(rows, x, ldx, k, W, bias, n, act, alpha, static_cast<float *>(z),
                                                                     ldz, static_cast<float *>(out), ldo, accumulate_out);
    HGIN_CHECK_LAUNCH("hgin_linear_fwd(thin)");
    return HGIN_OK;
}

template <typename TW>
static int32_t linear_bwd_t(int64_t rows, const TW *g, int64_t ldg, const TW *z, int64_t ldz, int act, const float *alpha,
                            const float *x, int64_t ldx, int k, const float *W, int n, int c0, int c1, const float *dot_x,
                            int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace, cudaStream_t s) {
    const int d = dot_x ? c1 - c0 : 0;
    const int ctas = thin_ctas(rows, n, 2);   // 128 registers/thread: two resident CTAs per SM, one wave
    float *part = static_cast<float *>(workspace);
    float *alpha_part = part + static_cast<int64_t>(ctas) * n * (k + d + 1);
    const bool want_alpha = dalpha && act == HGIN_ACT_PRELU;
    if (act != HGIN_ACT_NONE)
        thin_bwd_kernel<true, TW><<<ctas, THREADS, 0, s>>>(rows, g, ldg, z, ldz, act, alpha, x, ldx, k, dot_x, ld_dot, d, n, part,
                                                           want_alpha ? alpha_part : nullptr);
    else {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(thin_bwd_async_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
            attr_set = true;
        }
        thin_bwd_async_kernel<TW><<<ctas, THREADS, thin_bwd_async_smem(n, sizeof(TW)), s>>>(rows, g, ldg, x, ldx, k, dot_x, ld_dot, d,
                                                                                             n, thin_bwd_rif(n, sizeof(TW)), part);
    }
    float *tbuf = alpha_part + ctas;
    const int total = n * (k + d + 1);
    thin_finalize_kernel<<<static_cast<unsigned>(ceil_div(total, 8)), 256, 0, s>>>(part, ctas, n, k, d, dW, db, tbuf);
    if (ddot || dalpha)
        thin_scalars_kernel<<<1, 256, 0, s>>>(tbuf, n, k, d, W, c0, ddot, want_alpha ? alpha_part : nullptr, ctas,
                                              dalpha);
    HGIN_CHECK_LAUNCH("hgin_linear_bwd(thin)");
    return HGIN_OK;
}

int32_t linear_bwd(int64_t rows, const void *g, int64_t ldg, const void *z, int64_t ldz, int act, const float *alpha,
                   const float *x, int64_t ldx, int k, const float *W, int n, int c0, int c1, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace, int dtype,
                   cudaStream_t s) {
    if (dtype == HGIN_DTYPE_BF16)
        return linear_bwd_t<bf16>(rows, static_cast<const bf16 *>(g), ldg, static_cast<const bf16 *>(z), ldz, act, alpha, x, ldx, k,
                                  W, n, c0, c1, dot_x, ld_dot, ddot, dW, db, dalpha, workspace, s);
    return linear_bwd_t<float>(rows, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act, alpha, x, ldx, k,
                               W, n, c0, c1, dot_x, ld_dot, ddot, dW, db, dalpha, workspace, s);
}

bool head_fwd_eligible(const void *x1, int64_t ld1, int k1, int k2, int n, int dtype) {
    return n == 1 && k2 == 0 && k1 >= 4 && k1 <= 128 && k1 % 4 == 0 && pow2(k1 >> 2) && x1 && wide_ok(x1, ld1, dtype);
}

bool head_bwd_eligible(const void *x1, int64_t ld1, int k1, int k2, int n, int c0, int c1, const void *dx,
                       int64_t lddx, const void *dot_x, const float *W, int dtype) {
    return head_fwd_eligible(x1, ld1, k1, k2, n, dtype) && dot_x == nullptr && aligned16(W) &&
           (dx == nullptr || (c0 == 0 && c1 == k1 && wide_ok(dx, lddx, dtype))) && (c1 == c0 || dx != nullptr);
}

int32_t head_fwd(int64_t rows, const void *x, int64_t ldx, int k, const float *W, const float *bias, int act,
                 const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo, int accumulate_out, int dtype,
                 cudaStream_t s) {
    if (dtype == HGIN_DTYPE_BF16)
        head_fwd_kernel<bf16><<<thin_ctas(rows, k), THREADS, 0, s>>>(rows, static_cast<const bf16 *>(x), ldx, k, W, bias, act, alpha,
                                                                    z, ldz, out, ldo, accumulate_out);
    else
        head_fwd_kernel<float><<<thin_ctas(rows, k), THREADS, 0, s>>>(rows, static_cast<const float *>(x), ldx, k, W, bias, act,
                                                                     alpha, z, ldz, out, ldo, accumulate_out);
    HGIN_CHECK_LAUNCH("hgin_linear_fwd(head)");
    return HGIN_OK;
}

int32_t head_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *alpha,
                 const void *x, int64_t ldx, int k, const float *W, void *dx, int64_t lddx, float *dW, float *db,
                 float *dalpha, void *workspace, const void *post_z, int64_t ld_post, int post_act,
                 const float *post_alpha, float *post_dalpha, int dtype, cudaStream_t s) {
    const int ctas = thin_ctas(rows, k);
    float *part = static_cast<float *>(workspace);
    float *alpha_part = part + static_cast<int64_t>(ctas) * (k + 1);
    float *palpha_part = alpha_part + ctas;
    float *tbuf = palpha_part + ctas;
    const bool want_alpha = dalpha && act == HGIN_ACT_PRELU;
    const bool post_on = post_z && dx && post_act != HGIN_ACT_NONE;
    const bool want_palpha = post_on && post_dalpha && post_act == HGIN_ACT_PRELU;
    if (dtype == HGIN_DTYPE_BF16)
        head_bwd_kernel<bf16><<<ctas, THREADS, 0, s>>>(rows, g, ldg, z, ldz, act, alpha, static_cast<const bf16 *>(x), ldx, k, W,
                                                      static_cast<bf16 *>(dx), lddx, part, want_alpha ? alpha_part : nullptr,
                                                      post_on ? static_cast<const bf16 *>(post_z) : nullptr, ld_post, post_act,
                                                      post_alpha, want_palpha ? palpha_part : nullptr);
    else
        head_bwd_kernel<float><<<ctas, THREADS, 0, s>>>(rows, g, ldg, z, ldz, act, alpha, static_cast<const float *>(x), ldx, k, W,
                                                       static_cast<float *>(dx), lddx, part, want_alpha ? alpha_part : nullptr,
                                                       post_on ? static_cast<const float *>(post_z) : nullptr, ld_post, post_act,
                                                       post_alpha, want_palpha ? palpha_part : nullptr);
    thin_finalize_kernel<<<static_cast<unsigned>(ceil_div(k + 1, 8)), 256, 0, s>>>(part, ctas, 1, k, 0, dW, db, tbuf);
    if (dalpha) thin_scalars_kernel<<<1, 256, 0, s>>>(tbuf, 1, k, 0, W, 0, nullptr, want_alpha ? alpha_part : nullptr, ctas, dalpha);
    if (post_dalpha)
        thin_scalars_kernel<<<1, 256, 0, s>>>(tbuf, 1, k, 0, W, 0, nullptr, want_palpha ? palpha_part : nullptr, ctas,
                                              post_dalpha);
    HGIN_CHECK_LAUNCH("hgin_linear_bwd(head)");
    return HGIN_OK;
}

}  // namespace thin
}  // namespace hgin

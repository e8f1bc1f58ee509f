// The whole HeteroGIN train step of config.json's model as THREE kernels (Cfg-A: launch-bound regime).
//
// config.json: MP_LAYERS = 1, NODE_EMBEDDING_SIZE = 8, MLP_LAYERS = [128, 32], PReLU everywhere.  With one message
// passing layer only link -> path ('includes') reaches the readout (models.py:362-371), so after the neighbour sum the
// ENTIRE network is row-local in the path rows:
//     h = [sum_{l in path} x_link[l] | (1+eps) x_path]          models.py:208-215 (concat=True)
//     e = PReLU(h W0^T + b0)                                   models.py:217, 236-239
//     a1 = PReLU([e | x_path] W1^T + b1); a2 = PReLU(a1 W2^T + b2); out = a2 W3^T + b3       models.py:362-374
// Eager execution is 25 launches of a few microseconds each (0.30 ms per step even as one CUDA-graph replay, most of it
// in the SIMT weight-gradient kernels and their partial reductions).  Here:
//   small_fwd_loss_kernel   forward of every row, per-CTA partial of sum |(out - y) / y|          (train.py:12-13)
//   small_bwd_kernel        S = sum of the partials (every CTA, same fixed order); per row the forward's pre-activations
//                           are reloaded (1 KB per row, L2-resident), backward per row, weight gradients
//                           accumulated in registers over the CTA's row tiles (each thread owns fixed dW elements, so no
//                           atomics and a fixed order), per-CTA partials
//   small_reduce_kernel     partials -> the parameters' gradient tensors (fixed order over CTAs)
// followed by the usual hgin_adam_step on the flat bucket.  All weights live in shared memory (33 KB), a tile of 32 rows'
// activations is staged there for the outer products; warp per row, lanes across the output features.
// Neighbour sums are sequential in CSR order with __fadd_rn (bit-identical to K1); the dense parts use FMA chains in a
// different association than the SIMT engine, i.e. they agree with the reference to fp32 rounding (tests: rtol 1e-4).
#include <math.h>

#include "hgin_common.cuh"

namespace hgin {
namespace {

constexpr int SS_THREADS = 256;
constexpr int SS_WARPS = 8;
constexpr int SS_TILE = 32;          // staging slots per CTA iteration (8 warps x up to 4 rows)
constexpr int SS_MAX_EMB = 32;
constexpr int SS_MAX_K0 = 16;        // fl + fp
constexpr int SS_MAX_K1 = 32;        // emb + (concat ? fp : 0)
constexpr int SS_MAX_N1 = 128;
constexpr int SS_MAX_N2 = 32;

// per-CTA partial layout (floats)
constexpr int P_W2 = 0;                                   // [32][128]  dW2[n][k]
constexpr int P_W1 = P_W2 + SS_MAX_N2 * SS_MAX_N1;        // [128][32]  dW1[n][k]
constexpr int P_W0 = P_W1 + SS_MAX_N1 * SS_MAX_K1;        // [32][16]   dW0[n][k]
constexpr int P_B1 = P_W0 + SS_MAX_EMB * SS_MAX_K0;       // [128]
constexpr int P_B2 = P_B1 + SS_MAX_N1;                    // [32]
constexpr int P_B0 = P_B2 + SS_MAX_N2;                    // [32]
constexpr int P_W3 = P_B0 + SS_MAX_EMB;                   // [32]
constexpr int P_SC = P_W3 + SS_MAX_N2;                    // db3, dalphaR, dalpha0, deps
constexpr int P_TOTAL = P_SC + 4;

struct SmallParams {
    int np;
    const int32_t *rowptr, *col;       // destination-sorted CSR of link -> path
    const float *xp; int ldp; int fp; int pcol[8];
    const float *xl; int ldl; int fl; int lcol[8];
    const float *y;
    int emb, n1, n2, concat;
    const float *W0, *b0, *a0, *eps0, *W1, *b1, *aR, *W2, *b2, *W3, *b3;
};

struct SmallSmem {
    // every contraction reads BOTH operands four elements at a time along the contraction index (one LDS.128 of the
    // weights feeds 4 x R FMAs, one broadcast LDS.128 of the activations 4 x 4): row pitches of 36 / 132 floats spread
    // the lanes' rows over the banks
    float W1nk[SS_MAX_N1][SS_MAX_K1 + 4];  // [n][k]   z1 (contract k), d e (contract n, 4 k at a time)
    float W2nk[SS_MAX_N2][SS_MAX_N1 + 4];  // [n][k]   z2 (contract k)
    float W2kn[SS_MAX_N1][SS_MAX_N2 + 4];  // [k][n]   d a1 (contract n)
    float W0[SS_MAX_EMB][SS_MAX_K0 + 1];   // [n][k] (+1: no bank conflicts across n)
    float b1[SS_MAX_N1], b2[SS_MAX_N2], b0[SS_MAX_EMB], W3[SS_MAX_N2];
    // row staging
    float H[SS_TILE][SS_MAX_K0];
    float Xin[SS_TILE][SS_MAX_K1];
    float A1[SS_TILE][SS_MAX_N1];
    float Dz1[SS_TILE][SS_MAX_N1];
    float Dz2[SS_TILE][SS_MAX_N2];
    float Dz0[SS_TILE][SS_MAX_EMB];
    float red[SS_WARPS][36];
};

__device__ __forceinline__ float prelu(float z, float a) { return z > 0.f ? z : a * z; }
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Weights global -> shared.  Loads are issued in batches of 8 per thread before any store (a loop of load-then-store
// serialises ~16 global latencies per matrix, which at these sizes is a third of the kernel); W2 is read twice, once per
// layout, with the THREAD index running along the fast dimension of the destination so that the stores are conflict-free.
template <int COUNT, class Src, class Dst>
__device__ __forceinline__ void stage(Src src, Dst dst) {
    static_assert(COUNT % (SS_THREADS * 8) == 0 || COUNT < SS_THREADS * 8, "batches of 8 per thread");
    constexpr int BATCH = COUNT >= SS_THREADS * 8 ? 8 : (COUNT + SS_THREADS - 1) / SS_THREADS;
    for (int base = threadIdx.x; base < COUNT; base += SS_THREADS * BATCH) {
        float v[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) v[u] = (base + u * SS_THREADS < COUNT) ? src(base + u * SS_THREADS) : 0.f;
#pragma unroll
        for (int u = 0; u < BATCH; ++u)
            if (base + u * SS_THREADS < COUNT) dst(base + u * SS_THREADS, v[u]);
    }
}

__device__ void load_weights(const SmallParams &p, SmallSmem &s) {
    const int k1 = p.emb + (p.concat ? p.fp : 0), k0 = p.fl + p.fp;
    stage<SS_MAX_K1 * SS_MAX_N1>(
        [&](int i) { const int k = i / SS_MAX_N1, n = i % SS_MAX_N1; return (k < k1 && n < p.n1) ? __ldg(p.W1 + n * k1 + k) : 0.f; },
        [&](int i, float v) { s.W1nk[i % SS_MAX_N1][i / SS_MAX_N1] = v; });
    stage<SS_MAX_N1 * SS_MAX_N2>(
        [&](int i) { const int k = i / SS_MAX_N2, n = i % SS_MAX_N2; return (k < p.n1 && n < p.n2) ? __ldg(p.W2 + n * p.n1 + k) : 0.f; },
        [&](int i, float v) { s.W2kn[i / SS_MAX_N2][i % SS_MAX_N2] = v; });      // (pad columns are never read)
    stage<SS_MAX_N2 * SS_MAX_N1>(
        [&](int i) { const int n = i / SS_MAX_N1, k = i % SS_MAX_N1; return (k < p.n1 && n < p.n2) ? __ldg(p.W2 + n * p.n1 + k) : 0.f; },
        [&](int i, float v) { s.W2nk[i / SS_MAX_N1][i % SS_MAX_N1] = v; });
    stage<SS_MAX_EMB * SS_MAX_K0>(
        [&](int i) { const int n = i / SS_MAX_K0, k = i % SS_MAX_K0; return (n < p.emb && k < k0) ? __ldg(p.W0 + n * k0 + k) : 0.f; },
        [&](int i, float v) { s.W0[i / SS_MAX_K0][i % SS_MAX_K0] = v; });
    if (threadIdx.x < SS_MAX_N1) s.b1[threadIdx.x] = threadIdx.x < p.n1 ? __ldg(p.b1 + threadIdx.x) : 0.f;
    if (threadIdx.x >= 128 && threadIdx.x < 128 + SS_MAX_N2) {
        const int i = threadIdx.x - 128;
        s.b2[i] = i < p.n2 ? __ldg(p.b2 + i) : 0.f;
        s.W3[i] = i < p.n2 ? __ldg(p.W3 + i) : 0.f;
    }
    if (threadIdx.x >= 192 && threadIdx.x < 192 + SS_MAX_EMB) {
        const int i = threadIdx.x - 192;
        s.b0[i] = i < p.emb ? __ldg(p.b0 + i) : 0.f;
    }
}

// SS_R rows are carried TOGETHER by a warp (R independent FMA chains hide the LDS latency); a tile is 8 * R rows.  R = 2, 3
// or 4 is picked by the host so that the CTAs' row ranges split into whole rounds with little waste.

// Layer widths either as run-time values or — for config.json's own shape (emb 8, 3 + 3 raw columns, readout 128 x 32,
// concat_path) — as compile-time constants, so that the short contraction loops unroll completely and their bounds checks
// fold away (the run-time variant spends 3 of 4 issue slots on loop and address arithmetic).
struct DynDims {
    int emb_, fl_, fp_, n1_, n2_, concat_;
    __device__ explicit DynDims(const SmallParams &p) : emb_(p.emb), fl_(p.fl), fp_(p.fp), n1_(p.n1), n2_(p.n2), concat_(p.concat) {}
    __device__ __forceinline__ int emb() const { return emb_; }
    __device__ __forceinline__ int fl() const { return fl_; }
    __device__ __forceinline__ int fp() const { return fp_; }
    __device__ __forceinline__ int n1() const { return n1_; }
    __device__ __forceinline__ int n2() const { return n2_; }
    __device__ __forceinline__ int k0() const { return fl_ + fp_; }
    __device__ __forceinline__ int k1() const { return emb_ + (concat_ ? fp_ : 0); }
};
template <int EMB, int FL, int FP, int N1, int N2, int CONCAT>
struct FixedDims {
    __device__ explicit FixedDims(const SmallParams &) {}
    __device__ __forceinline__ constexpr int emb() const { return EMB; }
    __device__ __forceinline__ constexpr int fl() const { return FL; }
    __device__ __forceinline__ constexpr int fp() const { return FP; }
    __device__ __forceinline__ constexpr int n1() const { return N1; }
    __device__ __forceinline__ constexpr int n2() const { return N2; }
    __device__ __forceinline__ constexpr int k0() const { return FL + FP; }
    __device__ __forceinline__ constexpr int k1() const { return EMB + (CONCAT ? FP : 0); }
};
using DefaultDims = FixedDims<8, 3, 3, 128, 32, 1>;     // config.json

template <int SS_R>
struct RowBatch {
    float z0[SS_R], z1[SS_R][4], z2[SS_R], a2[SS_R], out[SS_R];
};

// The gather chain rowptr -> col -> x_link is three dependent (cold: DRAM) latencies and bounds the kernel, not the FMAs
// (ncu: the store of the gathered columns into the scratch is the top stall).  So the row bounds are requested two
// batches ahead and the first 16 neighbour ids one batch ahead; only the x_link fetch itself stays exposed.
template <int SS_R>
struct GatherPre {
    int e0[SS_R], len[SS_R], nb[SS_R];
};
template <int SS_R>
__device__ __forceinline__ void pre_bounds(const SmallParams &p, int row0, int end, GatherPre<SS_R> &g) {
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        const int row = min(row0 + q, p.np - 1);
        const bool on = row0 < end;            // (whole batch past the CTA's range: nothing to fetch)
        g.e0[q] = on ? __ldg(p.rowptr + row) : 0;
        g.len[q] = on ? __ldg(p.rowptr + row + 1) - g.e0[q] : 0;
    }
}
template <int SS_R>
__device__ __forceinline__ void pre_cols(const SmallParams &p, int lane, GatherPre<SS_R> &g) {
#pragma unroll
    for (int q = 0; q < SS_R; ++q) g.nb[q] = (lane < 16 && lane < g.len[q]) ? __ldg(p.col + g.e0[q] + lane) : -1;
}

// Forward of SS_R path rows (tile slots r0 .. r0 + SS_R - 1; rows past `np` are clamped to the last row and masked by the
// caller) by one warp; stages H, Xin, A1.  Returns the pre-activations the backward needs.
template <int SS_R, class D>
__device__ __forceinline__ RowBatch<SS_R> rows_forward(const SmallParams &p, const D &d, SmallSmem &s, int row0, int r0, int lane,
                                                       float ope, float a0, float aR, float b3, const GatherPre<SS_R> &pre) {
    RowBatch<SS_R> st;
    const int k0 = d.k0(), k1 = d.k1();
    int rows[SS_R];
#pragma unroll
    for (int q = 0; q < SS_R; ++q) rows[q] = min(row0 + q, p.np - 1);
    // Neighbour sums.  The chain rowptr -> col -> x_link costs three memory latencies; walking it neighbour by neighbour
    // (and row by row) made this the slowest part of the kernel.  So: all rows' bounds first, then ONE parallel load of up to
    // 16 neighbour ids per row (lane = neighbour), one parallel fetch of their feature columns into this row's scratch
    // (its Dz1 staging slot, unused until the backward part), and only then the sequential, left-to-right additions
    // (== zeros().scatter_add_ on the CPU) out of shared memory.
    int e0[SS_R], len[SS_R];
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        e0[q] = pre.e0[q];
        len[q] = pre.len[q];
    }
    float hv[SS_R];
#pragma unroll
    for (int q = 0; q < SS_R; ++q) hv[q] = 0.f;
    int max_len = 0;
#pragma unroll
    for (int q = 0; q < SS_R; ++q) max_len = max(max_len, len[q]);
    for (int base = 0; base < max_len; base += 16) {
        int nb[SS_R];
#pragma unroll
        for (int q = 0; q < SS_R; ++q)
            nb[q] = base == 0 ? pre.nb[q] : ((lane < 16 && base + lane < len[q]) ? __ldg(p.col + e0[q] + base + lane) : -1);
#pragma unroll
        for (int q = 0; q < SS_R; ++q) {
            if (nb[q] >= 0) {
                const float *src = p.xl + static_cast<int64_t>(nb[q]) * p.ldl;
#pragma unroll
                for (int c = 0; c < d.fl(); ++c) s.Dz1[r0 + q][lane * 8 + c] = __ldg(src + p.lcol[c]);
            }
        }
        __syncwarp();
        if (lane < d.fl()) {
#pragma unroll
            for (int q = 0; q < SS_R; ++q) {
                const int cnt = min(16, len[q] - base);
                for (int j = 0; j < cnt; ++j) hv[q] = __fadd_rn(hv[q], s.Dz1[r0 + q][j * 8 + lane]);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        float v = hv[q];
        if (lane >= d.fl() && lane < k0) v = __fmul_rn(ope, __ldg(p.xp + static_cast<int64_t>(rows[q]) * p.ldp + p.pcol[lane - d.fl()]));
        if (lane >= k0) v = 0.f;
        if (lane < SS_MAX_K0) s.H[r0 + q][lane] = v;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < SS_R; ++q) st.z0[q] = lane < d.emb() ? s.b0[lane] : 0.f;
    if (lane < d.emb()) {
#pragma unroll
        for (int k = 0; k < k0; ++k) {
            const float w = s.W0[lane][k];
#pragma unroll
            for (int q = 0; q < SS_R; ++q) st.z0[q] = fmaf(s.H[r0 + q][k], w, st.z0[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        float xin = 0.f;
        if (lane < d.emb()) xin = prelu(st.z0[q], a0);
        else if (lane < k1) xin = __ldg(p.xp + static_cast<int64_t>(rows[q]) * p.ldp + p.pcol[lane - d.emb()]);
        s.Xin[r0 + q][lane] = xin;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < SS_R; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) st.z1[q][j] = s.b1[lane + 32 * j];
#pragma unroll
    for (int k = 0; k < k1; k += 4) {          // (Xin and W1nk are zero beyond k1)
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4 *>(&s.W1nk[lane + 32 * j][k]);
#pragma unroll
        for (int q = 0; q < SS_R; ++q) {
            const float4 x = *reinterpret_cast<const float4 *>(&s.Xin[r0 + q][k]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                st.z1[q][j] = fmaf(x.w, w[j].w, fmaf(x.z, w[j].z, fmaf(x.y, w[j].y, fmaf(x.x, w[j].x, st.z1[q][j]))));
        }
    }
#pragma unroll
    for (int q = 0; q < SS_R; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) s.A1[r0 + q][lane + 32 * j] = (lane + 32 * j) < d.n1() ? prelu(st.z1[q][j], aR) : 0.f;
    __syncwarp();
    // z2: two partial chains per row (k mod 8 < 4 / >= 4): 2 * SS_R independent FMA chains
    float ze[SS_R], zo[SS_R];
#pragma unroll
    for (int q = 0; q < SS_R; ++q) { ze[q] = s.b2[lane]; zo[q] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < SS_MAX_N1; k += 8) {      // (A1 and W2nk are zero beyond n1)
        const float4 w0 = *reinterpret_cast<const float4 *>(&s.W2nk[lane][k]);
        const float4 w1 = *reinterpret_cast<const float4 *>(&s.W2nk[lane][k + 4]);
#pragma unroll
        for (int q = 0; q < SS_R; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(&s.A1[r0 + q][k]);
            const float4 b = *reinterpret_cast<const float4 *>(&s.A1[r0 + q][k + 4]);
            ze[q] = fmaf(a.w, w0.w, fmaf(a.z, w0.z, fmaf(a.y, w0.y, fmaf(a.x, w0.x, ze[q]))));
            zo[q] = fmaf(b.w, w1.w, fmaf(b.z, w1.z, fmaf(b.y, w1.y, fmaf(b.x, w1.x, zo[q]))));
        }
    }
    float part[SS_R];
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        st.z2[q] = ze[q] + zo[q];
        st.a2[q] = lane < d.n2() ? prelu(st.z2[q], aR) : 0.f;
        part[q] = st.a2[q] * s.W3[lane];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < SS_R; ++q) part[q] += __shfl_xor_sync(0xffffffffu, part[q], o);
#pragma unroll
    for (int q = 0; q < SS_R; ++q) st.out[q] = b3 + part[q];
    return st;
}

// What the forward kernel leaves per row for the backward kernel (1 KB, L2-resident at config.json's sizes): the three
// pre-activations, the readout input and the GIN layer input.  The backward kernel then reloads 8 coalesced words per lane
// and row instead of repeating the gather and the three contractions (the recompute was ~1/3 of the backward kernel).
constexpr int A_Z1 = 0, A_Z2 = 128, A_Z0 = 160, A_XIN = 192, A_H = 224, A_OUT = 240, A_ROW = 256;

template <int SS_R>
__device__ __forceinline__ void rows_store(const SmallSmem &s, const RowBatch<SS_R> &st, float *__restrict__ act, int row0, int end,
                                           int r0, int lane) {
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        if (row0 + q >= end) continue;
        float *a = act + static_cast<int64_t>(row0 + q) * A_ROW;
#pragma unroll
        for (int j = 0; j < 4; ++j) a[A_Z1 + lane + 32 * j] = st.z1[q][j];
        a[A_Z2 + lane] = st.z2[q];
        a[A_Z0 + lane] = st.z0[q];
        a[A_XIN + lane] = s.Xin[r0 + q][lane];
        if (lane < SS_MAX_K0) a[A_H + lane] = s.H[r0 + q][lane];
        if (lane == 0) a[A_OUT] = st.out[q];
    }
}

template <int SS_R, class D>
__device__ __forceinline__ RowBatch<SS_R> rows_reload(const D &d, SmallSmem &s, const float *__restrict__ act, int row0, int np,
                                                      int r0, int lane, float aR) {
    RowBatch<SS_R> st;
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
        const float *a = act + static_cast<int64_t>(min(row0 + q, np - 1)) * A_ROW;
#pragma unroll
        for (int j = 0; j < 4; ++j) st.z1[q][j] = __ldg(a + A_Z1 + lane + 32 * j);
        st.z2[q] = __ldg(a + A_Z2 + lane);
        st.z0[q] = __ldg(a + A_Z0 + lane);
        s.Xin[r0 + q][lane] = __ldg(a + A_XIN + lane);
        if (lane < SS_MAX_K0) s.H[r0 + q][lane] = __ldg(a + A_H + lane);
        st.out[q] = __ldg(a + A_OUT);
    }
#pragma unroll
    for (int q = 0; q < SS_R; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s.A1[r0 + q][lane + 32 * j] = (lane + 32 * j) < d.n1() ? prelu(st.z1[q][j], aR) : 0.f;
        st.a2[q] = lane < d.n2() ? prelu(st.z2[q], aR) : 0.f;
    }
    __syncwarp();
    return st;
}

// CTA c owns the contiguous rows [c * rows_per_cta, (c + 1) * rows_per_cta): cost is proportional to rows, not to tiles
__device__ __forceinline__ void cta_rows(const SmallParams &p, int &beg, int &end) {
    const int per = (p.np + gridDim.x - 1) / gridDim.x;
    beg = min(static_cast<int>(blockIdx.x) * per, p.np);
    end = min(beg + per, p.np);
}

template <int SS_R, class D>
__global__ void __launch_bounds__(SS_THREADS, 2)
small_fwd_loss_kernel(const SmallParams p, float *__restrict__ partial_s, float *__restrict__ out, float *__restrict__ act) {
    constexpr int TILE = SS_WARPS * SS_R;
    const D d(p);
    extern __shared__ __align__(16) uint8_t raw[];
    SmallSmem &s = *reinterpret_cast<SmallSmem *>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int beg, end;
    cta_rows(p, beg, end);
    GatherPre<SS_R> g1, g2;      // (requested before the weights: the two latencies overlap)
    pre_bounds<SS_R>(p, beg + warp * SS_R, end, g1);
    pre_bounds<SS_R>(p, beg + TILE + warp * SS_R, end, g2);
    const float ope = __fadd_rn(1.0f, __ldg(p.eps0)), a0 = __ldg(p.a0), aR = __ldg(p.aR), b3 = __ldg(p.b3);
    load_weights(p, s);
    pre_cols<SS_R>(p, lane, g1);
    __syncthreads();
    float sum = 0.f;
    for (int t0 = beg; t0 < end; t0 += TILE) {
        const int r0 = warp * SS_R, row0 = t0 + r0;
        const GatherPre<SS_R> cur = g1;
        g1 = g2;
        pre_bounds<SS_R>(p, t0 + 2 * TILE + r0, end, g2);
        pre_cols<SS_R>(p, lane, g1);
        if (row0 >= end) continue;
        const RowBatch<SS_R> st = rows_forward<SS_R, D>(p, d, s, row0, r0, lane, ope, a0, aR, b3, cur);
        rows_store<SS_R>(s, st, act, row0, end, r0, lane);
#pragma unroll
        for (int q = 0; q < SS_R; ++q) {
            if (row0 + q < end) {
                const float yi = __ldg(p.y + row0 + q);
                sum += fabsf(__fdiv_rn(__fsub_rn(st.out[q], yi), yi));
                if (out != nullptr && lane == 0) out[row0 + q] = st.out[q];
            }
        }
    }
    if (lane == 0) s.red[warp][0] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < SS_WARPS; ++w) t += s.red[w][0];
        partial_s[blockIdx.x] = t;
    }
}

template <int SS_R, class D>
__global__ void __launch_bounds__(SS_THREADS, 2)
small_bwd_kernel(const SmallParams p, const float *__restrict__ partial_s, int num_partial_s, float *__restrict__ sums,
                 float *__restrict__ loss_out, float *__restrict__ partials, const float *__restrict__ act) {
    constexpr int TILE = SS_WARPS * SS_R;
    const D d(p);
    extern __shared__ __align__(16) uint8_t raw[];
    SmallSmem &s = *reinterpret_cast<SmallSmem *>(raw);
    load_weights(p, s);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const float a0 = __ldg(p.a0), aR = __ldg(p.aR);
    // S = sum |(out - y) / y| over all rows: every CTA adds the forward kernel's per-CTA partials itself, in the same fixed
    // order (lane l takes partials l, l + 32, ..., then a butterfly), so no launch sits between forward and backward
    // (num_partial_s == 0: data-parallel step — `sums` already holds the GLOBAL (S, N), all-reduced by the caller)
    if (warp == 0) {
        float tsum = 0.f;
        for (int i = lane; i < num_partial_s; i += 32) tsum += partial_s[i];
        tsum = warp_sum_all(tsum);
        if (lane == 0) {
            s.red[0][0] = num_partial_s > 0 ? tsum : sums[0];
            s.red[0][1] = num_partial_s > 0 ? static_cast<float>(p.np) : sums[1];
        }
    }
    __syncthreads();
    const float S = s.red[0][0], N = s.red[0][1];
    __syncthreads();
    // seed of d sqrt(100 S / N) / d out_i  (hgin_sqrt_mape_bwd)
    const float mape = 100.0f * (S / N);
    const float L = sqrtf(mape);
    const float cseed = 50.0f / (N * L);
    if (blockIdx.x == 0 && t == 0) {
        if (num_partial_s > 0) {
            sums[0] = S;
            sums[1] = N;
        }
        loss_out[0] = mape;
        loss_out[1] = L;
    }

    // weight-gradient accumulators: every thread owns fixed elements for the whole kernel
    float acc2[16], acc1[16], acc0[2] = {0.f, 0.f}, accb = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc2[i] = acc1[i] = 0.f;
    const int k2 = t & 127, n2base = (t >> 7) * 16;       // dW2[n2base + i][k2]
    const int n1 = t & 127, k1base = (t >> 7) * 16;       // dW1[n1][k1base + i]
    // per-warp accumulators (reduced across warps in a fixed order at the end)
    float accW3 = 0.f, db3 = 0.f, daR = 0.f, da0 = 0.f, deps = 0.f;

    int beg, end;
    cta_rows(p, beg, end);
    for (int t0 = beg; t0 < end; t0 += TILE) {
        const int rows_here = min(TILE, end - t0);
        const int r0 = warp * SS_R, row0 = t0 + r0;
        if (row0 < end) {
            const RowBatch<SS_R> st = rows_reload<SS_R, D>(d, s, act, row0, p.np, r0, lane, aR);
            float g[SS_R];
#pragma unroll
            for (int q = 0; q < SS_R; ++q) {
                const int row = min(row0 + q, p.np - 1);
                const float yi = __ldg(p.y + row);
                const float u = (st.out[q] - yi) / yi;
                g[q] = (row0 + q < end) ? cseed * ((u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f)) / yi : 0.f;   // masked rows: 0
                // head and readout layer 2
                accW3 = fmaf(g[q], st.a2[q], accW3);
                if (lane == 0) db3 += g[q];
                const float da2 = lane < d.n2() ? g[q] * s.W3[lane] : 0.f;
                if (!(st.z2[q] > 0.f)) daR = fmaf(da2, st.z2[q], daR);
                s.Dz2[r0 + q][lane] = st.z2[q] > 0.f ? da2 : aR * da2;
            }
            __syncwarp();
            // readout layer 1: d a1[k] = sum_n dz2[n] W2[n][k],  k = lane + 32 j
            float dz1[SS_R][4];
#pragma unroll
            for (int q = 0; q < SS_R; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) dz1[q][j] = 0.f;
#pragma unroll 2
            for (int n = 0; n < SS_MAX_N2; n += 4) {      // (Dz2 is zero beyond n2)
                float4 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4 *>(&s.W2kn[lane + 32 * j][n]);
#pragma unroll
                for (int q = 0; q < SS_R; ++q) {
                    const float4 dv = *reinterpret_cast<const float4 *>(&s.Dz2[r0 + q][n]);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dz1[q][j] = fmaf(dv.w, w[j].w, fmaf(dv.z, w[j].z, fmaf(dv.y, w[j].y, fmaf(dv.x, w[j].x, dz1[q][j]))));
                }
            }
#pragma unroll
            for (int q = 0; q < SS_R; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float dv = (lane + 32 * j) < d.n1() ? dz1[q][j] : 0.f;
                    if (!(st.z1[q][j] > 0.f)) daR = fmaf(dv, st.z1[q][j], daR);
                    dv = st.z1[q][j] > 0.f ? dv : aR * dv;
                    dz1[q][j] = dv;
                    s.Dz1[r0 + q][lane + 32 * j] = dv;
                }
            // d e[j] = sum_n dz1[n] W1[n][j]  (the first emb input columns of readout layer 1)
            float de[SS_R];
#pragma unroll
            for (int q = 0; q < SS_R; ++q) de[q] = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < d.emb(); j4 += 4) {      // four input columns j at a time
                float part[SS_R][4];
#pragma unroll
                for (int q = 0; q < SS_R; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) part[q][c] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = *reinterpret_cast<const float4 *>(&s.W1nk[lane + 32 * i][j4]);
#pragma unroll
                    for (int q = 0; q < SS_R; ++q) {
                        part[q][0] = fmaf(dz1[q][i], w.x, part[q][0]);
                        part[q][1] = fmaf(dz1[q][i], w.y, part[q][1]);
                        part[q][2] = fmaf(dz1[q][i], w.z, part[q][2]);
                        part[q][3] = fmaf(dz1[q][i], w.w, part[q][3]);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int q = 0; q < SS_R; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) part[q][c] += __shfl_xor_sync(0xffffffffu, part[q][c], o);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (lane == j4 + c) {
#pragma unroll
                        for (int q = 0; q < SS_R; ++q) de[q] = part[q][c];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < SS_R; ++q) {
                if (lane < d.emb() && !(st.z0[q] > 0.f)) da0 = fmaf(de[q], st.z0[q], da0);
                s.Dz0[r0 + q][lane] = lane < d.emb() ? (st.z0[q] > 0.f ? de[q] : a0 * de[q]) : 0.f;
            }
            __syncwarp();
            // d eps = sum_c dh_self[c] * x_path[c],  dh_self[c] = sum_n dz0[n] W0[n][fl + c]
            if (lane < d.fp()) {
#pragma unroll
                for (int q = 0; q < SS_R; ++q) {
                    float dh = 0.f;
#pragma unroll
                    for (int n = 0; n < d.emb(); ++n) dh = fmaf(s.Dz0[r0 + q][n], s.W0[n][d.fl() + lane], dh);
                    const int row = min(row0 + q, p.np - 1);
                    deps = fmaf(dh, __ldg(p.xp + static_cast<int64_t>(row) * p.ldp + p.pcol[lane]), deps);
                }
            }
        }
        __syncthreads();
        // outer products over the staged rows (slots of warps that had no rows hold stale data: bounded by rows_here)
        for (int r = 0; r < rows_here; ++r) {
            const float a = s.A1[r][k2];
            const float dv = s.Dz1[r][n1];
#pragma unroll
            for (int i4 = 0; i4 < 16; i4 += 4) {          // the broadcast operands four at a time (LDS.128)
                const float4 dz = *reinterpret_cast<const float4 *>(&s.Dz2[r][n2base + i4]);
                const float4 xi = *reinterpret_cast<const float4 *>(&s.Xin[r][k1base + i4]);
                acc2[i4 + 0] = fmaf(dz.x, a, acc2[i4 + 0]);
                acc2[i4 + 1] = fmaf(dz.y, a, acc2[i4 + 1]);
                acc2[i4 + 2] = fmaf(dz.z, a, acc2[i4 + 2]);
                acc2[i4 + 3] = fmaf(dz.w, a, acc2[i4 + 3]);
                acc1[i4 + 0] = fmaf(dv, xi.x, acc1[i4 + 0]);
                acc1[i4 + 1] = fmaf(dv, xi.y, acc1[i4 + 1]);
                acc1[i4 + 2] = fmaf(dv, xi.z, acc1[i4 + 2]);
                acc1[i4 + 3] = fmaf(dv, xi.w, acc1[i4 + 3]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = t + i * SS_THREADS;
                acc0[i] = fmaf(s.Dz0[r][idx / SS_MAX_K0], s.H[r][idx % SS_MAX_K0], acc0[i]);
            }
            if (t < 128) accb += s.Dz1[r][t];
            else if (t < 160) accb += s.Dz2[r][t - 128];
            else if (t < 192) accb += s.Dz0[r][t - 160];
        }
        __syncthreads();
    }
    // ---- per-CTA partials ----
    float *dst = partials + static_cast<int64_t>(blockIdx.x) * P_TOTAL;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        dst[P_W2 + (n2base + i) * SS_MAX_N1 + k2] = acc2[i];
        dst[P_W1 + n1 * SS_MAX_K1 + k1base + i] = acc1[i];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) dst[P_W0 + t + i * SS_THREADS] = acc0[i];
    if (t < 128) dst[P_B1 + t] = accb;
    else if (t < 160) dst[P_B2 + t - 128] = accb;
    else if (t < 192) dst[P_B0 + t - 160] = accb;
    // per-warp accumulators: lanes -> warps -> CTA in a fixed order
    s.red[warp][lane] = accW3;
    const float v_db3 = warp_sum_all(db3), v_aR = warp_sum_all(daR), v_a0 = warp_sum_all(da0), v_eps = warp_sum_all(deps);
    if (lane == 0) {
        s.red[warp][32] = v_db3;
        s.red[warp][33] = v_aR;
        s.red[warp][34] = v_a0;
        s.red[warp][35] = v_eps;
    }
    __syncthreads();
    if (t < 36) {
        float v = 0.f;
        for (int w = 0; w < SS_WARPS; ++w) v += s.red[w][t];
        if (t < 32) dst[P_W3 + t] = v;
        else dst[P_SC + t - 32] = v;
    }
}

// data-parallel forward phase: this rank's (S, N) from the forward kernel's partials, to be all-reduced by the caller
__global__ void small_local_sums_kernel(const float *__restrict__ partial_s, int count, int np, float *__restrict__ sums) {
    float t = 0.f;
    for (int i = threadIdx.x; i < count; i += 32) t += partial_s[i];
    t = warp_sum_all(t);
    if (threadIdx.x == 0) {
        sums[0] = t;
        sums[1] = static_cast<float>(np);
    }
}

struct SmallGrads {
    float *dW0, *db0, *da0, *deps0, *dW1, *db1, *daR, *dW2, *db2, *dW3, *db3;
};

// element i of the padded partial layout -> the parameter gradient it belongs to.  One CTA per 32 consecutive elements:
// warp w adds the partials of CTAs w, w + 8, ... (lanes = elements, coalesced), the eight sums are combined in a fixed order.
__global__ void __launch_bounds__(256)
small_reduce_kernel(const float *__restrict__ partials, int ctas, int emb, int k0, int k1, int n1, int n2, SmallGrads g) {
    __shared__ float red[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (i < P_TOTAL) {
#pragma unroll 4
        for (int c = warp; c < ctas; c += 8) acc += partials[static_cast<int64_t>(c) * P_TOTAL + i];
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp != 0 || i >= P_TOTAL) return;
    float *dst = nullptr;
    if (i < P_W1) {
        const int n = (i - P_W2) / SS_MAX_N1, k = (i - P_W2) % SS_MAX_N1;
        if (n < n2 && k < n1) dst = g.dW2 + n * n1 + k;
    } else if (i < P_W0) {
        const int n = (i - P_W1) / SS_MAX_K1, k = (i - P_W1) % SS_MAX_K1;
        if (n < n1 && k < k1) dst = g.dW1 + n * k1 + k;
    } else if (i < P_B1) {
        const int n = (i - P_W0) / SS_MAX_K0, k = (i - P_W0) % SS_MAX_K0;
        if (n < emb && k < k0) dst = g.dW0 + n * k0 + k;
    } else if (i < P_B2) {
        if (i - P_B1 < n1) dst = g.db1 + (i - P_B1);
    } else if (i < P_B0) {
        if (i - P_B2 < n2) dst = g.db2 + (i - P_B2);
    } else if (i < P_W3) {
        if (i - P_B0 < emb) dst = g.db0 + (i - P_B0);
    } else if (i < P_SC) {
        if (i - P_W3 < n2) dst = g.dW3 + (i - P_W3);
    } else {
        const int j = i - P_SC;
        dst = j == 0 ? g.db3 : (j == 1 ? g.daR : (j == 2 ? g.da0 : g.deps0));
    }
    if (dst == nullptr) return;
    *dst = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
}

inline int small_ctas(int64_t np) {      // two CTAs per SM (100 KB of shared memory each), each with a contiguous row range
    const int64_t tiles = ceil_div(np, SS_TILE);
    return static_cast<int>(tiles < 1 ? 1 : (tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs));
}

}  // namespace
}  // namespace hgin

using namespace hgin;

extern "C" int64_t hgin_small_step_workspace_bytes(int64_t num_paths) {
    return align_up(static_cast<int64_t>(small_ctas(num_paths)) * P_TOTAL * 4, 256) + 4096 + num_paths * A_ROW * 4;
}

static int32_t small_step_impl(int phase, int64_t num_paths, const int32_t *rowptr, const int32_t *col, const float *x_path,
                                   int64_t ld_path, int32_t f_path, const int32_t *path_cols_host, const float *x_link,
                                   int64_t ld_link, int32_t f_link, const int32_t *link_cols_host, const float *y, int32_t emb,
                                   int32_t n1, int32_t n2, int32_t concat_path, const float *W0, const float *b0,
                                   const float *alpha0, const float *eps0, const float *W1, const float *b1,
                                   const float *alpha_r, const float *W2, const float *b2, const float *W3, const float *b3,
                                   float *dW0, float *db0, float *dalpha0, float *deps0, float *dW1, float *db1, float *dalpha_r,
                                   float *dW2, float *db2, float *dW3, float *db3, float *sums, float *loss_out, float *out,
                                   void *workspace, int64_t workspace_bytes, void *stream) {
    const int k0 = f_link + f_path, k1 = emb + (concat_path ? f_path : 0);
    if (!(num_paths > 0 && num_paths < INT32_MAX && f_path >= 1 && f_path <= 8 && f_link >= 1 && f_link <= 8 && k0 <= SS_MAX_K0 &&
          emb >= 1 && emb <= SS_MAX_EMB && k1 <= SS_MAX_K1 && n1 >= 1 && n1 <= SS_MAX_N1 && n2 >= 1 && n2 <= SS_MAX_N2))
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_small_step: needs f_path, f_link <= 8, emb <= %d, emb + f_path <= %d, n1 <= %d, "
                    "n2 <= %d", SS_MAX_EMB, SS_MAX_K1, SS_MAX_N1, SS_MAX_N2);
    HGIN_CHECK_ARG(rowptr && x_path && x_link && y && path_cols_host && link_cols_host && ld_path >= 1 && ld_link >= 1,
                   "hgin_small_step: null input");
    HGIN_CHECK_ARG(W0 && b0 && alpha0 && eps0 && W1 && b1 && alpha_r && W2 && b2 && W3 && b3, "hgin_small_step: null parameter");
    HGIN_CHECK_ARG(dW0 && db0 && dalpha0 && deps0 && dW1 && db1 && dalpha_r && dW2 && db2 && dW3 && db3 && sums && loss_out,
                   "hgin_small_step: null gradient / loss output");
    if (!workspace || workspace_bytes < hgin_small_step_workspace_bytes(num_paths))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_small_step: workspace %lld < %lld", (long long)workspace_bytes,
                    (long long)hgin_small_step_workspace_bytes(num_paths));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SmallParams p{};
    p.np = static_cast<int>(num_paths);
    p.rowptr = rowptr;
    p.col = col;
    p.xp = x_path; p.ldp = static_cast<int>(ld_path); p.fp = f_path;
    p.xl = x_link; p.ldl = static_cast<int>(ld_link); p.fl = f_link;
    for (int i = 0; i < 8; ++i) {
        p.pcol[i] = i < f_path ? path_cols_host[i] : 0;
        p.lcol[i] = i < f_link ? link_cols_host[i] : 0;
        HGIN_CHECK_ARG(p.pcol[i] >= 0 && p.pcol[i] < ld_path && p.lcol[i] >= 0 && p.lcol[i] < ld_link, "hgin_small_step: column index");
    }
    p.y = y;
    p.emb = emb; p.n1 = n1; p.n2 = n2; p.concat = concat_path;
    p.W0 = W0; p.b0 = b0; p.a0 = alpha0; p.eps0 = eps0; p.W1 = W1; p.b1 = b1; p.aR = alpha_r; p.W2 = W2; p.b2 = b2; p.W3 = W3; p.b3 = b3;
    static bool attr = false;
    if (!attr) {
        const int bytes = static_cast<int>(sizeof(SmallSmem));
#define HGIN_SMALL_ATTR(R, D)                                                                                  \
    cudaFuncSetAttribute(small_fwd_loss_kernel<R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);     \
    cudaFuncSetAttribute(small_bwd_kernel<R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
        HGIN_SMALL_ATTR(2, DynDims); HGIN_SMALL_ATTR(3, DynDims); HGIN_SMALL_ATTR(4, DynDims);
        HGIN_SMALL_ATTR(2, DefaultDims); HGIN_SMALL_ATTR(3, DefaultDims); HGIN_SMALL_ATTR(4, DefaultDims);
#undef HGIN_SMALL_ATTR
        attr = true;
    }
    const int ctas = small_ctas(num_paths);
    float *partials = static_cast<float *>(workspace);
    float *partial_s = partials + static_cast<int64_t>(ctas) * P_TOTAL;
    float *act = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) +
                                           align_up(static_cast<int64_t>(ctas) * P_TOTAL * 4, 256) + 4096);
    // rows a warp carries together: whole rounds of 8 * R rows over each CTA's range, weighted by the (sub-linear) cost of a round
    const int64_t per = ceil_div(num_paths, ctas);
    int best_r = 4;
    int64_t best_cost = INT64_MAX;
    for (int r = 4; r >= 2; --r) {
        const int64_t cost = ceil_div(per, SS_WARPS * r) * (r + 1);
        if (cost < best_cost) { best_cost = cost; best_r = r; }
    }
    const bool is_default = emb == 8 && f_link == 3 && f_path == 3 && n1 == 128 && n2 == 32 && concat_path;
    // phase 0: the whole step; 1: forward + this rank's (S, N) only; 2: backward with the global (S, N) given in `sums`
#define HGIN_SMALL(R, D)                                                                                     \
    do {                                                                                                     \
        if (phase != 2) small_fwd_loss_kernel<R, D><<<ctas, SS_THREADS, sizeof(SmallSmem), s>>>(p, partial_s, out, act); \
        if (phase == 1) small_local_sums_kernel<<<1, 32, 0, s>>>(partial_s, ctas, p.np, sums);               \
        if (phase != 1)                                                                                      \
            small_bwd_kernel<R, D><<<ctas, SS_THREADS, sizeof(SmallSmem), s>>>(p, partial_s, phase == 0 ? ctas : 0, sums, \
                                                                                loss_out, partials, act);     \
    } while (0)
#define HGIN_SMALL_R(D)                     \
    do {                                    \
        if (best_r == 4) HGIN_SMALL(4, D);  \
        else if (best_r == 3) HGIN_SMALL(3, D); \
        else HGIN_SMALL(2, D);              \
    } while (0)
    if (is_default) HGIN_SMALL_R(DefaultDims);
    else HGIN_SMALL_R(DynDims);
#undef HGIN_SMALL_R
#undef HGIN_SMALL
    if (phase != 1) {
        const SmallGrads g{dW0, db0, dalpha0, deps0, dW1, db1, dalpha_r, dW2, db2, dW3, db3};
        small_reduce_kernel<<<static_cast<int>(ceil_div(P_TOTAL, 32)), 256, 0, s>>>(partials, ctas, emb, k0, k1, n1, n2, g);
    }
    HGIN_CHECK_LAUNCH("hgin_small_step");
    return HGIN_OK;
}

extern "C" int32_t hgin_small_step(int64_t num_paths, const int32_t *rowptr, const int32_t *col, const float *x_path,
                                   int64_t ld_path, int32_t f_path, const int32_t *path_cols_host, const float *x_link,
                                   int64_t ld_link, int32_t f_link, const int32_t *link_cols_host, const float *y, int32_t emb,
                                   int32_t n1, int32_t n2, int32_t concat_path, const float *W0, const float *b0,
                                   const float *alpha0, const float *eps0, const float *W1, const float *b1,
                                   const float *alpha_r, const float *W2, const float *b2, const float *W3, const float *b3,
                                   float *dW0, float *db0, float *dalpha0, float *deps0, float *dW1, float *db1, float *dalpha_r,
                                   float *dW2, float *db2, float *dW3, float *db3, float *sums, float *loss_out, float *out,
                                   void *workspace, int64_t workspace_bytes, void *stream) {
    return small_step_impl(0, num_paths, rowptr, col, x_path, ld_path, f_path, path_cols_host, x_link, ld_link, f_link,
                           link_cols_host, y, emb, n1, n2, concat_path, W0, b0, alpha0, eps0, W1, b1, alpha_r, W2, b2, W3, b3, dW0,
                           db0, dalpha0, deps0, dW1, db1, dalpha_r, dW2, db2, dW3, db3, sums, loss_out, out, workspace,
                           workspace_bytes, stream);
}

extern "C" int32_t hgin_small_step_phase(int32_t phase, int64_t num_paths, const int32_t *rowptr, const int32_t *col,
                                         const float *x_path, int64_t ld_path, int32_t f_path, const int32_t *path_cols_host,
                                         const float *x_link, int64_t ld_link, int32_t f_link, const int32_t *link_cols_host,
                                         const float *y, int32_t emb, int32_t n1, int32_t n2, int32_t concat_path,
                                         const float *W0, const float *b0, const float *alpha0, const float *eps0,
                                         const float *W1, const float *b1, const float *alpha_r, const float *W2,
                                         const float *b2, const float *W3, const float *b3, float *dW0, float *db0,
                                         float *dalpha0, float *deps0, float *dW1, float *db1, float *dalpha_r, float *dW2,
                                         float *db2, float *dW3, float *db3, float *sums, float *loss_out, float *out,
                                         void *workspace, int64_t workspace_bytes, void *stream) {
    HGIN_CHECK_ARG(phase == 1 || phase == 2, "hgin_small_step_phase: phase must be 1 (forward) or 2 (backward)");
    return small_step_impl(phase, num_paths, rowptr, col, x_path, ld_path, f_path, path_cols_host, x_link, ld_link, f_link,
                           link_cols_host, y, emb, n1, n2, concat_path, W0, b0, alpha0, eps0, W1, b1, alpha_r, W2, b2, W3, b3, dW0,
                           db0, dalpha0, deps0, dW1, db1, dalpha_r, dW2, db2, dW3, db3, sums, loss_out, out, workspace,
                           workspace_bytes, stream);
}

// K1 / K4 — segmented neighbour sum over CSR rows fused with the GIN self term.
//
// Forward  (models.py:208-215):  h[d] = sum_{e in row d} x_src[col[e]]  (+|concat)  (1+eps) * x_dst[d]
// Backward (autograd of the same, train.py:43): the identical kernel over the TRANSPOSED CSR
//          gathers dh rows into dx_src, fused with the (1+eps)*dh self branch and, through
//          `accumulate`, with the sum over relations that share a node type.
//
// Roofline: HBM.  Algorithmic bytes per row = sum_e (F*4 + 4) + 4 + F_self*4 + F_out*4
// (SURVEY §8(d)).  No atomics, no materialised x_j: each row is owned by one group of LPR
// lanes that walks its neighbours left to right (CSR order == stable edge order, so the fp32
// result is bit-identical to the CPU reference's scatter_add_), each lane carrying VEC
// consecutive features so that a full warp issues one 512-byte row read per neighbour at F=128.
// Neighbour indices are fetched LPR at a time with one coalesced load and handed round by
// shuffle; the gathers of a batch are issued back to back (UNROLL in flight) before the
// dependent adds.
#include <cuda_bf16.h>

#include "hgin_common.cuh"
#include <cstdlib>

#include "gin_scatter_blocks.cuh"
#include "gin_stage_blocks.cuh"

namespace hgin {
namespace {

using bf16 = __nv_bfloat16;

// Storage type T of the row matrices (x_src, x_self, post.z, out): float, or bf16 (HGIN_DTYPE_BF16: rows are
// stored in 16 bits, every addition is still an fp32 addition in CSR order, the result is rounded once on
// the store).  Raw<T, VEC> is what ONE vector load returns — 16 bytes for the vector variants (4 floats or
// 8 bf16) — and stays packed in registers until it is consumed, so the gathers in flight cost the same
// registers in both types.
template <typename T, int VEC>
struct Raw;
template <>
struct Raw<float, 1> { float v[1]; };
template <>
struct Raw<float, 4> { float v[4]; };
template <>
struct Raw<bf16, 1> { unsigned short u; };
template <>
struct Raw<bf16, 8> { uint4 q; };

__device__ __forceinline__ void bf16x2_to_f32(uint32_t w, float &lo, float &hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t f32x2_to_bf16(float lo, float hi) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}

__device__ __forceinline__ void unpack(const Raw<float, 1> &r, float (&o)[1]) { o[0] = r.v[0]; }
__device__ __forceinline__ void unpack(const Raw<float, 4> &r, float (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = r.v[i];
}
__device__ __forceinline__ void unpack(const Raw<bf16, 1> &r, float (&o)[1]) { o[0] = __uint_as_float(static_cast<uint32_t>(r.u) << 16); }
__device__ __forceinline__ void unpack(const Raw<bf16, 8> &r, float (&o)[8]) {
    bf16x2_to_f32(r.q.x, o[0], o[1]);
    bf16x2_to_f32(r.q.y, o[2], o[3]);
    bf16x2_to_f32(r.q.z, o[4], o[5]);
    bf16x2_to_f32(r.q.w, o[6], o[7]);
}
__device__ __forceinline__ void pack(const float (&o)[1], Raw<float, 1> &r) { r.v[0] = o[0]; }
__device__ __forceinline__ void pack(const float (&o)[4], Raw<float, 4> &r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = o[i];
}
__device__ __forceinline__ void pack(const float (&o)[1], Raw<bf16, 1> &r) { r.u = __bfloat16_as_ushort(__float2bfloat16_rn(o[0])); }
__device__ __forceinline__ void pack(const float (&o)[8], Raw<bf16, 8> &r) {
    r.q.x = f32x2_to_bf16(o[0], o[1]);
    r.q.y = f32x2_to_bf16(o[2], o[3]);
    r.q.z = f32x2_to_bf16(o[4], o[5]);
    r.q.w = f32x2_to_bf16(o[6], o[7]);
}

__device__ __forceinline__ Raw<float, 1> load_raw(const float *p, Raw<float, 1> *) { return {{__ldg(p)}}; }
__device__ __forceinline__ Raw<float, 4> load_raw(const float *p, Raw<float, 4> *) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    return {{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ Raw<bf16, 1> load_raw(const bf16 *p, Raw<bf16, 1> *) {
    return {__ldg(reinterpret_cast<const unsigned short *>(p))};
}
__device__ __forceinline__ Raw<bf16, 8> load_raw(const bf16 *p, Raw<bf16, 8> *) { return {__ldg(reinterpret_cast<const uint4 *>(p))}; }

// plain (coherent) loads: the accumulate path reads `out`, which this kernel also writes
__device__ __forceinline__ Raw<float, 1> load_raw_coherent(const float *p, Raw<float, 1> *) { return {{*p}}; }
__device__ __forceinline__ Raw<float, 4> load_raw_coherent(const float *p, Raw<float, 4> *) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    return {{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ Raw<bf16, 1> load_raw_coherent(const bf16 *p, Raw<bf16, 1> *) {
    return {*reinterpret_cast<const unsigned short *>(p)};
}
__device__ __forceinline__ Raw<bf16, 8> load_raw_coherent(const bf16 *p, Raw<bf16, 8> *) { return {*reinterpret_cast<const uint4 *>(p)}; }

// Streaming variants for rows that are touched exactly once (self rows, post-activation rows, the
// output): they must not displace the gathered source rows, which ARE reused (each link row ~35
// times per topology), from L1.
__device__ __forceinline__ Raw<float, 1> load_raw_stream(const float *p, Raw<float, 1> *) { return {{__ldg(p)}}; }
__device__ __forceinline__ Raw<float, 4> load_raw_stream(const float *p, Raw<float, 4> *) {
    Raw<float, 4> r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
        : "l"(p));
    return r;
}
__device__ __forceinline__ Raw<bf16, 1> load_raw_stream(const bf16 *p, Raw<bf16, 1> *t) { return load_raw(p, t); }
__device__ __forceinline__ Raw<bf16, 8> load_raw_stream(const bf16 *p, Raw<bf16, 8> *) {
    Raw<bf16, 8> r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
        : "=r"(r.q.x), "=r"(r.q.y), "=r"(r.q.z), "=r"(r.q.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ void store_raw(float *p, const Raw<float, 1> &r) { p[0] = r.v[0]; }
__device__ __forceinline__ void store_raw(float *p, const Raw<float, 4> &r) {
    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void store_raw(bf16 *p, const Raw<bf16, 1> &r) { *reinterpret_cast<unsigned short *>(p) = r.u; }
__device__ __forceinline__ void store_raw(bf16 *p, const Raw<bf16, 8> &r) { *reinterpret_cast<uint4 *>(p) = r.q; }
__device__ __forceinline__ void store_raw_stream(float *p, const Raw<float, 1> &r) { p[0] = r.v[0]; }
__device__ __forceinline__ void store_raw_stream(float *p, const Raw<float, 4> &r) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                 "f"(r.v[3])
                 : "memory");
}
__device__ __forceinline__ void store_raw_stream(bf16 *p, const Raw<bf16, 1> &r) { store_raw(p, r); }
__device__ __forceinline__ void store_raw_stream(bf16 *p, const Raw<bf16, 8> &r) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(r.q.x), "r"(r.q.y), "r"(r.q.z),
                 "r"(r.q.w)
                 : "memory");
}

__device__ __forceinline__ float ld1(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const bf16 *p) {
    return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
}
__device__ __forceinline__ float ld1_coherent(const float *p) { return *p; }
__device__ __forceinline__ float ld1_coherent(const bf16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float *p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16 *p, float v) { *p = __float2bfloat16_rn(v); }

// LPR lanes per row, VEC features per lane per chunk, NC chunks per lane:
// covers f_src <= LPR * VEC * NC.
// POST: the row result is a gradient w.r.t. the OUTPUT of the layer below; the kernel multiplies
// it by that layer's activation derivative (reading its saved pre-activation `post.z`) and so
// writes dz directly, plus the per-CTA partial of dalpha = sum grad * min(z, 0) — the separate
// dz = g * act'(z) pass over the row-sized tensors of the layer below disappears.
struct PostAct {
    const void *z;            // same storage type as the rows
    int64_t ldz;
    int act;
    const float *alpha;
    float *dalpha_partials;   // [gridDim.x]
    // d(eps) of the relation whose self branch rides on this pass: sum x_self * act(z) — the rows of
    // x_self are dh_self and act(z) IS x_dst (the output of the layer below), so the input-gradient
    // GEMM above no longer has to read x_dst for it.
    float *ddot_partials;     // [gridDim.x] or NULL
    // MODE 2 (forward, "pre-activation inputs"): x_src / x_self hold the PRE-activations z of the layer
    // below and act(z) is applied to every element as it is loaded, so that layer never has to write
    // its activated output (hgin_gin_combine_pre).
    int src_act;
    const float *src_alpha;
    int self_act;
    const float *self_alpha;
};

// Row indices and leading dimensions are 32-bit inside the kernel (the dispatch checks the ranges), so that
// every row address is ONE widening multiply-add (IMAD.WIDE) instead of a 64 x 64-bit product — the
// short-row launches issue ~60 % of their scheduler slots, and integer address math was most of it.
// FULL: f_src == LPR * VEC * NC, the feature-range predicates fold away.
// Block-diagonal batches (one block per topology sample): block b owns source rows [in_ptr[b], in_ptr[b+1]) and output rows
// [out_ptr[b], out_ptr[b+1]).  gate: the counters of block_gate_kernel (gin_scatter_blocks.cuh).  gate_mode selects which
// schedule the counters must allow for this launch to be SKIPPED (the other kernel of the pair does the work then):
//   1  input-major streaming (scatter_blocks_kernel): containment, ascending rows, output rows of a block <= gate_cap
//   2  the TABLE variant of this kernel: containment, source rows of a block <= gate_cap
struct BlockInfo {
    int num_blocks;
    const int64_t *in_ptr;
    const int64_t *out_ptr;
    const int32_t *gate;
    int gate_cap;
    int gate_mode;
};
__device__ __forceinline__ bool gate_allows(const BlockInfo &b) {
    if (b.gate == nullptr) return false;
    if (__ldg(b.gate) != 0) return false;
    if (b.gate_mode == 1) return __ldg(b.gate + 3) == 0 && __ldg(b.gate + 1) <= b.gate_cap;
    return __ldg(b.gate + 2) <= b.gate_cap;
}

// TABLE (short rows on block-diagonal batches; the north star's "neighbour rows staged through shared memory"): one
// 1024-thread CTA per SM takes one block at a time, copies the block's source rows (200 link rows = 100 KB in fp32) into
// shared memory once — applying the source pre-activation there, once per source row instead of once per gathered
// element — and every gather of the block's ~2450 output rows is then a shared-memory load: no L1/L2 gather traffic, no
// col -> gather latency through L2; the kernel is left with the streaming of the self / post / output rows.  Same CSR,
// same order of additions: bit-identical to the global-gather variant, which runs behind the inverse gate otherwise.
template <typename T, int VEC, int LPR, int NC, bool CONTIG, int MINB, int MODE, bool FULL, bool TABLE = false>
__global__ void __launch_bounds__(TABLE ? 1024 : 256, MINB)
gin_combine_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const T *__restrict__ x_src, int ld_src, int f_src,
                   const T *__restrict__ x_self, int ld_self, int f_self,
                   const float *__restrict__ eps_ptr, int self_mode, int accumulate,
                   T *__restrict__ out, int ld_out, const PostAct post, const BlockInfo blk) {
    using R = Raw<T, VEC>;
    // exactly one kernel of a (TABLE / streaming, global-gather) pair does the work; the other leaves zero partials
    if (TABLE ? !gate_allows(blk) : gate_allows(blk)) {
        if (MODE == 1 && threadIdx.x == 0) {
            if (post.dalpha_partials) post.dalpha_partials[blockIdx.x] = 0.0f;
            if (post.ddot_partials) post.ddot_partials[blockIdx.x] = 0.0f;
        }
        return;
    }
    extern __shared__ __align__(16) uint8_t table_raw[];
    T *table = reinterpret_cast<T *>(table_raw);
    // gathers in flight per lane before the dependent adds; bounded by the batch (LPR) and by
    // the register budget when a lane carries several chunks
    // (pre-activation sources need a few registers for the on-the-fly act: two gathers fewer in flight)
    // (bf16 lanes carry 8 accumulators and unpack 8 values per gather: fewer in flight at 64 registers)
    constexpr int UNROLL_MAX = (NC >= 4) ? 2 : ((NC == 2) ? 4 : (VEC == 8 ? ((MODE == 0 || MODE == 3) ? (CONTIG ? 4 : 6) : 4)
                                                                          : ((MODE == 2 || MODE == 4) ? 6 : 8)));
    // (TABLE: gathers are shared-memory loads, two in flight cover their latency — and the 1024-thread CTA has 64 registers)
    constexpr int UNROLL = TABLE ? ((NC >= 2) ? (MODE == 1 ? 1 : 2) : 4) : ((LPR < UNROLL_MAX) ? LPR : UNROLL_MAX);
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;   // lane inside the row group
    const int grp = lane / LPR;   // row group inside the warp
    const unsigned full = 0xffffffffu;
    // fl(1 + eps): the reference computes (1 + self.eps) as an fp32 tensor op (models.py:213/215).
    const float ope = __fadd_rn(1.0f, eps_ptr ? __ldg(eps_ptr) : 0.0f);
    constexpr bool POST = MODE == 1;
    constexpr bool PRE_SRC = MODE == 2 || MODE == 4;    // x_src holds pre-activations
    constexpr bool PRE_SELF = MODE == 3 || MODE == 4;   // x_self holds pre-activations
    const float post_alpha = (POST && post.act == HGIN_ACT_PRELU) ? __ldg(post.alpha) : 0.0f;
    // pre-activation inputs: x = z > 0 ? z : a * z with a = slope (PReLU) or 0 (ReLU)
    const float src_alpha = !PRE_SRC ? 1.0f : (post.src_act == HGIN_ACT_PRELU ? __ldg(post.src_alpha) : 0.0f);
    const float self_alpha = !PRE_SELF ? 1.0f : (post.self_act == HGIN_ACT_PRELU ? __ldg(post.self_alpha) : 0.0f);
    float dalpha = 0.0f, ddot = 0.0f;
    const T *post_z = static_cast<const T *>(post.z);
    // Row addresses: base + row * (row pitch in BYTES), both factors unsigned 32-bit, so that every address is ONE
    // IMAD.WIDE.U32 with the 64-bit base as its addend (rows are >= 0 where an address is formed; the dispatch bounds
    // the pitches).  The signed element-index form cost four instructions per gather, in a kernel that is issue-bound.
    auto at = [](const T *base, int row, uint32_t pitch) {
        return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) + static_cast<uint64_t>(static_cast<uint32_t>(row)) * pitch);
    };
    const uint32_t pitch_src = static_cast<uint32_t>(ld_src) * static_cast<uint32_t>(sizeof(T));
    const uint32_t pitch_self = static_cast<uint32_t>(ld_self) * static_cast<uint32_t>(sizeof(T));
    const uint32_t pitch_post = static_cast<uint32_t>(post.ldz) * static_cast<uint32_t>(sizeof(T));
    const uint32_t pitch_out = static_cast<uint32_t>(ld_out) * static_cast<uint32_t>(sizeof(T));

    // Row -> warp mapping.
    // CONTIG (short rows): each CTA owns a CONTIGUOUS block of rows, its warps interleaving inside it.
    //   Batched datanet graphs are block-diagonal with contiguous ids, so neighbouring rows gather
    //   from the same few hundred source rows (one topology: 200 link rows = 100 KB), which then stay
    //   resident in this SM's L1 instead of being re-fetched from L2 ~36 times each.
    // otherwise (long, heavy-tailed rows): rows are dealt round-robin over all warps of the grid so
    //   that the tail is spread evenly.
    const int warps_per_cta = blockDim.x >> 5;
    int cta_beg, cta_end, stride;
    int tbl0 = 0;     // first source row held by the table
    if (TABLE) {
        cta_beg = cta_end = 0;
        stride = warps_per_cta * ROWS_PER_WARP;
    } else if (CONTIG) {
        const int unit = warps_per_cta * ROWS_PER_WARP;
        const int rows_per_cta = ((num_rows + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) + unit - 1) / unit * unit;
        const int64_t b64 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
        cta_beg = b64 < num_rows ? static_cast<int>(b64) : num_rows;
        cta_end = min(b64 + rows_per_cta, static_cast<int64_t>(num_rows));
        stride = unit;
    } else {
        cta_beg = static_cast<int>(blockIdx.x) * warps_per_cta * ROWS_PER_WARP;
        cta_end = num_rows;
        stride = static_cast<int>(gridDim.x) * warps_per_cta * ROWS_PER_WARP;
    }

    // Software pipeline over this warp's rows.  The dependent chain rowptr -> col -> x_src[col] costs
    // three memory latencies; with ~3 neighbours per row (link->path) that chain, not bandwidth,
    // bounds the kernel.  So the row bounds are fetched two iterations ahead and the first batch of
    // neighbour indices one iteration ahead, leaving only the gather itself exposed.
    for (int bi = TABLE ? static_cast<int>(blockIdx.x) : 0; bi < (TABLE ? blk.num_blocks : 1); bi += TABLE ? static_cast<int>(gridDim.x) : 1) {
    if constexpr (TABLE) {
        tbl0 = static_cast<int>(__ldg(blk.in_ptr + bi));
        const int tbl_rows = static_cast<int>(__ldg(blk.in_ptr + bi + 1)) - tbl0;
        cta_beg = static_cast<int>(__ldg(blk.out_ptr + bi));
        cta_end = static_cast<int>(__ldg(blk.out_ptr + bi + 1));
        __syncthreads();                       // the previous block's gathers are done with the table
        const int ppr = f_src / VEC;           // (TABLE launches are full-width vector launches: f_src % VEC == 0)
        for (int i = threadIdx.x; i < tbl_rows * ppr; i += blockDim.x) {
            const int r = i / ppr, c = (i - r * ppr) * VEC;
            R v = load_raw(at(x_src + c, tbl0 + r, pitch_src), (R *)nullptr);
            if (PRE_SRC) {                     // act(z) once per source element (the gathers then add plain values)
                float t[VEC];
                unpack(v, t);
#pragma unroll
                for (int k = 0; k < VEC; ++k) t[k] = t[k] > 0.f ? t[k] : src_alpha * t[k];
                pack(t, v);
            }
            store_raw(table + r * f_src + c, v);
        }
        __syncthreads();
    }
    int row0 = cta_beg + static_cast<int>(threadIdx.x >> 5) * ROWS_PER_WARP;
    auto load_bounds = [&](int r0, int32_t &b, int32_t &l) {
        const int r = r0 + grp;
        b = 0;
        l = 0;
        if (r < cta_end && rowptr != nullptr) {   // rowptr == NULL: no edges at all (self term only)
            b = __ldg(rowptr + r);
            l = __ldg(rowptr + r + 1) - b;
        }
    };
    int32_t beg, len, nbeg, nlen;
    load_bounds(row0, beg, len);
    load_bounds(row0 + stride, nbeg, nlen);
    int32_t mine = (sub < len) ? __ldg(col + (beg + sub)) : -1;

    for (; row0 < cta_end; row0 += stride) {
        const int row = row0 + grp;
        const bool live = row < cta_end;
        // issue the prefetches for the following rows before touching this row's data
        const int32_t nmine = (sub < nlen) ? __ldg(col + (nbeg + sub)) : -1;
        int32_t nnbeg, nnlen;
        load_bounds(row0 + 2 * stride, nnbeg, nnlen);
        R self_r[NC];
        if (self_mode == HGIN_SELF_ADD && live) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int f = (c * LPR + sub) * VEC;
                if (FULL || f < f_src)
                    self_r[c] = CONTIG ? load_raw_stream(at(x_self + f, row, pitch_self), (R *)nullptr)
                                       : load_raw(at(x_self + f, row, pitch_self), (R *)nullptr);
            }
        }
        R post_r[NC];
        if (POST && post.act != HGIN_ACT_NONE && live) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int f = (c * LPR + sub) * VEC;
                if (FULL || f < f_src)
                    post_r[c] = CONTIG ? load_raw_stream(at(post_z + f, row, pitch_post), (R *)nullptr)
                                       : load_raw(at(post_z + f, row, pitch_post), (R *)nullptr);
            }
        }
        // warp-uniform trip count so the shuffles below are always convergent
        int32_t max_len = len;
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) max_len = max(max_len, __shfl_xor_sync(full, max_len, o));

        float acc[NC][VEC];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[c][i] = 0.0f;

        for (int32_t base = 0; base < max_len; base += LPR) {
            // one coalesced index load per group, LPR neighbours at a time (the first batch was prefetched)
            if (base > 0) mine = (base + sub < len) ? __ldg(col + (beg + base + sub)) : -1;
            const int32_t batch = min(LPR, max_len - base);
            for (int32_t j0 = 0; j0 < batch; j0 += UNROLL) {
                R v[UNROLL][NC];
                int32_t nb[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if constexpr (LPR % UNROLL == 0) {
                        // slot j0 + u < LPR: lanes past the row's end hold -1 already (batch <= max_len - base), no mask needed
                        nb[u] = __shfl_sync(full, mine, grp * LPR + j0 + u);
                    } else {
                        // (j0+u) % LPR keeps the source lane in range; out-of-batch slots are masked by nb < 0
                        const int32_t s = __shfl_sync(full, mine, grp * LPR + ((j0 + u) % LPR));
                        nb[u] = (j0 + u < batch) ? s : -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int f = (c * LPR + sub) * VEC;
                        if (nb[u] >= 0 && (FULL || f < f_src)) {
                            if constexpr (TABLE) v[u][c] = load_raw_coherent(table + (nb[u] - tbl0) * f_src + f, (R *)nullptr);
                            else v[u][c] = load_raw(at(x_src + f, nb[u], pitch_src), (R *)nullptr);
                        } else {
                            v[u][c] = R{};   // (zero-filled: keeps the unpack below unconditional, no spills)
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (nb[u] >= 0) {  // strictly left-to-right; skipped slots add nothing (not even +0)
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            float t[VEC];
                            unpack(v[u][c], t);
#pragma unroll
                            for (int i = 0; i < VEC; ++i) {
                                float tv = t[i];
                                if (PRE_SRC && !TABLE) tv = tv > 0.f ? tv : src_alpha * tv;   // x = act(z), on the fly
                                acc[c][i] = __fadd_rn(acc[c][i], tv);
                            }
                        }
                    }
                }
            }
        }

        if (live) {
            T *orow = const_cast<T *>(at(out, row, pitch_out));
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int f = (c * LPR + sub) * VEC;
                if (!FULL && f >= f_src) continue;
                float r[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) r[i] = acc[c][i];
                float sv[VEC];
                if (self_mode == HGIN_SELF_ADD) {
                    unpack(self_r[c], sv);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float xs = sv[i];
                        if (PRE_SELF) xs = xs > 0.f ? xs : self_alpha * xs;
                        r[i] = __fadd_rn(r[i], __fmul_rn(ope, xs));
                    }
                }
                if (accumulate) {
                    float old[VEC];
                    unpack(load_raw_coherent(orow + f, (R *)nullptr), old);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) r[i] = __fadd_rn(old[i], r[i]);
                }
                if (POST && post.act != HGIN_ACT_NONE) {
                    float pv[VEC];
                    unpack(post_r[c], pv);
                    if (post.ddot_partials && self_mode == HGIN_SELF_ADD) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) ddot = fmaf(sv[i], act_forward(pv[i], post.act, post_alpha), ddot);
                    }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float zv = pv[i];
                        if (post.act == HGIN_ACT_PRELU && !(zv > 0.f)) dalpha = fmaf(r[i], zv, dalpha);
                        r[i] = act_backward(r[i], zv, post.act, post_alpha);
                    }
                }
                R packed;
                pack(r, packed);
                if constexpr (CONTIG) store_raw_stream(orow + f, packed);
                else store_raw(orow + f, packed);
            }
            if (self_mode == HGIN_SELF_CONCAT) {
                // [agg | (1+eps) x_self]: the self block starts at column f_src (rarely 16B aligned) -> scalar
                for (int f = sub; f < f_self; f += LPR) {
                    float xs = ld1(at(x_self + f, row, pitch_self));
                    if (PRE_SELF) xs = xs > 0.f ? xs : self_alpha * xs;
                    float t = __fmul_rn(ope, xs);
                    if (accumulate) t = __fadd_rn(ld1_coherent(orow + f_src + f), t);
                    st1(orow + f_src + f, t);
                }
            }
        }
        beg = nbeg; len = nlen; mine = nmine;
        nbeg = nnbeg; nlen = nnlen;
    }
    }   // blocks (TABLE) / single pass
    if (POST && (post.dalpha_partials || post.ddot_partials)) {   // fixed association: lanes -> warps -> CTA partial
        __shared__ float red[32];
        if (post.dalpha_partials) {
            dalpha = block_sum(dalpha, red);
            if (threadIdx.x == 0) post.dalpha_partials[blockIdx.x] = dalpha;
        }
        if (post.ddot_partials) {
            ddot = block_sum(ddot, red);
            if (threadIdx.x == 0) post.ddot_partials[blockIdx.x] = ddot;
        }
    }
}

__global__ void __launch_bounds__(1024) combine_reduce_scalar_kernel(const float *__restrict__ v, int count,
                                                                     float *__restrict__ out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

constexpr int kMaxCombineCtas = kNumSMs * 33;      // 32 waves of the gather kernel + one wave of its TABLE twin
constexpr int kTableSmem = 220 * 1024;              // source rows of one block held in shared memory by the TABLE variant

template <typename T, int VEC, int LPR, int NC, bool CONTIG = false, int MINB = ((NC <= 1) ? 4 : 1), bool TABLE = false>
int launch(int64_t num_rows64, const int32_t *rowptr, const int32_t *col, const T *x_src, int64_t ld_src64,
           int f_src, const T *x_self, int64_t ld_self64, int f_self, const float *eps, int self_mode,
           int accumulate, T *out, int64_t ld_out64, const PostAct *post, cudaStream_t s, const BlockInfo blk = BlockInfo{}) {
    const int num_rows = static_cast<int>(num_rows64), ld_src = static_cast<int>(ld_src64);
    const int ld_self = static_cast<int>(ld_self64), ld_out = static_cast<int>(ld_out64);
    constexpr int threads = TABLE ? 1024 : 256;
    constexpr int rows_per_cta = (256 / 32) * (32 / LPR);
    // Grid-stride over rows with whole waves of CTAs: enough CTAs (32 per SM) that the hardware
    // scheduler evens out the heavy-tailed row lengths of the path->link relation.
    // TABLE: one persistent CTA per SM, blocks dealt round-robin.
    static const int waves_probe = getenv("HGIN_PROBE_WAVES") ? atoi(getenv("HGIN_PROBE_WAVES")) : 0;   // EXPERIMENT
    const int grid = TABLE ? kNumSMs : grid_for(num_rows, rows_per_cta, (CONTIG && waves_probe > 0) ? waves_probe : 32);
    const size_t smem = TABLE ? static_cast<size_t>(kTableSmem) : 0;
    const PostAct pa = post ? *post : PostAct{};
    const int mode = !post ? 0 : (post->z ? 1 : ((post->src_act != HGIN_ACT_NONE && post->self_act != HGIN_ACT_NONE) ? 4
                                                 : (post->src_act != HGIN_ACT_NONE ? 2 : 3)));
    const bool full = f_src == LPR * VEC * NC;
#define HGIN_GO(M, F)                                                                                              \
    do {                                                                                                           \
        auto kfn = gin_combine_kernel<T, VEC, LPR, NC, CONTIG, TABLE ? 1 : MINB, M, F, TABLE>;                     \
        if (TABLE) {                                                                                               \
            static bool attr_set = false;                                                                          \
            if (!attr_set) {                                                                                       \
                cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableSmem);                \
                attr_set = true;                                                                                   \
            }                                                                                                      \
        }                                                                                                          \
        kfn<<<grid, threads, smem, s>>>(num_rows, rowptr, col, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,   \
                                        self_mode, accumulate, out, ld_out, pa, blk);                               \
    } while (0)
#define HGIN_GO_MODE(M)          \
    do {                         \
        if (full) HGIN_GO(M, true); \
        else HGIN_GO(M, false);  \
    } while (0)
    switch (mode) {
        case 0: HGIN_GO_MODE(0); break;
        case 1: HGIN_GO_MODE(1); break;
        case 2: HGIN_GO_MODE(2); break;
        case 3: HGIN_GO_MODE(3); break;
        default: HGIN_GO_MODE(4); break;
    }
#undef HGIN_GO_MODE
#undef HGIN_GO
    return grid;
}

template <typename T>
struct VecOf;
template <>
struct VecOf<float> { static constexpr int value = 4; };
template <>
struct VecOf<bf16> { static constexpr int value = 8; };

template <typename T>
int32_t combine_dispatch_t(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges,
                           const T *x_src, int64_t ld_src, int32_t f_src, const T *x_self, int64_t ld_self,
                           int32_t f_self, const float *eps, int32_t self_mode, int32_t accumulate, T *out,
                           int64_t ld_out, const T *post_z, int64_t ld_post, int32_t post_act,
                           const float *post_alpha, float *post_dalpha, float *post_ddot, void *workspace,
                           int64_t workspace_bytes, void *stream, const char *who, int32_t src_act,
                           const float *src_alpha, int32_t self_act, const float *self_alpha,
                           const BlockInfo blk_in = BlockInfo{}, bool table = false) {
    HGIN_CHECK_ARG(num_rows >= 0 && num_rows < INT32_MAX - (1 << 22), "%s: bad num_rows %lld", who, (long long)num_rows);
    HGIN_CHECK_ARG(ld_src < INT32_MAX && ld_self < INT32_MAX && ld_out < INT32_MAX && ld_post < INT32_MAX,
                   "%s: leading dimensions must fit 32 bits", who);
    HGIN_CHECK_ARG(f_src > 0 && f_src <= 512, "%s: f_src must be in [1,512], got %d", who, f_src);
    HGIN_CHECK_ARG(self_mode >= HGIN_SELF_NONE && self_mode <= HGIN_SELF_CONCAT, "%s: bad self_mode %d", who, self_mode);
    HGIN_CHECK_ARG(self_mode == HGIN_SELF_NONE || x_self != nullptr, "%s: x_self is null", who);
    HGIN_CHECK_ARG(self_mode != HGIN_SELF_ADD || f_self == f_src, "%s: SELF_ADD needs f_self == f_src", who);
    HGIN_CHECK_ARG(self_mode != HGIN_SELF_CONCAT || f_self > 0, "%s: SELF_CONCAT needs f_self > 0", who);
    const bool post_on = post_z != nullptr && post_act != HGIN_ACT_NONE;
    const bool pre_on = src_act != HGIN_ACT_NONE || self_act != HGIN_ACT_NONE;
    HGIN_CHECK_ARG(src_act >= HGIN_ACT_NONE && src_act <= HGIN_ACT_RELU && self_act >= HGIN_ACT_NONE && self_act <= HGIN_ACT_RELU,
                   "%s: bad input activation", who);
    HGIN_CHECK_ARG((src_act != HGIN_ACT_PRELU || src_alpha) && (self_act != HGIN_ACT_PRELU || self_alpha),
                   "%s: PReLU input activation needs its slope", who);
    HGIN_CHECK_ARG(!(post_on && pre_on), "%s: input activations and a post-activation cannot be combined", who);
    HGIN_CHECK_ARG(post_act >= HGIN_ACT_NONE && post_act <= HGIN_ACT_RELU, "%s: bad post_act %d", who, post_act);
    HGIN_CHECK_ARG(!post_on || self_mode != HGIN_SELF_CONCAT, "%s: post-activation needs a [rows, f_src] result", who);
    HGIN_CHECK_ARG(!post_on || post_act != HGIN_ACT_PRELU || post_alpha, "%s: PReLU post-activation needs alpha", who);
    HGIN_CHECK_ARG(!post_on || ld_post >= f_src, "%s: ld_post too small", who);
    const bool want_dalpha = post_dalpha != nullptr;
    HGIN_CHECK_ARG(!post_ddot || (post_on && self_mode == HGIN_SELF_ADD), "%s: post_ddot needs a post-activation and SELF_ADD", who);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((want_dalpha && post_on && post_act == HGIN_ACT_PRELU) || post_ddot) {
        if (!workspace || workspace_bytes < static_cast<int64_t>(kMaxCombineCtas) * 8)
            return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "%s: workspace %lld < %lld bytes", who, (long long)workspace_bytes,
                        (long long)kMaxCombineCtas * 8);
    }
    if (want_dalpha && (num_rows == 0 || !(post_on && post_act == HGIN_ACT_PRELU))) cudaMemsetAsync(post_dalpha, 0, sizeof(float), s);
    if (post_ddot && num_rows == 0) cudaMemsetAsync(post_ddot, 0, sizeof(float), s);
    if (num_rows == 0) return HGIN_OK;
    // `rowptr` may be null when the relation has no edges at all (self term only), `col` for an
    // edgeless relation (every row empty): neither is dereferenced then.
    HGIN_CHECK_ARG(x_src && out, "%s: null pointer", who);
    HGIN_CHECK_ARG(rowptr || self_mode != HGIN_SELF_NONE, "%s: no adjacency and no self term", who);
    const int width = f_src + (self_mode == HGIN_SELF_CONCAT ? f_self : 0);
    HGIN_CHECK_ARG(ld_src >= f_src && ld_out >= width, "%s: leading dimension too small", who);
    PostAct post{post_z, ld_post, post_act, post_alpha,
                 (want_dalpha && post_act == HGIN_ACT_PRELU) ? static_cast<float *>(workspace) : nullptr,
                 post_ddot ? static_cast<float *>(workspace) + kMaxCombineCtas : nullptr,
                 src_act, src_alpha, self_act, self_alpha};
    if (pre_on) post.z = nullptr;            // MODE 2 is selected by a PostAct without z
    const PostAct *pp = (post_on || pre_on) ? &post : nullptr;
    if (!rowptr) num_edges = 0;

    // 128-bit lanes need every row start 16-byte aligned.
    constexpr int V = VecOf<T>::value;
    bool vec = (f_src % V == 0) && (ld_src % V == 0) && (ld_out % V == 0) && aligned16(x_src) && aligned16(out);
    if (self_mode == HGIN_SELF_ADD) vec = vec && (ld_self % V == 0) && aligned16(x_self);
    if (post_on) vec = vec && (ld_post % V == 0) && aligned16(post_z);

    int grid = 0;
#define HGIN_LAUNCH(VV, L, N)                                                                                      \
    grid = launch<T, VV, L, N>(num_rows, rowptr, col, x_src, ld_src, f_src, x_self, ld_self, f_self, eps, self_mode, \
                               accumulate, out, ld_out, pp, s, table ? BlockInfo{} : blk_in)
#define HGIN_LAUNCH_CONTIG(VV, L, N)                                                                                        \
    do {                                                                                                                    \
        if (table) {    /* the TABLE twin first (one wave), then the global-gather kernel behind the inverse gate */         \
            BlockInfo bt = blk_in;                                                                                          \
            bt.gate_mode = 2;                                                                                               \
            bt.gate_cap = kTableSmem / (f_src * static_cast<int>(sizeof(T)));                                               \
            launch<T, VV, L, N, true, 1, true>(num_rows, rowptr, col, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,      \
                                               self_mode, accumulate, out, ld_out, pp, s, bt);                              \
            PostAct behind = post;                                                                                          \
            if (behind.dalpha_partials) behind.dalpha_partials += kNumSMs;                                                  \
            if (behind.ddot_partials) behind.ddot_partials += kNumSMs;                                                      \
            grid = kNumSMs + launch<T, VV, L, N, true, 4>(num_rows, rowptr, col, x_src, ld_src, f_src, x_self, ld_self, f_self, \
                                                          eps, self_mode, accumulate, out, ld_out, pp ? &behind : nullptr, s, bt); \
        } else {                                                                                                            \
            grid = launch<T, VV, L, N, true, 4>(num_rows, rowptr, col, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,      \
                                                self_mode, accumulate, out, ld_out, pp, s, blk_in);                          \
        }                                                                                                                   \
    } while (0)
    // Lanes per row: a full warp per row suits long rows (path->link, ~36 neighbours); for short
    // rows (link->path, ~3 neighbours) the per-row latency chain rowptr -> col -> gather dominates,
    // so several rows share a warp and each lane carries more 128-bit chunks.
    const double avg_len = (num_edges >= 0 && num_rows > 0) ? static_cast<double>(num_edges) / num_rows : 1e9;
    const bool short_rows = avg_len <= 8.0;
    if (vec) {
        const int chunks = f_src / V;
        if constexpr (V == 4) {
            // Measured on B200 (profiles/): 2.5 M rows x ~3 neighbours, F = 128: warp/row 1.10 ms ->
            // half-warp/row + contiguous CTA ranges + pipelined indices 0.60 ms; long rows keep warp/row.
            if (chunks == 32 && short_rows) HGIN_LAUNCH_CONTIG(4, 16, 2);
            else if (chunks == 16 && short_rows) HGIN_LAUNCH_CONTIG(4, 8, 2);
            else if (chunks <= 1) HGIN_LAUNCH(4, 1, 1);
            else if (chunks <= 2) HGIN_LAUNCH(4, 2, 1);
            else if (chunks <= 4) HGIN_LAUNCH(4, 4, 1);
            else if (chunks <= 8) HGIN_LAUNCH(4, 8, 1);
            else if (chunks <= 16) HGIN_LAUNCH(4, 16, 1);
            else if (chunks <= 32) HGIN_LAUNCH(4, 32, 1);
            else if (chunks <= 64) HGIN_LAUNCH(4, 32, 2);
            else HGIN_LAUNCH(4, 32, 4);
        } else {
            // bf16 rows: a 128-wide row is 16 chunks of 8 — half a warp per row with one chunk per lane
            // carries as many rows per warp as the fp32 short-row variant at half the bytes.
            if (chunks == 16 && short_rows) HGIN_LAUNCH_CONTIG(8, 16, 1);
            else if (chunks == 8 && short_rows) HGIN_LAUNCH_CONTIG(8, 8, 1);
            else if (chunks <= 1) HGIN_LAUNCH(8, 1, 1);
            else if (chunks <= 2) HGIN_LAUNCH(8, 2, 1);
            else if (chunks <= 4) HGIN_LAUNCH(8, 4, 1);
            else if (chunks <= 8) HGIN_LAUNCH(8, 8, 1);
            else if (chunks <= 16) HGIN_LAUNCH(8, 16, 1);
            else if (chunks <= 32) HGIN_LAUNCH(8, 32, 1);
            else HGIN_LAUNCH(8, 32, 2);
        }
    } else {
        if (f_src <= 1) HGIN_LAUNCH(1, 1, 1);
        else if (f_src <= 2) HGIN_LAUNCH(1, 2, 1);
        else if (f_src <= 4) HGIN_LAUNCH(1, 4, 1);
        else if (f_src <= 8) HGIN_LAUNCH(1, 8, 1);
        else if (f_src <= 16) HGIN_LAUNCH(1, 16, 1);
        else if (f_src <= 32) HGIN_LAUNCH(1, 32, 1);
        else if (f_src <= 64) HGIN_LAUNCH(1, 32, 2);
        else if (f_src <= 128) HGIN_LAUNCH(1, 32, 4);
        else if (f_src <= 256) HGIN_LAUNCH(1, 32, 8);
        else HGIN_LAUNCH(1, 32, 16);
    }
#undef HGIN_LAUNCH
#undef HGIN_LAUNCH_CONTIG
    if (post.dalpha_partials && post_on) combine_reduce_scalar_kernel<<<1, 1024, 0, s>>>(post.dalpha_partials, grid, post_dalpha);
    if (post.ddot_partials && post_on) combine_reduce_scalar_kernel<<<1, 1024, 0, s>>>(post.ddot_partials, grid, post_ddot);
    HGIN_CHECK_LAUNCH(who);
    return HGIN_OK;
}

int32_t combine_dispatch(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges,
                         const void *x_src, int64_t ld_src, int32_t f_src, const void *x_self, int64_t ld_self,
                         int32_t f_self, const float *eps, int32_t self_mode, int32_t accumulate, void *out,
                         int64_t ld_out, const void *post_z, int64_t ld_post, int32_t post_act,
                         const float *post_alpha, float *post_dalpha, float *post_ddot, void *workspace,
                         int64_t workspace_bytes, void *stream, const char *who, int32_t src_act = HGIN_ACT_NONE,
                         const float *src_alpha = nullptr, int32_t self_act = HGIN_ACT_NONE,
                         const float *self_alpha = nullptr, int32_t dtype = HGIN_DTYPE_F32,
                         const BlockInfo blk = BlockInfo{}, bool table = false) {
    if (dtype == HGIN_DTYPE_BF16)
        return combine_dispatch_t<bf16>(num_rows, rowptr, col, num_edges, static_cast<const bf16 *>(x_src), ld_src, f_src,
                                        static_cast<const bf16 *>(x_self), ld_self, f_self, eps, self_mode, accumulate,
                                        static_cast<bf16 *>(out), ld_out, static_cast<const bf16 *>(post_z), ld_post,
                                        post_act, post_alpha, post_dalpha, post_ddot, workspace, workspace_bytes, stream,
                                        who, src_act, src_alpha, self_act, self_alpha, blk, table);
    HGIN_CHECK_ARG(dtype == HGIN_DTYPE_F32, "%s: bad dtype %d", who, dtype);
    return combine_dispatch_t<float>(num_rows, rowptr, col, num_edges, static_cast<const float *>(x_src), ld_src, f_src,
                                     static_cast<const float *>(x_self), ld_self, f_self, eps, self_mode, accumulate,
                                     static_cast<float *>(out), ld_out, static_cast<const float *>(post_z), ld_post,
                                     post_act, post_alpha, post_dalpha, post_ddot, workspace, workspace_bytes, stream, who,
                                     src_act, src_alpha, self_act, self_alpha, blk, table);
}

// Streaming kernel + gated gather kernel for one long-row aggregation on a block-diagonal batch.
template <typename T>
int32_t combine_blocks_t(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges, int64_t num_in,
                         const int32_t *rowptr_in, const int32_t *col_in, int32_t num_blocks, const int64_t *in_ptr,
                         const int64_t *out_ptr, const int32_t *gate, const T *x_src, int64_t ld_src, int32_t f_src,
                         const T *x_self, int64_t ld_self, const float *eps, int32_t self_mode, int32_t accumulate, T *out,
                         int64_t ld_out, int32_t src_act, const float *src_alpha, int32_t self_act, const float *self_alpha,
                         void *stream) {
    const char *who = "hgin_gin_combine_blocks_t";
    constexpr int elem = static_cast<int>(sizeof(T));
    HGIN_CHECK_ARG(num_blocks > 0 && in_ptr && out_ptr && gate && rowptr && col && rowptr_in && col_in, "%s: null pointer", who);
    if (!(f_src % (16 / elem) == 0 && f_src >= 32 && f_src <= 128 && ld_src == f_src && ld_out % 4 == 0 && ld_out >= f_src &&
          (self_mode == HGIN_SELF_NONE || (self_mode == HGIN_SELF_ADD && ld_self % 4 == 0 && ld_self >= f_src)) &&
          aligned16(x_src) && aligned16(out) && aligned16(x_self) && num_rows < INT32_MAX && num_in < INT32_MAX))
        return fail(HGIN_ERR_UNSUPPORTED, "%s: needs contiguous input rows of 32..128 features (16-byte multiples), SELF_NONE / "
                    "SELF_ADD, 16-byte aligned rows", who);
    HGIN_CHECK_ARG(src_act >= HGIN_ACT_NONE && src_act <= HGIN_ACT_RELU && self_act >= HGIN_ACT_NONE && self_act <= HGIN_ACT_RELU &&
                   (src_act != HGIN_ACT_PRELU || src_alpha) && (self_act != HGIN_ACT_PRELU || self_alpha), "%s: input activation", who);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    scatter::SbParams p{};
    p.num_blocks = num_blocks;
    p.in_ptr = in_ptr;
    p.out_ptr = out_ptr;
    p.rowptr = rowptr_in;
    p.col = col_in;
    p.gate = gate;
    p.cap_rows = scatter::capacity_rows(f_src, elem);
    p.x_in = x_src;
    p.f = f_src;
    p.x_self = x_self;
    p.ld_self = static_cast<int>(ld_self);
    p.eps = eps;
    p.self_mode = self_mode;
    p.accumulate = accumulate;
    p.out = out;
    p.ld_out = static_cast<int>(ld_out);
    p.in_act = src_act;
    p.in_alpha = src_alpha;
    p.self_act = self_act;
    p.self_alpha = self_alpha;
    p.stages = scatter::stages_for(f_src, elem);
    p.stage_bytes = scatter::stage_bytes_for(f_src, elem);
    static bool attr_set[2] = {false, false};
    if (!attr_set[elem == 2]) {
        cudaFuncSetAttribute(scatter::scatter_blocks_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter::SB_SMEM);
        attr_set[elem == 2] = true;
    }
    const int grid = num_blocks < kNumSMs ? num_blocks : kNumSMs;
    scatter::scatter_blocks_kernel<T><<<grid, scatter::SB_THREADS, scatter::SB_SMEM, s>>>(p);
    HGIN_CHECK_LAUNCH(who);
    // the gather kernel behind it, with the inverse gate
    return combine_dispatch_t<T>(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_src, eps, self_mode,
                                 accumulate, out, ld_out, nullptr, 0, HGIN_ACT_NONE, nullptr, nullptr, nullptr, nullptr, 0, stream,
                                 who, src_act, src_alpha, self_act, self_alpha,
                                 BlockInfo{num_blocks, in_ptr, out_ptr, gate, p.cap_rows, 1});
}

// Staged-source kernel + gated gather kernel for one long-row aggregation on a block-diagonal batch.
template <typename T>
int32_t combine_staged_t(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges, int32_t num_blocks,
                         const int64_t *in_ptr, const int64_t *out_ptr, const int32_t *gate, const T *x_src, int64_t ld_src,
                         int32_t f_src, const T *x_self, int64_t ld_self, const float *eps, int32_t self_mode, int32_t accumulate,
                         T *out, int64_t ld_out, int32_t src_act, const float *src_alpha, int32_t self_act,
                         const float *self_alpha, void *stream) {
    const char *who = "hgin_gin_combine_staged_t";
    constexpr int elem = static_cast<int>(sizeof(T));
    HGIN_CHECK_ARG(num_blocks > 0 && in_ptr && out_ptr && gate && rowptr && col, "%s: null pointer", who);
    if (!(f_src % (16 / elem) == 0 && f_src >= 16 && f_src <= 128 && ld_src == f_src && ld_out % 4 == 0 && ld_out >= f_src &&
          (self_mode == HGIN_SELF_NONE || (self_mode == HGIN_SELF_ADD && ld_self % 4 == 0 && ld_self >= f_src)) &&
          aligned16(x_src) && aligned16(out) && aligned16(x_self) && num_rows < INT32_MAX && num_edges > 0 && num_edges < INT32_MAX))
        return fail(HGIN_ERR_UNSUPPORTED, "%s: needs contiguous input rows of 16..128 features (16-byte multiples), SELF_NONE / "
                    "SELF_ADD, 16-byte aligned rows, at least one edge", who);
    HGIN_CHECK_ARG(src_act >= HGIN_ACT_NONE && src_act <= HGIN_ACT_RELU && self_act >= HGIN_ACT_NONE && self_act <= HGIN_ACT_RELU &&
                   (src_act != HGIN_ACT_PRELU || src_alpha) && (self_act != HGIN_ACT_PRELU || self_alpha), "%s: input activation", who);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    staged::SgParams p{};
    p.num_blocks = num_blocks;
    p.in_ptr = in_ptr;
    p.out_ptr = out_ptr;
    p.rowptr = rowptr;
    p.col = col;
    p.num_edges = static_cast<int>(num_edges);
    p.gate = gate;
    p.x_in = x_src;
    p.f = f_src;
    p.x_self = x_self;
    p.ld_self = static_cast<int>(ld_self);
    p.eps = eps;
    p.self_mode = self_mode;
    p.accumulate = accumulate;
    p.out = out;
    p.ld_out = static_cast<int>(ld_out);
    p.in_act = src_act;
    p.in_alpha = src_alpha;
    p.self_act = self_act;
    p.self_alpha = self_alpha;
    const char *dbg = getenv("HGIN_SG_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    static bool attr_set[2] = {false, false};
    if (!attr_set[elem == 2]) {
        cudaFuncSetAttribute(staged::stage_blocks_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, staged::SG_SMEM);
        cudaFuncSetAttribute(staged::stage_blocks_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, staged::SG_SMEM);
        attr_set[elem == 2] = true;
    }
    const int grid = num_blocks < kNumSMs ? num_blocks : kNumSMs;
    if (src_act != HGIN_ACT_NONE) staged::stage_blocks_kernel<T, true><<<grid, staged::SG_THREADS, staged::SG_SMEM, s>>>(p);
    else staged::stage_blocks_kernel<T, false><<<grid, staged::SG_THREADS, staged::SG_SMEM, s>>>(p);
    HGIN_CHECK_LAUNCH(who);
    // the gather kernel behind it, with the inverse gate
    return combine_dispatch_t<T>(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_src, eps, self_mode,
                                 accumulate, out, ld_out, nullptr, 0, HGIN_ACT_NONE, nullptr, nullptr, nullptr, nullptr, 0, stream,
                                 who, src_act, src_alpha, self_act, self_alpha,
                                 BlockInfo{num_blocks, in_ptr, out_ptr, gate, staged::SG_CAP_ROWS, 1});
}

}  // namespace
}  // namespace hgin

extern "C" int32_t hgin_gin_combine_staged_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                             int64_t num_edges, int32_t num_blocks, const int64_t *in_ptr, const int64_t *out_ptr,
                                             const int32_t *gate, const void *x_src, int64_t ld_src, int32_t f_src,
                                             const void *x_self, int64_t ld_self, const float *eps, int32_t self_mode,
                                             int32_t accumulate, void *out, int64_t ld_out, int32_t src_act, const float *src_alpha,
                                             int32_t self_act, const float *self_alpha, void *stream) {
    using namespace hgin;
    if (dtype == HGIN_DTYPE_BF16)
        return combine_staged_t<bf16>(num_rows, rowptr, col, num_edges, num_blocks, in_ptr, out_ptr, gate,
                                      static_cast<const bf16 *>(x_src), ld_src, f_src, static_cast<const bf16 *>(x_self), ld_self,
                                      eps, self_mode, accumulate, static_cast<bf16 *>(out), ld_out, src_act, src_alpha, self_act,
                                      self_alpha, stream);
    HGIN_CHECK_ARG(dtype == HGIN_DTYPE_F32, "hgin_gin_combine_staged_t: bad dtype %d", dtype);
    return combine_staged_t<float>(num_rows, rowptr, col, num_edges, num_blocks, in_ptr, out_ptr, gate,
                                   static_cast<const float *>(x_src), ld_src, f_src, static_cast<const float *>(x_self), ld_self,
                                   eps, self_mode, accumulate, static_cast<float *>(out), ld_out, src_act, src_alpha, self_act,
                                   self_alpha, stream);
}

extern "C" int32_t hgin_block_gate(int64_t rows_a, const int32_t *rowptr_a, const int32_t *col_a, int64_t rows_b,
                                   const int32_t *rowptr_b, const int32_t *col_b, int32_t num_blocks, const int64_t *in_ptr,
                                   const int64_t *out_ptr, int32_t *gate, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(rows_a >= 0 && rows_b >= 0 && num_blocks >= 0 && gate, "hgin_block_gate: bad arguments");
    HGIN_CHECK_ARG(num_blocks == 0 || (in_ptr && out_ptr), "hgin_block_gate: null block pointers");
    HGIN_CHECK_ARG((rows_a == 0 || rowptr_a) && (rows_b == 0 || rowptr_b), "hgin_block_gate: null row pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(gate, 0, 4 * sizeof(int32_t), s);
    const int64_t work = rows_a > rows_b ? rows_a : rows_b;
    scatter::block_gate_kernel<<<grid_for(work > 0 ? work : 1, 256, 8), 256, 0, s>>>(rows_a, rowptr_a, col_a, rows_b, rowptr_b, col_b,
                                                                                    num_blocks, in_ptr, out_ptr, gate);
    HGIN_CHECK_LAUNCH("hgin_block_gate");
    return HGIN_OK;
}

extern "C" int32_t hgin_gin_combine_blocks_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                             int64_t num_edges, int64_t num_in, const int32_t *rowptr_in, const int32_t *col_in,
                                             int32_t num_blocks, const int64_t *in_ptr, const int64_t *out_ptr, const int32_t *gate,
                                             const void *x_src, int64_t ld_src, int32_t f_src, const void *x_self, int64_t ld_self,
                                             const float *eps, int32_t self_mode, int32_t accumulate, void *out, int64_t ld_out,
                                             int32_t src_act, const float *src_alpha, int32_t self_act, const float *self_alpha,
                                             void *stream) {
    using namespace hgin;
    if (dtype == HGIN_DTYPE_BF16)
        return combine_blocks_t<bf16>(num_rows, rowptr, col, num_edges, num_in, rowptr_in, col_in, num_blocks, in_ptr, out_ptr, gate,
                                      static_cast<const bf16 *>(x_src), ld_src, f_src, static_cast<const bf16 *>(x_self), ld_self,
                                      eps, self_mode, accumulate, static_cast<bf16 *>(out), ld_out, src_act, src_alpha, self_act,
                                      self_alpha, stream);
    HGIN_CHECK_ARG(dtype == HGIN_DTYPE_F32, "hgin_gin_combine_blocks_t: bad dtype %d", dtype);
    return combine_blocks_t<float>(num_rows, rowptr, col, num_edges, num_in, rowptr_in, col_in, num_blocks, in_ptr, out_ptr, gate,
                                   static_cast<const float *>(x_src), ld_src, f_src, static_cast<const float *>(x_self), ld_self,
                                   eps, self_mode, accumulate, static_cast<float *>(out), ld_out, src_act, src_alpha, self_act,
                                   self_alpha, stream);
}

extern "C" int32_t hgin_gin_combine_table_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                            int64_t num_edges, int32_t num_blocks, const int64_t *in_ptr,
                                            const int64_t *out_ptr, const int32_t *gate, const void *x_src, int64_t ld_src,
                                            int32_t f_src, const void *x_self, int64_t ld_self, int32_t f_self,
                                            const float *eps, int32_t self_mode, int32_t accumulate, void *out, int64_t ld_out,
                                            int32_t src_act, const float *src_alpha, int32_t self_act, const float *self_alpha,
                                            const void *post_z, int64_t ld_post, int32_t post_act, const float *post_alpha,
                                            float *post_dalpha, float *post_ddot, void *workspace, int64_t workspace_bytes,
                                            void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(num_blocks > 0 && in_ptr && out_ptr && gate && rowptr && col, "hgin_gin_combine_table_t: null pointer");
    return combine_dispatch(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_self, eps, self_mode,
                            accumulate, out, ld_out, post_z, ld_post, post_act, post_alpha, post_dalpha, post_ddot, workspace,
                            workspace_bytes, stream, "hgin_gin_combine_table_t", src_act, src_alpha, self_act, self_alpha, dtype,
                            BlockInfo{num_blocks, in_ptr, out_ptr, gate, 0, 2}, true);
}

extern "C" int32_t hgin_gin_combine(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges,
                                    const float *x_src, int64_t ld_src, int32_t f_src, const float *x_self, int64_t ld_self,
                                    int32_t f_self, const float *eps, int32_t self_mode, int32_t accumulate,
                                    float *out, int64_t ld_out, void *stream) {
    return hgin::combine_dispatch(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,
                                  self_mode, accumulate, out, ld_out, nullptr, 0, HGIN_ACT_NONE, nullptr, nullptr, nullptr,
                                  nullptr, 0, stream, "hgin_gin_combine");
}

extern "C" int64_t hgin_gin_combine_post_workspace_bytes(void) { return static_cast<int64_t>(hgin::kMaxCombineCtas) * 8; }

extern "C" int32_t hgin_gin_combine_post(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges,
                                         const float *x_src, int64_t ld_src, int32_t f_src, const float *x_self,
                                         int64_t ld_self, int32_t f_self, const float *eps, int32_t self_mode,
                                         int32_t accumulate, float *out, int64_t ld_out, const float *post_z,
                                         int64_t ld_post, int32_t post_act, const float *post_alpha, float *post_dalpha,
                                         float *post_ddot, void *workspace, int64_t workspace_bytes, void *stream) {
    return hgin::combine_dispatch(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,
                                  self_mode, accumulate, out, ld_out, post_z, ld_post, post_act, post_alpha, post_dalpha,
                                  post_ddot, workspace, workspace_bytes, stream, "hgin_gin_combine_post");
}

extern "C" int32_t hgin_gin_combine_pre(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_edges,
                                        const float *x_src, int64_t ld_src, int32_t f_src, const float *x_self,
                                        int64_t ld_self, int32_t f_self, const float *eps, int32_t self_mode,
                                        int32_t accumulate, float *out, int64_t ld_out, int32_t src_act,
                                        const float *src_alpha, int32_t self_act, const float *self_alpha, void *stream) {
    return hgin::combine_dispatch(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,
                                  self_mode, accumulate, out, ld_out, nullptr, 0, HGIN_ACT_NONE, nullptr, nullptr, nullptr,
                                  nullptr, 0, stream, "hgin_gin_combine_pre", src_act, src_alpha, self_act, self_alpha);
}

extern "C" int32_t hgin_gin_combine_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                      int64_t num_edges, const void *x_src, int64_t ld_src, int32_t f_src,
                                      const void *x_self, int64_t ld_self, int32_t f_self, const float *eps,
                                      int32_t self_mode, int32_t accumulate, void *out, int64_t ld_out, int32_t src_act,
                                      const float *src_alpha, int32_t self_act, const float *self_alpha,
                                      const void *post_z, int64_t ld_post, int32_t post_act, const float *post_alpha,
                                      float *post_dalpha, float *post_ddot, void *workspace, int64_t workspace_bytes,
                                      void *stream) {
    return hgin::combine_dispatch(num_rows, rowptr, col, num_edges, x_src, ld_src, f_src, x_self, ld_self, f_self, eps,
                                  self_mode, accumulate, out, ld_out, post_z, ld_post, post_act, post_alpha, post_dalpha,
                                  post_ddot, workspace, workspace_bytes, stream, "hgin_gin_combine_t", src_act, src_alpha,
                                  self_act, self_alpha, dtype);
}

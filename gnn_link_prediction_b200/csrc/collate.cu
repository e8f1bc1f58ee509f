// Device-side batch assembly: the dataset lives in HBM as flat per-field arenas
// (all samples' rows back to back, plus an int64 row-pointer table per field) and a batch is
// assembled ON the GPU from a list of sample ids — what PyG's collate (`Batch.from_data_list`,
// reached through torch_geometric.loader.DataLoader, dataset.py:242) and `sample.cuda()`
// (train.py:28) do on the host and over PCIe in the reference.
//
//   collate_offsets_kernel : offsets[c][b] = exclusive prefix sum over the batch of the size of
//                            sample ids[b] in size class c (path / link / node rows, edges per
//                            relation) — the per-type node offsets and per-relation edge offsets
//                            PyG adds to `edge_index` when it concatenates.
//   collate_gather_kernel  : every field of every sample copied to its place in the batch; int32
//                            index fields (CSR row pointers / columns) get their offset added on
//                            the way.  Integer work: bit-exact with the host collate.
//
// Roofline: HBM, 8 bytes moved per 4-byte element (one read, one write); a cfgC batch is ~230 MB.
#include "hgin_common.cuh"

namespace hgin {
namespace {

constexpr int kMaxFields = 32;
constexpr int kMaxClasses = 16;

struct FieldTable {
    hgin_collate_field f[kMaxFields];
};

// One CTA; batches of up to a few 10^4 samples.  Thread t scans chunk t of the id list for every
// class, CTA-level exclusive scan of the chunk totals, then a second walk writes the offsets.
__global__ void __launch_bounds__(1024)
collate_offsets_kernel(int batch, const int32_t *__restrict__ ids, int num_classes,
                       const int64_t *__restrict__ class_ptr, int64_t num_samples, int64_t *__restrict__ offsets,
                       int32_t *__restrict__ status) {
    __shared__ int64_t totals[1024];
    const int per = (batch + blockDim.x - 1) / blockDim.x;
    const int beg = min(static_cast<int>(threadIdx.x) * per, batch);
    const int end = min(beg + per, batch);
    for (int c = 0; c < num_classes; ++c) {
        const int64_t *ptr = class_ptr + static_cast<int64_t>(c) * (num_samples + 1);
        int64_t sum = 0;
        for (int b = beg; b < end; ++b) {
            const int32_t id = ids[b];
            if (id < 0 || id >= num_samples) {
                if (c == 0) atomicExch(status, 1);
                continue;
            }
            sum += ptr[id + 1] - ptr[id];
        }
        totals[threadIdx.x] = sum;
        __syncthreads();
        // Hillis-Steele inclusive scan over the chunk totals
        for (int o = 1; o < static_cast<int>(blockDim.x); o <<= 1) {
            const int64_t v = threadIdx.x >= static_cast<unsigned>(o) ? totals[threadIdx.x - o] : 0;
            __syncthreads();
            totals[threadIdx.x] += v;
            __syncthreads();
        }
        int64_t run = totals[threadIdx.x] - sum;   // exclusive prefix of this chunk
        int64_t *out = offsets + static_cast<int64_t>(c) * (batch + 1);
        for (int b = beg; b < end; ++b) {
            out[b] = run;
            const int32_t id = ids[b];
            if (id >= 0 && id < num_samples) run += ptr[id + 1] - ptr[id];
        }
        if (threadIdx.x == blockDim.x - 1) out[batch] = totals[threadIdx.x];
        __syncthreads();
    }
}

// grid = (ctas per sample, batch, fields).  4-byte elements; consecutive threads copy consecutive
// words, so both sides are coalesced whatever the (word-aligned) start of a sample.
__global__ void __launch_bounds__(256)
collate_gather_kernel(const __grid_constant__ FieldTable table, int batch, const int32_t *__restrict__ ids,
                      int64_t num_samples, const int64_t *__restrict__ offsets) {
    const hgin_collate_field &fd = table.f[blockIdx.z];
    const int b = blockIdx.y;
    const int32_t id = ids[b];
    if (id < 0 || id >= num_samples) return;
    const int64_t *off = offsets + static_cast<int64_t>(fd.size_class) * (batch + 1);
    const int64_t src_row = fd.ptr[id];
    int64_t rows = fd.ptr[id + 1] - src_row;
    if (fd.closing_row && b != batch - 1) rows -= 1;           // rowptr blocks end with a closing entry: only the batch's last one is kept
    const int64_t words = rows * fd.width;
    const uint32_t *src = static_cast<const uint32_t *>(fd.src) + src_row * fd.width;
    uint32_t *dst = static_cast<uint32_t *>(fd.dst) + off[b] * fd.width;
    const int32_t add = fd.add_class >= 0
        ? static_cast<int32_t>(offsets[static_cast<int64_t>(fd.add_class) * (batch + 1) + b]) : 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // widest vector both sides are aligned for (sample starts are only word-aligned in general)
    const uintptr_t mis = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
    if ((mis & 15u) == 0) {
        const int64_t nv = words >> 2;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (int64_t i = t0; i < nv; i += stride) {
            uint4 v = __ldg(s4 + i);
            v.x += add; v.y += add; v.z += add; v.w += add;
            d4[i] = v;
        }
        for (int64_t i = (nv << 2) + t0; i < words; i += stride) dst[i] = __ldg(src + i) + add;
    } else if ((mis & 7u) == 0) {
        const int64_t nv = words >> 1;
        const uint2 *s2 = reinterpret_cast<const uint2 *>(src);
        uint2 *d2 = reinterpret_cast<uint2 *>(dst);
        for (int64_t i = t0; i < nv; i += stride) {
            uint2 v = __ldg(s2 + i);
            v.x += add; v.y += add;
            d2[i] = v;
        }
        for (int64_t i = (nv << 1) + t0; i < words; i += stride) dst[i] = __ldg(src + i) + add;
    } else {
        for (int64_t i = t0; i < words; i += stride) dst[i] = __ldg(src + i) + add;
    }
}

}  // namespace
}  // namespace hgin

extern "C" int32_t hgin_collate_offsets(int32_t batch, const int32_t *ids, int32_t num_classes, const int64_t *class_ptr,
                                        int64_t num_samples, int64_t *offsets, int32_t *status, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(batch >= 0 && num_classes > 0 && num_classes <= kMaxClasses && num_samples >= 0,
                   "hgin_collate_offsets: bad sizes (batch %d, classes %d)", batch, num_classes);
    HGIN_CHECK_ARG(offsets && status && (batch == 0 || ids) && class_ptr, "hgin_collate_offsets: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(status, 0, sizeof(int32_t), s);
    collate_offsets_kernel<<<1, 1024, 0, s>>>(batch, ids, num_classes, class_ptr, num_samples, offsets, status);
    HGIN_CHECK_LAUNCH("hgin_collate_offsets");
    return HGIN_OK;
}

extern "C" int32_t hgin_collate_gather(int32_t batch, const int32_t *ids, int64_t num_samples, int32_t num_fields,
                                       const hgin_collate_field *fields_host, int32_t num_classes, const int64_t *offsets,
                                       int64_t max_words_per_sample, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(batch >= 0 && batch <= 65535, "hgin_collate_gather: batch must be in [0, 65535], got %d", batch);
    HGIN_CHECK_ARG(num_fields > 0 && num_fields <= kMaxFields, "hgin_collate_gather: 1..%d fields, got %d", kMaxFields,
                   num_fields);
    HGIN_CHECK_ARG(fields_host && offsets && (batch == 0 || ids), "hgin_collate_gather: null pointer");
    FieldTable table{};
    for (int i = 0; i < num_fields; ++i) {
        const hgin_collate_field &f = fields_host[i];
        HGIN_CHECK_ARG(f.ptr && f.width > 0, "hgin_collate_gather: field %d: null ptr table or bad width", i);   // src/dst may be null for a field that is empty in every sample
        HGIN_CHECK_ARG(f.size_class >= 0 && f.size_class < num_classes && f.add_class < num_classes,
                       "hgin_collate_gather: field %d: class out of range", i);
        table.f[i] = f;
    }
    if (batch == 0) return HGIN_OK;
    int64_t per = ceil_div(max_words_per_sample > 0 ? max_words_per_sample : 1, 256 * 8);
    if (per > 64) per = 64;
    if (per < 1) per = 1;
    dim3 grid(static_cast<unsigned>(per), static_cast<unsigned>(batch), static_cast<unsigned>(num_fields));
    collate_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(table, batch, ids, num_samples, offsets);
    HGIN_CHECK_LAUNCH("hgin_collate_gather");
    return HGIN_OK;
}

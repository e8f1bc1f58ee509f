// Host-side batch assembly: the CPU twin of collate.cu for datasets that stay in host memory.
//
// Same contract as hgin_collate_offsets + hgin_collate_gather (PyG's collate, dataset.py:242), but
// every pointer is a HOST pointer: the arenas of a SampleArena (typically memory-mapped from its
// file) are copied, with the node / edge offsets added to int32 index fields, straight into one
// pinned, packed batch buffer that crosses PCIe as a single DMA (data.PackedBatch).  A Python
// collate of 1024 samples costs ~1 s; this runs at memcpy speed on a few threads.  Host code only:
// no CUDA call, usable without a GPU.
#include <atomic>
#include <thread>
#include <vector>

#include <string.h>

#include "hgin_common.cuh"

extern "C" int32_t hgin_host_collate(int32_t batch, const int32_t *ids_host, int64_t num_samples, int32_t num_fields,
                                     const hgin_collate_field *fields_host, int32_t num_classes,
                                     const int64_t *class_ptr_host, int64_t *offsets_host, int32_t num_threads) {
    using namespace hgin;
    HGIN_CHECK_ARG(batch >= 0 && num_samples >= 0 && num_fields > 0 && num_classes > 0, "hgin_host_collate: bad sizes");
    HGIN_CHECK_ARG(fields_host && class_ptr_host && offsets_host && (batch == 0 || ids_host), "hgin_host_collate: null pointer");
    for (int b = 0; b < batch; ++b)
        HGIN_CHECK_ARG(ids_host[b] >= 0 && ids_host[b] < num_samples, "hgin_host_collate: sample id %d outside [0, %lld)",
                       ids_host[b], (long long)num_samples);
    for (int i = 0; i < num_fields; ++i) {
        const hgin_collate_field &f = fields_host[i];
        HGIN_CHECK_ARG(f.ptr && f.width > 0 && f.size_class >= 0 && f.size_class < num_classes && f.add_class < num_classes,
                       "hgin_host_collate: field %d: bad descriptor", i);
    }
    // offsets[c][b] = exclusive prefix sum of the size of sample ids[b] in class c
    for (int c = 0; c < num_classes; ++c) {
        const int64_t *ptr = class_ptr_host + static_cast<int64_t>(c) * (num_samples + 1);
        int64_t *out = offsets_host + static_cast<int64_t>(c) * (batch + 1);
        int64_t run = 0;
        for (int b = 0; b < batch; ++b) {
            out[b] = run;
            run += ptr[ids_host[b] + 1] - ptr[ids_host[b]];
        }
        out[batch] = run;
    }
    if (batch == 0) return HGIN_OK;
    const int64_t jobs = static_cast<int64_t>(num_fields) * batch;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int64_t j = next.fetch_add(1, std::memory_order_relaxed);
            if (j >= jobs) return;
            const hgin_collate_field &f = fields_host[j / batch];
            const int b = static_cast<int>(j % batch);
            const int32_t id = ids_host[b];
            const int64_t *off = offsets_host + static_cast<int64_t>(f.size_class) * (batch + 1);
            const int64_t *fptr = static_cast<const int64_t *>(f.ptr);
            int64_t rows = fptr[id + 1] - fptr[id];
            if (f.closing_row && b != batch - 1) rows -= 1;
            const int64_t words = rows * f.width;
            if (words <= 0) continue;
            const int32_t *src = static_cast<const int32_t *>(f.src) + fptr[id] * f.width;
            int32_t *dst = static_cast<int32_t *>(f.dst) + off[b] * f.width;
            if (f.add_class >= 0) {
                const int32_t add = static_cast<int32_t>(offsets_host[static_cast<int64_t>(f.add_class) * (batch + 1) + b]);
                for (int64_t i = 0; i < words; ++i) dst[i] = src[i] + add;
            } else {
                memcpy(dst, src, static_cast<size_t>(words) * 4);
            }
        }
    };
    int nt = num_threads > 0 ? num_threads : static_cast<int>(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (static_cast<int64_t>(nt) > jobs) nt = static_cast<int>(jobs);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return HGIN_OK;
}

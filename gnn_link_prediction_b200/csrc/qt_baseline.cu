// Queueing-theory baseline — the fixed-point iteration of QTBaseline.forward
// (models.py:42-158) whose outputs become the `bl_features` columns of link.x / path.x
// (dataset.py:86, 105-106).  The reference runs it on the CPU, one sample at a time, as a Python
// loop over hop positions with gather / scatter-sum per position; here a whole block-diagonal
// batch is processed by three grid-wide kernels per iteration-free formulation:
//
//   traffic (thread per path)  walks the path's links in route order (the source-sorted CSR of the
//                              path->link relation keeps it): the traffic entering hop k is
//                              A[p] * prod_{j<k} (1 - blocking[link_j])  (models.py:106-115), stored
//                              per edge;
//   link    (thread per link)  T[l] = sum of the per-edge traffic over the link's incoming edges
//                              (destination-sorted CSR, stable edge order — deterministic, no
//                              atomics), rho = T / capacity, blocking = M/M/1/B formula
//                              (models.py:127-134); after the last iteration also the expected
//                              queue occupancy series (models.py:141-148);
//   delay   (thread per path)  sum over the path's links of occupancy * 32000 / capacity
//                              (models.py:153-157).
//
// fp32 throughout like the reference; results agree to rounding (powf differs in the last ulp
// between libm and CUDA), not bit for bit.  Scalar-feature, latency-bound work: KB, not GB.
#include "hgin_common.cuh"

namespace hgin {
namespace {

constexpr int kBuffer = 32;   // buffer_size B of models.py:124

__global__ void __launch_bounds__(256)
qt_traffic_kernel(int64_t num_paths, const int32_t *__restrict__ rowptr_s, const int32_t *__restrict__ col_s,
                  const int32_t *__restrict__ perm_s, const float *__restrict__ avg_bw,
                  const float *__restrict__ blocking, float *__restrict__ edge_traffic) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= num_paths) return;
    float traffic = avg_bw[p];
    const int32_t beg = rowptr_s[p], end = rowptr_s[p + 1];
    for (int32_t j = beg; j < end; ++j) {
        edge_traffic[perm_s ? perm_s[j] : j] = traffic;
        traffic = __fmul_rn(traffic, __fsub_rn(1.0f, blocking[col_s[j]]));   // traffic[paths] *= (1 - p_block)
    }
}

__global__ void __launch_bounds__(256)
qt_link_kernel(int64_t num_links, const int32_t *__restrict__ rowptr_d, const int32_t *__restrict__ perm_d,
               const float *__restrict__ edge_traffic, const float *__restrict__ capacity /* L / 1000 */,
               float *__restrict__ blocking, float *__restrict__ link_out /* [num_links, 3] or NULL */) {
    const int64_t l = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l >= num_links) return;
    float t = 0.0f;
    for (int32_t j = rowptr_d[l]; j < rowptr_d[l + 1]; ++j) t = __fadd_rn(t, edge_traffic[perm_d[j]]);
    const float rho = t / capacity[l];
    const float num = (1.0f - rho) * powf(rho, static_cast<float>(kBuffer));
    const float den = 1.0f - powf(rho, static_cast<float>(kBuffer + 1));
    blocking[l] = num / (den + 1e-08f);
    if (link_out) {
        // models.py:141-148: pi_0 = (1-rho)/(1-rho^(B+1)); res = (pi_0 + sum_{j<32} (j+1) pi_0 rho^(j+1)) / 32;
        // the reference returns the pi_0 variable AFTER the loop has scaled it by rho^32.
        float pi0 = (1.0f - rho) / (1.0f - powf(rho, static_cast<float>(kBuffer + 1)));
        float res = 1.0f * pi0;
        for (int j = 0; j < 32; ++j) {
            pi0 = pi0 * rho;
            res += static_cast<float>(j + 1) * pi0;
        }
        res = res / 32.0f;
        link_out[l * 3 + 0] = res;
        link_out[l * 3 + 1] = rho;
        link_out[l * 3 + 2] = pi0;
    }
}

__global__ void __launch_bounds__(256)
qt_delay_kernel(int64_t num_paths, const int32_t *__restrict__ rowptr_s, const int32_t *__restrict__ col_s,
                const float *__restrict__ link_out, const float *__restrict__ capacity_raw,
                float *__restrict__ path_delay) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= num_paths) return;
    float d = 0.0f;
    for (int32_t j = rowptr_s[p]; j < rowptr_s[p + 1]; ++j) {
        const int32_t l = col_s[j];
        d = __fadd_rn(d, link_out[l * 3] * 32000.0f / capacity_raw[l]);
    }
    path_delay[p] = d;
}

__global__ void __launch_bounds__(256) qt_fill_kernel(int64_t n, float v, float *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

}  // namespace
}  // namespace hgin

extern "C" int64_t hgin_qt_baseline_workspace_bytes(int64_t num_links, int64_t num_edges) {
    if (num_links < 0 || num_edges < 0) return -1;
    return hgin::align_up(num_links * 4, 256) + hgin::align_up(num_edges * 4, 256) + 256;
}

extern "C" int32_t hgin_qt_baseline(int64_t num_paths, int64_t num_links, int64_t num_edges, const int32_t *rowptr_src,
                                    const int32_t *col_src, const int32_t *perm_src, const int32_t *rowptr_dst,
                                    const int32_t *perm_dst, const float *avg_bw, const float *capacity_scaled,
                                    const float *capacity_raw, int32_t num_iterations, float *path_delay,
                                    float *link_out, void *workspace, int64_t workspace_bytes, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(num_paths >= 0 && num_links >= 0 && num_edges >= 0 && num_paths < INT32_MAX && num_links < INT32_MAX &&
                       num_edges < INT32_MAX, "hgin_qt_baseline: bad sizes");
    HGIN_CHECK_ARG(num_iterations >= 1, "hgin_qt_baseline: num_iterations must be >= 1, got %d", num_iterations);
    if (num_paths == 0 && num_links == 0) return HGIN_OK;
    HGIN_CHECK_ARG(rowptr_src && rowptr_dst && path_delay && link_out && avg_bw && capacity_scaled && capacity_raw,
                   "hgin_qt_baseline: null pointer");
    HGIN_CHECK_ARG(num_edges == 0 || (col_src && perm_dst), "hgin_qt_baseline: null index array");
    const int64_t need = hgin_qt_baseline_workspace_bytes(num_links, num_edges);
    if (!workspace || workspace_bytes < need)
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_qt_baseline: workspace %lld < %lld bytes", (long long)workspace_bytes,
                    (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *blocking = static_cast<float *>(workspace);
    float *edge_traffic = reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(num_links * 4, 256));
    const unsigned gp = static_cast<unsigned>(ceil_div(num_paths > 0 ? num_paths : 1, 256));
    const unsigned gl = static_cast<unsigned>(ceil_div(num_links > 0 ? num_links : 1, 256));
    qt_fill_kernel<<<gl, 256, 0, s>>>(num_links, 0.5f, blocking);          // blocking_probs = 0.5 (models.py:96)
    for (int it = 0; it < num_iterations; ++it) {
        qt_traffic_kernel<<<gp, 256, 0, s>>>(num_paths, rowptr_src, col_src, perm_src, avg_bw, blocking, edge_traffic);
        qt_link_kernel<<<gl, 256, 0, s>>>(num_links, rowptr_dst, perm_dst, edge_traffic, capacity_scaled, blocking,
                                          it == num_iterations - 1 ? link_out : nullptr);
    }
    qt_delay_kernel<<<gp, 256, 0, s>>>(num_paths, rowptr_src, col_src, link_out, capacity_raw, path_delay);
    HGIN_CHECK_LAUNCH("hgin_qt_baseline");
    return HGIN_OK;
}

// sqrt(MAPE) loss forward/backward (train.py:12-13, 40-42), Adam/AdamW on a flat bucket
// (train.py:44, 140-148), and the library-level entry points (version, last error).
#include <math.h>

#include "hgin_common.cuh"

namespace hgin {

char *error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

namespace {

constexpr int kRedThreads = 256;

inline int reduce_ctas(int64_t n) { return grid_for(n, kRedThreads * 8, 2); }

__global__ void __launch_bounds__(kRedThreads)
mape_partial_kernel(int64_t n, const float *__restrict__ pred, const float *__restrict__ y,
                    float *__restrict__ partials) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float yi = __ldg(y + i);
        s += fabsf(__fdiv_rn(__fsub_rn(__ldg(pred + i), yi), yi));  // |(p - y) / y|, train.py:13
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
mape_final_kernel(const float *__restrict__ partials, int count, int64_t n, float *__restrict__ sums) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partials[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        sums[0] = s;
        sums[1] = static_cast<float>(n);
    }
}

__global__ void __launch_bounds__(256)
sqrt_mape_bwd_kernel(int64_t n, const float *__restrict__ pred, const float *__restrict__ y,
                     const float *__restrict__ sums, float gscale, float *__restrict__ loss_out,
                     float *__restrict__ dpred) {
    const float S = __ldg(sums), N = __ldg(sums + 1);
    const float mape = 100.0f * (S / N);
    const float L = sqrtf(mape);
    if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
        loss_out[0] = mape;
        loss_out[1] = L;
    }
    // dL/dp_i = 1/(2L) * 100/N * sign(u_i)/y_i, u = (p - y)/y   (chain of sqrt, mean, abs, div)
    const float c = gscale * 50.0f / (N * L);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float yi = __ldg(y + i);
        const float u = (__ldg(pred + i) - yi) / yi;
        const float sg = (u > 0.0f) ? 1.0f : ((u < 0.0f) ? -1.0f : 0.0f);
        dpred[i] = c * sg / yi;
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(int64_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, const int32_t *__restrict__ step_ptr, double lr, double b1, double b2, double eps,
            double wd, int decoupled) {
    // Scalar prologue in double, as torch.optim.Adam's single-tensor path computes it in Python floats.
    const double t = static_cast<double>(__ldg(step_ptr));
    const double bc1 = 1.0 - pow(b1, t);
    const double bc2 = 1.0 - pow(b2, t);
    const float neg_step_size = static_cast<float>(-(lr / bc1));
    const float bc2_sqrt = static_cast<float>(sqrt(bc2));
    const float w1 = static_cast<float>(1.0 - b1), w2 = static_cast<float>(1.0 - b2);
    const float b2f = static_cast<float>(b2), epsf = static_cast<float>(eps), wdf = static_cast<float>(wd);
    const float decay = static_cast<float>(1.0 - lr * wd);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float pi = p[i], gi = g[i];
        if (wd != 0.0) {
            if (decoupled) pi *= decay;          // AdamW: param.mul_(1 - lr * wd)
            else gi = gi + wdf * pi;             // Adam: grad.add(param, alpha=wd)
        }
        const float mi = m[i] + w1 * (gi - m[i]);           // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * b2f + (w2 * gi) * gi;       // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + epsf;
        p[i] = pi + (neg_step_size * mi) / denom;           // addcdiv_(exp_avg, denom, value=-step_size)
    }
}

__global__ void increment_kernel(int32_t *c) { *c += 1; }

}  // namespace
}  // namespace hgin

extern "C" int32_t hgin_version(void) { return HGIN_VERSION; }
extern "C" const char *hgin_last_error(void) { return hgin::error_buffer(); }

extern "C" int64_t hgin_reduce_workspace_bytes(int64_t n) {
    if (n < 0) return -1;
    return static_cast<int64_t>(hgin::reduce_ctas(n)) * 4 + 256;
}

extern "C" int32_t hgin_mape_sum(int64_t n, const float *pred, const float *y, float *sums, void *workspace,
                                 int64_t workspace_bytes, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(n >= 0 && sums, "hgin_mape_sum: bad arguments");
    HGIN_CHECK_ARG(n == 0 || (pred && y), "hgin_mape_sum: null pointer");
    if (!workspace || workspace_bytes < hgin_reduce_workspace_bytes(n))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_mape_sum: workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ctas = reduce_ctas(n);
    float *partials = static_cast<float *>(workspace);
    mape_partial_kernel<<<ctas, kRedThreads, 0, s>>>(n, pred, y, partials);
    mape_final_kernel<<<1, 1024, 0, s>>>(partials, ctas, n, sums);
    HGIN_CHECK_LAUNCH("hgin_mape_sum");
    return HGIN_OK;
}

extern "C" int32_t hgin_sqrt_mape_bwd(int64_t n, const float *pred, const float *y, const float *sums, float gscale,
                                      float *loss_out, float *dpred, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(n >= 0 && sums && (n == 0 || (pred && y && dpred)), "hgin_sqrt_mape_bwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    sqrt_mape_bwd_kernel<<<grid_for(n, 256 * 4, 4), 256, 0, s>>>(n, pred, y, sums, gscale, loss_out, dpred);
    HGIN_CHECK_LAUNCH("hgin_sqrt_mape_bwd");
    return HGIN_OK;
}

extern "C" int32_t hgin_adam_step(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                                  const int32_t *step, double lr, double beta1, double beta2, double eps,
                                  double weight_decay, int32_t decoupled, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(n >= 0 && step, "hgin_adam_step: bad arguments");
    if (n == 0) return HGIN_OK;
    HGIN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "hgin_adam_step: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    adam_kernel<<<grid_for(n, 256 * 4, 4), 256, 0, s>>>(n, param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2,
                                                       eps, weight_decay, decoupled);
    HGIN_CHECK_LAUNCH("hgin_adam_step");
    return HGIN_OK;
}

extern "C" int32_t hgin_increment(int32_t *counter, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(counter, "hgin_increment: null pointer");
    increment_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(counter);
    HGIN_CHECK_LAUNCH("hgin_increment");
    return HGIN_OK;
}

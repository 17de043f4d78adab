// Adam over ONE flat parameter vector (trainer.py:134-146 builds torch.optim.Adam over the trainable parameters; the foreach
// implementation is 7 multi-tensor launches of ~27 us each for GEDNet's 30 small tensors).  Here every parameter, its gradient and
// both moments are views into four flat fp32 buffers, and the whole update is one elementwise kernel (plus a one-thread step counter
// so that a captured CUDA graph advances the bias correction on replay):
//   t += 1;  g' = g + wd*p;  m += (g' - m)(1 - b1);  v = b2*v + (1 - b2) g'^2;  p -= lr/(1 - b1^t) * m / (sqrt(v)/sqrt(1 - b2^t) + eps)
// which is torch.optim.Adam's update (amsgrad=False, maximize=False) evaluated in fp32 with the bias corrections in double.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace optim {

__global__ void adam_advance_kernel(float* __restrict__ step) { step[0] += 1.0f; }

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float wd, float omb1, float b2, float omb2, float step_size,
                                      float inv_sqrt_bc2, float eps) {
    g = fmaf(wd, p, g);
    m = fmaf(g - m, omb1, m);
    v = fmaf(omb2 * g, g, b2 * v);
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ step, long long n, double lr, double b1d,
                                                   double b2d, float eps, float wd) {
    __shared__ float s_step_size, s_inv_sqrt_bc2;
    if (threadIdx.x == 0) {
        const double t = (double)step[0];
        s_step_size = (float)(lr / (1.0 - pow(b1d, t)));
        s_inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(b2d, t)));
    }
    __syncthreads();
    // 1 - beta in double, then rounded once (1.0f - 0.999f is off by 1.3e-5 relative)
    const float step_size = s_step_size, isb = s_inv_sqrt_bc2, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d), b2 = (float)b2d;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam1(pp.x, gg.x, mm.x, vv.x, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.y, gg.y, mm.y, vv.y, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.z, gg.z, mm.z, vv.z, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.w, gg.w, mm.w, vv.w, wd, omb1, b2, omb2, step_size, isb, eps);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        adam1(p[i], g[i], m[i], v[i], wd, omb1, b2, omb2, step_size, isb, eps);
}

// Device-resident hyper-parameters (adc_adam_step_dev): hyper = double[5] {lr, beta1, beta2, eps, weight_decay}; state = float[4]
// {t, step_size, 1/sqrt(1-b2^t), apply}.  One thread advances the step counter and derives the scalars of this update -- unless the
// guard value (the step's loss) is not finite, in which case `apply` = 0 and the update kernel leaves everything untouched
// (trainer.py:206-208 skips such a batch).  A captured CUDA graph therefore follows a learning-rate schedule (the host refreshes
// `hyper` between replays) and skips poisoned batches without a host branch.
__global__ void adam_prepare_kernel(float* __restrict__ state, const double* __restrict__ hyper, const float* __restrict__ guard) {
    const bool ok = guard == nullptr || isfinite(guard[0]);
    if (ok) state[0] += 1.0f;
    const double t = (double)state[0];
    state[1] = ok ? (float)(hyper[0] / (1.0 - pow(hyper[1], t))) : 0.f;
    state[2] = ok ? (float)(1.0 / sqrt(1.0 - pow(hyper[2], t))) : 0.f;
    state[3] = ok ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, const float* __restrict__ state,
                                                       const double* __restrict__ hyper, long long n) {
    if (state[3] == 0.f) return;                 // non-finite loss: parameters, moments and the step counter stay as they are
    const float step_size = state[1], isb = state[2];
    const double b1d = hyper[1], b2d = hyper[2];
    const float omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d), b2 = (float)b2d, eps = (float)hyper[3], wd = (float)hyper[4];
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam1(pp.x, gg.x, mm.x, vv.x, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.y, gg.y, mm.y, vv.y, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.z, gg.z, mm.z, vv.z, wd, omb1, b2, omb2, step_size, isb, eps);
        adam1(pp.w, gg.w, mm.w, vv.w, wd, omb1, b2, omb2, step_size, isb, eps);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        adam1(p[i], g[i], m[i], v[i], wd, omb1, b2, omb2, step_size, isb, eps);
}

}  // namespace optim

extern "C" {

int adc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step, long long n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, adc_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n <= 0 || ((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg |
                                                                          (uintptr_t)exp_avg_sq) & 15) {
        adc::set_error("adc_adam_step: bad argument (flat fp32 buffers, 16-byte aligned)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TRAIN, (double)n * 28.0);
    adc::count_launches(1);
    optim::adam_advance_kernel<<<1, 1, 0, st>>>(step);
    int rc = adc::check_cuda(cudaGetLastError(), "adam_advance_kernel launch");
    if (rc) return rc;
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    optim::adam_kernel<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step, n, lr, beta1, beta2, (float)eps, (float)weight_decay);
    return adc::check_cuda(cudaGetLastError(), "adam_kernel launch");
}

int adc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* state, const double* hyper,
                      const float* guard, long long n, adc_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !state || !hyper || n <= 0 ||
        ((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) {
        adc::set_error("adc_adam_step_dev: bad argument (flat fp32 buffers, 16-byte aligned; state float[4], hyper double[5])");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TRAIN, (double)n * 28.0);
    adc::count_launches(1);
    optim::adam_prepare_kernel<<<1, 1, 0, st>>>(state, hyper, guard);
    int rc = adc::check_cuda(cudaGetLastError(), "adam_prepare_kernel launch");
    if (rc) return rc;
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    optim::adam_dev_kernel<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, state, hyper, n);
    return adc::check_cuda(cudaGetLastError(), "adam_dev_kernel launch");
}

}  // extern "C"

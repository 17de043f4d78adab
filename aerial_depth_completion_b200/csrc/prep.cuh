// Weight preparation: EnforcePos softplus iteration (model_zoo/nconv.py:114-131), sum(w) (nconv.py:58-61) and the
// constant-memory images the kernels read (layouts in adc_common.cuh).
#pragma once
#include "adc_common.cuh"

namespace adc {

__device__ __forceinline__ float softplus10(float x) {
    // F.softplus(x, beta=10, threshold=20): x if 10x > 20 else log1p(exp(10x))/10
    const float bx = x * 10.0f;
    return bx > 20.0f ? x : log1pf(expf(bx)) / 10.0f;
}

struct PrepCall {
    const float* w;     // [Co,Ci,K,K] as stored
    const float* b;     // [Co]
    float* w_out;       // standard layout copy of the weight this call used (may be null)
    int Co, Ci, K;
    int times;          // softplus applications before use
    int flip;           // 0: forward image, 1: data-gradient image (transposed + flipped taps)
    const float* s_in;  // flip=1: image of the forward call (source of s, 1/s), else null
};

struct PrepArgs {
    PrepCall call[ADC_PYR_CALLS];
    float* img;         // [ncalls][kWStride] global staging copy of the constant image
};

__global__ void __launch_bounds__(128) prep_weights_kernel(PrepArgs a) {
    const PrepCall& c = a.call[blockIdx.x];
    float* img = a.img + (size_t)blockIdx.x * kWStride;
    __shared__ float ssum[2][4];
    const int KK = c.K * c.K, per_co = c.Ci * KK, n = c.Co * per_co;
    const int t = threadIdx.x;
    if (t < kWStride) img[t] = 0.f;
    __syncthreads();
    float part[2] = {0.f, 0.f};
    if (t < n) {
        float w = c.w[t];
        for (int i = 0; i < c.times; ++i) w = softplus10(w);
        if (c.w_out) c.w_out[t] = w;
        const int co = t / per_co, rem = t % per_co, ci = rem / KK, u = (rem % KK) / c.K, v = rem % c.K;
        part[co] = w;
        if (!c.flip)
            img[t] = w;
        else
            img[(ci * c.Co + co) * KK + (c.K - 1 - u) * c.K + (c.K - 1 - v)] = w;
    }
    const float s0 = warp_sum(part[0]), s1 = warp_sum(part[1]);
    if ((t & 31) == 0) {
        ssum[0][t >> 5] = s0;
        ssum[1][t >> 5] = s1;
    }
    __syncthreads();
    if (t < c.Co) {
        img[kWBias + t] = c.b[t];
        const float s = c.flip ? c.s_in[kWSum + t] : (ssum[t][0] + ssum[t][1] + ssum[t][2] + ssum[t][3]);
        img[kWSum + t] = s;
        img[kWInvSum + t] = 1.0f / s;
    }
}

__global__ void softplus_pow_kernel(const float* __restrict__ w, float* __restrict__ out, int n, int times) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = w[i];
    for (int k = 0; k < times; ++k) v = softplus10(v);
    out[i] = v;
}

}  // namespace adc

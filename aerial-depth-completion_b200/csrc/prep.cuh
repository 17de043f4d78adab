// Weight preparation: EnforcePos softplus iteration (model_zoo/nconv.py:114-131), sum(w) (nconv.py:58-61) and
// re-layout of the <=100 weights of each NConv2d call into the constant-memory image the kernels read.
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
    int flip;           // 0: forward image, 1: data-gradient image (flipped taps, ci-pair major)
    const float* s_in;  // flip=1: s of the forward call (image of the forward), else null
};

struct PrepArgs {
    PrepCall call[ADC_PYR_CALLS];
    float* img;         // [ncalls][kWStride] global staging copy of the constant image
};

// Forward image (Co == 2): float2 index (ci*K+u)*K+v = {w[0,ci,u,v], w[1,ci,u,v]}; (Co == 1): float index (ci*K+u)*K+v.
// Data-gradient image: for ci-pair q, float2 index ((q*Co+o)*K+u')*K+v' = {w[o,2q,K-1-u',K-1-v'], w[o,2q+1,...]}
// (an odd Ci is padded with a zero channel).  [kWBias..] = bias, [kWSum..] = s = sum_{ci,u,v} w (forward weight).
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
        if (!c.flip) {
            img[((ci * c.K + u) * c.K + v) * c.Co + co] = w;
        } else {
            const int q = ci >> 1, j = ci & 1;
            img[((((q * c.Co + co) * c.K) + (c.K - 1 - u)) * c.K + (c.K - 1 - v)) * 2 + j] = w;
        }
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

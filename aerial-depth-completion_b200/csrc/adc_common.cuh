// Shared device/host helpers for the sm_100a kernels of the NConv depth-completion hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/adc_b200.h"

namespace adc {

constexpr float kEps = 1e-20f;       // model_zoo/nconv.py:25
constexpr int kWStride = 128;        // floats per call in the constant-memory weight image
constexpr int kWBias = 100;          // [100..101] bias
constexpr int kWSum = 102;           // [102..103] s = sum(w) of the forward call
constexpr int kWInvSum = 104;        // [104..105] 1/s (IEEE division, computed once per call)
constexpr int kWMax = 100;           // max Co*Ci*k*k (2*2*25)

// Constant-memory weight images: [slot][fwd/bwd][call][kWStride].  Written with stream-ordered D2D copies,
// read through the uniform datapath (LDCU -> FFMA2 UR operands), so weights cost no vector registers.
__constant__ float c_wimg[ADC_WSLOTS][2][ADC_PYR_CALLS][kWStride];  // single-TU build (adc_nconv.cu)

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define ADC_CHECK_ARG(cond, code, ...)          \
    do {                                        \
        if (!(cond)) {                          \
            adc::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

#define ADC_CUDA(expr)                                        \
    do {                                                      \
        int _rc = adc::check_cuda((expr), #expr);             \
        if (_rc) return _rc;                                  \
    } while (0)

// ATen 'nearest' source index (aten/src/ATen/native/UpSample.h nearest_idx): identity if sizes match,
// dst>>1 for exact 2x, else min(floor(dst * float(in)/out), in-1).
struct NearestMap {
    int in, out, mode;  // mode 0 identity, 1 shift, 2 scaled
    float scale;
    __host__ static NearestMap make(int in, int out) {
        NearestMap m;
        m.in = in;
        m.out = out;
        m.mode = (in == out) ? 0 : (out == 2 * in ? 1 : 2);
        m.scale = (float)in / (float)out;
        return m;
    }
    __device__ __forceinline__ int operator()(int dst) const {
        if (mode == 0) return dst;
        if (mode == 1) return dst >> 1;
        int s = (int)floorf((float)dst * scale);
        return s < in - 1 ? s : in - 1;
    }
};

// One source of input channels as the kernels see it.
struct Src {
    const float* x;
    const float* c;
    long long bs;      // batch stride (floats)
    int ch, H, W;
    NearestMap mh, mw; // dst(row/col of the layer) -> src index
    int up;            // 1 if (H,W) differs from the layer size
};

__host__ inline Src make_src(const adc_src_t* s, int H, int W) {
    Src r;
    r.x = s->x;
    r.c = s->c;
    r.bs = s->batch_stride;
    r.ch = s->channels;
    r.H = s->H;
    r.W = s->W;
    r.mh = NearestMap::make(s->H, H);
    r.mw = NearestMap::make(s->W, W);
    r.up = (s->H != H || s->W != W) ? 1 : 0;
    return r;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 bcast(float a) { return make_float2(a, a); }

// 1/x by the SFU (MUFU.RCP, <= 1 ulp, IEEE special cases: rcp(0)=inf so q*rcp(0) behaves like q/0).  The operands
// here are never denormal (D+1e-20 >= 1e-20 for non-negative confidences).  An IEEE-rounded division costs ~11
// issue slots with a slow-path branch; at 64 quotients per thread that was 40% of the forward kernel.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace adc

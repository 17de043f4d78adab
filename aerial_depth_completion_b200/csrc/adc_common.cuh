// Shared device/host helpers for the sm_100a kernels of the NConv depth-completion hot path.
//
// Internal data layout ("pair tensors"): inside the pyramid every activation is stored as float2 {x, c}
// (data, confidence) per pixel, [N, C, H, W] of float2, and every gradient as float2 {GN, GD} (or raw {gx, gc}).
// A pair is exactly one operand of the packed sm_100 FFMA2 (fma.rn.f32x2): {Nm, D} += {x*c, c} * w, so the two
// convolutions of a normalized convolution (and the two correlations of its backward) ride in the two lanes of
// one instruction, one TMA box brings both planes, and loads/stores are 8/16-byte vectors by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>
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
// forward image : w[(co*Ci+ci)*K*K + u*K + v]            (= the Parameter's own layout)
// backward image: w[(ci*Co+co)*K*K + u'*K + v'] = W*[co][ci][K-1-u'][K-1-v']   (transposed + flipped)
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

// One source of input channels as the kernels see it: a pair tensor [N, ch, H, W] of float2 {x, c}.
struct SrcP {
    const float2* p;
    long long bs;      // batch stride in float2 units
    int ch, H, W;
    NearestMap mh, mw; // dst(row/col of the layer) -> src index
    int up;            // 1 if (H,W) differs from the layer size (nearest-upsampled on the fly)
    int tma;           // 1 if the tile is fetched with a TMA box copy (map passed separately)
    int up2;           // 1 if the source is EXACTLY 2x nearest-upsampled and TMA-able: the low-resolution box is fetched by TMA
                       // into the tail of the plane and expanded in shared memory (expand_up2_planes)
};

// Compile-time loop: f(std::integral_constant<int, 0>{}) ... f(<N-1>).  The row loops of the convolution cores MUST be
// fully unrolled (accumulator indices are registers, weights are uniform-register operands); `#pragma unroll` alone was
// observed to unroll the 12-iteration row loop only 4x, leaving run-time tap indices, branches and 500 dependent LDCUs.
template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
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

// Raw gradient (gy, gc) of a layer output whose forward values were (y, cout)  ->  {GN, GD} of that layer:
//   GN = gy/(D+eps),  GD = -gy*(y-b)/(D+eps) + gc/s,  D = cout*s   (SURVEY.md appendix A),
// and the per-pixel terms of the bias gradient (gb += gy) and of gs = -sum gc*cout/s.
__device__ __forceinline__ float2 grad_xform(float gy, float gc, float y, float cout, float b, float s, float inv_s, float& gb,
                                             float& gs) {
    const float inv = rcp_approx(fmaf(cout, s, kEps));
    const float gcs = gc * inv_s;
    gb += gy;
    gs = fmaf(-gcs, cout, gs);
    return make_float2(gy * inv, fmaf(-gy * (y - b), inv, gcs));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum four per-thread values over the CTA and write them as one row {a0,a1,a2,a3} (deterministic two-pass reductions).
template <int NT>
__device__ __forceinline__ void block_sum4_store(float a0, float a1, float a2, float a3, float* row) {
    __shared__ float red[4][NT / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; red[3][warp] = a3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) t += red[threadIdx.x][w];
        row[threadIdx.x] = t;
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace adc

// K3/K4: confidence-argmax 2x2 pooling with gather (model_zoo/nconv_sd.py:43-49) and its backward, plus the
// deterministic reduction that turns full-resolution gradients of a nearest-upsampled tensor into gradients of
// the low-resolution tensor (backward of F.interpolate(mode='nearest'), model_zoo/nconv_sd.py:81-93).
// All three are pure HBM streaming kernels: coalesced row-wise access, one pooled pixel (2x2 window) per thread.
#pragma once
#include "adc_common.cuh"

namespace adc {

// ATen max_pool2d rule (aten/src/ATen/native/cpu/MaxPoolKernel.cpp / cuda/DilatedMaxPool2d.cu): start with
// maxval=-inf, maxidx=first window element; take a candidate if (val > maxval) || isnan(val), row-major order.
__device__ __forceinline__ void argmax_step(float v, int id, float& mv, int& mi) {
    if (v > mv || v != v) {
        mv = v;
        mi = id;
    }
}

__global__ void __launch_bounds__(256) conf_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ c,
                                                            float* __restrict__ x_ds, float* __restrict__ c_ds,
                                                            long long* __restrict__ idx, int H, int W, int Ho, int Wo,
                                                            long long planes) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Ho * Wo;
    if (t >= total) return;
    const int ow = (int)(t % Wo);
    const int oh = (int)((t / Wo) % Ho);
    const long long pl = t / ((long long)Wo * Ho);
    const float* cp = c + pl * H * W;
    const float* xp = x + pl * H * W;
    const int i00 = (2 * oh) * W + 2 * ow;
    float mv = -INFINITY;
    int mi = i00;
    float v0, v1, v2, v3;
    if ((W & 1) == 0 && ((reinterpret_cast<uintptr_t>(cp) & 7) == 0)) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(cp + i00));
        const float2 b = __ldg(reinterpret_cast<const float2*>(cp + i00 + W));
        v0 = a.x; v1 = a.y; v2 = b.x; v3 = b.y;
    } else {
        v0 = __ldg(cp + i00); v1 = __ldg(cp + i00 + 1); v2 = __ldg(cp + i00 + W); v3 = __ldg(cp + i00 + W + 1);
    }
    argmax_step(v0, i00, mv, mi);
    argmax_step(v1, i00 + 1, mv, mi);
    argmax_step(v2, i00 + W, mv, mi);
    argmax_step(v3, i00 + W + 1, mv, mi);
    c_ds[t] = mv / 4.0f;
    x_ds[t] = __ldg(xp + mi);
    idx[t] = (long long)mi;
}

// gx[h,w] = (idx[h/2,w/2] == h*W+w ? g_xds[h/2,w/2] : 0) + add_x[h,w]; same for gc with g_cds/4.
// One thread per 2x2 window writes its four pixels; threads of the last pooled row/column also clear the
// unpooled odd row/column.
__global__ void __launch_bounds__(256) conf_pool_bwd_kernel(const float* __restrict__ g_xds, const float* __restrict__ g_cds,
                                                            const long long* __restrict__ idx, const float* __restrict__ add_x,
                                                            const float* __restrict__ add_c, float* __restrict__ gx,
                                                            float* __restrict__ gc, int H, int W, int Ho, int Wo,
                                                            long long planes) {
    const int Wc = (W + 1) >> 1, Hc = (H + 1) >> 1;  // windows incl. the partial ones of odd sizes
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Hc * Wc;
    if (t >= total) return;
    const int ow = (int)(t % Wc);
    const int oh = (int)((t / Wc) % Hc);
    const long long pl = t / ((long long)Wc * Hc);
    const bool pooled = oh < Ho && ow < Wo;
    int mi = -1;
    float gxv = 0.f, gcv = 0.f;
    if (pooled) {
        const long long o = (pl * Ho + oh) * Wo + ow;
        mi = (int)idx[o];
        gxv = __ldg(g_xds + o);
        gcv = __ldg(g_cds + o) / 4.0f;
    }
    const long long base = pl * H * W;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int h = 2 * oh + dy;
        if (h >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int w = 2 * ow + dx;
            if (w >= W) continue;
            const int id = h * W + w;
            float ax = add_x ? __ldg(add_x + base + id) : 0.f;
            float ac = add_c ? __ldg(add_c + base + id) : 0.f;
            if (id == mi) {
                ax += gxv;
                ac += gcv;
            }
            gx[base + id] = ax;
            gc[base + id] = ac;
        }
    }
}

// g_low[hs,ws] = sum of g_full[h,w] over the (contiguous) destination ranges that map to (hs,ws).
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ g_full, float* __restrict__ g_low,
                                                           int H, int W, int Hs, int Ws, NearestMap mh, NearestMap mw,
                                                           long long planes) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Hs * Ws;
    if (t >= total) return;
    const int ws = (int)(t % Ws);
    const int hs = (int)((t / Ws) % Hs);
    const long long pl = t / ((long long)Ws * Hs);
    // first destination index whose source is >= s: dst*in/out >= s  =>  start near s*out/in and adjust
    int h_lo = (int)(((long long)hs * H) / Hs) - 2;
    if (h_lo < 0) h_lo = 0;
    while (h_lo < H && mh(h_lo) < hs) ++h_lo;
    int h_hi = h_lo;
    while (h_hi < H && mh(h_hi) == hs) ++h_hi;
    int w_lo = (int)(((long long)ws * W) / Ws) - 2;
    if (w_lo < 0) w_lo = 0;
    while (w_lo < W && mw(w_lo) < ws) ++w_lo;
    int w_hi = w_lo;
    while (w_hi < W && mw(w_hi) == ws) ++w_hi;
    const float* g = g_full + pl * H * W;
    float acc = 0.f;
    for (int h = h_lo; h < h_hi; ++h)
        for (int w = w_lo; w < w_hi; ++w) acc += __ldg(g + (long long)h * W + w);
    g_low[t] = acc;
}

}  // namespace adc

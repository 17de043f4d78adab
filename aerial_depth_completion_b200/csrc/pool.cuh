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

// ---- pair-tensor variants used inside the pyramid ---------------------------------------------------------------
namespace adc {

// {x,c} pairs in, {x_sel, c_max/4} pairs + int64 argmax out.  One pooled pixel per thread: two 16-byte loads.
__global__ void __launch_bounds__(256) pair_pool_fwd_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                            long long* __restrict__ idx, int H, int W, int Ho, int Wo,
                                                            long long planes) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Ho * Wo;
    if (t >= total) return;
    const int ow = (int)(t % Wo);
    const int oh = (int)((t / Wo) % Ho);
    const long long pl = t / ((long long)Wo * Ho);
    const float2* p = in + pl * H * W;
    const int i00 = (2 * oh) * W + 2 * ow;
    float2 v0, v1, v2, v3;
    if ((W & 1) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + i00));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + i00 + W));
        v0 = make_float2(a.x, a.y); v1 = make_float2(a.z, a.w);
        v2 = make_float2(b.x, b.y); v3 = make_float2(b.z, b.w);
    } else {
        v0 = __ldg(p + i00); v1 = __ldg(p + i00 + 1); v2 = __ldg(p + i00 + W); v3 = __ldg(p + i00 + W + 1);
    }
    float mv = -INFINITY, xs = v0.x;
    int mi = i00;
    if (v0.y > mv || v0.y != v0.y) { mv = v0.y; mi = i00; xs = v0.x; }
    if (v1.y > mv || v1.y != v1.y) { mv = v1.y; mi = i00 + 1; xs = v1.x; }
    if (v2.y > mv || v2.y != v2.y) { mv = v2.y; mi = i00 + W; xs = v2.x; }
    if (v3.y > mv || v3.y != v3.y) { mv = v3.y; mi = i00 + W + 1; xs = v3.x; }
    out[t] = make_float2(xs, mv / 4.0f);
    idx[t] = (long long)mi;
}

// Backward of the pooling + the skip-connection sum + the transform for the layer that produced the pool input A:
//   raw(h,w) = (idx == h*W+w ? {g_xds, g_cds/4} : 0) + skip(h,w);   G(h,w) = grad_xform(raw, A(h,w)).
// One thread per 2x2 window (incl. the partial windows of odd sizes, whose extra row/column only gets the skip term).
__global__ void __launch_bounds__(256) pair_pool_bwd_kernel(const float2* __restrict__ g_ds, const long long* __restrict__ idx,
                                                            const float2* __restrict__ skip, const float2* __restrict__ act,
                                                            float2* __restrict__ G, float* __restrict__ bpart, int C, int H, int W,
                                                            int Ho, int Wo, long long planes, int ximg) {
    const int Wc = (W + 1) >> 1, Hc = (H + 1) >> 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Hc * Wc;
    float gb[2] = {0.f, 0.f}, gs[2] = {0.f, 0.f};
    if (t < total) {
        const int ow = (int)(t % Wc);
        const int oh = (int)((t / Wc) % Hc);
        const long long pl = t / ((long long)Wc * Hc);
        const int ch = (int)(pl % C);
        const float* xc = &c_wimg[0][0][0][0] + (size_t)ximg * kWStride;
        const float xb = xc[kWBias + ch], xs = xc[kWSum + ch], xis = xc[kWInvSum + ch];
        int mi = -1;
        float2 gv = make_float2(0.f, 0.f);
        if (oh < Ho && ow < Wo) {
            const long long o = (pl * Ho + oh) * Wo + ow;
            mi = (int)idx[o];
            gv = __ldg(g_ds + o);
            gv.y *= 0.25f;
        }
        const long long base = pl * H * W;
        float gbl = 0.f, gsl = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int h = 2 * oh + dy;
            if (h >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int w = 2 * ow + dx;
                if (w >= W) continue;
                const int id = h * W + w;
                float2 r = skip ? __ldg(skip + base + id) : make_float2(0.f, 0.f);
                if (id == mi) { r.x += gv.x; r.y += gv.y; }
                const float2 a = __ldg(act + base + id);
                G[base + id] = grad_xform(r.x, r.y, a.x, a.y, xb, xs, xis, gbl, gsl);
            }
        }
        // C == 2 in the pyramid: route the per-thread sums to their channel's slot
        if (ch == 0) { gb[0] = gbl; gs[0] = gsl; } else { gb[1] = gbl; gs[1] = gsl; }
    }
    block_sum4_store<256>(gb[0], gb[1], gs[0], gs[1], bpart + (size_t)blockIdx.x * 4);
}

// Backward of the on-the-fly nearest upsampling: sum the raw layer-resolution pairs over the destination block that maps
// to each low-resolution pixel, then transform for the layer that produced the low-resolution tensor.
__global__ void __launch_bounds__(256) pair_upsample_bwd_kernel(const float2* __restrict__ g_full, const float2* __restrict__ act,
                                                                float2* __restrict__ G, float* __restrict__ bpart, int C, int H,
                                                                int W, int Hs, int Ws, NearestMap mh, NearestMap mw,
                                                                long long planes, int ximg) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes * Hs * Ws;
    float gb[2] = {0.f, 0.f}, gs[2] = {0.f, 0.f};
    if (t < total) {
        const int ws = (int)(t % Ws);
        const int hs = (int)((t / Ws) % Hs);
        const long long pl = t / ((long long)Ws * Hs);
        const int ch = (int)(pl % C);
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
        const float2* g = g_full + pl * H * W;
        float2 acc = make_float2(0.f, 0.f);
        for (int h = h_lo; h < h_hi; ++h)
            for (int w = w_lo; w < w_hi; ++w) {
                const float2 v = __ldg(g + (long long)h * W + w);
                acc.x += v.x;
                acc.y += v.y;
            }
        if (ximg >= 0) {
            const float* xc = &c_wimg[0][0][0][0] + (size_t)ximg * kWStride;
            const float2 a = __ldg(act + t);
            float gbl = 0.f, gsl = 0.f;
            G[t] = grad_xform(acc.x, acc.y, a.x, a.y, xc[kWBias + ch], xc[kWSum + ch], xc[kWInvSum + ch], gbl, gsl);
            if (ch == 0) { gb[0] = gbl; gs[0] = gsl; } else { gb[1] = gbl; gs[1] = gsl; }
        } else {
            G[t] = acc;   // raw mode (single-layer API)
        }
    }
    if (ximg >= 0) block_sum4_store<256>(gb[0], gb[1], gs[0], gs[1], bpart + (size_t)blockIdx.x * 4);
}

// Exact 2x case (H = 2*Hs, W = 2*Ws: mode-1 maps, every low-resolution pixel owns a 2x2 block): 32-bit index arithmetic and two
// 16-byte loads per thread -- the general kernel above spends its time in 64-bit divisions and the child-range searches
// (sm 69 % busy at 2.2 TB/s for the 228x304 -> 114x152 launch).
__global__ void __launch_bounds__(256) pair_upsample2_bwd_kernel(const float2* __restrict__ g_full, const float2* __restrict__ act,
                                                                 float2* __restrict__ G, float* __restrict__ bpart, int C, int Hs,
                                                                 int Ws, unsigned total, int ximg) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    float gb[2] = {0.f, 0.f}, gs[2] = {0.f, 0.f};
    if (t < total) {
        const unsigned ws = t % (unsigned)Ws, r = t / (unsigned)Ws, hs = r % (unsigned)Hs, pl = r / (unsigned)Hs;
        const int ch = (int)(pl % (unsigned)C);
        const float4* g = reinterpret_cast<const float4*>(g_full + ((size_t)pl * 2 * Hs + 2 * hs) * (2 * (size_t)Ws)) + ws;
        const float4 v0 = __ldg(g), v1 = __ldg(g + Ws);          // rows 2hs and 2hs+1 (a row is 2*Ws pairs = Ws float4)
        const float2 acc = make_float2(((v0.x + v0.z) + v1.x) + v1.z, ((v0.y + v0.w) + v1.y) + v1.w);   // the general kernel's order
        if (ximg >= 0) {
            const float* xc = &c_wimg[0][0][0][0] + (size_t)ximg * kWStride;
            const float2 a = __ldg(act + t);
            float gbl = 0.f, gsl = 0.f;
            G[t] = grad_xform(acc.x, acc.y, a.x, a.y, xc[kWBias + ch], xc[kWSum + ch], xc[kWInvSum + ch], gbl, gsl);
            if (ch == 0) { gb[0] = gbl; gs[0] = gsl; } else { gb[1] = gbl; gs[1] = gsl; }
        } else {
            G[t] = acc;
        }
    }
    if (ximg >= 0) block_sum4_store<256>(gb[0], gb[1], gs[0], gs[1], bpart + (size_t)blockIdx.x * 4);
}

}  // namespace adc

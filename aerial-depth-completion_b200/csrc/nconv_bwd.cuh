// K2: NConv2d backward for sm_100a, as the reference's autograd graph computes it.
//
//   gNm = gy/(D+eps)      gD = -gy*(y-b)/(D+eps) + gcout/s      (D = cout*s, s = sum(w_fwd))
//   data : g_xc[i] = corr^T(gNm, W*)   g_c'[i] = corr^T(gD, W*)   gx = g_xc*c   gc = g_xc*x + g_c'
//   weight: gW[o,i,u,v] = sum_pix gNm[o]*xc[i](shifted) + gD[o]*c[i](shifted)  + gs[o],  gs = -sum gcout*cout/s
//   bias : gb[o] = sum gy
//
// Two kernels share the gNm/gD tile builder:
//  * nconv_bwd_data_kernel   -- same 64x32 tile / 4x4 patch / packed-FFMA2 core as the forward (conv_patch),
//    with the planes {gD,gNm} as inputs, the flipped backward-time weights W* and pairs of INPUT channels as the
//    packed lanes; the epilogue multiplies by c / x and stores float4.
//  * nconv_bwd_weight_kernel -- every warp owns one (o,i) weight slice; a lane streams a 4-pixel-wide column strip
//    down the tile keeping a K-row register window of {gNm,gD} pairs, so each {xc,c} pair read from shared memory
//    feeds K*K*... FFMA2 ({gNm*xc, gD*c} in the two packed lanes); the K*K accumulators are then reduced with
//    warp shuffles and written as per-CTA partials (deterministic two-pass reduction, no atomics).
#pragma once
#include "adc_common.cuh"
#include "nconv_fwd.cuh"

namespace adc {

struct GradIn {
    const float* y;
    const float* cout;
    const float* gy;
    const float* gcout;
};

// gNm / gD at one output pixel of channel o (zero outside the image).
__device__ __forceinline__ void grad_terms(const GradIn& g, long long off, float b, float s, float is, float& gNm, float& gD,
                                           float& gyv, float& gsv) {
    const float yv = __ldg(g.y + off), cv = __ldg(g.cout + off);
    gyv = __ldg(g.gy + off);
    const float gcv = __ldg(g.gcout + off);
    const float inv = rcp_approx(fmaf(cv, s, kEps));
    gNm = gyv * inv;
    const float gcs = gcv * is;
    gD = fmaf(-gyv * (yv - b), inv, gcs);
    gsv = -gcs * cv;
}

template <int CI, int K>
__global__ void __launch_bounds__(128) nconv_bwd_data_kernel(Src s0, Src s1, GradIn g, float* __restrict__ gx0,
                                                             float* __restrict__ gc0, float* __restrict__ gx1,
                                                             float* __restrict__ gc1, int H, int W, int wimg) {
    using C = FwdCfg<2, K>;
    constexpr int P = C::P;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    const long long plane = (long long)H * W;

    // planes: o -> {gD (A), gNm (B)}
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const float b = wc[kWBias + o], s = wc[kWSum + o], is = wc[kWInvSum + o];
        const long long base = ((long long)n * 2 + o) * plane;
        float* sa = smem + (o * 2) * C::PLANE;
        float* sb = sa + C::PLANE;
        for (int i = tid; i < C::SH * C::SW; i += C::NT) {
            const int r = i / C::SW, col = i % C::SW;
            const int gh = h0 - P + r, gw = w0 - P + col;
            float gNm = 0.f, gD = 0.f, t0, t1;
            if (gh >= 0 && gh < H && gw >= 0 && gw < W) grad_terms(g, base + (long long)gh * W + gw, b, s, is, gNm, gD, t0, t1);
            sa[r * C::PITCH + col] = gD;
            sb[r * C::PITCH + col] = gNm;
        }
    }
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;
    const int oh = h0 + ty * 4, ow = w0 + tx * 4;
    if (oh >= H || ow >= W) return;

    constexpr int NQ = (CI + 1) / 2;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float2 accC[4][4], accX[4][4];
        conv_patch<2, K>(smem, reinterpret_cast<const float2*>(wc) + q * (2 * K * K), tx, ty, accC, accX);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ci = 2 * q + j;
            if (ci >= CI) break;
            const bool first = ci < s0.ch;
            const Src& s = first ? s0 : s1;
            const int sc = first ? ci : ci - s0.ch;
            const float* xp = s.x + (long long)n * s.bs + (long long)sc * s.H * s.W;
            const float* cp = s.c + (long long)n * s.bs + (long long)sc * s.H * s.W;
            // gradients are stored at the layer resolution, channel-contiguous per source: [N, s.ch, H, W]
            float* gxo = (first ? gx0 : gx1) + ((long long)n * s.ch + sc) * plane;
            float* gco = (first ? gc0 : gc1) + ((long long)n * s.ch + sc) * plane;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int h = oh + r;
                if (h >= H) break;
                const int hsrc = s.mh(h);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int w = ow + p;
                    if (w >= W) break;
                    const long long so = (long long)hsrc * s.W + s.mw(w);
                    const float xv = __ldg(xp + so), cv = __ldg(cp + so);
                    const float gxc = j == 0 ? accX[r][p].x : accX[r][p].y;
                    const float gcc = j == 0 ? accC[r][p].x : accC[r][p].y;
                    gxo[(long long)h * W + w] = gxc * cv;
                    gco[(long long)h * W + w] = fmaf(gxc, xv, gcc);
                }
            }
        }
    }
}

template <int CI, int K>
struct WgCfg {
    static constexpr int P = K / 2;
    static constexpr int TW = 64, TH = 32, RG = 2, RH = TH / RG;   // two row groups of 16 rows
    static constexpr int ROLES = 2 * CI;                           // (o,i), Co = 2
    static constexpr int NT = ROLES * 32;
    static constexpr int XP = TW + 2 * P;                          // pitch in float2 pairs
    static constexpr int XR = TH + 2 * P;
    static constexpr int STEPS = RH + K - 1;                       // 20 (K=5), 18 (K=3), 16 (K=1): multiples of K
    static constexpr int SMEM_BYTES = (CI * XR * XP + 2 * TH * TW) * 8;
    static_assert(STEPS % K == 0, "row streaming is unrolled by K");
};

template <int CI, int K>
__global__ void __launch_bounds__(WgCfg<CI, K>::NT) nconv_bwd_weight_kernel(Src s0, Src s1, GradIn g, float* __restrict__ partial,
                                                                            int H, int W, int wimg) {
    using C = WgCfg<CI, K>;
    constexpr int P = C::P;
    extern __shared__ __align__(16) float smem[];
    float2* sX = reinterpret_cast<float2*>(smem);                    // [CI][XR][XP]  {x*c, c}
    float2* sG = sX + CI * C::XR * C::XP;                            // [2][TH][TW]   {gNm, gD}
    __shared__ float sred[4][C::NT / 32];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    const long long plane = (long long)H * W;

    float gb_acc[2] = {0.f, 0.f}, gs_acc[2] = {0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const float b = wc[kWBias + o], s = wc[kWSum + o], is = wc[kWInvSum + o];
        const long long base = ((long long)n * 2 + o) * plane;
        for (int i = tid; i < C::TH * C::TW; i += C::NT) {
            const int r = i / C::TW, col = i % C::TW;
            const int gh = h0 + r, gw = w0 + col;
            float gNm = 0.f, gD = 0.f, gyv = 0.f, gsv = 0.f;
            if (gh < H && gw < W) grad_terms(g, base + (long long)gh * W + gw, b, s, is, gNm, gD, gyv, gsv);
            sG[(o * C::TH + r) * C::TW + col] = make_float2(gNm, gD);
            gb_acc[o] += gyv;
            gs_acc[o] += gsv;
        }
    }
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
        const bool first = ci < s0.ch;
        const Src& s = first ? s0 : s1;
        const int sc = first ? ci : ci - s0.ch;
        const float* xp = s.x + (long long)n * s.bs + (long long)sc * s.H * s.W;
        const float* cp = s.c + (long long)n * s.bs + (long long)sc * s.H * s.W;
        for (int i = tid; i < C::XR * C::XP; i += C::NT) {
            const int r = i / C::XP, col = i % C::XP;
            const int gh = h0 - P + r, gw = w0 - P + col;
            float cv = 0.f, xv = 0.f;
            if (gh >= 0 && gh < H && gw >= 0 && gw < W) {
                const long long off = (long long)s.mh(gh) * s.W + s.mw(gw);
                cv = __ldg(cp + off);
                xv = __ldg(xp + off);
            }
            sX[(ci * C::XR + r) * C::XP + col] = make_float2(xv * cv, cv);
        }
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int o = warp / CI, ci = warp % CI;
    const int grp = lane >> 4, strip = lane & 15;
    const float2* gptr = sG + (o * C::TH + grp * C::RH) * C::TW + strip * 4;
    const float2* xptr = sX + (ci * C::XR + grp * C::RH) * C::XP + strip * 4;

    float2 acc[K][K];
    float2 win[K][4];
#pragma unroll
    for (int u = 0; u < K; ++u) {
#pragma unroll
        for (int v = 0; v < K; ++v) acc[u][v] = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 4; ++p) win[u][p] = make_float2(0.f, 0.f);
    }

    // step j reads X row (grp*RH + j) [tile row grp*RH + j - P] and G row grp*RH + j (zero once j >= RH)
#pragma unroll 1
    for (int jt = 0; jt < C::STEPS / K; ++jt) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const int j = jt * K + jj;
            if (j < C::RH) {
                const float4 a = *reinterpret_cast<const float4*>(gptr + j * C::TW);
                const float4 b = *reinterpret_cast<const float4*>(gptr + j * C::TW + 2);
                win[jj][0] = make_float2(a.x, a.y); win[jj][1] = make_float2(a.z, a.w);
                win[jj][2] = make_float2(b.x, b.y); win[jj][3] = make_float2(b.z, b.w);
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p) win[jj][p] = make_float2(0.f, 0.f);
            }
            float2 xr[4 + K - 1 + 1];
#pragma unroll
            for (int t = 0; t < (4 + K - 1 + 1) / 2; ++t) {
                const float4 a = *reinterpret_cast<const float4*>(xptr + j * C::XP + 2 * t);
                xr[2 * t] = make_float2(a.x, a.y);
                xr[2 * t + 1] = make_float2(a.z, a.w);
            }
#pragma unroll
            for (int u = 0; u < K; ++u) {
                // G row of this step minus u  ->  window slot (jj - u) mod K
                const int slot = ((jj - u) % K + K) % K;
#pragma unroll
                for (int v = 0; v < K; ++v)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[u][v] = ffma2(win[slot][p], xr[p + v], acc[u][v]);
            }
        }
    }

    float* out = partial + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kWStride;
#pragma unroll
    for (int u = 0; u < K; ++u)
#pragma unroll
        for (int v = 0; v < K; ++v) {
            const float r = warp_sum(acc[u][v].x + acc[u][v].y);
            if (lane == 0) out[((o * CI + ci) * K + u) * K + v] = r;
        }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float a = warp_sum(gb_acc[q]), b = warp_sum(gs_acc[q]);
        if (lane == 0) {
            sred[q][warp] = a;
            sred[2 + q][warp] = b;
        }
    }
    __syncthreads();
    if (tid < 4) {
        float t = 0.f;
        for (int w = 0; w < C::NT / 32; ++w) t += sred[tid][w];
        out[kWBias + tid] = t;   // [100,101] = gb partial, [102,103] = gs partial
    }
}

// nconv7 (2 -> 1, 1x1) backward: everything is per-pixel; one float4 of pixels per thread, block-reduced partials.
__global__ void __launch_bounds__(256) nconv_1x1_c2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ c,
                                                               long long bs, GradIn g, float* __restrict__ gx,
                                                               float* __restrict__ gc, float* __restrict__ partial,
                                                               long long plane, int wimg_f, int wimg_b) {
    const float* wf = &c_wimg[0][0][0][0] + (size_t)wimg_f * kWStride;  // forward image: b, s
    const float* wb = &c_wimg[0][0][0][0] + (size_t)wimg_b * kWStride;  // backward-time weights (index = ci)
    const float b = wf[kWBias], s = wf[kWSum], is = wf[kWInvSum];
    const float w0 = wb[0], w1 = wb[1];
    const int n = blockIdx.y;
    const float* xp = x + (long long)n * bs;
    const float* cp = c + (long long)n * bs;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float a_w0 = 0.f, a_w1 = 0.f, a_b = 0.f, a_s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long i = i0 + j;
        if (i >= plane) break;
        float gNm, gD, gyv, gsv;
        grad_terms(g, (long long)n * plane + i, b, s, is, gNm, gD, gyv, gsv);
        const float x0 = __ldg(xp + i), x1 = __ldg(xp + plane + i), c0 = __ldg(cp + i), c1 = __ldg(cp + plane + i);
        a_w0 += fmaf(gNm, x0 * c0, gD * c0);
        a_w1 += fmaf(gNm, x1 * c1, gD * c1);
        a_b += gyv;
        a_s += gsv;
        if (gx) {
            const float gxc0 = gNm * w0, gxc1 = gNm * w1;
            gx[((long long)n * 2) * plane + i] = gxc0 * c0;
            gx[((long long)n * 2 + 1) * plane + i] = gxc1 * c1;
            gc[((long long)n * 2) * plane + i] = fmaf(gxc0, x0, gD * w0);
            gc[((long long)n * 2 + 1) * plane + i] = fmaf(gxc1, x1, gD * w1);
        }
    }
    __shared__ float sred[4][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    a_w0 = warp_sum(a_w0); a_w1 = warp_sum(a_w1); a_b = warp_sum(a_b); a_s = warp_sum(a_s);
    if (lane == 0) { sred[0][warp] = a_w0; sred[1][warp] = a_w1; sred[2][warp] = a_b; sred[3][warp] = a_s; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sred[threadIdx.x][w];
        float* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kWStride;
        const int slot = threadIdx.x == 0 ? 0 : (threadIdx.x == 1 ? 1 : (threadIdx.x == 2 ? kWBias : kWSum));
        out[slot] = t;
        if (threadIdx.x == 0) { out[kWBias + 1] = 0.f; out[kWSum + 1] = 0.f; }
    }
}

// Second pass of the weight gradient: sum the per-CTA partials of every call of a module, add gs, write/accumulate.
struct WgReduceCall {
    const float* partial;
    int nblocks;
    int module;   // which output the call accumulates into
    int Co, n;    // n = Co*Ci*K*K
};
struct WgReduceArgs {
    WgReduceCall call[ADC_PYR_CALLS];
    int ncalls;
    float* gw[ADC_PYR_MODULES];
    float* gb[ADC_PYR_MODULES];
    int accumulate;
};

__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(WgReduceArgs a) {
    const int m = blockIdx.x;
    const int t = threadIdx.x & 127, slice = threadIdx.x >> 7;   // 8 slices of 128 entries
    __shared__ float sh[8][kWStride];
    float acc = 0.f;
    int n = 0, Co = 1;
    for (int k = 0; k < a.ncalls; ++k) {
        const WgReduceCall& c = a.call[k];
        if (c.module != m) continue;
        n = c.n;
        Co = c.Co;
        for (int blk = slice; blk < c.nblocks; blk += 8) acc += c.partial[(size_t)blk * kWStride + t];
    }
    sh[slice][t] = acc;
    __syncthreads();
    if (slice == 0 && n > 0) {
        float v = 0.f;
        for (int s = 0; s < 8; ++s) v += sh[s][t];
        sh[0][t] = v;
    }
    __syncthreads();
    if (slice == 0 && n > 0) {
        if (t < n) {
            const int co = t / (n / Co);
            const float v = sh[0][t] + sh[0][kWSum + co];
            a.gw[m][t] = a.accumulate ? a.gw[m][t] + v : v;
        }
        if (t < Co && a.gb[m]) {
            const float v = sh[0][kWBias + t];
            a.gb[m][t] = a.accumulate ? a.gb[m][t] + v : v;
        }
    }
}

template <int CI, int K>
int launch_nconv_bwd_data(const Src& s0, const Src& s1, const GradIn& g, float* gx0, float* gc0, float* gx1, float* gc1,
                          int N, int H, int W, int wimg, cudaStream_t st) {
    using C = FwdCfg<2, K>;
    static bool attr_done = false;
    if (!attr_done) {
        ADC_CUDA(cudaFuncSetAttribute(nconv_bwd_data_kernel<CI, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    dim3 grid(ceil_div(W, C::TW), ceil_div(H, C::TH), N);
    nconv_bwd_data_kernel<CI, K><<<grid, C::NT, C::SMEM_BYTES, st>>>(s0, s1, g, gx0, gc0, gx1, gc1, H, W, wimg);
    return check_cuda(cudaGetLastError(), "nconv_bwd_data_kernel launch");
}

template <int CI, int K>
int launch_nconv_bwd_weight(const Src& s0, const Src& s1, const GradIn& g, float* partial, int N, int H, int W, int wimg,
                            cudaStream_t st) {
    using C = WgCfg<CI, K>;
    static bool attr_done = false;
    if (!attr_done) {
        ADC_CUDA(cudaFuncSetAttribute(nconv_bwd_weight_kernel<CI, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    dim3 grid(ceil_div(W, C::TW), ceil_div(H, C::TH), N);
    nconv_bwd_weight_kernel<CI, K><<<grid, C::NT, C::SMEM_BYTES, st>>>(s0, s1, g, partial, H, W, wimg);
    return check_cuda(cudaGetLastError(), "nconv_bwd_weight_kernel launch");
}

}  // namespace adc

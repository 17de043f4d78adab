// The non-GEMM pieces of S2DResNet (model_zoo/s2d_resnet.py:159-242, resnet18 of BASELINE config 5) around the tensor-core
// convolutions of conv_tc.cu and the BatchNorm passes of bn_act.cu.  Layout as there: activations = zero-bordered NHWC bf16 (hi, lo)
// pairs, raw convolution results / gradients = zero-bordered NHWC fp32, every tensor addressed as a channel slice of a wider pixel.
//
//   im2col_kxk        the 7x7 / stride-2 / pad-3 stem (:175, in_channels 3..5) becomes ONE GEMM: channel (u*K + v)*C + c of the
//                     output-resolution pixel (the conv1 trick of ged_quad.cu, general window)
//   maxpool3x3s2      nn.MaxPool2d(3, 2, 1) (:181) forward with the ATen scan order / tie rule (first maximum, `>` or NaN), window
//                     position kept as one byte per element; backward as a gather (every input pixel looks at the <= 4 windows that
//                     contain it: no atomics, deterministic)
//   pixel (un)shuffle the 5x5 convolutions of UpProj (:105-143) act on a zero-stuffed 2x up-sampling (Unpool :11-24): output pixel
//                     (2i+py, 2j+px) only sees the taps u = py mod 2, v = px mod 2, i.e. a 3x3 stride-1 convolution of the LOW-resolution
//                     tensor with 4 parity blocks of output channels (3x3, 3x2, 2x3, 2x2 live taps: 25 of 36).  The convolution, its data
//                     and weight gradients run on the stride-1 kernels at a quarter of the pixels; these two kernels move between the
//                     parity-block layout [N][h][w][4*C] and the full-resolution tensor [N][2h][2w][C]
//   bilinear          F.interpolate(mode='bilinear', align_corners=True) (:230-234) to NCHW fp32 and its backward as a gather
// Algorithmic bytes: every kernel reads its input once and writes its output once.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace s2d {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void load8_bf16(const bf16* p, float (&v)[8]) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* rr = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __low2float(rr[k]); v[2 * k + 1] = __high2float(rr[k]); }
}
__device__ __forceinline__ void store8_pair(bf16* hi, bf16* lo, const float (&o)[8]) {
    uint4 h, l;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
    __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hh[j] = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
        ll[j] = __floats2bfloat162_rn(o[2 * j] - __low2float(hh[j]), o[2 * j + 1] - __high2float(hh[j]));
    }
    *reinterpret_cast<uint4*>(hi) = h;
    if (lo) *reinterpret_cast<uint4*>(lo) = l;
}
__device__ __forceinline__ long long bpix(long long n, int y, int x, int H, int W) {      // pixel (y, x) of a zero-bordered tensor
    return (n * (H + 2) + y + 1) * (W + 2) + x + 1;
}

static int grid_for(long long total, int threads = 256) {
    long long b = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---- im2col of a K x K window ----------------------------------------------------------------------------------------------------------
// thread = (output pixel, 8-channel group of the destination); channels >= K*K*C are written as zeros up to `groups` * 8
__global__ void __launch_bounds__(256) im2col_kxk_kernel(const float* __restrict__ src, long long bs, int C, int K, int stride, int pad,
                                                         bf16* __restrict__ dst, bf16* __restrict__ dst_lo, int dst_stride, int groups, int N,
                                                         int H, int W, int Ho, int Wo) {
    // (channel, tap row, tap column) of every destination channel, once per CTA: the first version derived them with five runtime
    // integer divisions per ELEMENT in 64-bit index arithmetic and ran at 1.1 TB/s (1.15 ms for the 7x7 stem at B = 64)
    __shared__ int tab_off[1024];          // c * H * W  (or -1: a pad channel)
    __shared__ short tab_u[1024], tab_w[1024];
    const int KKC = K * K * C, Q = groups * 8;
    const int HW = H * W;                  // (host: H * W * C < 2^31)
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        if (q < KKC) {
            const int c = q % C, tap = q / C, u = tap / K;
            tab_off[q] = c * HW;
            tab_u[q] = (short)(u - pad);
            tab_w[q] = (short)(tap - u * K - pad);
        } else {
            tab_off[q] = -1;
            tab_u[q] = tab_w[q] = 0;
        }
    }
    __syncthreads();
    const unsigned total = (unsigned)N * Ho * Wo * groups;      // (host: < 2^31)
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned g = t % (unsigned)groups, r = t / (unsigned)groups;
        const unsigned j = r % (unsigned)Wo, r2 = r / (unsigned)Wo, i = r2 % (unsigned)Ho, n = r2 / (unsigned)Ho;
        const float* sn = src + (long long)n * bs;
        const int y0 = stride * (int)i, x0 = stride * (int)j;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int q = (int)g * 8 + k;
            const int off = tab_off[q], y = y0 + tab_u[q], x = x0 + tab_w[q];
            const bool ok = off >= 0 && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
            o[k] = ok ? __ldg(sn + off + y * W + x) : 0.f;
        }
        const long long p = bpix(n, i, j, Ho, Wo);
        store8_pair(dst + p * dst_stride + g * 8, dst_lo ? dst_lo + p * dst_stride + g * 8 : nullptr, o);
    }
}

// Adjoint of im2col_kxk: gx[n][c][y][x] = sum over the taps (u, v) and output pixels (i, j) with stride*i + u - pad == y, stride*j + v - pad == x
// of gcol[n][i][j][(u*K + v)*C + c].  A gather (every input pixel visits the <= ceil(K/stride)^2 windows that contain it): no atomics.
// thread = one input pixel (x fastest: coalesced NCHW stores), loops over the C planes.
__global__ void __launch_bounds__(256) col2im_kxk_kernel(const float* __restrict__ gcol, int gs, int C, int K, int stride, int pad,
                                                         float* __restrict__ gx, int N, int H, int W, int Ho, int Wo) {
    const long long total = (long long)N * H * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % W), y = (int)((t / W) % H);
        const long long n = t / ((long long)W * H);
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            for (int u = (y + pad) % stride; u < K; u += stride) {
                const int i = (y + pad - u) / stride;
                if (y + pad - u < 0 || i >= Ho) continue;
                for (int v = (x + pad) % stride; v < K; v += stride) {
                    const int j = (x + pad - v) / stride;
                    if (x + pad - v < 0 || j >= Wo) continue;
                    acc += __ldg(gcol + bpix(n, i, j, Ho, Wo) * gs + (u * K + v) * C + c);
                }
            }
            gx[((n * C + c) * H + y) * W + x] = acc;
        }
    }
}

// ---- MaxPool2d(3, 2, 1) ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const bf16* __restrict__ xh, const bf16* __restrict__ xl, int xs,
                                                          bf16* __restrict__ oh, bf16* __restrict__ ol, int os, unsigned char* __restrict__ idx,
                                                          int C8, int N, int H, int W, int Ho, int Wo) {
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % C8);
        const long long r = t / C8;
        const int j = (int)(r % Wo), i = (int)((r / Wo) % Ho);
        const long long n = r / ((long long)Wo * Ho);
        float best[8];
        int bi[8];
        bool first = true;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int y = 2 * i + u - 1;
            if (y < 0 || y >= H) continue;
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const int x = 2 * j + w - 1;
                if (x < 0 || x >= W) continue;
                const long long p = bpix(n, y, x, H, W);
                float v[8];
                load8_bf16(xh + p * xs + g * 8, v);
                if (xl) {
                    float v2[8];
                    load8_bf16(xl + p * xs + g * 8, v2);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += v2[k];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (first || v[k] > best[k] || v[k] != v[k]) { best[k] = v[k]; bi[k] = u * 3 + w; }
                first = false;
            }
        }
        const long long po = bpix(n, i, j, Ho, Wo);
        store8_pair(oh + po * os + g * 8, ol ? ol + po * os + g * 8 : nullptr, best);
        unsigned long long packed = 0ull;
#pragma unroll
        for (int k = 0; k < 8; ++k) packed |= (unsigned long long)(unsigned)bi[k] << (8 * k);
        *reinterpret_cast<unsigned long long*>(idx + (r * C8 + g) * 8) = packed;
    }
}

// gin[y][x][c] = sum over the windows (oy, ox) containing (y, x) whose recorded position is (y, x) of g[oy][ox][c]
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ g, int gs, const unsigned char* __restrict__ idx,
                                                          float* __restrict__ gin, int gis, int C8, int N, int H, int W, int Ho, int Wo) {
    const long long total = (long long)N * H * W * C8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int gq = (int)(t % C8);
        const long long r = t / C8;
        const int x = (int)(r % W), y = (int)((r / W) % H);
        const long long n = r / ((long long)W * H);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int oy0 = y / 2, ox0 = x / 2;                 // windows oy with 2*oy - 1 <= y <= 2*oy + 1: floor(y/2) and, for odd y, +1
        for (int a = 0; a < 2; ++a) {
            const int oy = oy0 + a;
            if (oy >= Ho || (a == 1 && (y & 1) == 0)) continue;
            const int u = y - (2 * oy - 1);
            for (int b = 0; b < 2; ++b) {
                const int ox = ox0 + b;
                if (ox >= Wo || (b == 1 && (x & 1) == 0)) continue;
                const int w = x - (2 * ox - 1);
                const unsigned long long packed =
                    __ldg(reinterpret_cast<const unsigned long long*>(idx + (((n * Ho + oy) * Wo + ox) * C8 + gq) * 8));
                const long long po = bpix(n, oy, ox, Ho, Wo);
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + po * gs + gq * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + po * gs + gq * 8) + 1);
                const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((int)((packed >> (8 * k)) & 0xff) == u * 3 + w) acc[k] += gv[k];
            }
        }
        const long long p = bpix(n, y, x, H, W);
        float4* o = reinterpret_cast<float4*>(gin + p * gis + gq * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ---- parity-block layout <-> full resolution ---------------------------------------------------------------------------------------------
// Y[n][2i+py][2j+px][c] (+)= Q[n][i][j][(2py+px)*C + c]   (fp32, 4 channels per thread); Y is H x W with H in {2h-1, 2h}, W likewise
// (the larger tensor of a stride-2 layer with an odd size)
__global__ void __launch_bounds__(256) shuffle_f32_kernel(const float* __restrict__ q, int qs, float* __restrict__ y, int ys, int C4, int N,
                                                          int h, int w, int H, int W, int accumulate) {
    const long long total = (long long)N * H * W * C4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % C4);
        const long long r = t / C4;
        const int X = (int)(r % W), Y = (int)((r / W) % H);
        const long long n = r / ((long long)W * H);
        const int par = 2 * (Y & 1) + (X & 1);
        float4 v = __ldg(reinterpret_cast<const float4*>(q + bpix(n, Y >> 1, X >> 1, h, w) * qs + (par * C4 + g) * 4));
        float4* o = reinterpret_cast<float4*>(y + bpix(n, Y, X, H, W) * ys + g * 4);
        if (accumulate) {
            const float4 a = *o;
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        *o = v;
    }
}
// Q[n][i][j][(2py+px)*C + c] = D[n][2i+py][2j+px][c]   (bf16 pair, 8 channels per thread)
__global__ void __launch_bounds__(256) unshuffle_pair_kernel(const bf16* __restrict__ dh, const bf16* __restrict__ dl, int ds,
                                                             bf16* __restrict__ qh, bf16* __restrict__ ql, int qs, int C8, int N, int h, int w) {
    const int H = 2 * h, W = 2 * w;
    const long long total = (long long)N * H * W * C8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % C8);
        const long long r = t / C8;
        const int X = (int)(r % W), Y = (int)((r / W) % H);
        const long long n = r / ((long long)W * H);
        const int par = 2 * (Y & 1) + (X & 1);
        const long long ps = bpix(n, Y, X, H, W), pd = bpix(n, Y >> 1, X >> 1, h, w);
        *reinterpret_cast<uint4*>(qh + pd * qs + (par * C8 + g) * 8) = __ldg(reinterpret_cast<const uint4*>(dh + ps * ds + g * 8));
        if (dl) *reinterpret_cast<uint4*>(ql + pd * qs + (par * C8 + g) * 8) = __ldg(reinterpret_cast<const uint4*>(dl + ps * ds + g * 8));
    }
}

// ---- bilinear, align_corners = True --------------------------------------------------------------------------------------------------------
// ATen's arithmetic (UpSampleBilinear2d.cu): scale = (in - 1) / (out - 1) in fp32 (0 if out == 1), src = scale * dst, i0 = int(src),
// l1 = src - i0, i1 = i0 + (i0 < in - 1), out = h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11).
__device__ __forceinline__ float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

template <bool PAIR>
__device__ __forceinline__ float src_at(const void* hi, const void* lo, long long off) {
    if (PAIR) {
        float v = __bfloat162float(__ldg(reinterpret_cast<const bf16*>(hi) + off));
        if (lo) v += __bfloat162float(__ldg(reinterpret_cast<const bf16*>(lo) + off));
        return v;
    }
    return __ldg(reinterpret_cast<const float*>(hi) + off);
}

// thread = one destination pixel (x fastest: coalesced NCHW stores), loops over the C channels
template <bool PAIR>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const void* __restrict__ sh, const void* __restrict__ sl, int ss, int C, int N,
                                                           int hs, int ws, float* __restrict__ dst, int Hd, int Wd) {
    const long long total = (long long)N * Hd * Wd;
    const float sy = ac_scale(hs, Hd), sx = ac_scale(ws, Wd);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(t % Wd), Y = (int)((t / Wd) % Hd);
        const long long n = t / ((long long)Wd * Hd);
        const float fy = sy * Y, fx = sx * X;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
        const float h1 = fy - y0, h0 = 1.f - h1, w1 = fx - x0, w0 = 1.f - w1;
        const long long p00 = bpix(n, y0, x0, hs, ws) * ss, p01 = bpix(n, y0, x1, hs, ws) * ss;
        const long long p10 = bpix(n, y1, x0, hs, ws) * ss, p11 = bpix(n, y1, x1, hs, ws) * ss;
        for (int c = 0; c < C; ++c) {
            const float v = h0 * (w0 * src_at<PAIR>(sh, sl, p00 + c) + w1 * src_at<PAIR>(sh, sl, p01 + c)) +
                            h1 * (w0 * src_at<PAIR>(sh, sl, p10 + c) + w1 * src_at<PAIR>(sh, sl, p11 + c));
            dst[((n * C + c) * Hd + Y) * Wd + X] = v;
        }
    }
}

// Destination rows whose footprint {y0, y1} contains source row i lie in a short range around i / scale; every candidate recomputes
// ATen's indices and weights, so the gather visits exactly the (dst, src) pairs of the forward.  thread = (channel, source pixel).
__device__ __forceinline__ void dst_range(int i, int in, int out, float scale, int& lo, int& hi) {
    if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
    const float inv = 1.f / scale;
    lo = (int)floorf((i - 1) * inv) - 1;
    hi = (int)ceilf((i + 1) * inv) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > out - 1 ? out - 1 : hi;
}

__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ g, int C, int N, int Hd, int Wd, float* __restrict__ gs,
                                                           int gss, int hs, int ws, int accumulate) {
    const long long total = (long long)N * hs * ws * C;
    const float sy = ac_scale(hs, Hd), sx = ac_scale(ws, Wd);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        // x fastest: neighbouring threads read neighbouring columns of the same NCHW gradient plane
        const int j = (int)(t % ws), i = (int)((t / ws) % hs);
        const int c = (int)((t / ((long long)ws * hs)) % C);
        const long long n = t / ((long long)ws * hs * C);
        int ylo, yhi, xlo, xhi;
        dst_range(i, hs, Hd, sy, ylo, yhi);
        dst_range(j, ws, Wd, sx, xlo, xhi);
        const float* gp = g + (n * C + c) * (long long)Hd * Wd;
        float acc = 0.f;
        for (int Y = ylo; Y <= yhi; ++Y) {
            const float fy = sy * Y;
            const int y0 = (int)fy, y1 = y0 + (y0 < hs - 1 ? 1 : 0);
            const float h1 = fy - y0, h0 = 1.f - h1;
            const float wy = (y0 == i ? h0 : 0.f) + (y1 == i ? h1 : 0.f);
            if (wy == 0.f && y0 != i && y1 != i) continue;
            float row = 0.f;
            for (int X = xlo; X <= xhi; ++X) {
                const float fx = sx * X;
                const int x0 = (int)fx, x1 = x0 + (x0 < ws - 1 ? 1 : 0);
                if (x0 != j && x1 != j) continue;
                const float w1 = fx - x0, w0 = 1.f - w1;
                row += ((x0 == j ? w0 : 0.f) + (x1 == j ? w1 : 0.f)) * __ldg(gp + (long long)Y * Wd + X);
            }
            acc += wy * row;
        }
        float* o = gs + bpix(n, i, j, hs, ws) * gss + c;
        *o = accumulate ? *o + acc : acc;
    }
}

}  // namespace s2d

using namespace s2d;

extern "C" {

int adc_im2col_kxk_pack(const float* src, long long batch_stride, int C, int K, int stride, int pad, void* dst, void* dst_lo, int dst_stride,
                        int dst_channels, int N, int H, int W, adc_stream_t stream) {
    if (!src || !dst || C <= 0 || K <= 0 || K > 11 || (stride != 1 && stride != 2) || pad < 0 || dst_channels % 8 != 0 ||
        dst_channels < K * K * C || dst_channels > 1024 || dst_stride < dst_channels || dst_stride % 8 != 0 || N <= 0 || H <= 0 || W <= 0 ||
        (long long)H * W * C >= (1LL << 31)) {
        adc::set_error("adc_im2col_kxk_pack: bad argument (dst_channels a multiple of 8 in [K*K*C, 1024], stride 1 or 2)");
        return ADC_E_BADARG;
    }
    const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
    const int groups = dst_channels / 8;
    const long long total = (long long)N * Ho * Wo * groups;
    if (total >= (1LL << 31)) {
        adc::set_error("adc_im2col_kxk_pack: more than 2^31 destination groups");
        return ADC_E_UNSUPP;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)N * H * W * C * 4.0 + (double)total * (dst_lo ? 32.0 : 16.0));
    im2col_kxk_kernel<<<grid_for(total), 256, 0, st>>>(src, batch_stride, C, K, stride, pad, reinterpret_cast<bf16*>(dst),
                                                      reinterpret_cast<bf16*>(dst_lo), dst_stride, groups, N, H, W, Ho, Wo);
    return adc::check_cuda(cudaGetLastError(), "im2col_kxk_kernel launch");
}

int adc_col2im_kxk(const float* gcol, int gcol_stride, int C, int K, int stride, int pad, float* gx, int N, int H, int W, adc_stream_t stream) {
    if (!gcol || !gx || C <= 0 || K <= 0 || K > 11 || (stride != 1 && stride != 2) || pad < 0 || gcol_stride < K * K * C || N <= 0 || H <= 0 ||
        W <= 0) {
        adc::set_error("adc_col2im_kxk: bad argument");
        return ADC_E_BADARG;
    }
    const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
    const long long total = (long long)N * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)N * Ho * Wo * K * K * C * 4.0 + (double)total * C * 4.0);
    col2im_kxk_kernel<<<grid_for(total), 256, 0, st>>>(gcol, gcol_stride, C, K, stride, pad, gx, N, H, W, Ho, Wo);
    return adc::check_cuda(cudaGetLastError(), "col2im_kxk_kernel launch");
}

int adc_maxpool3x3s2_fwd(const void* x, const void* x_lo, int x_stride, int x_off, void* out, void* out_lo, int out_stride, int out_off,
                         unsigned char* idx, int N, int C, int H, int W, adc_stream_t stream) {
    if (!x || !out || !idx || C <= 0 || C % 8 != 0 || x_off % 8 != 0 || out_off % 8 != 0 || x_stride % 8 != 0 || out_stride % 8 != 0 ||
        N <= 0 || H <= 0 || W <= 0 || (x_lo == nullptr) != (out_lo == nullptr)) {
        adc::set_error("adc_maxpool3x3s2_fwd: bad argument (C, offsets and strides multiples of 8)");
        return ADC_E_BADARG;
    }
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, C8 = C / 8;
    const long long total = (long long)N * Ho * Wo * C8;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_BN, (double)N * H * W * C * (x_lo ? 4.0 : 2.0) + (double)total * 8 * (x_lo ? 5.0 : 3.0));
    maxpool_fwd_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(x) + x_off, x_lo ? reinterpret_cast<const bf16*>(x_lo) + x_off : nullptr,
                                                       x_stride, reinterpret_cast<bf16*>(out) + out_off,
                                                       out_lo ? reinterpret_cast<bf16*>(out_lo) + out_off : nullptr, out_stride, idx, C8, N, H, W, Ho, Wo);
    return adc::check_cuda(cudaGetLastError(), "maxpool_fwd_kernel launch");
}

int adc_maxpool3x3s2_bwd(const float* g, int g_stride, int g_off, const unsigned char* idx, float* gin, int gin_stride, int gin_off, int N, int C,
                         int H, int W, adc_stream_t stream) {
    if (!g || !idx || !gin || C <= 0 || C % 8 != 0 || g_off % 4 != 0 || gin_off % 4 != 0 || g_stride % 4 != 0 || gin_stride % 4 != 0 || N <= 0 ||
        H <= 0 || W <= 0) {
        adc::set_error("adc_maxpool3x3s2_bwd: bad argument");
        return ADC_E_BADARG;
    }
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, C8 = C / 8;
    const long long total = (long long)N * H * W * C8;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_BN, (double)N * Ho * Wo * C * 5.0 + (double)total * 32.0);
    maxpool_bwd_kernel<<<grid_for(total), 256, 0, st>>>(g + g_off, g_stride, idx, gin + gin_off, gin_stride, C8, N, H, W, Ho, Wo);
    return adc::check_cuda(cudaGetLastError(), "maxpool_bwd_kernel launch");
}

int adc_pixel_shuffle_f32(const float* q, int q_stride, int q_off, float* y, int y_stride, int y_off, int N, int C, int h, int w, int H, int W,
                          int accumulate, adc_stream_t stream) {
    if (!q || !y || C <= 0 || C % 4 != 0 || q_off % 4 != 0 || y_off % 4 != 0 || q_stride % 4 != 0 || y_stride % 4 != 0 || N <= 0 || h <= 0 || w <= 0 ||
        (H != 2 * h && H != 2 * h - 1) || (W != 2 * w && W != 2 * w - 1)) {
        adc::set_error("adc_pixel_shuffle_f32: bad argument (C, offsets and strides multiples of 4; H in {2h-1, 2h}, W in {2w-1, 2w})");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * H * W * (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)total * (accumulate ? 48.0 : 32.0));
    shuffle_f32_kernel<<<grid_for(total), 256, 0, st>>>(q + q_off, q_stride, y + y_off, y_stride, C / 4, N, h, w, H, W, accumulate);
    return adc::check_cuda(cudaGetLastError(), "shuffle_f32_kernel launch");
}

int adc_pixel_unshuffle_pair(const void* d, const void* d_lo, int d_stride, int d_off, void* q, void* q_lo, int q_stride, int q_off, int N, int C,
                             int h, int w, adc_stream_t stream) {
    if (!d || !q || C <= 0 || C % 8 != 0 || d_off % 8 != 0 || q_off % 8 != 0 || d_stride % 8 != 0 || q_stride % 8 != 0 || N <= 0 || h <= 0 ||
        w <= 0 || (d_lo == nullptr) != (q_lo == nullptr)) {
        adc::set_error("adc_pixel_unshuffle_pair: bad argument (C, offsets and strides multiples of 8)");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * 4 * h * w * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)total * (d_lo ? 64.0 : 32.0));
    unshuffle_pair_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(d) + d_off,
                                                          d_lo ? reinterpret_cast<const bf16*>(d_lo) + d_off : nullptr, d_stride,
                                                          reinterpret_cast<bf16*>(q) + q_off, q_lo ? reinterpret_cast<bf16*>(q_lo) + q_off : nullptr,
                                                          q_stride, C / 8, N, h, w);
    return adc::check_cuda(cudaGetLastError(), "unshuffle_pair_kernel launch");
}

int adc_bilinear_ac_fwd(const void* src, const void* src_lo, int src_is_pair, int src_stride, int src_off, int N, int C, int hs, int ws, float* dst,
                        int Hd, int Wd, adc_stream_t stream) {
    if (!src || !dst || C <= 0 || N <= 0 || hs <= 0 || ws <= 0 || Hd <= 0 || Wd <= 0 || src_stride < src_off + C) {
        adc::set_error("adc_bilinear_ac_fwd: bad argument");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * Hd * Wd;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)N * hs * ws * C * 4.0 + (double)total * C * 4.0);
    if (src_is_pair)
        bilinear_fwd_kernel<true><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(src) + src_off,
                                                                  src_lo ? reinterpret_cast<const bf16*>(src_lo) + src_off : nullptr, src_stride, C, N,
                                                                  hs, ws, dst, Hd, Wd);
    else
        bilinear_fwd_kernel<false><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float*>(src) + src_off, nullptr, src_stride, C, N, hs, ws,
                                                                   dst, Hd, Wd);
    return adc::check_cuda(cudaGetLastError(), "bilinear_fwd_kernel launch");
}

int adc_bilinear_ac_bwd(const float* g, int N, int C, int Hd, int Wd, float* gsrc, int gsrc_stride, int gsrc_off, int hs, int ws, int accumulate,
                        adc_stream_t stream) {
    if (!g || !gsrc || C <= 0 || N <= 0 || hs <= 0 || ws <= 0 || Hd <= 0 || Wd <= 0 || gsrc_stride < gsrc_off + C) {
        adc::set_error("adc_bilinear_ac_bwd: bad argument");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * hs * ws * C;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)N * Hd * Wd * C * 4.0 + (double)total * 4.0);
    bilinear_bwd_kernel<<<grid_for(total), 256, 0, st>>>(g, C, N, Hd, Wd, gsrc + gsrc_off, gsrc_stride, hs, ws, accumulate);
    return adc::check_cuda(cudaGetLastError(), "bilinear_bwd_kernel launch");
}

}  // extern "C"

// BatchNorm2d (+ residual add) + ReLU / LeakyReLU(0.2) behind a tensor-core convolution, forward and backward, for the BN nets
// (model_zoo/s2d_u_resnet.py:26-61 conv_bn_relu / convt_bn_relu, torchvision BasicBlock as used by :99-107, the cbr3-c1 head
// confidence_depth_framework.py:299-329) -- and, with the identity affine, the plain bias + ReLU layers of GMSNet (:536-607).
//
// Tensors: Y = raw convolution output, fp32, zero-bordered NHWC [N][H+2][W+2][y_stride]; activations / gradient operands = zero-bordered
// NHWC bf16 (hi, lo) pairs; raw gradients = fp32 zero-bordered NHWC (several consumers accumulate into one buffer).  Every tensor is
// addressed as a channel slice [off, off+C) of a wider pixel stride, so torch.cat is a slice write and its backward a slice read.
//
//   bn_stats      per-channel sum / sum of squares of Y -> coef rows {a = gamma*rstd, b = beta - mean*a, mean, rstd}; running statistics
//                 updated with momentum and the unbiased variance (train) or read (eval: no pass over Y at all)
//   bn_apply      out = act(a*Y + b [+ residual]) as a (hi, lo) pair into the slice of the buffer that consumes it
//   bn_bwd_sums   ga = g * act'(out);  sum ga (= dbeta, or the bias gradient of a plain layer), sum ga*yhat (= dgamma)
//   bn_bwd_apply  dY = a*(ga - mean(ga) - yhat*mean(ga*yhat)) as the (hi, lo) operand of the data- and weight-gradient GEMMs; the
//                 residual branch receives ga (written or accumulated into its fp32 gradient slice)
// Reductions are two-stage and deterministic: fp32 per thread (a few hundred pixels), double across the CTA and across CTAs.
// Algorithmic bytes per pixel and channel: stats 4, apply 4 (+4 residual) + 4, bwd 2 x (4 g + 4 Y + 2 out) + 4 dY (+4/8 residual).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace bnact {

constexpr int kParts = 148 * 4;      // first-stage CTAs (4 per SM)
constexpr int kCMax = 1024;          // channel limit of one call
constexpr float kSlope = 0.2f;

// coef rows (fp32 [6][Cp], Cp = C rounded up to 4): 0 a, 1 b, 2 mean, 3 rstd, 4 k1 = mean(ga), 5 k2 = mean(ga*yhat)

__device__ __forceinline__ long long pix_of(long long r, int H, int W) {
    const int x = (int)(r % W), y = (int)((r / W) % H);
    const long long n = r / ((long long)W * H);
    return (n * (H + 2) + y + 1) * (W + 2) + x + 1;
}

__device__ __forceinline__ void reduce_slots(double* red, float v, bool valid, int C4, int groups, double* out, int out_idx_stride,
                                             int out_base) {
    red[threadIdx.x] = valid ? (double)v : 0.0;
    __syncthreads();
    if ((int)threadIdx.x < C4) {
        double t = 0.0;
        for (int q = 0; q < groups; ++q) t += red[q * C4 + threadIdx.x];
        out[out_base + threadIdx.x * out_idx_stride] = t;
    }
    __syncthreads();
}

__device__ __forceinline__ double col_sum(const double* __restrict__ part, int rows, int width, int col) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int r = 0;
    for (; r + 3 < rows; r += 4) {
        t0 += part[(long long)r * width + col];
        t1 += part[(long long)(r + 1) * width + col];
        t2 += part[(long long)(r + 2) * width + col];
        t3 += part[(long long)(r + 3) * width + col];
    }
    for (; r < rows; ++r) t0 += part[(long long)r * width + col];
    return (t0 + t1) + (t2 + t3);
}

// The same sum by one WARP per column (lane l takes rows l, l + 32, ...; fixed-order shuffle tree): the second stage of the statistics
// used to be one thread per channel walking all 592 partial rows, a 150-load dependent chain that put a ~100-150 us floor under every
// BatchNorm pass of the small deep layers (8x10 ... 29x38 pixels at 128-512 channels: most launches of the ResNet bodies).
__device__ __forceinline__ double col_sum_warp(const double* __restrict__ part, int rows, int width, int col, int lane) {
    double t0 = 0.0, t1 = 0.0;
    int r = lane;
    for (; r + 32 < rows; r += 64) {
        t0 += part[(long long)r * width + col];
        t1 += part[(long long)(r + 32) * width + col];
    }
    if (r < rows) t0 += part[(long long)r * width + col];
    double t = t0 + t1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// First-stage CTAs for `pixels` pixels at C4 four-channel groups: every pixel slot of a CTA should see >= 8 pixels, so small tensors
// produce few partial rows (1 ... kParts).
static int parts_for(long long pixels, int C4);

// part[blk][2][Cp]: sum y, sum y^2 of channels [0,Cp) over this CTA's pixels.  Thread = (pixel slot, 4-channel group).
__global__ void __launch_bounds__(256) stats_kernel(const float* __restrict__ Y, int stride, int N, int C4, int H, int W,
                                                    double* __restrict__ part) {
    __shared__ double red[256];
    const int groups = 256 / C4, cl = threadIdx.x % C4, pg = threadIdx.x / C4;
    const long long pixels = (long long)N * H * W;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (pg < groups)
        for (long long r = (long long)blockIdx.x * groups + pg; r < pixels; r += (long long)gridDim.x * groups) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(Y + pix_of(r, H, W) * stride + cl * 4));
            s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
            s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]); s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
        }
    const int Cp = C4 * 4;
    double* out = part + (long long)blockIdx.x * 2 * Cp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        reduce_slots(red, s1[j], pg < groups, C4, groups, out, 4, j);
        reduce_slots(red, s2[j], pg < groups, C4, groups, out, 4, Cp + j);
    }
}

__global__ void __launch_bounds__(128) finalize_kernel(const double* __restrict__ part, int rows, int Cp, int C, double M,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                                       float eps, int training, float* __restrict__ coef) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;      // one warp per channel
    if (c >= Cp) return;
    float a = 0.f, b = 0.f, mean = 0.f, rstd = 0.f;
    if (c < C) {
        if (!gamma) {                // identity affine: a plain (bias + activation) layer rides on the same kernels
            a = 1.f;
        } else {
            double mu, var;
            if (training) {
                mu = col_sum_warp(part, rows, 2 * Cp, c, lane) / M;
                var = col_sum_warp(part, rows, 2 * Cp, Cp + c, lane) / M - mu * mu;
                if (var < 0.0) var = 0.0;
                if (running_mean && lane == 0) {
                    running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mu);
                    running_var[c] =
                        (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * var * (M / (M > 1.0 ? M - 1.0 : 1.0)));
                }
            } else {
                mu = (double)running_mean[c];
                var = (double)running_var[c];
            }
            mean = (float)mu;
            rstd = (float)(1.0 / sqrt(var + (double)eps));
            a = gamma[c] * rstd;
            b = beta[c] - mean * a;
        }
    }
    if (lane != 0) return;
    coef[c] = a;
    coef[Cp + c] = b;
    coef[2 * Cp + c] = mean;
    coef[3 * Cp + c] = rstd;
}

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* rr = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __low2float(rr[k]); v[2 * k + 1] = __high2float(rr[k]); }
}
__device__ __forceinline__ void store8_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&o)[8]) {
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

// thread = (pixel, 8-channel group): out[c] = act(a[c]*Y[c] + b[c] + R[c]) for the C8 = ceil(C/8) groups (channels past C written as 0)
template <int ACT>
__global__ void __launch_bounds__(256) apply_kernel(const float* __restrict__ Y, int ystride, const float* __restrict__ coef, int Cp,
                                                    const __nv_bfloat16* __restrict__ rh, const __nv_bfloat16* __restrict__ rl, int rstride,
                                                    __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int ostride, int N, int C,
                                                    int H, int W) {
    extern __shared__ float sc[];                 // a[Cp8] | b[Cp8]
    const int C8 = (C + 7) / 8, Cp8 = C8 * 8;
    for (int c = threadIdx.x; c < Cp8; c += blockDim.x) {
        sc[c] = c < C ? coef[c] : 0.f;
        sc[Cp8 + c] = c < C ? coef[Cp + c] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)N * H * W * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % C8);
        const long long p = pix_of(i / C8, H, W);
        float v[8], o[8];
        load8_f32(Y + p * ystride + g * 8, v);
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (rh) {
            load8_bf16(rh + p * rstride + g * 8, r);
            if (rl) {
                float r2[8];
                load8_bf16(rl + p * rstride + g * 8, r2);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] += r2[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            float t = fmaf(sc[c], v[j], sc[Cp8 + c]) + r[j];
            if (ACT == 1) t = fmaxf(t, 0.f);
            else if (ACT == 2) t = t > 0.f ? t : kSlope * t;
            o[j] = c < C ? t : 0.f;
        }
        store8_pair(oh + p * ostride + g * 8, ol ? ol + p * ostride + g * 8 : nullptr, o);
    }
}

// part[blk][2][Cp]: per channel sum ga, sum ga*yhat;  ga = g * act'(out)
template <int ACT>
__global__ void __launch_bounds__(256) bwd_sums_kernel(const float* __restrict__ G, int gstride, const float* __restrict__ Y, int ystride,
                                                       const __nv_bfloat16* __restrict__ O, int ostride, const float* __restrict__ coef,
                                                       int Cp, int N, int C4, int H, int W, double* __restrict__ part) {
    __shared__ double red[256];
    const int groups = 256 / C4, cl = threadIdx.x % C4, pg = threadIdx.x / C4;
    const long long pixels = (long long)N * H * W;
    float mu[4], rs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mu[j] = coef[2 * Cp + cl * 4 + j];
        rs[j] = coef[3 * Cp + cl * 4 + j];
    }
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (pg < groups)
        for (long long r = (long long)blockIdx.x * groups + pg; r < pixels; r += (long long)gridDim.x * groups) {
            const long long p = pix_of(r, H, W);
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(G + p * gstride + cl * 4));
            const float4 y4 = __ldg(reinterpret_cast<const float4*>(Y + p * ystride + cl * 4));
            float g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float y[4] = {y4.x, y4.y, y4.z, y4.w};
            if (ACT != 0) {
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(O + p * ostride + cl * 4));
                const __nv_bfloat162* oo = reinterpret_cast<const __nv_bfloat162*>(&raw);
                const float o[4] = {__low2float(oo[0]), __high2float(oo[0]), __low2float(oo[1]), __high2float(oo[1])};
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] *= o[j] > 0.f ? 1.f : (ACT == 1 ? 0.f : kSlope);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sg[j] += g[j];
                sh[j] = fmaf(g[j], (y[j] - mu[j]) * rs[j], sh[j]);
            }
        }
    double* out = part + (long long)blockIdx.x * 2 * Cp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        reduce_slots(red, sg[j], pg < groups, C4, groups, out, 4, j);
        reduce_slots(red, sh[j], pg < groups, C4, groups, out, 4, Cp + j);
    }
}

__global__ void __launch_bounds__(128) bwd_finalize_kernel(const double* __restrict__ part, int rows, int Cp, int C, double M, int training,
                                                           float* __restrict__ ggamma, float* __restrict__ gbeta, float* __restrict__ coef) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;      // one warp per channel
    if (c >= Cp) return;
    float k1 = 0.f, k2 = 0.f;
    if (c < C) {
        const double g = col_sum_warp(part, rows, 2 * Cp, c, lane), h = col_sum_warp(part, rows, 2 * Cp, Cp + c, lane);
        if (lane != 0) return;
        if (gbeta) gbeta[c] = (float)g;
        if (ggamma) ggamma[c] = (float)h;
        if (training) {
            k1 = (float)(g / M);
            k2 = (float)(h / M);
        }
    }
    if (lane != 0) return;
    coef[4 * Cp + c] = k1;
    coef[5 * Cp + c] = k2;
}

// thread = (pixel, 8-channel group) of the dY slice (all C8*8 channels written, zeros past C); the residual branch's gradient = ga
template <int ACT>
__global__ void __launch_bounds__(256) bwd_apply_kernel(const float* __restrict__ G, int gstride, const float* __restrict__ Y, int ystride,
                                                        const __nv_bfloat16* __restrict__ O, int ostride, const float* __restrict__ coef,
                                                        int Cp, __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl, int dstride,
                                                        float* __restrict__ GR, int grstride, int gr_accum, int N, int C, int H, int W) {
    extern __shared__ float sc[];                 // a | mean | rstd | k1 | k2, Cp8 each
    const int C8 = (C + 7) / 8, Cp8 = C8 * 8;
    for (int c = threadIdx.x; c < Cp8; c += blockDim.x) {
        const bool in = c < C;
        sc[c] = in ? coef[c] : 0.f;
        sc[Cp8 + c] = in ? coef[2 * Cp + c] : 0.f;
        sc[2 * Cp8 + c] = in ? coef[3 * Cp + c] : 0.f;
        sc[3 * Cp8 + c] = in ? coef[4 * Cp + c] : 0.f;
        sc[4 * Cp8 + c] = in ? coef[5 * Cp + c] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)N * H * W * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g8 = (int)(i % C8);
        const long long p = pix_of(i / C8, H, W);
        float g[8], y[8], o[8];
        load8_f32(G + p * gstride + g8 * 8, g);
        load8_f32(Y + p * ystride + g8 * 8, y);
        if (ACT != 0) {
            float ov[8];
            load8_bf16(O + p * ostride + g8 * 8, ov);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= ov[j] > 0.f ? 1.f : (ACT == 1 ? 0.f : kSlope);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g8 * 8 + j;
            const float yh = (y[j] - sc[Cp8 + c]) * sc[2 * Cp8 + c];
            o[j] = c < C ? sc[c] * (g[j] - sc[3 * Cp8 + c] - yh * sc[4 * Cp8 + c]) : 0.f;
        }
        store8_pair(dh + p * dstride + g8 * 8, dl ? dl + p * dstride + g8 * 8 : nullptr, o);
        if (GR) {
            float* q = GR + p * grstride + g8 * 8;
            if (g8 * 8 + 8 <= C) {
                float4 a = make_float4(g[0], g[1], g[2], g[3]), b = make_float4(g[4], g[5], g[6], g[7]);
                if (gr_accum) {
                    const float4 pa = *reinterpret_cast<const float4*>(q), pb = *reinterpret_cast<const float4*>(q + 4);
                    a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w; b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
                }
                *reinterpret_cast<float4*>(q) = a;
                *reinterpret_cast<float4*>(q + 4) = b;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (g8 * 8 + j < C) q[j] = gr_accum ? q[j] + g[j] : g[j];
            }
        }
    }
}

// y[n][0][h][w] = bias + sum_c w[c] * x[c] over a (hi, lo) NHWC slice: the 1x1 head of the BN nets (s2d_u_resnet.py:128 convtf, 128 -> 1)
__global__ void __launch_bounds__(256) dot_head_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int xstride,
                                                       const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                                       int relu, int N, int C, int H, int W, __nv_bfloat16* __restrict__ yh,
                                                       __nv_bfloat16* __restrict__ yl, int ystride) {
    extern __shared__ float sw[];
    const int C8 = (C + 7) / 8;
    for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) sw[c] = c < C ? w[c] : 0.f;
    __syncthreads();
    const float b0 = bias ? bias[0] : 0.f;
    const long long pixels = (long long)N * H * W;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < pixels; r += (long long)gridDim.x * blockDim.x) {
        const long long p = pix_of(r, H, W);
        float s0 = 0.f, s1 = 0.f;
        for (int g = 0; g < C8; ++g) {
            float v[8];
            load8_bf16(xh + p * xstride + g * 8, v);
            if (xl) {
                float v2[8];
                load8_bf16(xl + p * xstride + g * 8, v2);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += v2[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                s0 = fmaf(sw[g * 8 + j], v[j], s0);
                s1 = fmaf(sw[g * 8 + j + 1], v[j + 1], s1);
            }
        }
        const float s = (s0 + s1) + b0;
        y[r] = relu ? fmaxf(s, 0.f) : s;
        if (yh) {        // the raw result also as one (hi, lo) channel of an NHWC tensor: the 129th feature channel of the confidence head
            const __nv_bfloat16 h = __float2bfloat16(s);
            yh[p * ystride] = h;
            if (yl) yl[p * ystride] = __float2bfloat16(s - __bfloat162float(h));
        }
    }
}

// backward of dot_head: GX[p][c] (+)= gy[p] * w[c] (fp32 gradient slice), part[blk][Cp + 4] = sum_p gy[p]*x[p][c] | sum_p gy[p]
__global__ void __launch_bounds__(256) dot_head_bwd_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                                           int xstride, const float* __restrict__ w, const float* __restrict__ gy,
                                                           float* __restrict__ GX, int gxstride, int gx_accum, int N, int C, int C4, int H,
                                                           int W, double* __restrict__ part, const float* __restrict__ gy2, int gy2stride) {
    __shared__ double red[256];
    const int groups = 256 / C4, cl = threadIdx.x % C4, pg = threadIdx.x / C4;
    const long long pixels = (long long)N * H * W;
    float wv[4], sx[4] = {0.f, 0.f, 0.f, 0.f}, sg = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) wv[j] = cl * 4 + j < C ? w[cl * 4 + j] : 0.f;
    if (pg < groups)
        for (long long r = (long long)blockIdx.x * groups + pg; r < pixels; r += (long long)gridDim.x * groups) {
            const long long p = pix_of(r, H, W);
            float g = __ldg(gy + r);
            if (gy2) g += gy2[p * gy2stride];      // + the gradient another consumer left in an NHWC gradient channel
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(xh + p * xstride + cl * 4));
            const __nv_bfloat162* xx = reinterpret_cast<const __nv_bfloat162*>(&raw);
            float x[4] = {__low2float(xx[0]), __high2float(xx[0]), __low2float(xx[1]), __high2float(xx[1])};
            if (xl) {
                const uint2 raw2 = __ldg(reinterpret_cast<const uint2*>(xl + p * xstride + cl * 4));
                const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&raw2);
                x[0] += __low2float(x2[0]); x[1] += __high2float(x2[0]); x[2] += __low2float(x2[1]); x[3] += __high2float(x2[1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sx[j] = fmaf(g, x[j], sx[j]);
            if (cl == 0) sg += g;
            if (GX) {
                float* q = GX + p * gxstride + cl * 4;
                float4 o = make_float4(g * wv[0], g * wv[1], g * wv[2], g * wv[3]);
                if (gx_accum) {
                    const float4 pv = *reinterpret_cast<const float4*>(q);
                    o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
                }
                *reinterpret_cast<float4*>(q) = o;
            }
        }
    const int Cp = C4 * 4;
    double* out = part + (long long)blockIdx.x * (Cp + 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) reduce_slots(red, sx[j], pg < groups, C4, groups, out, 4, j);
    red[threadIdx.x] = (pg < groups && cl == 0) ? (double)sg : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < groups; ++q) t += red[q * C4];
        out[Cp] = t;
    }
}

__global__ void __launch_bounds__(128) dot_head_bwd_finalize_kernel(const double* __restrict__ part, int rows, int Cp, int C,
                                                                    float* __restrict__ gw, float* __restrict__ gb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && gb) gb[0] = (float)col_sum(part, rows, Cp + 4, Cp);
    if (c < C) gw[c] = (float)col_sum(part, rows, Cp + 4, c);
}

// fp32 NCHW planes -> fp32 zero-bordered NHWC gradient slice (write or accumulate): how a gradient that autograd hands over per plane
// (the loss's gradient of the depth, the confidence head's gradient of the features) enters the NHWC backward
__global__ void __launch_bounds__(256) nchw_to_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int dstride, int accum,
                                                               int N, int C, int H, int W) {
    const long long total = (long long)N * C * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
        const long long n = i / ((long long)W * H * C);
        float* q = dst + ((n * (H + 2) + y + 1) * (W + 2) + x + 1) * dstride + c;
        *q = accum ? *q + src[i] : src[i];
    }
}

// im2col of a 3x3 / pad-1 window (stride 1 or 2) over up to 7 fp32 planes gathered from up to three NCHW sources, at the OUTPUT
// resolution: channel (3u+v)*C + c of a 64-channel (hi, lo) NHWC pixel (zeros past 9C).  A thin first layer (C <= 7 input channels:
// the stems 3 -> 48 and 2 -> 16 of s2d_u_resnet.py:91-97) becomes ONE 1x1 / GEMM launch of the row-segment kernel.
struct PackSrc { const float* p[3]; long long bs[3]; int ch[3]; };
template <int C>
__global__ void __launch_bounds__(128) im2col_pack_kernel(const __grid_constant__ PackSrc src, __nv_bfloat16* __restrict__ dst,
                                                          __nv_bfloat16* __restrict__ dst_lo, int dst_stride, int N, int H, int W, int stride) {
    const int H1 = (H - 1) / stride + 1, W1 = (W - 1) / stride + 1;
    const long long HW = (long long)H * W, total = (long long)N * H1 * W1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        const float* pl[C];
        {
            int c = 0;
#pragma unroll
            for (int s = 0; s < 3; ++s)
                for (int k = 0; k < src.ch[s] && c < C; ++k, ++c) pl[c] = src.p[s] + n * src.bs[s] + k * HW;
        }
        float v[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) v[q] = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int y = stride * i + u - 1;
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const int x = stride * j + w - 1;
                const bool ok = y >= 0 && y < H && x >= 0 && x < W;
                const long long o = ok ? (long long)y * W + x : 0;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float f = __ldg(pl[c] + o);
                    v[(u * 3 + w) * C + c] = ok ? f : 0.f;
                }
            }
        }
        const long long pix = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float o8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o8[k] = v[q * 8 + k];
            store8_pair(dst + pix * dst_stride + q * 8, dst_lo ? dst_lo + pix * dst_stride + q * 8 : nullptr, o8);
        }
    }
}

static int parts_for(long long pixels, int C4) {
    const int groups = 256 / C4 > 0 ? 256 / C4 : 1;
    long long p = (pixels + (long long)groups * 8 - 1) / ((long long)groups * 8);
    return (int)(p < 1 ? 1 : (p > kParts ? kParts : p));
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static bool bad_slice(const void* p, int stride, int off, int C, int align) {
    return !p || stride <= 0 || off < 0 || off % align != 0 || stride % align != 0 || off + (C + align - 1) / align * align > stride;
}

}  // namespace bnact

using namespace bnact;

extern "C" {

size_t adc_bn_workspace_bytes(int C) {
    const int Cp = (C + 3) / 4 * 4;
    return (size_t)kParts * (2 * Cp + 4) * sizeof(double);
}

int adc_bn_coef_floats(int C) { return 6 * ((C + 3) / 4 * 4); }

int adc_bn_stats(const float* y, int y_stride, int y_off, int N, int C, int H, int W, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps, int training, float* coef, void* workspace,
                 adc_stream_t stream) {
    if (bad_slice(y, y_stride, y_off, C, 4) || C <= 0 || C > kCMax || N <= 0 || H <= 0 || W <= 0 || !coef || !workspace ||
        ((gamma == nullptr) != (beta == nullptr)) || (gamma && !training && (!running_mean || !running_var)) ||
        ((running_mean == nullptr) != (running_var == nullptr))) {
        adc::set_error("adc_bn_stats: bad argument (C <= 1024, fp32 NHWC slice aligned to 4 channels, eval mode needs the running statistics)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = (C + 3) / 4, Cp = C4 * 4;
    const double M = (double)N * H * W;
    const bool pass = gamma && training;
    const int parts = parts_for((long long)N * H * W, C4);
    adc::Prof prof(st, adc::KIND_BN, pass ? M * C * 4.0 : 0.0);
    int rc = 0;
    if (pass) {
        adc::count_launches(1);
        stats_kernel<<<parts, 256, 0, st>>>(y + y_off, y_stride, N, C4, H, W, reinterpret_cast<double*>(workspace));
        if ((rc = adc::check_cuda(cudaGetLastError(), "bn stats_kernel launch"))) return rc;
    }
    finalize_kernel<<<(Cp + 3) / 4, 128, 0, st>>>(reinterpret_cast<const double*>(workspace), parts, Cp, C, M, gamma, beta, running_mean,
                                                  running_var, momentum, eps, training, coef);
    return adc::check_cuda(cudaGetLastError(), "bn finalize_kernel launch");
}

int adc_bn_apply(const float* y, int y_stride, int y_off, const float* coef, const void* res_hi, const void* res_lo, int res_stride, int res_off,
                 int act, void* out_hi, void* out_lo, int out_stride, int out_off, int N, int C, int H, int W, adc_stream_t stream) {
    if (bad_slice(y, y_stride, y_off, C, 8) || bad_slice(out_hi, out_stride, out_off, C, 8) || (res_hi && bad_slice(res_hi, res_stride, res_off, C, 8)) ||
        !coef || act < 0 || act > 2 || C <= 0 || C > kCMax || N <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_bn_apply: bad argument (slices aligned to 8 channels and padded to a multiple of 8; act 0 none, 1 relu, 2 leaky 0.2)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int Cp = (C + 3) / 4 * 4, C8 = (C + 7) / 8;
    const long long total = (long long)N * H * W * C8;
    adc::Prof prof(st, adc::KIND_BN, (double)N * H * W * C * (4.0 + (res_hi ? (res_lo ? 4.0 : 2.0) : 0.0) + (out_lo ? 4.0 : 2.0)));
    const __nv_bfloat16* rh = res_hi ? reinterpret_cast<const __nv_bfloat16*>(res_hi) + res_off : nullptr;
    const __nv_bfloat16* rl = (res_hi && res_lo) ? reinterpret_cast<const __nv_bfloat16*>(res_lo) + res_off : nullptr;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi) + out_off;
    __nv_bfloat16* ol = out_lo ? reinterpret_cast<__nv_bfloat16*>(out_lo) + out_off : nullptr;
    const size_t sm = 2 * C8 * 8 * sizeof(float);
    const int grid = grid_for(total);
    if (act == 0) apply_kernel<0><<<grid, 256, sm, st>>>(y + y_off, y_stride, coef, Cp, rh, rl, res_stride, oh, ol, out_stride, N, C, H, W);
    else if (act == 1) apply_kernel<1><<<grid, 256, sm, st>>>(y + y_off, y_stride, coef, Cp, rh, rl, res_stride, oh, ol, out_stride, N, C, H, W);
    else apply_kernel<2><<<grid, 256, sm, st>>>(y + y_off, y_stride, coef, Cp, rh, rl, res_stride, oh, ol, out_stride, N, C, H, W);
    return adc::check_cuda(cudaGetLastError(), "bn apply_kernel launch");
}

int adc_bn_bwd(const float* g, int g_stride, int g_off, const float* y, int y_stride, int y_off, const void* out_hi, int out_stride, int out_off,
               float* coef, int act, int training, void* dy_hi, void* dy_lo, int dy_stride, int dy_off, float* gres, int gres_stride,
               int gres_off, int gres_accum, float* ggamma, float* gbeta, void* workspace, int N, int C, int H, int W, adc_stream_t stream) {
    if (bad_slice(g, g_stride, g_off, C, 8) || bad_slice(y, y_stride, y_off, C, 8) || bad_slice(dy_hi, dy_stride, dy_off, C, 8) ||
        (act != 0 && bad_slice(out_hi, out_stride, out_off, C, 8)) || (gres && bad_slice(gres, gres_stride, gres_off, C, 8)) || !coef ||
        !workspace || act < 0 || act > 2 || C <= 0 || C > kCMax || N <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_bn_bwd: bad argument (slices aligned to 8 channels and padded to a multiple of 8)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = (C + 3) / 4, Cp = C4 * 4, C8 = (C + 7) / 8;
    const double M = (double)N * H * W;
    adc::Prof prof(st, adc::KIND_BN, M * C * (2.0 * (8.0 + (act ? 2.0 : 0.0)) + (dy_lo ? 4.0 : 2.0) + (gres ? (gres_accum ? 8.0 : 4.0) : 0.0)));
    adc::count_launches(2);
    const float* G = g + g_off;
    const float* Y = y + y_off;
    const __nv_bfloat16* O = act ? reinterpret_cast<const __nv_bfloat16*>(out_hi) + out_off : nullptr;
    double* part = reinterpret_cast<double*>(workspace);
    const int parts = parts_for((long long)N * H * W, C4);
    if (act == 0) bwd_sums_kernel<0><<<parts, 256, 0, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, N, C4, H, W, part);
    else if (act == 1) bwd_sums_kernel<1><<<parts, 256, 0, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, N, C4, H, W, part);
    else bwd_sums_kernel<2><<<parts, 256, 0, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, N, C4, H, W, part);
    int rc = adc::check_cuda(cudaGetLastError(), "bn bwd_sums_kernel launch");
    if (rc) return rc;
    bwd_finalize_kernel<<<(Cp + 3) / 4, 128, 0, st>>>(part, parts, Cp, C, M, training, ggamma, gbeta, coef);
    if ((rc = adc::check_cuda(cudaGetLastError(), "bn bwd_finalize_kernel launch"))) return rc;
    __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(dy_hi) + dy_off;
    __nv_bfloat16* dl = dy_lo ? reinterpret_cast<__nv_bfloat16*>(dy_lo) + dy_off : nullptr;
    float* GR = gres ? gres + gres_off : nullptr;
    const size_t sm = 5 * C8 * 8 * sizeof(float);
    const int grid = grid_for((long long)N * H * W * C8);
    if (act == 0)
        bwd_apply_kernel<0><<<grid, 256, sm, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, dh, dl, dy_stride, GR, gres_stride, gres_accum, N, C, H, W);
    else if (act == 1)
        bwd_apply_kernel<1><<<grid, 256, sm, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, dh, dl, dy_stride, GR, gres_stride, gres_accum, N, C, H, W);
    else
        bwd_apply_kernel<2><<<grid, 256, sm, st>>>(G, g_stride, Y, y_stride, O, out_stride, coef, Cp, dh, dl, dy_stride, GR, gres_stride, gres_accum, N, C, H, W);
    return adc::check_cuda(cudaGetLastError(), "bn bwd_apply_kernel launch");
}

int adc_dot_head_fwd(const void* x_hi, const void* x_lo, int x_stride, int x_off, const float* w, const float* bias, float* y, int relu, int N,
                     int C, int H, int W, void* y_hi, void* y_lo, int y_stride, int y_off, adc_stream_t stream) {
    if (bad_slice(x_hi, x_stride, x_off, C, 8) || !w || !y || C <= 0 || C > kCMax || N <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_dot_head_fwd: bad argument");
        return ADC_E_BADARG;
    }
    const long long pixels = (long long)N * H * W;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_BN, (double)pixels * (C * (x_lo ? 4.0 : 2.0) + 4.0));
    const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(x_hi) + x_off;
    const __nv_bfloat16* xl = x_lo ? reinterpret_cast<const __nv_bfloat16*>(x_lo) + x_off : nullptr;
    __nv_bfloat16* yh = y_hi ? reinterpret_cast<__nv_bfloat16*>(y_hi) + y_off : nullptr;
    __nv_bfloat16* yl = (y_hi && y_lo) ? reinterpret_cast<__nv_bfloat16*>(y_lo) + y_off : nullptr;
    dot_head_kernel<<<grid_for(pixels), 256, (C + 7) / 8 * 8 * sizeof(float), (cudaStream_t)stream>>>(xh, xl, x_stride, w, bias, y, relu, N, C, H, W,
                                                                                                     yh, yl, y_stride);
    return adc::check_cuda(cudaGetLastError(), "dot_head_kernel launch");
}

int adc_dot_head_bwd(const void* x_hi, const void* x_lo, int x_stride, int x_off, const float* w, const float* gy, float* gx, int gx_stride,
                     int gx_off, int gx_accum, float* gw, float* gb, void* workspace, int N, int C, int H, int W, const float* gy_extra,
                     int gy_extra_stride, adc_stream_t stream) {
    if (bad_slice(x_hi, x_stride, x_off, C, 4) || (gx && bad_slice(gx, gx_stride, gx_off, C, 4)) || !w || !gy || !gw || !workspace || C <= 0 ||
        C > kCMax || N <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_dot_head_bwd: bad argument");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = (C + 3) / 4, Cp = C4 * 4;
    const long long pixels = (long long)N * H * W;
    adc::Prof prof(st, adc::KIND_BN, (double)pixels * (C * ((x_lo ? 4.0 : 2.0) + (gx ? (gx_accum ? 8.0 : 4.0) : 0.0)) + 4.0));
    adc::count_launches(1);
    const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(x_hi) + x_off;
    const __nv_bfloat16* xl = x_lo ? reinterpret_cast<const __nv_bfloat16*>(x_lo) + x_off : nullptr;
    double* part = reinterpret_cast<double*>(workspace);
    dot_head_bwd_kernel<<<kParts, 256, 0, st>>>(xh, xl, x_stride, w, gy, gx ? gx + gx_off : nullptr, gx_stride, gx_accum, N, C, C4, H, W, part,
                                                gy_extra, gy_extra_stride);
    int rc = adc::check_cuda(cudaGetLastError(), "dot_head_bwd_kernel launch");
    if (rc) return rc;
    dot_head_bwd_finalize_kernel<<<(Cp + 127) / 128, 128, 0, st>>>(part, kParts, Cp, C, gw, gb);
    return adc::check_cuda(cudaGetLastError(), "dot_head_bwd_finalize_kernel launch");
}

int adc_im2col3x3_pack(const float* s0, int c0, long long bs0, const float* s1, int c1, long long bs1, const float* s2, int c2, long long bs2,
                       void* dst, void* dst_lo, int dst_stride, int N, int H, int W, int stride, adc_stream_t stream) {
    const int C = c0 + (s1 ? c1 : 0) + (s2 ? c2 : 0);
    if (!s0 || !dst || c0 <= 0 || C < 1 || C > 7 || dst_stride < 64 || dst_stride % 8 != 0 || N <= 0 || H <= 0 || W <= 0 || (stride != 1 && stride != 2)) {
        adc::set_error("adc_im2col3x3_pack: bad argument (1..7 input planes in total, destination slice of 64 channels)");
        return ADC_E_BADARG;
    }
    PackSrc src;
    src.p[0] = s0; src.ch[0] = c0; src.bs[0] = bs0;
    src.p[1] = s1; src.ch[1] = s1 ? c1 : 0; src.bs[1] = bs1;
    src.p[2] = s2; src.ch[2] = s2 ? c2 : 0; src.bs[2] = bs2;
    const int H1 = (H - 1) / stride + 1, W1 = (W - 1) / stride + 1;
    const long long total = (long long)N * H1 * W1;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)total * (36.0 * C + (dst_lo ? 256.0 : 128.0)));
    long long b = (total + 127) / 128;
    const int grid = (int)(b < 148LL * 32 ? b : 148LL * 32);
    __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(dst);
    __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(dst_lo);
    switch (C) {
        case 1: im2col_pack_kernel<1><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        case 2: im2col_pack_kernel<2><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        case 3: im2col_pack_kernel<3><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        case 4: im2col_pack_kernel<4><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        case 5: im2col_pack_kernel<5><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        case 6: im2col_pack_kernel<6><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
        default: im2col_pack_kernel<7><<<grid, 128, 0, st>>>(src, dh, dl, dst_stride, N, H, W, stride); break;
    }
    return adc::check_cuda(cudaGetLastError(), "im2col_pack_kernel launch");
}

int adc_nchw_to_nhwc_f32(const float* src, float* dst, int dst_stride, int dst_off, int accumulate, int N, int C, int H, int W,
                         adc_stream_t stream) {
    if (!src || !dst || dst_stride < dst_off + C || dst_off < 0 || N <= 0 || C <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_nchw_to_nhwc_f32: bad argument");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * C * H * W;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_BN, (double)total * (accumulate ? 12.0 : 8.0));
    nchw_to_nhwc_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(src, dst + dst_off, dst_stride, accumulate, N, C, H, W);
    return adc::check_cuda(cudaGetLastError(), "nchw_to_nhwc_f32_kernel launch");
}

}  // extern "C"

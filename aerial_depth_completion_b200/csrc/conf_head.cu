// cbr3-c1 confidence head (confidence_depth_framework.py:299-329): conv3x3 (no bias) -> BatchNorm2d -> LeakyReLU(0.2) ->
// conv1x1 (+bias) -> sigmoid.  The 3x3 convolution, its data gradient and its weight gradient run on the tcgen05 kernels
// (conv_tc.cu / conv_tc_bwd.cu); this file is everything behind it, fused into four HBM-bound passes over the convolution
// output Y (fp32, zero-bordered NHWC [N][H+2][W+2][stride]):
//   conf_bn_stats     per-channel sum / sum of squares of Y -> (mean, 1/sqrt(var+eps)), the folded affine a = gamma*rstd,
//                     b = beta - mean*a, and the running-statistics update of train mode (momentum, unbiased variance)
//   conf_apply        conf = sigmoid(b2 + sum_c w2[c] * lrelu(a[c]*Y[c] + b[c])): BN + LeakyReLU + 1x1 + sigmoid in ONE read of Y
//   conf_bwd_sums     the four reductions of the backward (dW2, db2, dgamma, dbeta) from Y, conf and the incoming gradient
//   conf_bwd_apply    dY = a*(ga - mean(ga) - yhat*mean(ga*yhat)) as a zero-bordered NHWC bf16 (hi, lo) tensor: the operand of
//                     the tensor-core data- and weight-gradient kernels of the 3x3 layer
// Reductions are two-stage and deterministic: fp32 per thread (<= ~100 pixels), double across the CTA and across CTAs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace confhead {

constexpr int kParts = 148 * 4;      // first-stage CTAs (4 per SM)
constexpr int kCMax = 128;           // channel limit (coef rows are kCMax wide)
constexpr float kSlope = 0.2f;

// coef rows (fp32 [6][kCMax]): 0 a = gamma*rstd, 1 b = beta - mean*a, 2 mean, 3 rstd, 4 k1 = mean(ga), 5 k2 = mean(ga*yhat)

__device__ __forceinline__ long long pix_of(long long r, int H, int W) {
    const int x = (int)(r % W), y = (int)((r / W) % H);
    const long long n = r / ((long long)W * H);
    return (n * (H + 2) + y + 1) * (W + 2) + x + 1;
}

// CTA-wide sum of one value per thread over the pixel slots of each 4-channel group: out[cl] = sum_pg v(pg, cl), fixed order.
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

// part[blk][2][Cp]: sum y, sum y^2 of channels [0,Cp) over this CTA's pixels
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ Y, int stride, int N, int C4, int H, int W,
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

// fixed-order sum of rows [0, rows) of column `col` of a [rows][width] double matrix (4 interleaved accumulators)
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

__global__ void __launch_bounds__(kCMax) bn_finalize_kernel(const double* __restrict__ part, int rows, int Cp, int C, double M,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ running_mean, float* __restrict__ running_var,
                                                            float momentum, float eps, int training, float* __restrict__ coef) {
    const int c = threadIdx.x;
    if (c >= Cp) return;
    float a = 0.f, b = 0.f, mean = 0.f, rstd = 0.f;
    if (c < C) {
        double mu, var;
        if (training) {
            mu = col_sum(part, rows, 2 * Cp, c) / M;
            var = col_sum(part, rows, 2 * Cp, Cp + c) / M - mu * mu;
            if (var < 0.0) var = 0.0;
            if (running_mean) {
                running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mu);
                running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * var * (M / (M > 1.0 ? M - 1.0 : 1.0)));
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
    coef[c] = a;
    coef[kCMax + c] = b;
    coef[2 * kCMax + c] = mean;
    coef[3 * kCMax + c] = rstd;
}

// one thread per pixel: conf[n][0][y][x] = sigmoid(b2 + sum_c w2[c] * lrelu(a[c]*Y[c] + b[c]))
__global__ void __launch_bounds__(256) conf_apply_kernel(const float* __restrict__ Y, int stride, const float* __restrict__ coef,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ conf, int N, int C, int C4, int H, int W) {
    __shared__ float sa[kCMax], sb[kCMax], sw[kCMax];
    for (int c = threadIdx.x; c < C4 * 4; c += blockDim.x) {
        sa[c] = coef[c];
        sb[c] = coef[kCMax + c];
        sw[c] = c < C ? w2[c] : 0.f;
    }
    __syncthreads();
    const float bias = b2[0];
    const long long pixels = (long long)N * H * W;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < pixels; r += (long long)gridDim.x * blockDim.x) {
        const float4* yp = reinterpret_cast<const float4*>(Y + pix_of(r, H, W) * stride);
        float s0 = 0.f, s1 = 0.f;
        for (int g = 0; g < C4; ++g) {
            const float4 v = __ldg(yp + g);
            float t;
            t = fmaf(sa[4 * g], v.x, sb[4 * g]);         s0 = fmaf(sw[4 * g], t > 0.f ? t : kSlope * t, s0);
            t = fmaf(sa[4 * g + 1], v.y, sb[4 * g + 1]); s1 = fmaf(sw[4 * g + 1], t > 0.f ? t : kSlope * t, s1);
            t = fmaf(sa[4 * g + 2], v.z, sb[4 * g + 2]); s0 = fmaf(sw[4 * g + 2], t > 0.f ? t : kSlope * t, s0);
            t = fmaf(sa[4 * g + 3], v.w, sb[4 * g + 3]); s1 = fmaf(sw[4 * g + 3], t > 0.f ? t : kSlope * t, s1);
        }
        const float s = (s0 + s1) + bias;
        conf[r] = 1.f / (1.f + expf(-s));
    }
}

// part[blk][3*Cp + 4]: per channel sum gs*z (dW2), sum ga (dbeta), sum ga*yhat (dgamma); [3*Cp] = sum gs (db2)
// gs = gconf*conf*(1-conf); t = a*y+b; z = lrelu(t); ga = gs*w2*lrelu'(t); yhat = (y-mean)*rstd
__global__ void __launch_bounds__(256) conf_bwd_sums_kernel(const float* __restrict__ Y, int stride, const float* __restrict__ coef,
                                                            const float* __restrict__ w2, const float* __restrict__ gconf,
                                                            const float* __restrict__ conf, int N, int C, int C4, int H, int W,
                                                            double* __restrict__ part) {
    __shared__ double red[256];
    const int groups = 256 / C4, cl = threadIdx.x % C4, pg = threadIdx.x / C4;
    const long long pixels = (long long)N * H * W;
    float a[4], b[4], mu[4], rs[4], w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = cl * 4 + j;
        a[j] = coef[c]; b[j] = coef[kCMax + c]; mu[j] = coef[2 * kCMax + c]; rs[j] = coef[3 * kCMax + c];
        w[j] = c < C ? w2[c] : 0.f;
    }
    float sz[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, ss = 0.f;
    if (pg < groups)
        for (long long r = (long long)blockIdx.x * groups + pg; r < pixels; r += (long long)gridDim.x * groups) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(Y + pix_of(r, H, W) * stride + cl * 4));
            const float cf = __ldg(conf + r);
            const float gs = __ldg(gconf + r) * cf * (1.f - cf);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t = fmaf(a[j], v[j], b[j]);
                const float z = t > 0.f ? t : kSlope * t;
                const float ga = gs * w[j] * (t > 0.f ? 1.f : kSlope);
                sz[j] = fmaf(gs, z, sz[j]);
                sg[j] += ga;
                sh[j] = fmaf(ga, (v[j] - mu[j]) * rs[j], sh[j]);
            }
            if (cl == 0) ss += gs;
        }
    const int Cp = C4 * 4;
    double* out = part + (long long)blockIdx.x * (3 * Cp + 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        reduce_slots(red, sz[j], pg < groups, C4, groups, out, 4, j);
        reduce_slots(red, sg[j], pg < groups, C4, groups, out, 4, Cp + j);
        reduce_slots(red, sh[j], pg < groups, C4, groups, out, 4, 2 * Cp + j);
    }
    red[threadIdx.x] = (pg < groups && cl == 0) ? (double)ss : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < groups; ++q) t += red[q * C4];
        out[3 * Cp] = t;
    }
}

__global__ void __launch_bounds__(kCMax) conf_bwd_finalize_kernel(const double* __restrict__ part, int rows, int Cp, int C, double M,
                                                                  float* __restrict__ gw2, float* __restrict__ gb2,
                                                                  float* __restrict__ ggamma, float* __restrict__ gbeta,
                                                                  float* __restrict__ coef) {
    const int c = threadIdx.x, width = 3 * Cp + 4;
    if (c == 0) gb2[0] = (float)col_sum(part, rows, width, 3 * Cp);
    if (c >= Cp) return;
    float k1 = 0.f, k2 = 0.f;
    if (c < C) {
        const double z = col_sum(part, rows, width, c), g = col_sum(part, rows, width, Cp + c), h = col_sum(part, rows, width, 2 * Cp + c);
        gw2[c] = (float)z;
        gbeta[c] = (float)g;
        ggamma[c] = (float)h;
        k1 = (float)(g / M);
        k2 = (float)(h / M);
    }
    coef[4 * kCMax + c] = k1;
    coef[5 * kCMax + c] = k2;
}

// thread = (pixel, 8-channel group) of the OUTPUT tensor (all gy_stride channels are written, zeros past C)
__global__ void __launch_bounds__(256) conf_bwd_apply_kernel(const float* __restrict__ Y, int stride, const float* __restrict__ coef,
                                                             const float* __restrict__ w2, const float* __restrict__ gconf,
                                                             const float* __restrict__ conf, int training, __nv_bfloat16* __restrict__ gh,
                                                             __nv_bfloat16* __restrict__ gl, int gstride, int N, int C, int H, int W) {
    __shared__ float sa[kCMax], sb[kCMax], sm[kCMax], sr[kCMax], sw[kCMax], s1[kCMax], s2[kCMax];
    for (int c = threadIdx.x; c < kCMax; c += blockDim.x) {
        sa[c] = coef[c]; sb[c] = coef[kCMax + c]; sm[c] = coef[2 * kCMax + c]; sr[c] = coef[3 * kCMax + c];
        s1[c] = training ? coef[4 * kCMax + c] : 0.f;
        s2[c] = training ? coef[5 * kCMax + c] : 0.f;
        sw[c] = c < C ? w2[c] : 0.f;
    }
    __syncthreads();
    const int G8 = gstride / 8;
    const long long total = (long long)N * H * W * G8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G8);
        const long long r = i / G8;
        const long long p = pix_of(r, H, W);
        uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
        if (g * 8 < C) {
            const float cf = __ldg(conf + r);
            const float gs = __ldg(gconf + r) * cf * (1.f - cf);
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(Y + p * stride + g * 8));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(Y + p * stride + g * 8 + 4));
            const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = g * 8 + j;
                const float t = fmaf(sa[c], v[j], sb[c]);
                const float ga = gs * sw[c] * (t > 0.f ? 1.f : kSlope);
                const float yh = (v[j] - sm[c]) * sr[c];
                o[j] = c < C ? sa[c] * (ga - s1[c] - yh * s2[c]) : 0.f;
            }
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
            __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hh[j] = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
                ll[j] = __floats2bfloat162_rn(o[2 * j] - __low2float(hh[j]), o[2 * j + 1] - __high2float(hh[j]));
            }
        }
        *reinterpret_cast<uint4*>(gh + p * gstride + g * 8) = h;
        if (gl) *reinterpret_cast<uint4*>(gl + p * gstride + g * 8) = l;
    }
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static bool bad_common(const void* y, int stride, int N, int C, int H, int W) {
    const int Cp = (C + 3) / 4 * 4;
    return !y || C <= 0 || C > kCMax || stride % 4 != 0 || (C + 7) / 8 * 8 > stride || Cp > stride || N <= 0 || H <= 0 || W <= 0;
}

}  // namespace confhead

using namespace confhead;

extern "C" {

size_t adc_conf_head_workspace_bytes(int C) {
    const int Cp = (C + 3) / 4 * 4;
    return (size_t)kParts * (3 * Cp + 4) * sizeof(double);
}

int adc_conf_head_coef_floats(void) { return 6 * kCMax; }

int adc_conf_bn_stats(const float* y, int y_stride, int N, int C, int H, int W, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, int training, float* coef, void* workspace,
                      adc_stream_t stream) {
    if (bad_common(y, y_stride, N, C, H, W) || !gamma || !beta || !coef || !workspace || (!training && (!running_mean || !running_var)) ||
        ((running_mean == nullptr) != (running_var == nullptr))) {
        adc::set_error("adc_conf_bn_stats: bad argument (C <= 128, fp32 NHWC with stride %% 4 == 0, eval mode needs the running statistics)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = (C + 3) / 4;
    const double M = (double)N * H * W;
    adc::Prof prof(st, adc::KIND_CONF_HEAD, training ? M * C * 4.0 : 0.0);
    int rc = 0;
    if (training) {
        adc::count_launches(1);
        bn_stats_kernel<<<kParts, 256, 0, st>>>(y, y_stride, N, C4, H, W, reinterpret_cast<double*>(workspace));
        rc = adc::check_cuda(cudaGetLastError(), "bn_stats_kernel launch");
        if (rc) return rc;
    }
    bn_finalize_kernel<<<1, kCMax, 0, st>>>(reinterpret_cast<const double*>(workspace), kParts, C4 * 4, C, M, gamma, beta, running_mean,
                                            running_var, momentum, eps, training, coef);
    return adc::check_cuda(cudaGetLastError(), "bn_finalize_kernel launch");
}

int adc_conf_head_apply(const float* y, int y_stride, const float* coef, const float* w2, const float* b2, float* conf, int N, int C, int H,
                        int W, adc_stream_t stream) {
    if (bad_common(y, y_stride, N, C, H, W) || !coef || !w2 || !b2 || !conf) {
        adc::set_error("adc_conf_head_apply: bad argument");
        return ADC_E_BADARG;
    }
    const long long pixels = (long long)N * H * W;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_CONF_HEAD, (double)pixels * (C + 1) * 4.0);
    conf_apply_kernel<<<grid_for(pixels), 256, 0, (cudaStream_t)stream>>>(y, y_stride, coef, w2, b2, conf, N, C, (C + 3) / 4, H, W);
    return adc::check_cuda(cudaGetLastError(), "conf_apply_kernel launch");
}

int adc_conf_head_bwd(const float* y, int y_stride, float* coef, const float* w2, const float* gconf, const float* conf, int training,
                      void* gy, void* gy_lo, int gy_stride, float* gw2, float* gb2, float* ggamma, float* gbeta, void* workspace, int N, int C,
                      int H, int W, adc_stream_t stream) {
    if (bad_common(y, y_stride, N, C, H, W) || !coef || !w2 || !gconf || !conf || !gy || gy_stride % 8 != 0 || gy_stride < C || gy_stride > y_stride ||
        !gw2 || !gb2 || !ggamma || !gbeta || !workspace) {
        adc::set_error("adc_conf_head_bwd: bad argument (gy_stride %% 8 == 0, C <= gy_stride <= y_stride)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = (C + 3) / 4;
    const long long pixels = (long long)N * H * W;
    adc::Prof prof(st, adc::KIND_CONF_HEAD, (double)pixels * (2.0 * C * 4.0 + 16.0 + gy_stride * (gy_lo ? 4.0 : 2.0)));
    adc::count_launches(2);
    conf_bwd_sums_kernel<<<kParts, 256, 0, st>>>(y, y_stride, coef, w2, gconf, conf, N, C, C4, H, W, reinterpret_cast<double*>(workspace));
    int rc = adc::check_cuda(cudaGetLastError(), "conf_bwd_sums_kernel launch");
    if (rc) return rc;
    conf_bwd_finalize_kernel<<<1, kCMax, 0, st>>>(reinterpret_cast<const double*>(workspace), kParts, C4 * 4, C, (double)pixels, gw2, gb2, ggamma,
                                                  gbeta, coef);
    rc = adc::check_cuda(cudaGetLastError(), "conf_bwd_finalize_kernel launch");
    if (rc) return rc;
    conf_bwd_apply_kernel<<<grid_for(pixels * (gy_stride / 8)), 256, 0, st>>>(y, y_stride, coef, w2, gconf, conf, training, (__nv_bfloat16*)gy,
                                                                             (__nv_bfloat16*)gy_lo, gy_stride, N, C, H, W);
    return adc::check_cuda(cudaGetLastError(), "conf_bwd_apply_kernel launch");
}

}  // extern "C"

// Thin last layer of GEDNet (confidence_depth_framework.py:504-506): conv10 = 3x3 convolution of cat(up(x9) [32 ch], x0_d [1 ch])
// to 1 (or 2) channels + LeakyReLU(0.2), forward and backward, as direct CUDA-core kernels.
//
// Why not the tensor-core kernels: with one output channel the layer is pure bandwidth.  Routed through the implicit-GEMM
// path it needed the up-sampled x9 materialised at full resolution in a 64-channel padded NHWC pair buffer (577 MB at batch
// 32) that was then read by the forward, the data-gradient and the weight-gradient kernels: ~2.8 ms of a 9.2 ms training step
// (profiles/r1_ged_tc_step.txt).  Here the nearest up-sampling is folded into the addressing (x9 is read at its own
// resolution through the ATen nearest map), x0_d and the output stay NCHW fp32, and the backward produces directly the
// pre-activation gradient G9 = leaky'(x9) * sum_children(dgrad) that the tensor-core kernels of conv9 consume.
//   x9 : zero-bordered NHWC bf16 (hi, lo pair or hi only) [N][H1+2][W1+2][stride9], channels [0,32)
//   W  : fp32 [oc][33][3][3] (channel 32 = x0_d), bias fp32 [oc], oc in {1, 2}
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace head {

constexpr int kC = 32;      // channels of x9
constexpr int kMaxOc = 2;

struct Near {               // ATen 'nearest' source index (see adc_common.cuh NearestMap)
    int in, out, mode;
    float scale;
    __host__ static Near make(int in, int out) {
        Near m;
        m.in = in; m.out = out;
        m.mode = (in == out) ? 0 : (out == 2 * in ? 1 : 2);
        m.scale = (float)in / (float)out;
        return m;
    }
    __device__ __forceinline__ int operator()(int dst) const {
        if (mode == 0) return dst;
        if (mode == 1) return dst >> 1;
        const int s = (int)floorf((float)dst * scale);
        return s < in - 1 ? s : in - 1;
    }
    // [lo, hi) = the dst indices whose source is s (contiguous range)
    __device__ __forceinline__ void children(int s, int& lo, int& hi) const {
        if (mode == 0) { lo = s; hi = s + 1; return; }
        if (mode == 1) { lo = 2 * s; hi = 2 * s + 2; return; }
        int g = (int)(((long long)s * out) / in) - 1;
        if (g < 0) g = 0;
        while (g < out && (*this)(g) < s) ++g;
        lo = g;
        while (g < out && (*this)(g) == s) ++g;
        hi = g;
    }
};

struct X9 {
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;    // may be null
    int stride, H1, W1;
};

__device__ __forceinline__ void load32(const X9& x, long long pix, float (&v)[kC]) {
    const uint4* ph = reinterpret_cast<const uint4*>(x.hi + pix * x.stride);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 h = ph[q];
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[q * 8 + 2 * j] = __low2float(hh[j]);
            v[q * 8 + 2 * j + 1] = __high2float(hh[j]);
        }
    }
    if (x.lo) {
        const uint4* pl = reinterpret_cast<const uint4*>(x.lo + pix * x.stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 l = pl[q];
            const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[q * 8 + 2 * j] += __low2float(ll[j]);
                v[q * 8 + 2 * j + 1] += __high2float(ll[j]);
            }
        }
    }
}

// shared weight image: ws[tap][c][o] (c = 0..32), 9*33*oc floats
__device__ __forceinline__ void stage_weights(float* ws, const float* __restrict__ w, int oc) {
    for (int i = threadIdx.x; i < 9 * 33 * oc; i += blockDim.x) {
        const int o = i % oc, c = (i / oc) % 33, tap = i / (oc * 33);
        ws[i] = w[((long long)o * 33 + c) * 9 + tap];
    }
}

template <int OC>
__global__ void __launch_bounds__(128) ged_head_fwd_kernel(X9 x9, const float* __restrict__ x0d, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                                                           Near mh, Near mw) {
    __shared__ float ws[9 * 33 * OC];
    stage_weights(ws, w, OC);
    __syncthreads();
    const long long total = (long long)N * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
        float acc[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[o] = bias[o];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int yy = y + u - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const int xx = x + v - 1;
                if (xx < 0 || xx >= W) continue;
                const float* wt = ws + (u * 3 + v) * 33 * OC;
                float xv[kC];
                load32(x9, ((long long)n * (x9.H1 + 2) + mh(yy) + 1) * (x9.W1 + 2) + mw(xx) + 1, xv);
#pragma unroll
                for (int c = 0; c < kC; ++c)
#pragma unroll
                    for (int o = 0; o < OC; ++o) acc[o] = fmaf(wt[c * OC + o], xv[c], acc[o]);
                const float d = x0d[((long long)n * H + yy) * W + xx];
#pragma unroll
                for (int o = 0; o < OC; ++o) acc[o] = fmaf(wt[32 * OC + o], d, acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            const float r = acc[o];
            out[(((long long)n * OC + o) * H + y) * W + x] = r > 0.f ? r : 0.2f * r;
        }
    }
}

// pre-activation gradient of conv10 at an in-bounds output pixel
template <int OC>
__device__ __forceinline__ void gpre_at(const float* __restrict__ gy, const float* __restrict__ x10, int n, int y, int x, int H, int W,
                                        float (&g)[OC]) {
#pragma unroll
    for (int o = 0; o < OC; ++o) {
        const long long p = (((long long)n * OC + o) * H + y) * W + x;
        g[o] = gy[p] * (x10[p] > 0.f ? 1.f : 0.2f);
    }
}

// G9[n][i][j][c] = leaky'(x9) * ( sum over children (y',x') of (i,j), taps (u,v), o of W[o][c][u][v] * gpre[o][y'-u+1][x'-v+1]
//                                 + sum over children of glli[n][c][y'][x'] )
template <int OC>
__global__ void __launch_bounds__(128) ged_head_bwd_data_kernel(X9 x9, const float* __restrict__ gy, const float* __restrict__ x10,
                                                                const float* __restrict__ glli, int glli_ch, const float* __restrict__ w,
                                                                __nv_bfloat16* __restrict__ g9, __nv_bfloat16* __restrict__ g9_lo,
                                                                int g9_stride, int N, int H, int W, Near mh, Near mw) {
    __shared__ float ws[9 * 33 * OC];
    stage_weights(ws, w, OC);
    __syncthreads();
    const int H1 = x9.H1, W1 = x9.W1;
    const long long total = (long long)N * H1 * W1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        int y0, y1, x0, x1;
        mh.children(i, y0, y1);
        mw.children(j, x0, x1);
        // S[tap][o] = sum over children of gpre shifted by the tap
        float S[9][OC];
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int o = 0; o < OC; ++o) S[tp][o] = 0.f;
        for (int yc = y0; yc < y1; ++yc)
            for (int xc = x0; xc < x1; ++xc) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int yy = yc - u + 1;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        const int xx = xc - v + 1;
                        if (xx < 0 || xx >= W) continue;
                        float g[OC];
                        gpre_at<OC>(gy, x10, n, yy, xx, H, W, g);
#pragma unroll
                        for (int o = 0; o < OC; ++o) S[u * 3 + v][o] += g[o];
                    }
                }
            }
        float acc[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) acc[c] = 0.f;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
            const float* wt = ws + tp * 33 * OC;
#pragma unroll
            for (int c = 0; c < kC; ++c)
#pragma unroll
                for (int o = 0; o < OC; ++o) acc[c] = fmaf(wt[c * OC + o], S[tp][o], acc[c]);
        }
        if (glli) {
            for (int yc = y0; yc < y1; ++yc)
                for (int xc = x0; xc < x1; ++xc)
#pragma unroll
                    for (int c = 0; c < kC; ++c) acc[c] += glli[(((long long)n * glli_ch + c) * H + yc) * W + xc];
        }
        const long long pix = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1;
        const uint4* ph = reinterpret_cast<const uint4*>(x9.hi + pix * x9.stride);
        uint4* oh = reinterpret_cast<uint4*>(g9 + pix * g9_stride);
        uint4* ol = g9_lo ? reinterpret_cast<uint4*>(g9_lo + pix * g9_stride) : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 a = ph[q];
            const __nv_bfloat162* aa = reinterpret_cast<const __nv_bfloat162*>(&a);
            uint4 h, l;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
            __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v0 = acc[q * 8 + 2 * k] * (__low2float(aa[k]) > 0.f ? 1.f : 0.2f);
                const float v1 = acc[q * 8 + 2 * k + 1] * (__high2float(aa[k]) > 0.f ? 1.f : 0.2f);
                hh[k] = __floats2bfloat162_rn(v0, v1);
                ll[k] = __floats2bfloat162_rn(v0 - __low2float(hh[k]), v1 - __high2float(hh[k]));
            }
            oh[q] = h;
            if (ol) ol[q] = l;
        }
    }
}

// gradient of the x0_d channel (needed when the net is the loss net and x0_d is a prediction):
// gx0d[n][y][x] = sum_{u,v,o} W[o][32][u][v] * gpre[o][y-u+1][x-v+1]  (+ glli[n][32][y][x])
template <int OC>
__global__ void __launch_bounds__(256) ged_head_bwd_x0d_kernel(const float* __restrict__ gy, const float* __restrict__ x10,
                                                               const float* __restrict__ glli, int glli_ch, const float* __restrict__ w,
                                                               float* __restrict__ gx0d, int N, int H, int W) {
    __shared__ float wd[9 * OC];
    if (threadIdx.x < 9 * OC) wd[threadIdx.x] = w[((long long)(threadIdx.x % OC) * 33 + 32) * 9 + threadIdx.x / OC];
    __syncthreads();
    const long long total = (long long)N * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
        float acc = glli ? glli[(((long long)n * glli_ch + 32) * H + y) * W + x] : 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int yy = y - u + 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const int xx = x - v + 1;
                if (xx < 0 || xx >= W) continue;
                float g[OC];
                gpre_at<OC>(gy, x10, n, yy, xx, H, W, g);
#pragma unroll
                for (int o = 0; o < OC; ++o) acc = fmaf(wd[(u * 3 + v) * OC + o], g[o], acc);
            }
        }
        gx0d[i] = acc;
    }
}

// Weight and bias gradient.  Warp = a strided set of x9 pixels (i,j); lane = channel c of x9.
//   dW[o][c][u][v]  = sum_{(i,j)} x9[i][j][c] * S_{uv,o}(i,j),   S = sum over children (y',x') of gpre[o][y'-u+1][x'-v+1]
//   dW[o][32][u][v] = sum_{(y,x)} gpre[o][y][x] * x0d[y+u-1][x+v-1]      (lanes 0..8 = tap, over the children of (i,j))
//   db[o]           = sum gpre[o]                                          (lane 9)
// Per-CTA partial rows [OC][9*33 + 1], reduced by ged_head_wgrad_reduce_kernel in a fixed order.
template <int OC>
__global__ void __launch_bounds__(256) ged_head_wgrad_kernel(X9 x9, const float* __restrict__ x0d, const float* __restrict__ gy,
                                                             const float* __restrict__ x10, float* __restrict__ part, int N, int H, int W,
                                                             Near mh, Near mw) {
    constexpr int ROW = 9 * 33 + 1;
    __shared__ float red[8][OC][ROW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H1 = x9.H1, W1 = x9.W1;
    float acc[9][OC];       // lane = c: dW[o][c][tap]
    float accd[OC];         // lanes 0..8: dW[o][32][tap = lane]; lane 9: db[o]
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[tp][o] = 0.f;
#pragma unroll
    for (int o = 0; o < OC; ++o) accd[o] = 0.f;
    const long long total = (long long)N * H1 * W1;
    const int u = lane / 3, v = lane % 3;          // for lanes 0..8
    for (long long t = (long long)blockIdx.x * 8 + warp; t < total; t += (long long)gridDim.x * 8) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        int y0, y1, x0, x1;
        mh.children(i, y0, y1);
        mw.children(j, x0, x1);
        float S[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) S[o] = 0.f;
        if (lane < 10) {
            for (int yc = y0; yc < y1; ++yc)
                for (int xc = x0; xc < x1; ++xc) {
                    if (lane < 9) {
                        // S for dW of x9: gpre at child - tap + 1
                        const int yy = yc - u + 1, xx = xc - v + 1;
                        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                            float g[OC];
                            gpre_at<OC>(gy, x10, n, yy, xx, H, W, g);
#pragma unroll
                            for (int o = 0; o < OC; ++o) S[o] += g[o];
                        }
                        // x0_d channel: gpre at the child times x0d at child + tap - 1
                        const int ys = yc + u - 1, xs = xc + v - 1;
                        if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
                            float g[OC];
                            gpre_at<OC>(gy, x10, n, yc, xc, H, W, g);
                            const float d = x0d[((long long)n * H + ys) * W + xs];
#pragma unroll
                            for (int o = 0; o < OC; ++o) accd[o] = fmaf(g[o], d, accd[o]);
                        }
                    } else {
                        float g[OC];
                        gpre_at<OC>(gy, x10, n, yc, xc, H, W, g);
#pragma unroll
                        for (int o = 0; o < OC; ++o) accd[o] += g[o];
                    }
                }
        }
        const long long pix = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1;
        float xv = __bfloat162float(x9.hi[pix * x9.stride + lane]);
        if (x9.lo) xv += __bfloat162float(x9.lo[pix * x9.stride + lane]);
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int o = 0; o < OC; ++o) acc[tp][o] = fmaf(xv, __shfl_sync(0xffffffffu, S[o], tp), acc[tp][o]);
    }
#pragma unroll
    for (int o = 0; o < OC; ++o) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) red[warp][o][lane * 9 + tp] = acc[tp][o];
        if (lane < 9) red[warp][o][32 * 9 + lane] = accd[o];
        if (lane == 9) red[warp][o][ROW - 1] = accd[o];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < OC * ROW; k += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][k / ROW][k % ROW];
        part[(long long)blockIdx.x * OC * ROW + k] = s;
    }
}

__global__ void ged_head_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int oc, float* __restrict__ dw, float* __restrict__ db) {
    const int ROW = 9 * 33 + 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= oc * ROW) return;
    float s = 0.f;
    for (int q = 0; q < nparts; ++q) s += part[(long long)q * oc * ROW + k];
    const int o = k / ROW, r = k % ROW;
    if (r == ROW - 1) db[o] = s;
    else dw[(long long)o * 33 * 9 + r] = s;          // r = c*9 + tap = the Parameter layout [o][c][u][v]
}

static int grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

constexpr int kWgradCtas = 148 * 4;

}  // namespace head

using namespace head;

extern "C" {

int adc_ged_head_fwd(const void* x9, const void* x9_lo, int x9_stride, int H1, int W1, const float* x0d, const float* w, const float* bias,
                     float* out, int N, int oc, int H, int W, adc_stream_t stream) {
    if (!x9 || !x0d || !w || !bias || !out || N <= 0 || H <= 0 || W <= 0 || H1 <= 0 || W1 <= 0 || H1 > H || W1 > W || oc < 1 || oc > kMaxOc ||
        x9_stride < kC || x9_stride % 8 != 0) {
        adc::set_error("adc_ged_head_fwd: bad argument");
        return ADC_E_BADARG;
    }
    X9 x{(const __nv_bfloat16*)x9, (const __nv_bfloat16*)x9_lo, x9_stride, H1, W1};
    const long long total = (long long)N * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)total * (4.0 + 4.0 * oc) + (double)N * H1 * W1 * kC * (x9_lo ? 4.0 : 2.0));
    if (oc == 1) ged_head_fwd_kernel<1><<<grid_for(total, 128), 128, 0, st>>>(x, x0d, w, bias, out, N, H, W, Near::make(H1, H), Near::make(W1, W));
    else ged_head_fwd_kernel<2><<<grid_for(total, 128), 128, 0, st>>>(x, x0d, w, bias, out, N, H, W, Near::make(H1, H), Near::make(W1, W));
    return adc::check_cuda(cudaGetLastError(), "ged_head_fwd_kernel launch");
}

size_t adc_ged_head_bwd_workspace_bytes(int oc) { return (size_t)kWgradCtas * oc * (9 * 33 + 1) * sizeof(float); }

// gy: gradient of the (post-LeakyReLU) output x10, both NCHW fp32 [N][oc][H][W]; glli: optional gradient of the feature tensor
// cat(up(x9), x0_d) (NCHW fp32, glli_ch >= 33 channels per image); g9/g9_lo: G9 = pre-activation gradient of conv9's output,
// zero-bordered NHWC bf16 (pair), channels [0,32); gx0d: optional [N][1][H][W]; dw [oc][33][3][3], db [oc].
int adc_ged_head_bwd(const void* x9, const void* x9_lo, int x9_stride, int H1, int W1, const float* x0d, const float* w, const float* gy,
                     const float* x10, const float* glli, int glli_ch, void* g9, void* g9_lo, int g9_stride, float* gx0d, float* dw,
                     float* db, void* workspace, int N, int oc, int H, int W, adc_stream_t stream) {
    if (!x9 || !x0d || !w || !gy || !x10 || !g9 || !dw || !db || !workspace || N <= 0 || H <= 0 || W <= 0 || H1 <= 0 || W1 <= 0 || H1 > H ||
        W1 > W || oc < 1 || oc > kMaxOc || x9_stride < kC || x9_stride % 8 != 0 || g9_stride < kC || g9_stride % 8 != 0 ||
        (glli && glli_ch < 33) || ((x9_lo == nullptr) != (g9_lo == nullptr))) {
        adc::set_error("adc_ged_head_bwd: bad argument");
        return ADC_E_BADARG;
    }
    X9 x{(const __nv_bfloat16*)x9, (const __nv_bfloat16*)x9_lo, x9_stride, H1, W1};
    const Near mh = Near::make(H1, H), mw = Near::make(W1, W);
    cudaStream_t st = (cudaStream_t)stream;
    const long long t1 = (long long)N * H1 * W1, t0 = (long long)N * H * W;
    float* part = reinterpret_cast<float*>(workspace);
    int rc;
    {
        adc::Prof prof(st, adc::KIND_HEAD10, (double)t0 * 8.0 * oc + (double)t1 * kC * (x9_lo ? 8.0 : 4.0));
        if (oc == 1) ged_head_bwd_data_kernel<1><<<grid_for(t1, 128), 128, 0, st>>>(x, gy, x10, glli, glli_ch, w, (__nv_bfloat16*)g9, (__nv_bfloat16*)g9_lo, g9_stride, N, H, W, mh, mw);
        else ged_head_bwd_data_kernel<2><<<grid_for(t1, 128), 128, 0, st>>>(x, gy, x10, glli, glli_ch, w, (__nv_bfloat16*)g9, (__nv_bfloat16*)g9_lo, g9_stride, N, H, W, mh, mw);
        if ((rc = adc::check_cuda(cudaGetLastError(), "ged_head_bwd_data_kernel launch"))) return rc;
    }
    if (gx0d) {
        adc::Prof prof(st, adc::KIND_HEAD10, (double)t0 * (4.0 + 8.0 * oc));
        if (oc == 1) ged_head_bwd_x0d_kernel<1><<<grid_for(t0, 256), 256, 0, st>>>(gy, x10, glli, glli_ch, w, gx0d, N, H, W);
        else ged_head_bwd_x0d_kernel<2><<<grid_for(t0, 256), 256, 0, st>>>(gy, x10, glli, glli_ch, w, gx0d, N, H, W);
        if ((rc = adc::check_cuda(cudaGetLastError(), "ged_head_bwd_x0d_kernel launch"))) return rc;
    }
    {
        adc::Prof prof(st, adc::KIND_HEAD10, (double)t0 * (4.0 + 8.0 * oc) + (double)t1 * kC * (x9_lo ? 4.0 : 2.0));
        adc::count_launches(1);
        if (oc == 1) ged_head_wgrad_kernel<1><<<kWgradCtas, 256, 0, st>>>(x, x0d, gy, x10, part, N, H, W, mh, mw);
        else ged_head_wgrad_kernel<2><<<kWgradCtas, 256, 0, st>>>(x, x0d, gy, x10, part, N, H, W, mh, mw);
        if ((rc = adc::check_cuda(cudaGetLastError(), "ged_head_wgrad_kernel launch"))) return rc;
        const int tot = oc * (9 * 33 + 1);
        ged_head_wgrad_reduce_kernel<<<(tot + 127) / 128, 128, 0, st>>>(part, kWgradCtas, oc, dw, db);
        if ((rc = adc::check_cuda(cudaGetLastError(), "ged_head_wgrad_reduce_kernel launch"))) return rc;
    }
    return 0;
}

// gradient of the x0_d channel alone (the quad path computes everything else on the tensor cores): gx0d [N][1][H][W]
int adc_ged_head_bwd_x0d(const float* gy, const float* x10, const float* glli, int glli_ch, const float* w, float* gx0d, int N, int oc,
                         int H, int W, adc_stream_t stream) {
    if (!gy || !x10 || !w || !gx0d || N <= 0 || H <= 0 || W <= 0 || oc < 1 || oc > kMaxOc || (glli && glli_ch < 33)) {
        adc::set_error("adc_ged_head_bwd_x0d: bad argument");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long t0 = (long long)N * H * W;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)t0 * (4.0 + 8.0 * oc));
    if (oc == 1) ged_head_bwd_x0d_kernel<1><<<grid_for(t0, 256), 256, 0, st>>>(gy, x10, glli, glli_ch, w, gx0d, N, H, W);
    else ged_head_bwd_x0d_kernel<2><<<grid_for(t0, 256), 256, 0, st>>>(gy, x10, glli, glli_ch, w, gx0d, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "ged_head_bwd_x0d_kernel launch");
}

}  // extern "C"

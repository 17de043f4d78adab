// "Quad" (pixel-shuffle) forms of the two full-resolution thin layers of GEDNet (confidence_depth_framework.py:442,452,
// 469-506) and their glue.
//
// conv10 reads cat(up(x9), x0_d) at full resolution, where up() is the exact 2x nearest up-sampling of x9.  The 3x3 taps of an
// output pixel (2i+py, 2j+px) then fall on only 2x2 distinct pixels of x9, so
//     conv10(up(x9))[2i+py][2j+px] = sum_{a,b in 3x3} Wq[(py,px)][c][a][b] * x9[i+a-1][j+b-1][c]
// with Wq = the taps of W that land on the same x9 pixel added together: a stride-1 3x3 convolution of x9 AT ITS OWN RESOLUTION
// with 4 output channels per original output channel (one per parity).  That convolution (and its data / weight gradients)
// runs on the tcgen05 kernels on a quarter of the pixels; what is left at full resolution is elementwise and lives here:
//   * quad_head_finish:   un-shuffle the 4 parity channels + the x0_d channel's own 3x3 taps + bias + LeakyReLU(0.2) -> NCHW fp32
//   * quad_head_bwd_prep: gpre = gy * leaky'(x10) shuffled into the 4-parity NHWC bf16 (hi, lo) gradient tensor the tensor-core
//                         kernels consume; x0_d-channel weight gradient and bias gradient partial sums
// The same identity read backwards gives the data gradient of conv1 (3x3, stride 2): gin[2i+py][2j+px][c] is a stride-1 3x3
// convolution of G1 (conv1's output gradient, low resolution) with 4 parity output channels per input channel:
//   * quad_unshuffle:     [N][H1+2][W1+2][4*K] fp32 -> K planar [N][H][W] fp32 gradients
// and the entry of the dense path:
//   * ged_pack_input:     (xout_d, rgb, cout_d) NCHW fp32 -> zero-bordered NHWC bf16 (hi, lo) [.., 16] = torch.cat + layout in one pass
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace quad {

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.2f * v; }

// 4x4 neighbourhood of a planar fp32 image around the 2x2 block (2i, 2j): rows 2i-1..2i+2, columns 2j-1..2j+2, zeros outside
__device__ __forceinline__ void load_nb4(const float* __restrict__ img, int H, int W, int i, int j, float (&d)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = 2 * i - 1 + r;
        const bool yok = y >= 0 && y < H;
        const float* row = img + (long long)(yok ? y : 0) * W;
        const int x0 = 2 * j - 1;
        d[r][0] = (yok && x0 >= 0) ? __ldg(row + x0) : 0.f;
        if (yok && ((W & 1) == 0)) {        // 2j, 2j+1 are both in range and 8-byte aligned when W is even
            const float2 m = __ldg(reinterpret_cast<const float2*>(row + x0 + 1));
            d[r][1] = m.x; d[r][2] = m.y;
        } else {
            d[r][1] = (yok && x0 + 1 < W) ? __ldg(row + x0 + 1) : 0.f;
            d[r][2] = (yok && x0 + 2 < W) ? __ldg(row + x0 + 2) : 0.f;
        }
        d[r][3] = (yok && x0 + 3 < W) ? __ldg(row + x0 + 3) : 0.f;
    }
}

// out[n][o][2i+py][2j+px] = leaky( q[n][i+1][j+1][(py*2+px)*OC + o] + bias[o] + sum_{u,v} w[o][32][u][v] * x0d[n][2i+py+u-1][2j+px+v-1] )
template <int OC>
__global__ void __launch_bounds__(256) quad_head_finish_kernel(const float* __restrict__ q, int q_stride, const float* __restrict__ x0d,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               float* __restrict__ out, int N, int H, int W) {
    const int H1 = H >> 1, W1 = W >> 1;
    float wd[OC][9], b[OC];
#pragma unroll
    for (int o = 0; o < OC; ++o) {
        b[o] = __ldg(bias + o);
#pragma unroll
        for (int t = 0; t < 9; ++t) wd[o][t] = __ldg(w + ((long long)o * 33 + 32) * 9 + t);
    }
    const long long total = (long long)N * H1 * W1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        float d[4][4];
        load_nb4(x0d + (long long)n * H * W, H, W, i, j, d);
        const float* qp = q + (((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1) * q_stride;
        float qv[4 * OC];
#pragma unroll
        for (int k = 0; k < OC; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(qp) + k);
            qv[4 * k] = v.x; qv[4 * k + 1] = v.y; qv[4 * k + 2] = v.z; qv[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            float r[2][2];
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    float acc = qv[(py * 2 + px) * OC + o] + b[o];
#pragma unroll
                    for (int u = 0; u < 3; ++u)
#pragma unroll
                        for (int v = 0; v < 3; ++v) acc = fmaf(wd[o][u * 3 + v], d[py + u][px + v], acc);
                    r[py][px] = leaky(acc);
                }
            float* op = out + (((long long)n * OC + o) * H + 2 * i) * W + 2 * j;
            *reinterpret_cast<float2*>(op) = make_float2(r[0][0], r[0][1]);
            *reinterpret_cast<float2*>(op + W) = make_float2(r[1][0], r[1][1]);
        }
    }
}

// gq[n][i+1][j+1][(py*2+px)*OC + o] = gpre[n][o][2i+py][2j+px] = gy * (x10 > 0 ? 1 : 0.2)   (fp32, 4*OC channels per pixel)
// part[block][o][t]  t < 9: sum gpre[o][y][x] * x0d[y+u-1][x+v-1] (weight gradient of the x0_d channel);  t = 9: sum gpre[o]
template <int OC>
__global__ void __launch_bounds__(256) quad_head_bwd_prep_kernel(const float* __restrict__ gy, const float* __restrict__ x10,
                                                                 const float* __restrict__ x0d, float* __restrict__ gq, float* __restrict__ part,
                                                                 int N, int H, int W) {
    __shared__ float red[8][OC * 10];
    const int H1 = H >> 1, W1 = W >> 1;
    float acc[OC][10];
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
        for (int t = 0; t < 10; ++t) acc[o][t] = 0.f;
    const long long total = (long long)N * H1 * W1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        float d[4][4];
        load_nb4(x0d + (long long)n * H * W, H, W, i, j, d);
        float g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = 0.f;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            const long long p = (((long long)n * OC + o) * H + 2 * i) * W + 2 * j;
            const float2 g0 = __ldg(reinterpret_cast<const float2*>(gy + p)), g1 = __ldg(reinterpret_cast<const float2*>(gy + p + W));
            const float2 y0 = __ldg(reinterpret_cast<const float2*>(x10 + p)), y1 = __ldg(reinterpret_cast<const float2*>(x10 + p + W));
            float gp[2][2];
            gp[0][0] = g0.x * (y0.x > 0.f ? 1.f : 0.2f);
            gp[0][1] = g0.y * (y0.y > 0.f ? 1.f : 0.2f);
            gp[1][0] = g1.x * (y1.x > 0.f ? 1.f : 0.2f);
            gp[1][1] = g1.y * (y1.y > 0.f ? 1.f : 0.2f);
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    g[(py * 2 + px) * OC + o] = gp[py][px];
                    acc[o][9] += gp[py][px];
#pragma unroll
                    for (int u = 0; u < 3; ++u)
#pragma unroll
                        for (int v = 0; v < 3; ++v) acc[o][u * 3 + v] = fmaf(gp[py][px], d[py + u][px + v], acc[o][u * 3 + v]);
                }
        }
        const long long pix = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1;
        float4* o4 = reinterpret_cast<float4*>(gq + pix * (4 * OC));
#pragma unroll
        for (int k = 0; k < OC; ++k) o4[k] = make_float4(g[4 * k], g[4 * k + 1], g[4 * k + 2], g[4 * k + 3]);
    }
    // deterministic block reduction: lanes (xor shuffles), then the 8 warps in index order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
        for (int t = 0; t < 10; ++t) {
            float v = acc[o][t];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if (lane == 0) red[warp][o * 10 + t] = v;
        }
    __syncthreads();
    if (threadIdx.x < OC * 10) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        part[(long long)blockIdx.x * OC * 10 + threadIdx.x] = s;
    }
}

// ---- backward of the quad conv10 on the CUDA cores ----------------------------------------------------------------------------
// With 4*OC <= 8 "output channels" the two gradient contractions are far too thin for the tensor pipe (an M = 128 MMA costs ~90
// cycles whatever its N: measured 298 us / 133 us as tcgen05 GEMMs) -- as direct kernels they are 36*OC FMAs per (pixel, channel):
// warp = a run of 32 pixels of one row of x9's grid, lane = one of x9's 32 channels, a 3x3 register window slides along the row.
struct X9v {
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;    // may be null
    int stride;
};
__device__ __forceinline__ float ld_x9(const X9v& x, long long pix, int c) {
    float v = __bfloat162float(x.hi[pix * x.stride + c]);
    if (x.lo) v += __bfloat162float(x.lo[pix * x.stride + c]);
    return v;
}

// Both kernels work on tiles of 8 rows x 32 pixels of x9's grid (warp = tile row, lane = channel) staged through shared memory with
// 16-byte loads issued back to back.  The first version read every operand with per-lane 2-byte loads inside the pixel loop: 64-byte
// requests with one dependent load per iteration, DRAM-latency bound at 0.4-0.75 TB/s (206 / 182 us per launch at B=32, 228x304).
constexpr int kQT_R = 8, kQT_C = 32, kQT_XR = kQT_R + 2, kQT_XC = kQT_C + 2;

__device__ __forceinline__ void bf16x8_to_f32(const uint4& h, float (&v)[8]) {
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[2 * k] = __low2float(hh[k]);
        v[2 * k + 1] = __high2float(hh[k]);
    }
}

// dWq[qo][c][a*3+b] = sum_p gq[p][qo] * x9[p + (a-1, b-1)][c]   -> part[block][4*OC][32][9]
// shared memory: sx[10][34][32] fp32 (hi + lo of x9's 32 real channels, halo included) | sg[8][32][Q] (gq of the tile's pixels)
template <int OC>
__global__ void __launch_bounds__(256, OC == 1 ? 2 : 1) quad_head_wgrad_kernel(X9v x9, const float* __restrict__ gq, float* __restrict__ part, int N, int H1,
                                                              int W1) {
    constexpr int Q = 4 * OC;
    extern __shared__ __align__(16) float qsm[];
    float* sx = qsm;
    float* sg = qsm + kQT_XR * kQT_XC * 32;
    __shared__ float red[8][9][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[Q][9];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[q][t] = 0.f;
    const int tx = (W1 + kQT_C - 1) / kQT_C, ty = (H1 + kQT_R - 1) / kQT_R;
    const int ntiles = N * tx * ty;
    constexpr int XITEMS = kQT_XR * kQT_XC * 4, XPER = (XITEMS + 255) / 256;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j0 = (tile % tx) * kQT_C, i0 = ((tile / tx) % ty) * kQT_R, n = tile / (tx * ty);
        __syncthreads();                                   // the previous tile has been consumed
        uint4 vh[XPER], vl[XPER];
#pragma unroll
        for (int k = 0; k < XPER; ++k) {                   // all loads first (memory-level parallelism), conversion afterwards
            const int e = threadIdx.x + k * 256;
            const int r = e / (kQT_XC * 4), col = (e >> 2) % kQT_XC, ch = (e & 3) * 8;
            const bool ok = e < XITEMS && i0 + r <= H1 + 1 && j0 + col <= W1 + 1;
            const long long off = (((long long)n * (H1 + 2) + i0 + r) * (W1 + 2) + j0 + col) * x9.stride + ch;
            vh[k] = ok ? __ldg(reinterpret_cast<const uint4*>(x9.hi + off)) : make_uint4(0, 0, 0, 0);
            vl[k] = (ok && x9.lo) ? __ldg(reinterpret_cast<const uint4*>(x9.lo + off)) : make_uint4(0, 0, 0, 0);
        }
        for (int e = threadIdx.x; e < kQT_R * kQT_C * OC; e += 256) {
            const int k = e % OC, col = (e / OC) % kQT_C, r = e / (OC * kQT_C);
            const bool ok = i0 + r < H1 && j0 + col < W1;  // pixels outside the image contribute nothing
            const long long pix = ((long long)n * (H1 + 2) + i0 + r + 1) * (W1 + 2) + j0 + col + 1;
            reinterpret_cast<float4*>(sg)[e] = ok ? __ldg(reinterpret_cast<const float4*>(gq + pix * Q) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < XPER; ++k) {
            const int e = threadIdx.x + k * 256;
            if (e < XITEMS) {
                float a[8], b[8];
                bf16x8_to_f32(vh[k], a);
                bf16x8_to_f32(vl[k], b);
                float4* d = reinterpret_cast<float4*>(sx + (e >> 2) * 32 + (e & 3) * 8);
                d[0] = make_float4(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]);
                d[1] = make_float4(a[4] + b[4], a[5] + b[5], a[6] + b[6], a[7] + b[7]);
            }
        }
        __syncthreads();
        const float* xr = sx + (warp * kQT_XC) * 32 + lane;     // x9 row (tile row warp - 1 + a), column j + b: xr[(a*XC + j + b)*32]
        const float* gr = sg + (warp * kQT_C) * Q;
        float xw[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            xw[a][0] = xr[(a * kQT_XC) * 32];
            xw[a][1] = xr[(a * kQT_XC + 1) * 32];
        }
#pragma unroll 4
        for (int j = 0; j < kQT_C; ++j) {
#pragma unroll
            for (int a = 0; a < 3; ++a) xw[a][2] = xr[(a * kQT_XC + j + 2) * 32];
            float g[Q];
#pragma unroll
            for (int k = 0; k < OC; ++k) {
                const float4 t = reinterpret_cast<const float4*>(gr + j * Q)[k];
                g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) acc[q][a * 3 + b] = fmaf(g[q], xw[a][b], acc[q][a * 3 + b]);
#pragma unroll
            for (int a = 0; a < 3; ++a) { xw[a][0] = xw[a][1]; xw[a][1] = xw[a][2]; }
        }
    }
    // the 8 warps of the CTA in index order, one parity channel at a time
    for (int q = 0; q < Q; ++q) {
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < Q; ++k) v = (k == q) ? acc[k][t] : v;      // keeps acc in registers (no dynamic indexing)
            red[warp][t][lane] = v;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 9 * 32; e += 256) {
            const int t = e / 32, c = e % 32;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += red[k][t][c];
            part[(((long long)blockIdx.x * Q + q) * 32 + c) * 9 + t] = sum;
        }
    }
}

// G9[p][c] = leaky'(x9[p][c]) * ( sum_{qo,a,b} Wq[qo][c][a][b] * gq[p - (a-1, b-1)][qo]  +  extra[p][c] )     (bf16 hi, lo)
// part[block][32]: per-channel sums of G9 (the bias gradient of conv9).  Wq: fp32 [4*OC][32][3][3].
// shared memory: sg[10][34][Q] = gq at padded rows i0.., columns j0.. (the tile's pixels and their 3x3 neighbourhood)
template <int OC>
__global__ void __launch_bounds__(256, OC == 1 ? 2 : 1) quad_head_dgrad_kernel(X9v x9, const float* __restrict__ gq, const float* __restrict__ wq, X9v extra,
                                                              __nv_bfloat16* __restrict__ g9, __nv_bfloat16* __restrict__ g9_lo, int g9_stride,
                                                              float* __restrict__ part, int N, int H1, int W1) {
    constexpr int Q = 4 * OC;
    __shared__ __align__(16) float sg[kQT_XR * kQT_XC * Q];
    __shared__ __align__(16) __nv_bfloat16 sxh[kQT_R * kQT_C * 32];     // x9 (hi halves: only the sign is needed) of the tile's pixels
    __shared__ float red[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float w[Q][9];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) w[q][t] = __ldg(wq + ((long long)q * 32 + lane) * 9 + t);
    float csum = 0.f;
    const int tx = (W1 + kQT_C - 1) / kQT_C, ty = (H1 + kQT_R - 1) / kQT_R;
    const int ntiles = N * tx * ty;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j0 = (tile % tx) * kQT_C, i0 = ((tile / tx) % ty) * kQT_R, n = tile / (tx * ty);
        __syncthreads();
        for (int e = threadIdx.x; e < kQT_XR * kQT_XC * OC; e += 256) {
            const int k = e % OC, col = (e / OC) % kQT_XC, r = e / (OC * kQT_XC);
            const bool ok = i0 + r <= H1 + 1 && j0 + col <= W1 + 1;
            const long long pix = ((long long)n * (H1 + 2) + i0 + r) * (W1 + 2) + j0 + col;
            reinterpret_cast<float4*>(sg)[e] = ok ? __ldg(reinterpret_cast<const float4*>(gq + pix * Q) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {   // x9 tile: 8 x 32 pixels x 4 chunks of 8 channels = 1024 16-byte loads, 4 per thread, all in flight together
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = threadIdx.x + k * 256, r = e / (kQT_C * 4), col = (e >> 2) % kQT_C, ch = (e & 3) * 8;
                const bool ok = i0 + r < H1 && j0 + col < W1;
                const long long pix = ((long long)n * (H1 + 2) + i0 + r + 1) * (W1 + 2) + j0 + col + 1;
                v[k] = ok ? __ldg(reinterpret_cast<const uint4*>(x9.hi + pix * x9.stride + ch)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(sxh)[threadIdx.x + k * 256] = v[k];
        }
        __syncthreads();
        const int i = i0 + warp;
        if (i >= H1) continue;                             // (the barriers above are reached by every warp: `continue` re-enters the loop)
        const long long rowpix = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j0 + 1;
        // window gw[r][col][q] = gq at padded (i + r, j + col): output pixel (i, j) = padded (i+1, j+1) reads padded (i + 2 - a, j + 2 - b)
        float gw[3][3][Q];
        auto load_col = [&](int col, int jj) {             // jj = tile column of the window's left edge
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < OC; ++k) {
                    const float4 t = reinterpret_cast<const float4*>(sg + (((warp + r) * kQT_XC) + jj + col) * Q)[k];
                    gw[r][col][4 * k] = t.x; gw[r][col][4 * k + 1] = t.y; gw[r][col][4 * k + 2] = t.z; gw[r][col][4 * k + 3] = t.w;
                }
        };
        load_col(0, 0);
        load_col(1, 0);
        // running pointers: one 64-bit add per pixel and tensor (a fresh pix * stride + lane per access cost ~50 integer instructions
        // per pixel, and the kernel is issue-bound)
        const bool has_lo = g9_lo != nullptr, has_ex = extra.hi != nullptr, has_exlo = has_ex && extra.lo != nullptr;
        const long long goff = rowpix * g9_stride + lane, eoff = rowpix * extra.stride + lane;
        __nv_bfloat16* hp = g9 + goff;
        __nv_bfloat16* lp = (has_lo ? g9_lo : g9) + goff;
        const __nv_bfloat16* ehp = (has_ex ? extra.hi : g9) + (has_ex ? eoff : 0);
        const __nv_bfloat16* elp = (has_exlo ? extra.lo : g9) + (has_exlo ? eoff : 0);
        constexpr int GRP = 8;
        for (int jb = 0; jb < kQT_C; jb += GRP) {
            // the group's global operands are requested before the first pixel is computed
            float xv[GRP], ex[GRP];
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const bool ok = j0 + jb + u < W1;
                xv[u] = __bfloat162float(sxh[(warp * kQT_C + jb + u) * 32 + lane]);
                ex[u] = 0.f;
                if (has_ex) {
                    if (ok) ex[u] = __bfloat162float(*ehp) + (has_exlo ? __bfloat162float(*elp) : 0.f);
                    ehp += extra.stride;
                    elp += extra.stride;
                }
            }
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int j = jb + u;
                load_col(2, j);
                float pa[3] = {ex[u], 0.f, 0.f};           // three independent chains (one per kernel row) instead of 36 dependent FMAs
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) pa[a] = fmaf(w[q][a * 3 + b], gw[2 - a][2 - b][q], pa[a]);
                float acc = (pa[0] + pa[1]) + pa[2];
                acc *= xv[u] > 0.f ? 1.f : 0.2f;
                if (j0 + j < W1) {
                    csum += acc;
                    const __nv_bfloat16 h = __float2bfloat16(acc);
                    *hp = h;
                    if (has_lo) *lp = __float2bfloat16(acc - __bfloat162float(h));
                }
                hp += g9_stride;
                lp += g9_stride;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < Q; ++q) { gw[r][0][q] = gw[r][1][q]; gw[r][1][q] = gw[r][2][q]; }
            }
        }
    }
    red[warp][lane] = csum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][threadIdx.x];
        part[(long long)blockIdx.x * 32 + threadIdx.x] = sum;
    }
}

// out[k][n][2i+py][2j+px] = q[n][i+1][j+1][(py*2+px)*K + k]     (H1 = ceil(H/2); odd H / W: the missing row / column is skipped)
template <int K>
__global__ void __launch_bounds__(256) quad_unshuffle_kernel(const float* __restrict__ q, int q_stride, float* __restrict__ out, int N, int H,
                                                             int W) {
    const int H1 = (H + 1) >> 1, W1 = (W + 1) >> 1;
    const long long total = (long long)N * H1 * W1, plane = (long long)N * H * W;
    const bool vec = (W & 1) == 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % W1), i = (int)((t / W1) % H1), n = (int)(t / ((long long)W1 * H1));
        const float* qp = q + (((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1) * q_stride;
        float qv[4 * K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(qp) + k);
            qv[4 * k] = v.x; qv[4 * k + 1] = v.y; qv[4 * k + 2] = v.z; qv[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int y = 2 * i + py;
                if (y >= H) continue;
                float* op = out + k * plane + ((long long)n * H + y) * W + 2 * j;
                const float a = qv[(py * 2) * K + k], b = qv[(py * 2 + 1) * K + k];
                if (vec) {
                    *reinterpret_cast<float2*>(op) = make_float2(a, b);
                } else {
                    op[0] = a;
                    if (2 * j + 1 < W) op[1] = b;
                }
            }
    }
}

// in0[n][y+1][x+1][0..15] = {xout_d, r, g, b, cout_d, 0 x 11} as bf16 (hi, lo)
__global__ void __launch_bounds__(256) ged_pack_input_kernel(const float* __restrict__ xout, const float* __restrict__ cout,
                                                             const float* __restrict__ rgb, long long rgb_bs, __nv_bfloat16* __restrict__ dst,
                                                             __nv_bfloat16* __restrict__ dst_lo, int dst_stride, int N, int H, int W) {
    const long long HW = (long long)H * W, total = (long long)N * HW;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long p = t % HW;
        const int n = (int)(t / HW);
        const int x = (int)(p % W), y = (int)(p / W);
        float v[8];
        v[0] = __ldg(xout + t);
        v[1] = __ldg(rgb + n * rgb_bs + p);
        v[2] = __ldg(rgb + n * rgb_bs + HW + p);
        v[3] = __ldg(rgb + n * rgb_bs + 2 * HW + p);
        v[4] = __ldg(cout + t);
        v[5] = v[6] = v[7] = 0.f;
        uint4 h, l;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
        __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hh[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            ll[k] = __floats2bfloat162_rn(v[2 * k] - __low2float(hh[k]), v[2 * k + 1] - __high2float(hh[k]));
        }
        const long long pix = ((long long)n * (H + 2) + y + 1) * (W + 2) + x + 1;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4* o = reinterpret_cast<uint4*>(dst + pix * dst_stride);
        o[0] = h;
        o[1] = z;
        if (dst_lo) {
            uint4* ol = reinterpret_cast<uint4*>(dst_lo + pix * dst_stride);
            ol[0] = l;
            ol[1] = z;
        }
    }
}

// conv1 (3x3, stride 2, 5 -> 80 channels) as a GEMM: im2col of its input at the OUTPUT resolution.
// x45[n][i+1][j+1][(u*3+v)*5 + c] = in[n][c][2i+u-1][2j+v-1] (0 outside), c = {xout_d, r, g, b, cout_d}; 45 of 48 written channels
// are real, bf16 (hi, lo).  The layer's forward is then a 1x1 convolution of x45 and its weight gradient a plain GEMM of the
// output gradient with x45, both at a quarter of the pixels and without strided tensor-map boxes.
// Thread = one output pixel (45 gathered loads, bf16 (hi, lo) split in registers); the 96 + 96 bytes of a pixel then go through a
// per-warp shared-memory stage so that consecutive lanes store consecutive 16-byte groups (6 lanes per pixel row: ~6 lines per
// store instruction instead of 32 half-written sectors 128 bytes apart -- the LSU-to-L2 request rate, not DRAM, bounded the
// direct form: 208 us for 585 MB at B = 128).  Stage rows are 7 groups apart (conflict-free 16-byte accesses).
__global__ void __launch_bounds__(128) ged_pack_im2col_kernel(const float* __restrict__ xout, const float* __restrict__ cout,
                                                              const float* __restrict__ rgb, long long rgb_bs, __nv_bfloat16* __restrict__ dst,
                                                              __nv_bfloat16* __restrict__ dst_lo, int dst_stride, int N, int H, int W) {
    __shared__ uint4 stage[4][2][32 * 7];
    __shared__ long long spix[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H1 = (H + 1) >> 1, W1 = (W + 1) >> 1;
    const long long HW = (long long)H * W;
    const unsigned total = (unsigned)N * H1 * W1;                 // host: < 2^31
    const unsigned wstep = gridDim.x * 4u * 32u;
    for (unsigned base = (blockIdx.x * 4u + warp) * 32u; base < total; base += wstep) {
        const unsigned t = base + lane;
        if (t < total) {
            const unsigned j = t % (unsigned)W1, r = t / (unsigned)W1, i = r % (unsigned)H1, n = r / (unsigned)H1;
            const float* pl[5] = {xout + n * HW, rgb + n * rgb_bs, rgb + n * rgb_bs + HW, rgb + n * rgb_bs + 2 * HW, cout + n * HW};
            float v[48];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int y = 2 * (int)i + u - 1;
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int x = 2 * (int)j + w - 1;
                    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
                    const int o = ok ? y * W + x : 0;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const float f = __ldg(pl[c] + o);
                        v[(u * 3 + w) * 5 + c] = ok ? f : 0.f;
                    }
                }
            }
            v[45] = v[46] = v[47] = 0.f;
            spix[warp][lane] = ((long long)n * (H1 + 2) + i + 1) * (W1 + 2) + j + 1;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                uint4 h, l;
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
                __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = v[q * 8 + 2 * k], b = v[q * 8 + 2 * k + 1];
                    hh[k] = __floats2bfloat162_rn(a, b);
                    ll[k] = __floats2bfloat162_rn(a - __low2float(hh[k]), b - __high2float(hh[k]));
                }
                stage[warp][0][lane * 7 + q] = h;
                stage[warp][1][lane * 7 + q] = l;
            }
        }
        __syncwarp();
        const unsigned left = total - base;
        const int groups = (int)(left < 32u ? left : 32u) * 6;
        for (int k = lane; k < groups; k += 32) {
            const int p = k / 6, q = k - p * 6;
            const long long pix = spix[warp][p];
            reinterpret_cast<uint4*>(dst + pix * dst_stride)[q] = stage[warp][0][p * 7 + q];
            if (dst_lo) reinterpret_cast<uint4*>(dst_lo + pix * dst_stride)[q] = stage[warp][1][p * 7 + q];
        }
        __syncwarp();
    }
}

static int grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    const long long cap = 148LL * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

constexpr int kPrepCtas = 148 * 4;

}  // namespace quad

using namespace quad;

extern "C" {

int adc_ged_pack_input(const float* xout, const float* cout, const float* rgb, long long rgb_batch_stride, void* dst, void* dst_lo,
                       int dst_stride, int N, int H, int W, adc_stream_t stream) {
    if (!xout || !cout || !rgb || !dst || N <= 0 || H <= 0 || W <= 0 || dst_stride < 16 || dst_stride % 8 != 0) {
        adc::set_error("adc_ged_pack_input: bad argument");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)total * (20.0 + (dst_lo ? 64.0 : 32.0)));
    ged_pack_input_kernel<<<grid_for(total, 256), 256, 0, st>>>(xout, cout, rgb, rgb_batch_stride, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_lo,
                                                                 dst_stride, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "ged_pack_input_kernel launch");
}

int adc_ged_pack_im2col(const float* xout, const float* cout, const float* rgb, long long rgb_batch_stride, void* dst, void* dst_lo,
                        int dst_stride, int N, int H, int W, adc_stream_t stream) {
    if (!xout || !cout || !rgb || !dst || N <= 0 || H <= 0 || W <= 0 || dst_stride < 48 || dst_stride % 8 != 0) {
        adc::set_error("adc_ged_pack_im2col: bad argument");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2);
    if (total >= (1LL << 31) || (long long)H * W >= (1LL << 31)) {
        adc::set_error("adc_ged_pack_im2col: more than 2^31 pixels");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)N * H * W * 20.0 + (double)total * (dst_lo ? 192.0 : 96.0));
    ged_pack_im2col_kernel<<<grid_for(total, 128) * 2, 128, 0, st>>>(xout, cout, rgb, rgb_batch_stride, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_lo,
                                                                      dst_stride, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "ged_pack_im2col_kernel launch");
}

int adc_quad_head_finish(const float* q, int q_stride, const float* x0d, const float* w, const float* bias, float* out, int N, int oc, int H,
                         int W, adc_stream_t stream) {
    if (!q || !x0d || !w || !bias || !out || N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || oc < 1 || oc > 2 || q_stride < 4 * oc ||
        q_stride % 4 != 0) {
        adc::set_error("adc_quad_head_finish: bad argument (needs even H, W; oc in {1,2})");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * (H / 2) * (W / 2);
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)total * (16.0 * oc + 16.0 + 16.0 * oc));
    if (oc == 1) quad_head_finish_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(q, q_stride, x0d, w, bias, out, N, H, W);
    else quad_head_finish_kernel<2><<<grid_for(total, 256), 256, 0, st>>>(q, q_stride, x0d, w, bias, out, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "quad_head_finish_kernel launch");
}

int adc_quad_head_bwd_prep_parts(void) { return kPrepCtas; }

int adc_quad_head_bwd_prep(const float* gy, const float* x10, const float* x0d, float* gq, float* part, int N, int oc, int H, int W,
                           adc_stream_t stream) {
    if (!gy || !x10 || !x0d || !gq || !part || N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || oc < 1 || oc > 2) {
        adc::set_error("adc_quad_head_bwd_prep: bad argument (needs even H, W; oc in {1,2})");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * (H / 2) * (W / 2);
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)total * (32.0 * oc + 16.0 + 16.0 * oc));
    if (oc == 1) quad_head_bwd_prep_kernel<1><<<kPrepCtas, 256, 0, st>>>(gy, x10, x0d, gq, part, N, H, W);
    else quad_head_bwd_prep_kernel<2><<<kPrepCtas, 256, 0, st>>>(gy, x10, x0d, gq, part, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "quad_head_bwd_prep_kernel launch");
}

int adc_quad_head_wgrad(const void* x9, const void* x9_lo, int x9_stride, const float* gq, float* part, int N, int oc, int H1, int W1,
                        adc_stream_t stream) {
    if (!x9 || !gq || !part || N <= 0 || H1 <= 0 || W1 <= 0 || oc < 1 || oc > 2 || x9_stride < 32) {
        adc::set_error("adc_quad_head_wgrad: bad argument");
        return ADC_E_BADARG;
    }
    X9v x{(const __nv_bfloat16*)x9, (const __nv_bfloat16*)x9_lo, x9_stride};
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)N * H1 * W1 * (32.0 * (x9_lo ? 4.0 : 2.0) + 16.0 * oc));
    static bool attr_done = false;
    const int smem1 = (kQT_XR * kQT_XC * 32 + kQT_R * kQT_C * 4) * 4, smem2 = (kQT_XR * kQT_XC * 32 + kQT_R * kQT_C * 8) * 4;
    if (!attr_done) {
        int rc = adc::check_cuda(cudaFuncSetAttribute(quad_head_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1), "smem attr");
        if (!rc) rc = adc::check_cuda(cudaFuncSetAttribute(quad_head_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2), "smem attr");
        if (rc) return rc;
        attr_done = true;
    }
    if (oc == 1) quad_head_wgrad_kernel<1><<<kPrepCtas, 256, smem1, st>>>(x, gq, part, N, H1, W1);
    else quad_head_wgrad_kernel<2><<<kPrepCtas, 256, smem2, st>>>(x, gq, part, N, H1, W1);
    return adc::check_cuda(cudaGetLastError(), "quad_head_wgrad_kernel launch");
}

int adc_quad_head_dgrad(const void* x9, int x9_stride, const float* gq, const float* wq, const void* extra, const void* extra_lo,
                        int extra_stride, void* g9, void* g9_lo, int g9_stride, float* part, int N, int oc, int H1, int W1,
                        adc_stream_t stream) {
    if (!x9 || !gq || !wq || !g9 || !part || N <= 0 || H1 <= 0 || W1 <= 0 || oc < 1 || oc > 2 || x9_stride < 32 || g9_stride < 32 ||
        (extra && extra_stride < 32)) {
        adc::set_error("adc_quad_head_dgrad: bad argument");
        return ADC_E_BADARG;
    }
    X9v x{(const __nv_bfloat16*)x9, nullptr, x9_stride};
    X9v e{(const __nv_bfloat16*)extra, (const __nv_bfloat16*)extra_lo, extra_stride};
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_HEAD10, (double)N * H1 * W1 * (32.0 * (2.0 + (g9_lo ? 4.0 : 2.0)) + 16.0 * oc));
    if (oc == 1)
        quad_head_dgrad_kernel<1><<<kPrepCtas, 256, 0, st>>>(x, gq, wq, e, (__nv_bfloat16*)g9, (__nv_bfloat16*)g9_lo, g9_stride, part, N, H1, W1);
    else
        quad_head_dgrad_kernel<2><<<kPrepCtas, 256, 0, st>>>(x, gq, wq, e, (__nv_bfloat16*)g9, (__nv_bfloat16*)g9_lo, g9_stride, part, N, H1, W1);
    return adc::check_cuda(cudaGetLastError(), "quad_head_dgrad_kernel launch");
}

int adc_quad_unshuffle(const float* q, int q_stride, float* out, int N, int K, int H, int W, adc_stream_t stream) {
    if (!q || !out || N <= 0 || H <= 0 || W <= 0 || K < 1 || K > 2 || q_stride < 4 * K || q_stride % 4 != 0) {
        adc::set_error("adc_quad_unshuffle: bad argument (K in {1,2})");
        return ADC_E_BADARG;
    }
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2);
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_LAYOUT, (double)total * 32.0 * K);
    if (K == 1) quad_unshuffle_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(q, q_stride, out, N, H, W);
    else quad_unshuffle_kernel<2><<<grid_for(total, 256), 256, 0, st>>>(q, q_stride, out, N, H, W);
    return adc::check_cuda(cudaGetLastError(), "quad_unshuffle_kernel launch");
}

}  // extern "C"

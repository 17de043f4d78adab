// NHWC glue of the tensor-core training path (backward of GEDNet / GMSNet, confidence_depth_framework.py:464-518,630-669):
// nearest up-sampling into a channel slice, the gradient "assemble" step (sum over the up-sampling children + skip gradient,
// times the activation derivative), and per-channel sums (bias gradients).  Activations and gradients are zero-bordered NHWC
// bf16 tensors [N][H+2][W+2][C], single (plain bf16 mode) or as a (hi, lo) pair whose sum carries 16 mantissa bits; all
// arithmetic here is fp32 on hi + lo and results are re-split.  HBM-bound elementwise kernels, 16-byte accesses.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace glue {

// ATen 'nearest' source index (same rule as adc_common.cuh NearestMap; restated here because that header owns the
// constant-memory weight images of another translation unit)
struct Near {
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
};

struct Slice {            // channel slice of a zero-bordered NHWC tensor
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;   // may be null
    int stride, off;
};
struct SliceW {
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    int stride, off;
};

__device__ __forceinline__ void load8(const Slice& s, long long pix, int c, float (&v)[8], bool accumulate) {
    const uint4 h = *reinterpret_cast<const uint4*>(s.hi + pix * s.stride + s.off + c);
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&h);
    float t[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t[2 * j] = __low2float(hh[j]);
        t[2 * j + 1] = __high2float(hh[j]);
    }
    if (s.lo) {
        const uint4 l = *reinterpret_cast<const uint4*>(s.lo + pix * s.stride + s.off + c);
        const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t[2 * j] += __low2float(ll[j]);
            t[2 * j + 1] += __high2float(ll[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = accumulate ? v[j] + t[j] : t[j];
}
__device__ __forceinline__ void store8(const SliceW& s, long long pix, int c, const float (&v)[8]) {
    uint4 h, l;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
    __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        ll[j] = __floats2bfloat162_rn(v[2 * j] - __low2float(hh[j]), v[2 * j + 1] - __high2float(hh[j]));
    }
    *reinterpret_cast<uint4*>(s.hi + pix * s.stride + s.off + c) = h;
    if (s.lo) *reinterpret_cast<uint4*>(s.lo + pix * s.stride + s.off + c) = l;
}

// dst[n][yd][xd][c] = src[n][mh(yd)][mw(xd)][c]   (bit copy of hi and lo)
// 32-bit index arithmetic (host: N*Hd*Wd*C8 < 2^31 -- the 64-bit divisions cost more issue slots than the copies) and two
// 16-byte groups per thread with all loads issued before the first store.
__global__ void __launch_bounds__(256) nhwc_upsample_kernel(Slice src, SliceW dst, int N, int C8, int Hs, int Ws, int Hd, int Wd, Near mh,
                                                            Near mw) {
    const unsigned total = (unsigned)N * Hd * Wd * C8, step = gridDim.x * blockDim.x;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * step) {
        long long sp[2], dp[2];
        int c[2];
        bool ok[2];
        uint4 h[2], l[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned i = i0 + q * step;
            ok[q] = i < total;
            const unsigned ii = ok[q] ? i : 0u;
            c[q] = (int)(ii % (unsigned)C8) * 8;
            unsigned r = ii / (unsigned)C8;
            const unsigned xd = r % (unsigned)Wd;
            r /= (unsigned)Wd;
            const unsigned yd = r % (unsigned)Hd, n = r / (unsigned)Hd;
            sp[q] = ((long long)n * (Hs + 2) + mh((int)yd) + 1) * (Ws + 2) + mw((int)xd) + 1;
            dp[q] = ((long long)n * (Hd + 2) + yd + 1) * (Wd + 2) + xd + 1;
            h[q] = *reinterpret_cast<const uint4*>(src.hi + sp[q] * src.stride + src.off + c[q]);
            if (dst.lo) l[q] = *reinterpret_cast<const uint4*>(src.lo + sp[q] * src.stride + src.off + c[q]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!ok[q]) continue;
            *reinterpret_cast<uint4*>(dst.hi + dp[q] * dst.stride + dst.off + c[q]) = h[q];
            if (dst.lo) *reinterpret_cast<uint4*>(dst.lo + dp[q] * dst.stride + dst.off + c[q]) = l[q];
        }
    }
}

// out[n][y][x][c] = d(act)[n][y][x][c] * ( sum over the up-sampling children (yd,xd) of (y,x) of A[n][yd][xd][c]  +  B[n][y][x][c] )
// d(act) = 1 where act > 0, else slope (ReLU: 0, LeakyReLU: 0.2); A lives at (Hd,Wd) >= (H,W) (identity map if equal), A/B/act
// optional.  Fixed summation order (deterministic).  A thread keeps ONE 8-channel group for its whole life (thread = (pixel
// slot, group)), so the per-channel sum of the output -- the bias gradient of the layer that produced `act` -- falls out of the
// same pass: part[block][C] partial rows, summed by the caller in index order.
// Three resident CTAs per SM (80 registers, 16 bytes of spills): measured on one box against the 116-register / two-CTA build,
// 551 vs 589 us for G1 at B = 128 (glue per step 1 390 -> 1 310 us); four CTAs (64 registers, 114 bytes of spills): 688 us.
__global__ void __launch_bounds__(256, 3) nhwc_grad_assemble_kernel(Slice A, int hasA, Slice B, int hasB, Slice act, int hasAct, float slope,
                                                                 SliceW out, int N, int C8, int H, int W, int Hd, int Wd, Near mh, Near mw,
                                                                 float* __restrict__ part, int a_quad) {
    __shared__ float red[256][8];
    const int groups = 256 / C8;                               // C8 <= 32
    const int cl = threadIdx.x % C8, pg = threadIdx.x / C8, c = cl * 8;
    const unsigned pixels = (unsigned)N * H * W;               // (host: < 2^31) -- 32-bit index arithmetic: the 64-bit divisions of the
    const unsigned step = gridDim.x * groups;                  // first version cost more issue slots than the loads
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    // single-source forms of A (quad block or same-size tensor): two pixels per iteration, all ten 16-byte loads issued before the
    // first use; the multi-child forms already have 8+ loads in flight per pixel
    const bool fast = !hasA || a_quad || (Hd == H && Wd == W);
    // general up-sampling ratios (odd sizes: Hd = 2H - 1): the children of row y / column x are a contiguous range, found ONCE per CTA
    // (the first version searched 5 x 5 candidates with a floorf each for every pixel and 8-channel group: 100+ us on the small levels)
    __shared__ short ytab[2][1024], xtab[2][1024];
    const bool tables = !fast && !(Hd == 2 * H && Wd == 2 * W) && H <= 1024 && W <= 1024;
    if (tables) {
        for (int i = threadIdx.x; i < H + W; i += blockDim.x) {
            const bool isy = i < H;
            const int q = isy ? i : i - H, in = isy ? H : W, out = isy ? Hd : Wd;
            const Near& m = isy ? mh : mw;
            int d0 = (int)(((long long)q * out) / in) - 1, first = -1, cnt = 0;
            for (int d = d0 < 0 ? 0 : d0; d < out && d <= d0 + 4; ++d)
                if (m(d) == q) {
                    if (first < 0) first = d;
                    ++cnt;
                }
            (isy ? ytab : xtab)[0][q] = (short)(first < 0 ? 0 : first);
            (isy ? ytab : xtab)[1][q] = (short)cnt;
        }
        __syncthreads();
    }
    if (pg < groups && fast) {
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (unsigned r0 = blockIdx.x * groups + pg; r0 < pixels; r0 += 2 * step) {
            long long po[2];
            uint4 ah[2], al[2], bh[2], bl[2], tt[2];
            bool ok[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const unsigned r = r0 + q * step;
                ok[q] = r < pixels;
                const unsigned rr = ok[q] ? r : 0u;
                const unsigned x = rr % (unsigned)W, t = rr / (unsigned)W, y = t % (unsigned)H, n = t / (unsigned)H;
                po[q] = ((long long)n * (H + 2) + y + 1) * (W + 2) + x + 1;
                ah[q] = al[q] = bh[q] = bl[q] = tt[q] = zero;
                if (hasA) {
                    const long long pa = a_quad ? ((long long)n * (Hd + 2) + (y >> 1) + 1) * (Wd + 2) + (x >> 1) + 1 : po[q];
                    const int ca = a_quad ? (int)((y & 1) * 2 + (x & 1)) * C8 * 8 + c : c;
                    ah[q] = *reinterpret_cast<const uint4*>(A.hi + pa * A.stride + A.off + ca);
                    if (A.lo) al[q] = *reinterpret_cast<const uint4*>(A.lo + pa * A.stride + A.off + ca);
                }
                if (hasB) {
                    bh[q] = *reinterpret_cast<const uint4*>(B.hi + po[q] * B.stride + B.off + c);
                    if (B.lo) bl[q] = *reinterpret_cast<const uint4*>(B.lo + po[q] * B.stride + B.off + c);
                }
                if (hasAct) tt[q] = *reinterpret_cast<const uint4*>(act.hi + po[q] * act.stride + act.off + c);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!ok[q]) continue;
                float v[8];
                const __nv_bfloat162* p0 = reinterpret_cast<const __nv_bfloat162*>(&ah[q]);
                const __nv_bfloat162* p1 = reinterpret_cast<const __nv_bfloat162*>(&al[q]);
                const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&bh[q]);
                const __nv_bfloat162* p3 = reinterpret_cast<const __nv_bfloat162*>(&bl[q]);
                const __nv_bfloat162* pt = reinterpret_cast<const __nv_bfloat162*>(&tt[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // same association as load8: (A_hi + A_lo) + (B_hi + B_lo)
                    const float a0 = __low2float(p0[j]) + __low2float(p1[j]), a1 = __high2float(p0[j]) + __high2float(p1[j]);
                    const float b0 = __low2float(p2[j]) + __low2float(p3[j]), b1 = __high2float(p2[j]) + __high2float(p3[j]);
                    v[2 * j] = hasA ? (hasB ? a0 + b0 : a0) : b0;
                    v[2 * j + 1] = hasA ? (hasB ? a1 + b1 : a1) : b1;
                    if (hasAct) {
                        v[2 * j] *= __low2float(pt[j]) > 0.f ? 1.f : slope;
                        v[2 * j + 1] *= __high2float(pt[j]) > 0.f ? 1.f : slope;
                    }
                }
                store8(out, po[q], c, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += v[j];
            }
        }
    } else if (pg < groups)
        for (unsigned r = blockIdx.x * groups + pg; r < pixels; r += step) {
            const int x = (int)(r % (unsigned)W), y = (int)((r / (unsigned)W) % (unsigned)H), n = (int)(r / ((unsigned)W * (unsigned)H));
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (Hd == 2 * H && Wd == 2 * W) {              // exact 2x: the four children, row-major
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx)
                        load8(A, ((long long)n * (Hd + 2) + 2 * y + dy + 1) * (Wd + 2) + 2 * x + dx + 1, c, v, true);
            } else if (tables) {
                const int yb = ytab[0][y], yn = ytab[1][y], xb = xtab[0][x], xn = xtab[1][x];
                for (int yd = yb; yd < yb + yn; ++yd)
                    for (int xd = xb; xd < xb + xn; ++xd)
                        load8(A, ((long long)n * (Hd + 2) + yd + 1) * (Wd + 2) + xd + 1, c, v, true);
            } else {
                // children of y: the yd with mh(yd) == y form a contiguous range around y*Hd/H
                int y0 = (int)(((long long)y * Hd) / H) - 1, x0 = (int)(((long long)x * Wd) / W) - 1;
                for (int yd = y0 < 0 ? 0 : y0; yd < Hd && yd <= y0 + 4; ++yd) {
                    if (mh(yd) != y) continue;
                    for (int xd = x0 < 0 ? 0 : x0; xd < Wd && xd <= x0 + 4; ++xd) {
                        if (mw(xd) != x) continue;
                        load8(A, ((long long)n * (Hd + 2) + yd + 1) * (Wd + 2) + xd + 1, c, v, true);
                    }
                }
            }
            const long long p = ((long long)n * (H + 2) + y + 1) * (W + 2) + x + 1;
            if (hasB) load8(B, p, c, v, true);
            if (hasAct) {
                const uint4 a = *reinterpret_cast<const uint4*>(act.hi + p * act.stride + act.off + c);
                const __nv_bfloat162* aa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[2 * j] *= __low2float(aa[j]) > 0.f ? 1.f : slope;
                    v[2 * j + 1] *= __high2float(aa[j]) > 0.f ? 1.f : slope;
                }
            }
            store8(out, p, c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += v[j];
        }
    if (part == nullptr) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = pg < groups ? s[j] : 0.f;
    __syncthreads();
    if (threadIdx.x < C8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        for (int q = 0; q < groups; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] += red[q * C8 + threadIdx.x][j];
#pragma unroll
        for (int j = 0; j < 8; ++j) part[(long long)blockIdx.x * C8 * 8 + threadIdx.x * 8 + j] = t[j];
    }
}

// per-channel sums over all pixels (bias gradient): stage 1 = per-CTA partials [grid][C], stage 2 = fixed-order sum
__global__ void __launch_bounds__(256) nhwc_colsum1_kernel(Slice g, long long F, int C8, float* __restrict__ part) {
    __shared__ float red[256][8];
    const int lanes = C8, groups = 256 / lanes;                // C8 <= 32
    const int cl = threadIdx.x % lanes, pg = threadIdx.x / lanes;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (pg < groups)
        for (long long p = (long long)blockIdx.x * groups + pg; p < F; p += (long long)gridDim.x * groups) load8(g, p, cl * 8, v, true);
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = v[j];
    __syncthreads();
    if (threadIdx.x < lanes) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        for (int q = 0; q < groups; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] += red[q * lanes + threadIdx.x][j];
#pragma unroll
        for (int j = 0; j < 8; ++j) part[(long long)blockIdx.x * C8 * 8 + threadIdx.x * 8 + j] = t[j];
    }
}
__global__ void nhwc_colsum2_kernel(const float* __restrict__ part, int nparts, int Cp, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t = 0.f;
    for (int q = 0; q < nparts; ++q) t += part[(long long)q * Cp + c];
    out[c] = t;
}

// out[c] = sum_r part[r][c] in a fixed order (rows strided over 32 groups per thread column, then the groups in index order):
// the second stage of every deterministic reduction of this library (a library reduce over dim 0 of a [~1000, C] tensor
// launches 1-2 CTAs and measured 26 us).
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ part, int rows, int C, float* __restrict__ out) {
    __shared__ float red[32][33];
    const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5, c = blockIdx.x * 32 + cl;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // four loads in flight per thread (the serial version was latency-bound: 18 us)
    if (c < C) {
        int r = rg;
        for (; r + 96 < rows; r += 128) {
            s0 += part[(long long)r * C + c];
            s1 += part[(long long)(r + 32) * C + c];
            s2 += part[(long long)(r + 64) * C + c];
            s3 += part[(long long)(r + 96) * C + c];
        }
        for (; r < rows; r += 32) s0 += part[(long long)r * C + c];
    }
    red[rg][cl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (rg == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][cl];
        out[c] = t;
    }
}

// Quad (parity-block) forms of a 3x3 weight tensor, one thread per output element (the broadcast product + reduction these replace
// materialised a [4][A][B][9][9] tensor per layer and step through three library launches):
//  which 0 'up': conv3x3(up2(x)) at output parity (py,px): original tap (u,v) falls on low-resolution tap ((py+u-1)>>1)+1, ((px+v-1)>>1)+1
//  which 1 's2': data gradient of a stride-2 conv at input parity (py,px): low-resolution tap (a,b) holds original tap (py+3-2a, px+3-2b)
// forward: out[q][a][b][s] (or [q][b][a][s] when transposed) = sum over the original taps t that map to s of w[a][b][t]
__device__ __forceinline__ int quad_up_slot(int py, int px, int t) {
    const int u = t / 3, v = t - 3 * u;
    return (((py + u - 1) >> 1) + 1) * 3 + ((px + v - 1) >> 1) + 1;
}
__global__ void __launch_bounds__(256) quad_tap_mix_kernel(const float* __restrict__ w, float* __restrict__ out, int A, int B, int which,
                                                           int transposed) {
    const int total = 4 * A * B * 9;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = i % 9, l = (i / 9) % (A * B), q = i / (9 * A * B);
    const int py = q >> 1, px = q & 1;
    const float* wl = w + (size_t)l * 9;
    float r = 0.f;
    if (which == 0) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
            if (quad_up_slot(py, px, t) == s) r += __ldg(wl + t);
    } else {
        const int a = s / 3, b = s - 3 * a, u = py + 3 - 2 * a, v = px + 3 - 2 * b;
        if (u >= 0 && u < 3 && v >= 0 && v < 3) r = __ldg(wl + u * 3 + v);
    }
    const int ia = l / B, ib = l - ia * B;
    const size_t o = transposed ? ((size_t)q * B + ib) * A + ia : (size_t)q * A * B + l;
    out[o * 9 + s] = r;
}
// adjoint of the 'up' form: out[l][t] = sum_q in[q][l][slot(q, t)]
__global__ void __launch_bounds__(256) quad_tap_mix_adjoint_kernel(const float* __restrict__ in, float* __restrict__ out, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * 9) return;
    const int t = i % 9, l = i / 9;
    float r = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) r += __ldg(in + ((size_t)q * L + l) * 9 + quad_up_slot(q >> 1, q & 1, t));
    out[i] = r;
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

constexpr int kAssembleCtas = 148 * 9;      // three resident CTAs per SM (80 registers), three waves

static bool bad_slice(const void* hi, int stride, int off, int C) { return !hi || stride % 8 != 0 || off % 8 != 0 || C % 8 != 0 || off + C > stride; }

}  // namespace glue

using namespace glue;

extern "C" {

int adc_sum_rows(const float* part, int rows, int C, float* out, adc_stream_t stream) {
    if (!part || !out || rows <= 0 || C <= 0) {
        adc::set_error("adc_sum_rows: bad argument");
        return ADC_E_BADARG;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_GLUE, (double)rows * C * 4.0);
    sum_rows_kernel<<<(C + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(part, rows, C, out);
    return adc::check_cuda(cudaGetLastError(), "sum_rows_kernel launch");
}

int adc_nhwc_upsample_nearest(const void* src, const void* src_lo, int src_stride, int src_off, void* dst, void* dst_lo, int dst_stride,
                              int dst_off, int N, int C, int Hs, int Ws, int Hd, int Wd, adc_stream_t stream) {
    if (bad_slice(src, src_stride, src_off, C) || bad_slice(dst, dst_stride, dst_off, C) || N <= 0 || Hs <= 0 || Ws <= 0 || Hd <= 0 || Wd <= 0 ||
        ((src_lo == nullptr) != (dst_lo == nullptr))) {
        adc::set_error("adc_nhwc_upsample_nearest: bad argument (channel slices must be multiples of 8 channels, 16-byte aligned)");
        return ADC_E_BADARG;
    }
    Slice s{(const __nv_bfloat16*)src, (const __nv_bfloat16*)src_lo, src_stride, src_off};
    SliceW d{(__nv_bfloat16*)dst, (__nv_bfloat16*)dst_lo, dst_stride, dst_off};
    const long long total = (long long)N * Hd * Wd * (C / 8);
    if (total >= (1LL << 31)) {
        adc::set_error("adc_nhwc_upsample_nearest: more than 2^31 16-byte groups");
        return ADC_E_BADARG;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_GLUE, (double)total * 16.0 * (src_lo ? 4 : 2));
    nhwc_upsample_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(s, d, N, C / 8, Hs, Ws, Hd, Wd, Near::make(Hs, Hd), Near::make(Ws, Wd));
    return adc::check_cuda(cudaGetLastError(), "nhwc_upsample_kernel launch");
}

int adc_nhwc_grad_assemble_parts(void) { return kAssembleCtas; }

int adc_nhwc_grad_assemble(const void* a, const void* a_lo, int a_stride, int a_off, int Hd, int Wd, const void* b, const void* b_lo,
                           int b_stride, int b_off, const void* act, int act_stride, int act_off, float slope, void* out, void* out_lo,
                           int out_stride, int out_off, int N, int C, int H, int W, float* colsum_part, int a_quad, adc_stream_t stream) {
    if (C > 256 || (!a && !b) || bad_slice(out, out_stride, out_off, C) || (a && bad_slice(a, a_stride, a_off, a_quad ? 4 * C : C)) ||
        (b && bad_slice(b, b_stride, b_off, C)) || (act && bad_slice(act, act_stride, act_off, C)) || N <= 0 || H <= 0 || W <= 0 ||
        (a && !a_quad && (Hd < H || Wd < W)) || (a && a_quad && (Hd != (H + 1) / 2 || Wd != (W + 1) / 2)) ||
        (long long)N * H * W >= (1LL << 31)) {
        adc::set_error("adc_nhwc_grad_assemble: bad argument");
        return ADC_E_BADARG;
    }
    Slice A{(const __nv_bfloat16*)a, (const __nv_bfloat16*)a_lo, a_stride, a_off};
    Slice B{(const __nv_bfloat16*)b, (const __nv_bfloat16*)b_lo, b_stride, b_off};
    Slice T{(const __nv_bfloat16*)act, nullptr, act_stride, act_off};
    SliceW O{(__nv_bfloat16*)out, (__nv_bfloat16*)out_lo, out_stride, out_off};
    if (!a) { Hd = H; Wd = W; }
    const long long total = (long long)N * H * W * (C / 8);
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_GLUE, (double)total * 16.0 * (out_lo ? 2 : 1) * (1 + (a ? 1 : 0) + (b ? 1 : 0)));
    // Small levels: one wave of CTAs instead of the full grid (at 116 registers two CTAs shared an SM, so the fixed grid of 1184 ran 4 latency-bound waves
    // of ~8 us each whatever the tensor size: 33-43 us for the 29x38 and smaller levels at B = 32).  The partial rows of the CTAs that are
    // not launched are zeroed (the caller sums all adc_nhwc_grad_assemble_parts() rows).
    const int groups_per_cta = 256 / (C / 8);
    long long want = ((long long)N * H * W + 2LL * groups_per_cta - 1) / (2LL * groups_per_cta);
    int grid = (int)(want < 148 ? 148 : (want > kAssembleCtas ? kAssembleCtas : want));
    if (grid < kAssembleCtas && colsum_part) {
        int rc0 = adc::check_cuda(cudaMemsetAsync(colsum_part + (size_t)grid * C, 0, (size_t)(kAssembleCtas - grid) * C * sizeof(float),
                                                  (cudaStream_t)stream), "cudaMemsetAsync(colsum partial rows)");
        if (rc0) return rc0;
    }
    nhwc_grad_assemble_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, a != nullptr, B, b != nullptr, T, act != nullptr, slope, O, N,
                                                                             C / 8, H, W, Hd, Wd, Near::make(H, Hd), Near::make(W, Wd), colsum_part, a_quad);
    return adc::check_cuda(cudaGetLastError(), "nhwc_grad_assemble_kernel launch");
}

size_t adc_nhwc_colsum_workspace_bytes(int C) { return (size_t)148 * 4 * ((C + 7) / 8 * 8) * sizeof(float); }

int adc_nhwc_colsum(const void* g, const void* g_lo, int g_stride, int g_off, int N, int C, int H, int W, float* out, void* workspace,
                    adc_stream_t stream) {
    const int Cp = (C + 7) / 8 * 8;
    if (!g || !out || !workspace || g_stride % 8 != 0 || g_off % 8 != 0 || g_off + Cp > g_stride || Cp > 256 || N <= 0 || H <= 0 || W <= 0) {
        adc::set_error("adc_nhwc_colsum: bad argument");
        return ADC_E_BADARG;
    }
    Slice G{(const __nv_bfloat16*)g, (const __nv_bfloat16*)g_lo, g_stride, g_off};
    const long long F = (long long)N * (H + 2) * (W + 2);      // border pixels are zeros
    const int nparts = 148 * 4;
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_GLUE, (double)F * Cp * (g_lo ? 4.0 : 2.0));
    adc::count_launches(1);
    nhwc_colsum1_kernel<<<nparts, 256, 0, st>>>(G, F, Cp / 8, reinterpret_cast<float*>(workspace));
    int rc = adc::check_cuda(cudaGetLastError(), "nhwc_colsum1_kernel launch");
    if (rc) return rc;
    nhwc_colsum2_kernel<<<(C + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float*>(workspace), nparts, Cp, C, out);
    return adc::check_cuda(cudaGetLastError(), "nhwc_colsum2_kernel launch");
}

int adc_quad_tap_mix(const float* w, float* out, int A, int B, int which, int transposed, adc_stream_t stream) {
    if (!w || !out || A <= 0 || B <= 0 || which < 0 || which > 1 || (long long)A * B * 36 >= (1LL << 31)) {
        adc::set_error("adc_quad_tap_mix: bad argument");
        return ADC_E_BADARG;
    }
    const int total = 4 * A * B * 9;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_PREP, (double)total * 8.0);
    quad_tap_mix_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, out, A, B, which, transposed);
    return adc::check_cuda(cudaGetLastError(), "quad_tap_mix_kernel launch");
}

int adc_quad_tap_mix_adjoint(const float* in, float* out, int L, adc_stream_t stream) {
    if (!in || !out || L <= 0 || (long long)L * 36 >= (1LL << 31)) {
        adc::set_error("adc_quad_tap_mix_adjoint: bad argument");
        return ADC_E_BADARG;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_PREP, (double)L * 9 * 20.0);
    quad_tap_mix_adjoint_kernel<<<(L * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, out, L);
    return adc::check_cuda(cudaGetLastError(), "quad_tap_mix_adjoint_kernel launch");
}

}  // extern "C"

// K1: fused NConv2d forward (model_zoo/nconv.py:37-67) for sm_100a.
//
// One CTA = one 64x32 output tile of one image, all channels.  The c and x*c tiles (+halo) of every
// input channel are staged in shared memory as separate planes; each thread owns a 4x4 output patch and
// keeps, for both output channels, the confidence sum D and the numerator Nm in registers as float2
// {co0,co1} pairs, so one packed FFMA2 (sm_100 fma.rn.f32x2) does two FMAs: the input scalar is the
// broadcast operand, the weight pair {w[co0],w[co1]} comes from constant memory through a uniform register.
// Epilogue: y = Nm*rcp(D+eps)+b (SFU reciprocal, <=1 ulp; D+eps >= 1e-20 is never denormal) and cout = D*(1/s)
// with 1/s computed once per call by IEEE division -- monotone in D, so the pooled argmax downstream sees the
// same ordering as D/s; coalesced float4 stores.
// A source whose (H,W) differs from the layer size is read through ATen's nearest-neighbour map, which
// folds F.interpolate + torch.cat (model_zoo/nconv_sd.py:81-94) into the tile loader.
#pragma once
#include "adc_common.cuh"

namespace adc {

template <int CI, int K>
struct FwdCfg {
    static constexpr int P = K / 2;
    static constexpr int TW = 64, TH = 32, NT = 128;
    static constexpr int SW = TW + 2 * P;
    static constexpr int PITCH = (SW + 3) & ~3;
    static constexpr int SH = TH + 2 * P;
    static constexpr int PLANE = PITCH * SH;
    static constexpr int SMEM_BYTES = CI * 2 * PLANE * 4;
};

template <int P>
__device__ __forceinline__ void sts4(float* dst, float4 v) {
    if constexpr (P % 4 == 0) {
        *reinterpret_cast<float4*>(dst) = v;
    } else if constexpr (P % 2 == 0) {
        *reinterpret_cast<float2*>(dst) = make_float2(v.x, v.y);
        *reinterpret_cast<float2*>(dst + 2) = make_float2(v.z, v.w);
    } else {
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
}

// Stage planes {c, x*c} of one source channel into shared memory (zero outside the image = conv zero padding).
template <int K, int CI>
__device__ __forceinline__ void load_channel_tile(const Src& s, int sc, int n, int h0, int w0, int H, int W,
                                                  float* smc, float* smx, int tid) {
    using C = FwdCfg<CI, K>;
    constexpr int P = C::P;
    const float* xp = s.x + (long long)n * s.bs + (long long)sc * s.H * s.W;
    const float* cp = s.c + (long long)n * s.bs + (long long)sc * s.H * s.W;
    const bool vec = !s.up && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(cp) & 15) == 0);
    if (vec) {
        constexpr int Q = C::TW / 4;
        for (int i = tid; i < C::SH * Q; i += C::NT) {
            const int r = i / Q, q = i % Q;
            const int gh = h0 - P + r, gw = w0 + q * 4;
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), xv = cv;
            if (gh >= 0 && gh < H && gw < W) {
                cv = __ldg(reinterpret_cast<const float4*>(cp + (long long)gh * W + gw));
                xv = __ldg(reinterpret_cast<const float4*>(xp + (long long)gh * W + gw));
            }
            sts4<P>(smc + r * C::PITCH + P + q * 4, cv);
            sts4<P>(smx + r * C::PITCH + P + q * 4, make_float4(xv.x * cv.x, xv.y * cv.y, xv.z * cv.z, xv.w * cv.w));
        }
        if constexpr (P > 0) {
            for (int i = tid; i < C::SH * 2 * P; i += C::NT) {
                const int r = i / (2 * P), j = i % (2 * P);
                const int col = j < P ? j : C::TW + j;
                const int gh = h0 - P + r, gw = w0 - P + col;
                float cv = 0.f, xv = 0.f;
                if (gh >= 0 && gh < H && gw >= 0 && gw < W) {
                    cv = __ldg(cp + (long long)gh * W + gw);
                    xv = __ldg(xp + (long long)gh * W + gw);
                }
                smc[r * C::PITCH + col] = cv;
                smx[r * C::PITCH + col] = xv * cv;
            }
        }
    } else {
        for (int i = tid; i < C::SH * C::SW; i += C::NT) {
            const int r = i / C::SW, col = i % C::SW;
            const int gh = h0 - P + r, gw = w0 - P + col;
            float cv = 0.f, xv = 0.f;
            if (gh >= 0 && gh < H && gw >= 0 && gw < W) {
                const long long off = (long long)s.mh(gh) * s.W + s.mw(gw);
                cv = __ldg(cp + off);
                xv = __ldg(xp + off);
            }
            smc[r * C::PITCH + col] = cv;
            smx[r * C::PITCH + col] = xv * cv;
        }
    }
}

// The shared compute core: for a thread's 4x4 patch, accumulate two correlations at once --
// accA += sum_{n,u,v} planeA[n] * w2[n,u,v] and accB += sum planeB[n] * w2[n,u,v] -- where each accumulator
// and each weight is a float2 over a PAIR of output channels (packed FFMA2) and the plane values are the
// broadcast scalar operand.  Forward: planes {c, x*c}, pairs {co0,co1}.  Data-gradient: planes {gD, gNm},
// flipped weights, pairs {ci0,ci1}.  Plane n lives at smem + (2n)*PLANE (A) and (2n+1)*PLANE (B).
template <int NIN, int K>
__device__ __forceinline__ void conv_patch(const float* smem, const float2* __restrict__ wp, int tx, int ty,
                                           float2 (&accA)[4][4], float2 (&accB)[4][4]) {
    using C = FwdCfg<NIN, K>;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int p = 0; p < 4; ++p) accA[r][p] = accB[r][p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ci = 0; ci < NIN; ++ci) {
        const float* sa = smem + (ci * 2) * C::PLANE + (ty * 4) * C::PITCH + tx * 4;
        const float* sb = sa + C::PLANE;
#pragma unroll
        for (int ir = 0; ir < 4 + K - 1; ++ir) {
            float aw[8], bw[8];
            {
                const float4 a = *reinterpret_cast<const float4*>(sa + ir * C::PITCH);
                const float4 b = *reinterpret_cast<const float4*>(sb + ir * C::PITCH);
                aw[0] = a.x; aw[1] = a.y; aw[2] = a.z; aw[3] = a.w;
                bw[0] = b.x; bw[1] = b.y; bw[2] = b.z; bw[3] = b.w;
                if constexpr (K == 5) {
                    const float4 a2 = *reinterpret_cast<const float4*>(sa + ir * C::PITCH + 4);
                    const float4 b2 = *reinterpret_cast<const float4*>(sb + ir * C::PITCH + 4);
                    aw[4] = a2.x; aw[5] = a2.y; aw[6] = a2.z; aw[7] = a2.w;
                    bw[4] = b2.x; bw[5] = b2.y; bw[6] = b2.z; bw[7] = b2.w;
                } else if constexpr (K == 3) {
                    const float2 a2 = *reinterpret_cast<const float2*>(sa + ir * C::PITCH + 4);
                    const float2 b2 = *reinterpret_cast<const float2*>(sb + ir * C::PITCH + 4);
                    aw[4] = a2.x; aw[5] = a2.y;
                    bw[4] = b2.x; bw[5] = b2.y;
                }
            }
#pragma unroll
            for (int orow = 0; orow < 4; ++orow) {
                const int u = ir - orow;
                if (u >= 0 && u < K) {
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        const float2 w2 = wp[(ci * K + u) * K + v];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            accA[orow][p] = ffma2(bcast(aw[p + v]), w2, accA[orow][p]);
                            accB[orow][p] = ffma2(bcast(bw[p + v]), w2, accB[orow][p]);
                        }
                    }
                }
            }
        }
    }
}

// CO is fixed to 2 here (float2 output-channel pairs); the 1x1 2->1 head has its own kernel below.
template <int CI, int K>
__global__ void __launch_bounds__(128) nconv_fwd_kernel(Src s0, Src s1, float* __restrict__ y, float* __restrict__ cout,
                                                        int H, int W, int wimg) {
    using C = FwdCfg<CI, K>;
    constexpr int P = C::P;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;

#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
        const bool first = ci < s0.ch;
        load_channel_tile<K, CI>(first ? s0 : s1, first ? ci : ci - s0.ch, n, h0, w0, H, W,
                                 smem + (ci * 2) * C::PLANE, smem + (ci * 2 + 1) * C::PLANE, tid);
    }
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;
    const int oh = h0 + ty * 4, ow = w0 + tx * 4;
    if (oh >= H || ow >= W) return;

    float2 accD[4][4], accN[4][4];
    conv_patch<CI, K>(smem, reinterpret_cast<const float2*>(wc), tx, ty, accD, accN);

    const float b0 = wc[kWBias], b1 = wc[kWBias + 1];
    const float is0 = wc[kWInvSum], is1 = wc[kWInvSum + 1];
    const long long plane = (long long)H * W;
    float* y0 = y + ((long long)n * 2) * plane;
    float* c0 = cout + ((long long)n * 2) * plane;
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(cout) & 15) == 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int h = oh + r;
        if (h >= H) break;
        float yv0[4], yv1[4], cv0[4], cv1[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            yv0[p] = fmaf(accN[r][p].x, rcp_approx(accD[r][p].x + kEps), b0);
            yv1[p] = fmaf(accN[r][p].y, rcp_approx(accD[r][p].y + kEps), b1);
            cv0[p] = accD[r][p].x * is0;
            cv1[p] = accD[r][p].y * is1;
        }
        const long long o = (long long)h * W + ow;
        if (vec) {  // W%4==0 and ow%4==0 => the whole float4 is inside the row
            *reinterpret_cast<float4*>(y0 + o) = make_float4(yv0[0], yv0[1], yv0[2], yv0[3]);
            *reinterpret_cast<float4*>(y0 + plane + o) = make_float4(yv1[0], yv1[1], yv1[2], yv1[3]);
            *reinterpret_cast<float4*>(c0 + o) = make_float4(cv0[0], cv0[1], cv0[2], cv0[3]);
            *reinterpret_cast<float4*>(c0 + plane + o) = make_float4(cv1[0], cv1[1], cv1[2], cv1[3]);
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (ow + p < W) {
                    y0[o + p] = yv0[p];
                    y0[plane + o + p] = yv1[p];
                    c0[o + p] = cv0[p];
                    c0[plane + o + p] = cv1[p];
                }
            }
        }
    }
}

// nconv7: 2 -> 1, 1x1 (model_zoo/nconv_sd.py:32).  Purely HBM-bound (AI 0.3): one float4 per thread per plane.
__global__ void __launch_bounds__(256) nconv_1x1_c2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ c,
                                                               long long bs, float* __restrict__ y,
                                                               float* __restrict__ cout, long long plane, int wimg) {
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    const float w0 = wc[0], w1 = wc[1], b = wc[kWBias], is = wc[kWInvSum];
    const int n = blockIdx.y;
    const float* x0 = x + (long long)n * bs;
    const float* c0 = c + (long long)n * bs;
    float* yo = y + (long long)n * plane;
    float* co = cout + (long long)n * plane;
    const bool vec = (plane % 4 == 0) && (bs % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(y) |
                        reinterpret_cast<uintptr_t>(cout)) & 15) == 0);
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= plane) return;
    float xa[4], xb[4], ca[4], cb[4], yv[4], cv[4];
    if (vec) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(x0 + i4));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(x0 + plane + i4));
        const float4 t2 = __ldg(reinterpret_cast<const float4*>(c0 + i4));
        const float4 t3 = __ldg(reinterpret_cast<const float4*>(c0 + plane + i4));
        xa[0] = t0.x; xa[1] = t0.y; xa[2] = t0.z; xa[3] = t0.w;
        xb[0] = t1.x; xb[1] = t1.y; xb[2] = t1.z; xb[3] = t1.w;
        ca[0] = t2.x; ca[1] = t2.y; ca[2] = t2.z; ca[3] = t2.w;
        cb[0] = t3.x; cb[1] = t3.y; cb[2] = t3.z; cb[3] = t3.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = i4 + j < plane;
            xa[j] = ok ? __ldg(x0 + i4 + j) : 0.f;
            xb[j] = ok ? __ldg(x0 + plane + i4 + j) : 0.f;
            ca[j] = ok ? __ldg(c0 + i4 + j) : 0.f;
            cb[j] = ok ? __ldg(c0 + plane + i4 + j) : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float D = fmaf(w1, cb[j], w0 * ca[j]);
        const float Nm = fmaf(w1, xb[j] * cb[j], w0 * (xa[j] * ca[j]));
        yv[j] = fmaf(Nm, rcp_approx(D + kEps), b);
        cv[j] = D * is;
    }
    if (vec) {
        *reinterpret_cast<float4*>(yo + i4) = make_float4(yv[0], yv[1], yv[2], yv[3]);
        *reinterpret_cast<float4*>(co + i4) = make_float4(cv[0], cv[1], cv[2], cv[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i4 + j < plane) {
                yo[i4 + j] = yv[j];
                co[i4 + j] = cv[j];
            }
    }
}

template <int CI, int K>
int launch_nconv_fwd(const Src& s0, const Src& s1, float* y, float* cout, int N, int H, int W, int wimg, cudaStream_t st) {
    using C = FwdCfg<CI, K>;
    static bool attr_done = false;
    if (!attr_done) {
        ADC_CUDA(cudaFuncSetAttribute(nconv_fwd_kernel<CI, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    dim3 grid(ceil_div(W, C::TW), ceil_div(H, C::TH), N);
    nconv_fwd_kernel<CI, K><<<grid, C::NT, C::SMEM_BYTES, st>>>(s0, s1, y, cout, H, W, wimg);
    return check_cuda(cudaGetLastError(), "nconv_fwd_kernel launch");
}

}  // namespace adc

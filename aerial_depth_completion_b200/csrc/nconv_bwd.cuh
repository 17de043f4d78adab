// K2: NConv2d backward for sm_100a, as the reference's autograd graph computes it, on pair tensors.
//
// The gradient that reaches a layer is kept in the transformed form G = {GN, GD} per output pixel and channel,
//   GN = gy/(D+eps)      GD = -gy*(y-b)/(D+eps) + gcout/s      (D = cout*s, s = sum(w_fwd)),
// produced by whichever kernel FINALISES the raw gradient (gy, gcout) of that layer's output (grad_xform in
// adc_common.cuh): the data-gradient epilogue of the next layer, the pooling backward, the upsample reduction or the
// 1x1 head.  The same kernel accumulates the layer's bias gradient gb = sum gy and gs = -sum gcout*cout/s.  With G in
// hand both backward kernels of a layer are pure TMA-tile + FFMA2 loops, like the forward:
//  * nconv_bwd_data_kernel   g_xc[i] = corr^T(GN, W*), g_c'[i] = corr^T(GD, W*) with the forward's conv_patch core on
//    the planes {GN,GD}, flipped backward-time weights W*; epilogue gx = g_xc*c, gc = g_xc*x + g_c' (optionally
//    transformed for the previous layer), STG.128.
//  * nconv_bwd_weight_kernel gW[o,i,u,v] = sum_pix GN[o]*xc[i](shift) + GD[o]*c[i](shift): every warp owns one (o,i)
//    slice and one 16-row group; a lane streams a 2-pixel-wide column strip down the tile keeping a K-row register
//    window of G pairs, so one {xc,c} row segment read from shared memory feeds K*K*2 packed FFMA2; accumulators are
//    reduced with warp shuffles and written as per-CTA partial rows (deterministic two-pass reduction, no atomics).
#pragma once
#include "adc_common.cuh"
#include "nconv_fwd.cuh"

namespace adc {

// 4 CTAs per SM (<= 128 registers) for the single-pass instantiations: measured 39.3 -> 37.6 us for <2,5> at L0 (28 bytes of spills);
// the two-pass <4,3> body spills 444 bytes under that cap and slows down (47.5 -> 64 us), so it keeps 3 CTAs per SM.
template <int CI, int K>
__global__ void __launch_bounds__(128, (CI > 2 ? 3 : 4)) nconv_bwd_data_kernel(const __grid_constant__ Maps2 maps, SrcP g, SrcP s0, SrcP s1, int has1,
                                                             float2* __restrict__ out0, float2* __restrict__ out1,
                                                             float* __restrict__ bpart, int H, int W, int wimg, int ximg) {
    using C = TileCfg<K>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2* planes = align128(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(planes + 2 * C::PLANE);
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;

    // TMA source: one box and one mbarrier per G plane, so the FFMA2 loop over plane 0 starts while plane 1 is in flight (as in the
    // forward kernel); other sources: staged together.
    const bool per_plane = g.tma != 0;
    if (per_plane) {
        if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
        __syncthreads();
        if (tid == 0) {
            const int x0 = 2 * (w0 - C::P - C::XOFF), y0 = h0 - C::P;
            for (int c = 0; c < 2; ++c) {
                mbar_expect_tx(bar + c, (uint32_t)(C::SH * C::SWP * 8));
                tma_load_plane(planes + c * C::PLANE, &maps.m0, bar + c, x0, y0, c, n);
            }
        }
    } else {
        const bool pending = stage_sources<K, 128>(maps, g, g, false, n, h0, w0, H, W, planes, bar, tid);
        __syncthreads();
        if (pending) mbar_wait(bar, 0);
    }

    const int tx = tid & 31, ty = tid >> 5;
    const int oh = h0 + ty * 8, ow = w0 + tx * 2;
    const bool active = oh < H && ow < W;
    const long long plane = (long long)H * W;
    float gb[2] = {0.f, 0.f}, gs[2] = {0.f, 0.f};

    if (active) {
        constexpr int NQ = (CI + 1) / 2, NO = CI == 1 ? 1 : 2;
        // The epilogue multiplies the accumulators with the forward {x, c} of the same pixels, read straight from global memory
        // AFTER the FFMA2 loop (no registers to hold them across it at the 128-register cap): request the lines into L2 now so
        // those loads see L2 latency instead of HBM latency.
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
            const bool first = ci < s0.ch;
            const SrcP& s = first ? s0 : s1;
            if (s.up || (W & 1)) continue;
            const float2* xp = s.p + (long long)n * s.bs + (long long)(first ? ci : ci - s0.ch) * s.H * s.W + ow;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int h = oh + r < H ? oh + r : H - 1;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + (long long)h * W));
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float2 acc[NO][8][2];
            if (per_plane) {
#pragma unroll
                for (int no = 0; no < NO; ++no)
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[no][r][0] = acc[no][r][1] = make_float2(0.f, 0.f);
#pragma unroll 1
                for (int ni = 0; ni < 2; ++ni) {
                    if (q == 0) mbar_wait(bar + ni, 0);
                    conv_patch_plane<2, NO, K>(planes + ni * C::PLANE, wc + q * (2 * 2 * K * K), ni, tx, ty, acc);
                }
            } else {
                conv_patch<2, NO, K>(planes, wc + q * (2 * 2 * K * K), tx, ty, acc);
            }
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                const int ci = 2 * q + j;
                const bool first = ci < s0.ch;
                const SrcP& s = first ? s0 : s1;
                const int sc = first ? ci : ci - s0.ch;
                const float2* xp = s.p + (long long)n * s.bs + (long long)sc * s.H * s.W;
                float2* go = (first ? out0 : out1) + ((long long)n * s.ch + sc) * plane + ow;
                float xb = 0.f, xs = 1.f, xis = 1.f;
                if (ximg >= 0) {
                    const float* xc = &c_wimg[0][0][0][0] + (size_t)ximg * kWStride;
                    xb = xc[kWBias + ci]; xs = xc[kWSum + ci]; xis = xc[kWInvSum + ci];
                }
                const bool vec = !s.up && (W & 1) == 0;
                // all eight rows of {x, c} are requested before the first one is consumed (the loop was latency-bound:
                // long_scoreboard 7-13 stall cycles per issue with load -> use -> store per row)
                float2 a0[8], a1[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int h = oh + r < H ? oh + r : H - 1;
                    if (vec) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(xp + (long long)h * W + ow));
                        a0[r] = make_float2(t.x, t.y); a1[r] = make_float2(t.z, t.w);
                    } else {
                        const long long ro = (long long)s.mh(h) * s.W;
                        a0[r] = __ldg(xp + ro + s.mw(ow));
                        a1[r] = __ldg(xp + ro + s.mw(ow + 1 < W ? ow + 1 : ow));
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int h = oh + r;
                    if (h >= H) break;
                    float2 r0 = make_float2(acc[j][r][0].x * a0[r].y, fmaf(acc[j][r][0].x, a0[r].x, acc[j][r][0].y));
                    float2 r1 = make_float2(acc[j][r][1].x * a1[r].y, fmaf(acc[j][r][1].x, a1[r].x, acc[j][r][1].y));
                    if (ximg >= 0) {
                        r0 = grad_xform(r0.x, r0.y, a0[r].x, a0[r].y, xb, xs, xis, gb[j], gs[j]);
                        if (ow + 1 < W) r1 = grad_xform(r1.x, r1.y, a1[r].x, a1[r].y, xb, xs, xis, gb[j], gs[j]);
                    }
                    if ((W & 1) == 0) {
                        *reinterpret_cast<float4*>(go + (long long)h * W) = make_float4(r0.x, r0.y, r1.x, r1.y);
                    } else {
                        go[(long long)h * W] = r0;
                        if (ow + 1 < W) go[(long long)h * W + 1] = r1;
                    }
                }
            }
        }
    }
    if (ximg >= 0) {   // uniform: bias / gs partials of the previous layer (single-source layers only, CI == 2)
        const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        block_sum4_store<128>(gb[0], gb[1], gs[0], gs[1], bpart + blk * 4);
    }
}

template <int CI, int K>
struct WgCfg {
    // 16-row tiles (one row group): 38 KB (2->2, 5x5) / 56 KB (4->2, 3x3) of shared memory per CTA instead of 72 / 106 KB, so 5 / 4 CTAs
    // share an SM and the TMA phase of one overlaps the FFMA2 phase of the others (with 2 CTAs per SM load time ~ compute time and
    // the FMA pipe sat at ~50 %)
    static constexpr int TH = 16;
    using T = TileCfg<K, TH>;
    static constexpr int P = K / 2;
    static constexpr int TW = 64, RG = 1, RH = TH / RG;
    static constexpr int ROLES = 2 * CI;                           // (o,i), Co = 2
    static constexpr int NT = 32 * ROLES * RG;
    static constexpr int STEPS = RH + K - 1;                       // 20 (K=5), 18 (K=3): multiples of K
    static constexpr int GPLANE = TH * TW;                         // pairs; 16 KB, a multiple of 128 B
    static constexpr int smem_bytes() { return CI * T::PLANE_BYTES + 2 * GPLANE * 8 + 128 + 16; }
    static_assert(STEPS % K == 0, "row streaming is unrolled by K");
};

template <int CI, int K>
__global__ void __launch_bounds__(WgCfg<CI, K>::NT) nconv_bwd_weight_kernel(const __grid_constant__ Maps2 maps,
                                                                            const __grid_constant__ CUtensorMap gmap, SrcP g,
                                                                            SrcP s0, SrcP s1, int has1,
                                                                            float* __restrict__ partial, int H, int W) {
    using C = WgCfg<CI, K>;
    using T = typename C::T;
    constexpr int P = C::P, NT = C::NT;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2* sX = align128(smem_raw);                 // [CI] planes of {x, c} with halo
    float2* sG = sX + CI * T::PLANE;                 // [2][TH][TW] {GN, GD}
    uint64_t* bar = reinterpret_cast<uint64_t*>(sG + 2 * C::GPLANE);
    uint64_t* bar2 = bar + 1;
    __shared__ float sred[C::RG][C::ROLES][K * K];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;

    // G interior tile
    if (g.tma) {
        if (tid == 0) mbar_init(bar2, 1);
    }
    const bool pending = stage_sources<K, NT, C::TH>(maps, s0, s1, has1 != 0, n, h0, w0, H, W, sX, bar, tid);
    if (g.tma) {
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar2, 2 * C::GPLANE * 8);
            tma_load_plane(sG, &gmap, bar2, 2 * w0, h0, 0, n);
            tma_load_plane(sG + C::GPLANE, &gmap, bar2, 2 * w0, h0, 1, n);
        }
    } else {
        for (int o = 0; o < 2; ++o) {
            const float2* base = g.p + (long long)n * g.bs + (long long)o * H * W;
            for (int i = tid; i < C::GPLANE; i += NT) {
                const int r = i / C::TW, col = i - r * C::TW;
                const int gh = h0 + r, gw = w0 + col;
                sG[o * C::GPLANE + i] = (gh < H && gw < W) ? __ldg(base + (long long)gh * W + gw) : make_float2(0.f, 0.f);
            }
        }
    }
    __syncthreads();
    if (pending) mbar_wait(bar, 0);
    expand_up2_planes<K, NT, C::TH>(s0, s1, has1 != 0, sX, tid);
    premul_planes<K, NT, C::TH>(sX, CI, tid);      // {x, c} -> {x*c, c}
    if (g.tma) mbar_wait(bar2, 0);
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int role = warp % C::ROLES, grp = warp / C::ROLES;
    const int o = role / CI, ci = role % CI;
    const float2* gptr = sG + o * C::GPLANE + (grp * C::RH) * C::TW + lane * 2;
    const float2* xptr = sX + ci * T::PLANE + (grp * C::RH) * T::SWP + lane * 2;

    float2 acc[K][K];
    float2 win[K][2];
#pragma unroll
    for (int u = 0; u < K; ++u) {
#pragma unroll
        for (int v = 0; v < K; ++v) acc[u][v] = make_float2(0.f, 0.f);
        win[u][0] = win[u][1] = make_float2(0.f, 0.f);
    }

    // step j reads X row (grp*RH + j) [tile row grp*RH + j - P] and G row grp*RH + j; tap row u of step j pairs that X row with
    // G row j - u, which exists for 0 <= j - u < RH.  The first and the last block of K steps are peeled so the (step, tap row)
    // pairs without a G row cost nothing (the uniform loop multiplied zero rows there: 20 % of the FFMA2 at K = 5, 11 % at K = 3).
    constexpr int XO = T::XOFF;
    constexpr int NP = (2 + K - 1 + XO + 1) & ~1;
    constexpr int NB = C::STEPS / K;
    static_assert(C::RH >= 2 * K - 1 && NB >= 2, "peeled first / last blocks must not overlap");
    auto block = [&](auto mode_c, int jt) {            // mode 0: first block (j < K), 1: middle, 2: last (j >= RH - 1)
        constexpr int MODE = decltype(mode_c)::value;
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const int j = jt * K + jj;
            if (MODE != 2 || jj == 0) {                // the last block starts at j = RH - 1, the last G row
                const float4 a = *reinterpret_cast<const float4*>(gptr + j * C::TW);
                win[jj][0] = make_float2(a.x, a.y);
                win[jj][1] = make_float2(a.z, a.w);
            }
            float2 xr[NP];
#pragma unroll
            for (int t = 0; t < NP / 2; ++t) {
                const float4 a = *reinterpret_cast<const float4*>(xptr + j * T::SWP + 2 * t);
                xr[2 * t] = make_float2(a.x, a.y);
                xr[2 * t + 1] = make_float2(a.z, a.w);
            }
#pragma unroll
            for (int u = 0; u < K; ++u) {
                if ((MODE == 0 && u > jj) || (MODE == 2 && u < jj)) continue;
                const int slot = ((jj - u) % K + K) % K;   // G row of step j-u
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    acc[u][v] = ffma2(win[slot][0], xr[XO + v], acc[u][v]);
                    acc[u][v] = ffma2(win[slot][1], xr[XO + v + 1], acc[u][v]);
                }
            }
        }
    };
    block(std::integral_constant<int, 0>{}, 0);
#pragma unroll 1
    for (int jt = 1; jt < NB - 1; ++jt) block(std::integral_constant<int, 1>{}, jt);
    block(std::integral_constant<int, 2>{}, NB - 1);

#pragma unroll
    for (int u = 0; u < K; ++u)
#pragma unroll
        for (int v = 0; v < K; ++v) {
            const float r = warp_sum(acc[u][v].x + acc[u][v].y);
            if (lane == 0) sred[grp][role][u * K + v] = r;
        }
    __syncthreads();
    float* out = partial + (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kWStride;
    for (int i = tid; i < C::ROLES * K * K; i += NT) {
        const int r = i / (K * K), t = i - r * (K * K);
        float v = 0.f;
#pragma unroll
        for (int gq = 0; gq < C::RG; ++gq) v += sred[gq][r][t];
        out[i] = v;    // role-major (o,i) then tap = the Parameter's own [Co][Ci][K][K] order
    }
}

// nconv7 (2 -> 1, 1x1) backward.  Takes the RAW planar gradients of (xout, cout) from autograd, transforms them on the
// fly, accumulates the head's own weight/bias/gs gradients and emits the gradient of its input A9 -- transformed for
// the layer that produced A9 when ximg >= 0 (pyramid), raw otherwise.  Two pixels per thread, 16-byte accesses.
__global__ void __launch_bounds__(256) nconv_1x1_bwd_kernel(const float2* __restrict__ in, const float* __restrict__ y,
                                                            const float* __restrict__ cout, const float* __restrict__ gy,
                                                            const float* __restrict__ gcout, float2* __restrict__ gout,
                                                            float* __restrict__ wpart, float* __restrict__ bpart_self,
                                                            float* __restrict__ bpart_prev, long long plane, int wimg_f,
                                                            int wimg_b, int ximg) {
    const float* wf = &c_wimg[0][0][0][0] + (size_t)wimg_f * kWStride;
    const float* wb = &c_wimg[0][0][0][0] + (size_t)wimg_b * kWStride;
    const float b = wf[kWBias], s = wf[kWSum], is = wf[kWInvSum];
    const float w0 = wb[0], w1 = wb[1];
    float xb[2] = {0.f, 0.f}, xs[2] = {1.f, 1.f}, xis[2] = {1.f, 1.f};
    if (ximg >= 0) {
        const float* xc = &c_wimg[0][0][0][0] + (size_t)ximg * kWStride;
#pragma unroll
        for (int j = 0; j < 2; ++j) { xb[j] = xc[kWBias + j]; xs[j] = xc[kWSum + j]; xis[j] = xc[kWInvSum + j]; }
    }
    const int n = blockIdx.y;
    const float2* p0 = in + (long long)n * 2 * plane;
    const float2* p1 = p0 + plane;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    float a_w0 = 0.f, a_w1 = 0.f, a_b = 0.f, a_s = 0.f, pb[2] = {0.f, 0.f}, ps[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const long long i = i0 + j;
        if (i >= plane) break;
        const long long o = (long long)n * plane + i;
        const float2 G = grad_xform(__ldg(gy + o), __ldg(gcout + o), __ldg(y + o), __ldg(cout + o), b, s, is, a_b, a_s);
        const float2 A = __ldg(p0 + i), B = __ldg(p1 + i);
        a_w0 += fmaf(G.x, A.x * A.y, G.y * A.y);
        a_w1 += fmaf(G.x, B.x * B.y, G.y * B.y);
        if (gout) {
            const float gxc0 = G.x * w0, gxc1 = G.x * w1;
            float2 r0 = make_float2(gxc0 * A.y, fmaf(gxc0, A.x, G.y * w0));
            float2 r1 = make_float2(gxc1 * B.y, fmaf(gxc1, B.x, G.y * w1));
            if (ximg >= 0) {
                r0 = grad_xform(r0.x, r0.y, A.x, A.y, xb[0], xs[0], xis[0], pb[0], ps[0]);
                r1 = grad_xform(r1.x, r1.y, B.x, B.y, xb[1], xs[1], xis[1], pb[1], ps[1]);
            }
            gout[((long long)n * 2) * plane + i] = r0;
            gout[((long long)n * 2 + 1) * plane + i] = r1;
        }
    }
    const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    block_sum4_store<256>(a_w0, a_w1, 0.f, 0.f, wpart + blk * kWStride);
    __syncthreads();
    block_sum4_store<256>(a_b, 0.f, a_s, 0.f, bpart_self + blk * 4);
    if (ximg >= 0) {
        __syncthreads();
        block_sum4_store<256>(pb[0], pb[1], ps[0], ps[1], bpart_prev + blk * 4);
    }
}

// Single-layer API entry: raw planar (gy, gcout) + planar forward outputs (y, cout) -> G pair tensor + bias/gs partials.
template <int CO>
__global__ void __launch_bounds__(256) grad_transform_kernel(const float* __restrict__ gy, const float* __restrict__ gcout,
                                                             const float* __restrict__ y, const float* __restrict__ cout,
                                                             float2* __restrict__ G, float* __restrict__ bpart, long long plane,
                                                             int wimg) {
    const float* wf = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    const int n = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float gb[2] = {0.f, 0.f}, gs[2] = {0.f, 0.f};
    if (i < plane) {
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            const long long off = ((long long)n * CO + o) * plane + i;
            G[off] = grad_xform(__ldg(gy + off), __ldg(gcout + off), __ldg(y + off), __ldg(cout + off), wf[kWBias + o],
                                wf[kWSum + o], wf[kWInvSum + o], gb[o], gs[o]);
        }
    }
    block_sum4_store<256>(gb[0], gb[1], gs[0], gs[1], bpart + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4);
}

// Second pass of the weight gradient: sum the per-CTA partial rows of every call of a module, add gs, write/accumulate.
struct WgReduceCall {
    const float* wpart;   // [wblocks][kWStride], entries [0, n)
    const float* bpart;   // [bblocks][4] = {gb0, gb1, gs0, gs1}
    int wblocks, bblocks;
    int module;           // which output the call accumulates into
    int Co, n;            // n = Co*Ci*K*K
};
struct WgReduceArgs {
    WgReduceCall call[ADC_PYR_CALLS];
    int ncalls;
    float* gw[ADC_PYR_MODULES];
    float* gb[ADC_PYR_MODULES];
    int accumulate;
};

constexpr int kRedChunks = 64;   // first-stage blocks per call (16 left the four full-resolution calls of the pyramid on 64 CTAs)

// Stage 1: grid (ncalls, kRedChunks) x 1024 threads = 8 row slices x 128 entries.  Each thread sums its rows with four
// independent accumulators so the loads overlap (a single serial loop was DRAM-latency bound: 116 us for 1280 rows).
__global__ void __launch_bounds__(1024) wgrad_reduce1_kernel(WgReduceArgs a, float* __restrict__ stage) {
    const WgReduceCall& c = a.call[blockIdx.x];
    const int t = threadIdx.x & 127, slice = threadIdx.x >> 7;
    __shared__ float sh[8][kWStride];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int first = blockIdx.y * 8 + slice, step = kRedChunks * 8;
    if (t < c.n) {
        const float* p = c.wpart + t;
        int r = first;
        for (; r + 3 * step < c.wblocks; r += 4 * step) {
            acc0 += p[(size_t)r * kWStride];
            acc1 += p[(size_t)(r + step) * kWStride];
            acc2 += p[(size_t)(r + 2 * step) * kWStride];
            acc3 += p[(size_t)(r + 3 * step) * kWStride];
        }
        for (; r < c.wblocks; r += step) acc0 += p[(size_t)r * kWStride];
    } else if (t >= kWBias && t < kWBias + 4) {
        const float* p = c.bpart + (t - kWBias);
        int r = first;
        for (; r + 3 * step < c.bblocks; r += 4 * step) {
            acc0 += p[(size_t)r * 4];
            acc1 += p[(size_t)(r + step) * 4];
            acc2 += p[(size_t)(r + 2 * step) * 4];
            acc3 += p[(size_t)(r + 3 * step) * 4];
        }
        for (; r < c.bblocks; r += step) acc0 += p[(size_t)r * 4];
    }
    sh[slice][t] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (slice == 0) {
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) v += sh[s][t];
        stage[((size_t)blockIdx.x * kRedChunks + blockIdx.y) * kWStride + t] = v;
    }
}

// Stage 2: one block per module: sum the chunks of every call of the module, add gs, write / accumulate.
__global__ void __launch_bounds__(128) wgrad_reduce2_kernel(WgReduceArgs a, const float* __restrict__ stage) {
    const int m = blockIdx.x, t = threadIdx.x;
    __shared__ float sh[kWStride];
    float v = 0.f;
    int n = 0, Co = 1;
    for (int k = 0; k < a.ncalls; ++k) {
        if (a.call[k].module != m) continue;
        n = a.call[k].n;
        Co = a.call[k].Co;
#pragma unroll
        for (int ch = 0; ch < kRedChunks; ++ch) v += stage[((size_t)k * kRedChunks + ch) * kWStride + t];
    }
    sh[t] = v;
    __syncthreads();
    if (n > 0) {
        if (t < n) {
            const int co = t / (n / Co);
            const float r = sh[t] + sh[kWBias + 2 + co];     // + gs[co]
            a.gw[m][t] = a.accumulate ? a.gw[m][t] + r : r;
        }
        if (t < Co && a.gb[m]) {
            const float r = sh[kWBias + t];
            a.gb[m][t] = a.accumulate ? a.gb[m][t] + r : r;
        }
    }
}

}  // namespace adc

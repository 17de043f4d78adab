// K1: fused NConv2d forward (model_zoo/nconv.py:37-67) for sm_100a, on pair tensors (see adc_common.cuh).
//
// One CTA = one 64x32 output tile of one image, all channels.  The {x,c} tile (+halo) of every input channel is
// brought into shared memory by ONE TMA box copy per channel (cp.async.bulk.tensor, hardware zero fill outside the
// image = the convolution's zero padding; mbarrier completion), or -- for a source that is read through ATen's
// nearest-neighbour map, which folds F.interpolate + torch.cat (model_zoo/nconv_sd.py:81-94) into the loader, or whose
// row pitch TMA cannot describe -- by batched 8-byte gathers.  Each thread owns a 2-wide x 8-tall output patch:
// a warp then reads 512 contiguous bytes per LDS.128 (no bank conflicts) and keeps {Nm, D} of both output channels
// in 32 float2 accumulators; one packed FFMA2 per (tap, pixel, channel) with the weight as the uniform-register
// broadcast operand.  Epilogue: y = Nm*rcp(D+eps)+b (SFU reciprocal, <=1 ulp; D+eps >= 1e-20 is never denormal)
// and cout = D*(1/s), 1/s computed once per call by IEEE division -- monotone in D, so the pooled argmax downstream
// sees the ordering D/s would give; one coalesced STG.128 {y0,c0,y1,c1} per (channel,row) and thread.
#pragma once
#include "adc_common.cuh"
#include "tma.cuh"

namespace adc {

template <int K, int TH_ = 32>      // TH_: tile height (32 for the forward / data-gradient kernels, 16 for the weight gradient)
struct TileCfg {
    static constexpr int P = K / 2;
    static constexpr int TW = 64, TH = TH_, NT = 128;
    // A TMA box must start on a 16-byte boundary of the innermost dimension (measured: an odd pair offset raises
    // 'illegal instruction'), so for odd halos the tile keeps XOFF spare pair(s) on the left.
    static constexpr int XOFF = P & 1;
    static constexpr int SWP = (TW + 2 * P + 2 * XOFF + 3) & ~3;            // pairs per tile row (= TMA box width)
    static constexpr int SH = TH + 2 * P;
    static constexpr int PLANE_BYTES = ((SH * SWP * 8 + 127) / 128) * 128;  // TMA destinations are 128-B aligned
    static constexpr int PLANE = PLANE_BYTES / 8;                           // in pairs
    static constexpr int smem_bytes(int planes) { return planes * PLANE_BYTES + 128 + 64; }      // + alignment slack + up to 8 mbarriers
};

struct Maps2 {
    CUtensorMap m0, m1;
};

// 128-byte aligned start of the dynamic shared memory, computed as an OFFSET into the array so the pointer keeps its
// shared-memory address space (a reinterpret_cast through an integer degrades every access to a generic LD).
__device__ __forceinline__ float2* align128(uint8_t* smem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return reinterpret_cast<float2*>(smem + ((128u - (a & 127u)) & 127u));
}

// Batched gather of one channel tile (generic path: upsampled / unaligned sources).  Four independent 8-byte loads are
// issued per thread before any is consumed, so the loop is not serialised on DRAM latency.
template <int K, int NT, int TH_ = 32>
__device__ __forceinline__ void gather_pair_tile(const SrcP& s, int sc, int n, int h0, int w0, int H, int W, float2* dst,
                                                 int tid) {
    using C = TileCfg<K, TH_>;
    constexpr int P = C::P, TOTAL = C::SH * C::SWP, U = 4;
    const float2* base = s.p + (long long)n * s.bs + (long long)sc * s.H * s.W;
    for (int i0 = tid; i0 < TOTAL; i0 += NT * U) {
        float2 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int i = i0 + j * NT;
            const int r = i / C::SWP, col = i - r * C::SWP;
            const int gh = h0 - P + r, gw = w0 - P - C::XOFF + col;
            const bool ok = i < TOTAL && gh >= 0 && gh < H && gw >= 0 && gw < W;
            const long long off = ok ? (long long)s.mh(gh) * s.W + s.mw(gw) : 0;
            const float2 t = __ldg(base + off);
            v[j] = ok ? t : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int i = i0 + j * NT;
            if (i < TOTAL) dst[i] = v[j];
        }
    }
}

// Exact 2x nearest up-sampling through TMA: tile row r / pair column c of the layer reads low-resolution row
// ((h0-P+r) >> 1) and column ((w0-P-XOFF+c) >> 1).  h0 and w0-P-XOFF are even, so relative to the box origin
// (floor((h0-P)/2), (w0-P-XOFF)/2) that is row (r + (P&1)) >> 1, column c >> 1.  Rows/columns outside the low-resolution
// image are the hardware's zero fill and coincide with the layer's own zero padding because H = 2*Hs, W = 2*Ws.
template <int K, int TH_ = 32>
struct Up2Cfg {
    using C = TileCfg<K, TH_>;
    // The box must start on an even pair (16-byte) offset of the row like every TMA box here; (w0-P-XOFF)/2 = w0/2 - 1 is odd, so
    // the box starts one low-resolution pair further left (LXO) and is two pairs wider.
    static constexpr int LXO = 1;
    static constexpr int LW = C::SWP / 2 + 2;                              // pairs per low-resolution box row (16-byte multiple)
    static constexpr int LH = ((C::SH - 1 + (C::P & 1)) >> 1) + 1;         // low-resolution rows
    static constexpr int OFF = (C::PLANE - LW * LH) & ~15;                 // box lands in the plane's tail (128-byte aligned)
    static_assert((C::SWP & 3) == 0 && ((LW * 8) % 16) == 0 && OFF >= 0 && OFF + LW * LH <= C::PLANE, "TMA box rows are 16-byte multiples");
};

// Stage all input-channel tiles of up to two sources.  Returns true if a TMA transaction is pending on `bar`.
template <int K, int NT, int TH_ = 32>
__device__ __forceinline__ bool stage_sources(const Maps2& maps, const SrcP& s0, const SrcP& s1, bool has1, int n, int h0, int w0,
                                              int H, int W, float2* planes, uint64_t* bar, int tid) {
    using C = TileCfg<K, TH_>;
    using U = Up2Cfg<K, TH_>;
    const bool t0 = s0.tma || s0.up2, t1 = has1 && (s1.tma || s1.up2);
    const bool any_tma = t0 || t1;
    if (any_tma) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            const int full = C::SH * C::SWP * 8, low = U::LW * U::LH * 8;
            const int bytes = (s0.tma ? s0.ch * full : 0) + (s0.up2 ? s0.ch * low : 0) +
                              ((has1 && s1.tma) ? s1.ch * full : 0) + ((has1 && s1.up2) ? s1.ch * low : 0);
            mbar_expect_tx(bar, (uint32_t)bytes);
            const int x0 = 2 * (w0 - C::P - C::XOFF), y0 = h0 - C::P;
            if (s0.tma)
                for (int c = 0; c < s0.ch; ++c) tma_load_plane(planes + c * C::PLANE, &maps.m0, bar, x0, y0, c, n);
            if (s0.up2)      // low-resolution coordinates: pairs (w0-P-XOFF)/2 -> floats w0-P-XOFF; rows floor(y0/2)
                for (int c = 0; c < s0.ch; ++c) tma_load_plane(planes + c * C::PLANE + U::OFF, &maps.m0, bar, (x0 >> 1) - 2 * U::LXO, y0 >> 1, c, n);
            if (has1 && s1.tma)
                for (int c = 0; c < s1.ch; ++c) tma_load_plane(planes + (s0.ch + c) * C::PLANE, &maps.m1, bar, x0, y0, c, n);
            if (has1 && s1.up2)
                for (int c = 0; c < s1.ch; ++c)
                    tma_load_plane(planes + (s0.ch + c) * C::PLANE + U::OFF, &maps.m1, bar, (x0 >> 1) - 2 * U::LXO, y0 >> 1, c, n);
        }
    }
    if (!t0)
        for (int c = 0; c < s0.ch; ++c) gather_pair_tile<K, NT, TH_>(s0, c, n, h0, w0, H, W, planes + c * C::PLANE, tid);
    if (has1 && !t1)
        for (int c = 0; c < s1.ch; ++c) gather_pair_tile<K, NT, TH_>(s1, c, n, h0, w0, H, W, planes + (s0.ch + c) * C::PLANE, tid);
    return any_tma;
}

// In-place expansion of the up2 planes after their TMA boxes have landed: every thread first reads ALL its low-resolution
// pairs into registers, the CTA synchronises, then each pair is written twice (one 16-byte store) at its two full-resolution
// columns -- reads complete before any write, so the box may overlap the rows it expands into.  Must be called by all
// threads of the CTA (contains a barrier); no-op when no source is up2.
template <int K, int NT, int TH_ = 32>
__device__ __forceinline__ void expand_up2_planes(const SrcP& s0, const SrcP& s1, bool has1, float2* planes, int tid) {
    using C = TileCfg<K, TH_>;
    using U = Up2Cfg<K, TH_>;
    const int n0 = s0.up2 ? s0.ch : 0, n1 = (has1 && s1.up2) ? s1.ch : 0;
    if (n0 + n1 == 0) return;
    const int first = s0.up2 ? 0 : s0.ch;              // the up2 planes are contiguous: [first, first + nup)
    const int nup = n0 + n1;
    // (s0 only: planes [0, s0.ch); s1 only: planes [s0.ch, s0.ch + s1.ch); both: all of them)
    constexpr int LC = C::SWP / 2;                     // low-resolution columns a plane row expands from
    constexpr int ITEMS = C::SH * LC;                  // (row, low column) items per plane
    constexpr int PER = (ITEMS + NT - 1) / NT;
    for (int q = 0; q < nup; ++q) {
        float2* pl = planes + (first + q) * C::PLANE;
        float2 v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * NT;
            if (i < ITEMS) {
                const int r = i / LC, j = i - r * LC;
                v[k] = pl[U::OFF + ((r + (C::P & 1)) >> 1) * U::LW + U::LXO + j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * NT;
            if (i < ITEMS) {
                const int r = i / LC, j = i - r * LC;
                *reinterpret_cast<float4*>(pl + r * C::SWP + 2 * j) = make_float4(v[k].x, v[k].y, v[k].x, v[k].y);
            }
        }
    }
    __syncthreads();
}

// The shared compute core.  For a thread's 2x8 patch and NOUT output maps:
//   acc[no][row][px] += sum_{ni,u,v} in[ni](row+u, px+v) * w[(no*NIN+ni)*K*K + u*K + v]
// where `in` and `acc` are float2 pairs (two correlations per packed FFMA2) and w is a broadcast scalar from constant
// memory.  The forward feeds {x*c, c} (premul_planes below), the backward {GN, GD}.
template <int NIN, int NOUT, int K>
__device__ __forceinline__ void conv_patch(const float2* planes, const float* __restrict__ w, int tx, int ty,
                                           float2 (&acc)[NOUT][8][2]) {
    using C = TileCfg<K>;
    constexpr int XO = C::XOFF;
    constexpr int NP = (2 + K - 1 + XO + 1) & ~1;   // pairs loaded per window row, from an even (16-byte aligned) offset
#pragma unroll
    for (int no = 0; no < NOUT; ++no)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[no][r][0] = acc[no][r][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ni = 0; ni < NIN; ++ni) {
        const float2* row0 = planes + ni * C::PLANE + (ty * 8) * C::SWP + tx * 2;
        static_for<8 + K - 1>([&](auto ir_c) {
            constexpr int ir = decltype(ir_c)::value;
            float2 in[NP];
#pragma unroll
            for (int t = 0; t < NP / 2; ++t) {
                const float4 a = *reinterpret_cast<const float4*>(row0 + ir * C::SWP + 2 * t);
                in[2 * t] = make_float2(a.x, a.y);
                in[2 * t + 1] = make_float2(a.z, a.w);
            }
            static_for<8>([&](auto orow_c) {
                constexpr int orow = decltype(orow_c)::value;
                constexpr int u = ir - orow;
                if constexpr (u >= 0 && u < K) {
#pragma unroll
                    for (int v = 0; v < K; ++v)
#pragma unroll
                        for (int no = 0; no < NOUT; ++no) {
                            const float wv = w[(no * NIN + ni) * K * K + u * K + v];
                            acc[no][orow][0] = ffma2(in[XO + v], bcast(wv), acc[no][orow][0]);
                            acc[no][orow][1] = ffma2(in[XO + v + 1], bcast(wv), acc[no][orow][1]);
                        }
                }
            });
        });
    }
}

// {x, c} -> {x*c, c} in place over `nplanes` staged planes.  Done once per CTA in shared memory (two pairs per 16-byte
// access) rather than per thread in registers: multiplying one half of a register pair in the FFMA2 loop made ptxas copy
// the pairs around (1183 MOVs per 1600 FFMA2).
template <int K, int NT, int TH_ = 32>
__device__ __forceinline__ void premul_planes(float2* planes, int nplanes, int tid) {
    using C = TileCfg<K, TH_>;
    float4* p = reinterpret_cast<float4*>(planes);
    const int total = nplanes * C::PLANE / 2;
    for (int i = tid; i < total; i += NT) {
        float4 a = p[i];
        a.x *= a.y;
        a.z *= a.w;
        p[i] = a;
    }
}

// One input plane of the conv core: acc[no][row][px] += sum_{u,v} plane(row+u, px+v) * w[(no*NIN + ni)*K*K + u*K + v].
template <int NIN, int NOUT, int K>
__device__ __forceinline__ void conv_patch_plane(const float2* plane, const float* __restrict__ w, int ni, int tx, int ty,
                                                 float2 (&acc)[NOUT][8][2]) {
    using C = TileCfg<K>;
    constexpr int XO = C::XOFF;
    constexpr int NP = (2 + K - 1 + XO + 1) & ~1;
    const float2* row0 = plane + (ty * 8) * C::SWP + tx * 2;
    const float* wn = w + ni * K * K;
    static_for<8 + K - 1>([&](auto ir_c) {
        constexpr int ir = decltype(ir_c)::value;
        float2 in[NP];
#pragma unroll
        for (int t = 0; t < NP / 2; ++t) {
            const float4 a = *reinterpret_cast<const float4*>(row0 + ir * C::SWP + 2 * t);
            in[2 * t] = make_float2(a.x, a.y);
            in[2 * t + 1] = make_float2(a.z, a.w);
        }
        static_for<8>([&](auto orow_c) {
            constexpr int orow = decltype(orow_c)::value;
            constexpr int u = ir - orow;
            if constexpr (u >= 0 && u < K) {
#pragma unroll
                for (int v = 0; v < K; ++v)
#pragma unroll
                    for (int no = 0; no < NOUT; ++no) {
                        const float wv = wn[no * NIN * K * K + u * K + v];
                        acc[no][orow][0] = ffma2(in[XO + v], bcast(wv), acc[no][orow][0]);
                        acc[no][orow][1] = ffma2(in[XO + v + 1], bcast(wv), acc[no][orow][1]);
                    }
            }
        });
    });
}

// Co = 2 (the float4 store packs two pixels of one channel); the 1x1 2->1 head has its own kernel below.
template <int CI, int K>
__global__ void __launch_bounds__(128) nconv_fwd_kernel(const __grid_constant__ Maps2 maps, SrcP s0, SrcP s1, int has1,
                                                        float2* __restrict__ out, int H, int W, int wimg) {
    using C = TileCfg<K>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2* planes = align128(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(planes + CI * C::PLANE);
    const int tid = threadIdx.x;
    const int n = blockIdx.z, h0 = blockIdx.y * C::TH, w0 = blockIdx.x * C::TW;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;

    const int tx = tid & 31, ty = tid >> 5;
    const int oh = h0 + ty * 8, ow = w0 + tx * 2;
    float2 acc[2][8][2];
    if (s0.tma && (!has1 || s1.tma)) {
        // Every plane arrives by its own TMA box on its own mbarrier: the warps start accumulating plane 0 while the boxes of the
        // other planes are still in flight (the five resident CTAs of an SM tend to run their load / compute / store phases in step,
        // so the overlap has to come from inside the CTA too).
        if (tid == 0)
            for (int c = 0; c < CI; ++c) mbar_init(bar + c, 1);
        __syncthreads();
        if (tid == 0) {
            const int x0 = 2 * (w0 - C::P - C::XOFF), y0 = h0 - C::P;
            for (int c = 0; c < CI; ++c) {
                mbar_expect_tx(bar + c, (uint32_t)(C::SH * C::SWP * 8));
                const bool first = c < s0.ch;
                tma_load_plane(planes + c * C::PLANE, first ? &maps.m0 : &maps.m1, bar + c, x0, y0, first ? c : c - s0.ch, n);
            }
        }
        const bool active = oh < H && ow < W;
#pragma unroll
        for (int co = 0; co < 2; ++co)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[co][r][0] = acc[co][r][1] = make_float2(0.f, 0.f);
        // a real loop over the planes (weights indexed through a uniform register): half the code of the unrolled form, which
        // spent 0.8 stall cycles per issue waiting for instructions (ncu no_instruction) with five CTAs in different phases
#pragma unroll 1
        for (int ci = 0; ci < CI; ++ci) {
            mbar_wait(bar + ci, 0);
            premul_planes<K, 128>(planes + ci * C::PLANE, 1, tid);
            __syncthreads();
            if (active) conv_patch_plane<CI, 2, K>(planes + ci * C::PLANE, wc, ci, tx, ty, acc);
        }
        if (!active) return;
    } else {
        const bool pending = stage_sources<K, 128>(maps, s0, s1, has1 != 0, n, h0, w0, H, W, planes, bar, tid);
        __syncthreads();
        if (pending) mbar_wait(bar, 0);
        expand_up2_planes<K, 128>(s0, s1, has1 != 0, planes, tid);
        premul_planes<K, 128>(planes, CI, tid);
        __syncthreads();
        if (oh >= H || ow >= W) return;
        conv_patch<CI, 2, K>(planes, wc, tx, ty, acc);
    }

    const long long plane = (long long)H * W;
    const bool vec = (W & 1) == 0;
#pragma unroll
    for (int co = 0; co < 2; ++co) {
        const float b = wc[kWBias + co], is = wc[kWInvSum + co];
        float2* o = out + ((long long)n * 2 + co) * plane + ow;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int h = oh + r;
            if (h >= H) break;
            const float2 a0 = acc[co][r][0], a1 = acc[co][r][1];
            const float2 r0 = make_float2(fmaf(a0.x, rcp_approx(a0.y + kEps), b), a0.y * is);
            const float2 r1 = make_float2(fmaf(a1.x, rcp_approx(a1.y + kEps), b), a1.y * is);
            if (vec) {
                *reinterpret_cast<float4*>(o + (long long)h * W) = make_float4(r0.x, r0.y, r1.x, r1.y);
            } else {
                o[(long long)h * W] = r0;
                if (ow + 1 < W) o[(long long)h * W + 1] = r1;
            }
        }
    }
}

// ---- persistent, double-buffered variant ---------------------------------------------------------------------------------
// One CTA per SM slot loops over tiles handed out by an atomic counter (SM-level load balance: 1280 tiles over 148 SMs is
// 8.65 per SM -> 9 rounds instead of the 2 partially filled waves of a one-tile-per-CTA grid) and prefetches the NEXT
// tile (TMA boxes on a per-stage mbarrier, cp.async 8-byte gathers for mapped sources) while it computes the current one,
// so the FMA pipe does not wait for HBM.
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool valid) {
    const int sz = valid ? 8 : 0;   // src-size 0 => the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K, int NT>
__device__ __forceinline__ void gather_pair_tile_async(const SrcP& s, int sc, int n, int h0, int w0, int H, int W, float2* dst,
                                                       int tid) {
    using C = TileCfg<K>;
    constexpr int P = C::P, TOTAL = C::SH * C::SWP;
    const float2* base = s.p + (long long)n * s.bs + (long long)sc * s.H * s.W;
    for (int i = tid; i < TOTAL; i += NT) {
        const int r = i / C::SWP, col = i - r * C::SWP;
        const int gh = h0 - P + r, gw = w0 - P - C::XOFF + col;
        const bool ok = gh >= 0 && gh < H && gw >= 0 && gw < W;
        const long long off = ok ? (long long)s.mh(gh) * s.W + s.mw(gw) : 0;
        cp_async8(dst + i, base + off, ok);
    }
}

struct TileSched {
    int* counters;   // [0] next tile, [1] finished CTAs (self-resetting: the last CTA to leave zeroes both)
    int tiles_x, tiles_y, ntiles;
};

// Persistent forward kernel, pipelined at the granularity of ONE input-channel plane of one tile: two plane buffers per
// CTA (39 KB for any Ci, so 5 CTAs = 20 warps per SM stay resident); while the warps accumulate plane q, the TMA box (or the
// cp.async gather) of plane q+1 -- the next channel of this tile or the first channel of the CTA's next tile -- is in flight.
template <int CI, int K>
__global__ void __launch_bounds__(128) nconv_fwd_persist_kernel(const __grid_constant__ Maps2 maps, SrcP s0, SrcP s1, int has1,
                                                                float2* __restrict__ out, int H, int W, int wimg, TileSched sch) {
    using C = TileCfg<K>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2* buf0 = align128(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(buf0 + 2 * C::PLANE);
    __shared__ int s_next[2];
    const int tid = threadIdx.x;
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        s_next[0] = atomicAdd(&sch.counters[0], 1);
    }
    __syncthreads();

    // plane `ci` of tile `tile` -> buffer `b`; returns true if the copy went through cp.async (per-thread group committed)
    auto issue = [&](int tile, int ci, int b) -> bool {
        const int tx_ = tile % sch.tiles_x, ty_ = (tile / sch.tiles_x) % sch.tiles_y, n = tile / (sch.tiles_x * sch.tiles_y);
        const int h0 = ty_ * C::TH, w0 = tx_ * C::TW;
        const bool first = ci < s0.ch;
        const SrcP& s = first ? s0 : s1;
        const int sc = first ? ci : ci - s0.ch;
        float2* dst = buf0 + b * C::PLANE;
        if (s.tma) {
            if (tid == 0) {
                mbar_expect_tx(&bars[b], (uint32_t)(C::SH * C::SWP * 8));
                tma_load_plane(dst, first ? &maps.m0 : &maps.m1, &bars[b], 2 * (w0 - C::P - C::XOFF), h0 - C::P, sc, n);
            }
            return false;
        }
        gather_pair_tile_async<K, 128>(s, sc, n, h0, w0, H, W, dst, tid);
        cp_async_commit();
        return true;
    };
    auto plane_is_tma = [&](int ci) { return (ci < s0.ch ? s0.tma : s1.tma) != 0; };

    int tile = s_next[0];
    int q = 0;                     // global plane-step counter of this CTA: buffer = q & 1, mbarrier parity = (q >> 1) & 1
    int uses0 = 0, uses1 = 0;      // completed TMA phases per buffer (parity tracking when TMA and gather planes alternate)
    if (tile < sch.ntiles) issue(tile, 0, 0);
    while (tile < sch.ntiles) {
        if (tid == 0) s_next[1] = atomicAdd(&sch.counters[0], 1);
        int next = -1;             // read after the first barrier of this tile
        const int tx_ = tile % sch.tiles_x, ty_ = (tile / sch.tiles_x) % sch.tiles_y, n = tile / (sch.tiles_x * sch.tiles_y);
        const int tx = tid & 31, ty = tid >> 5;
        const int oh = ty_ * C::TH + ty * 8, ow = tx_ * C::TW + tx * 2;
        const bool active = oh < H && ow < W;
        float2 acc[2][8][2];
#pragma unroll
        for (int co = 0; co < 2; ++co)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[co][r][0] = acc[co][r][1] = make_float2(0.f, 0.f);

        static_for<CI>([&](auto ci_c) {
            constexpr int ci = decltype(ci_c)::value;
            const int b = q & 1;
            if constexpr (ci == 0) {
                __syncthreads();                     // publishes s_next[1] (next write is a whole tile later)
                next = s_next[1];
            }
            // prefetch the following plane into the other buffer (released by the barrier that ended the previous step)
            bool nxt_async = false;
            if (ci + 1 < CI) nxt_async = issue(tile, ci + 1, b ^ 1);
            else if (next < sch.ntiles) nxt_async = issue(next, 0, b ^ 1);
            // wait for this step's plane
            if (plane_is_tma(ci)) {
                if (b == 0) { mbar_wait(&bars[0], uses0 & 1); ++uses0; }
                else        { mbar_wait(&bars[1], uses1 & 1); ++uses1; }
            } else {
                if (nxt_async) cp_async_wait<1>(); else cp_async_wait<0>();
            }
            __syncthreads();
            premul_planes<K, 128>(buf0 + b * C::PLANE, 1, tid);
            __syncthreads();
            if (active) conv_patch_plane<CI, 2, K>(buf0 + b * C::PLANE, wc, ci, tx, ty, acc);
            __syncthreads();                         // all reads of buffer b done -> it may be refilled by the next step's prefetch
            ++q;
        });

        if (active) {
            const long long plane = (long long)H * W;
            const bool vec = (W & 1) == 0;
#pragma unroll
            for (int co = 0; co < 2; ++co) {
                const float b = wc[kWBias + co], is = wc[kWInvSum + co];
                float2* o = out + ((long long)n * 2 + co) * plane + ow;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int h = oh + r;
                    if (h >= H) break;
                    const float2 a0 = acc[co][r][0], a1 = acc[co][r][1];
                    const float2 r0 = make_float2(fmaf(a0.x, rcp_approx(a0.y + kEps), b), a0.y * is);
                    const float2 r1 = make_float2(fmaf(a1.x, rcp_approx(a1.y + kEps), b), a1.y * is);
                    if (vec) {
                        *reinterpret_cast<float4*>(o + (long long)h * W) = make_float4(r0.x, r0.y, r1.x, r1.y);
                    } else {
                        o[(long long)h * W] = r0;
                        if (ow + 1 < W) o[(long long)h * W + 1] = r1;
                    }
                }
            }
        }
        tile = next;
    }
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&sch.counters[1], 1) == (int)gridDim.x - 1) {   // last CTA out: leave the counters clean for the next launch
            sch.counters[0] = 0;
            sch.counters[1] = 0;
            __threadfence();
        }
    }
}

// nconv7: 2 -> 1, 1x1 (model_zoo/nconv_sd.py:32).  Reads the {x,c} pairs of both channels, writes the planar
// xout / cout tensors the caller sees.  Purely HBM-bound (AI 0.3): two pixels per thread, 16-byte loads.
__global__ void __launch_bounds__(256) nconv_1x1_fwd_kernel(const float2* __restrict__ in, float* __restrict__ y,
                                                            float* __restrict__ cout, long long plane, int wimg) {
    const float* wc = &c_wimg[0][0][0][0] + (size_t)wimg * kWStride;
    const float w0 = wc[0], w1 = wc[1], b = wc[kWBias], is = wc[kWInvSum];
    const int n = blockIdx.y;
    const float2* p0 = in + (long long)n * 2 * plane;
    const float2* p1 = p0 + plane;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= plane) return;
    const bool two = i + 1 < plane;
    const bool even = (plane & 1) == 0;
    float2 a0, a1, b0, b1;
    if (two && even) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(p0 + i));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(p1 + i));
        a0 = make_float2(t0.x, t0.y); a1 = make_float2(t0.z, t0.w);
        b0 = make_float2(t1.x, t1.y); b1 = make_float2(t1.z, t1.w);
    } else {
        a0 = __ldg(p0 + i); b0 = __ldg(p1 + i);
        a1 = two ? __ldg(p0 + i + 1) : make_float2(0.f, 0.f);
        b1 = two ? __ldg(p1 + i + 1) : make_float2(0.f, 0.f);
    }
    const float D0 = fmaf(w1, b0.y, w0 * a0.y), N0 = fmaf(w1, b0.x * b0.y, w0 * (a0.x * a0.y));
    const float D1 = fmaf(w1, b1.y, w0 * a1.y), N1 = fmaf(w1, b1.x * b1.y, w0 * (a1.x * a1.y));
    float* yo = y + (long long)n * plane + i;
    float* co = cout + (long long)n * plane + i;
    const float y0 = fmaf(N0, rcp_approx(D0 + kEps), b), y1 = fmaf(N1, rcp_approx(D1 + kEps), b);
    if (two && even && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(cout)) & 7) == 0) {
        *reinterpret_cast<float2*>(yo) = make_float2(y0, y1);
        *reinterpret_cast<float2*>(co) = make_float2(D0 * is, D1 * is);
    } else {
        yo[0] = y0; co[0] = D0 * is;
        if (two) { yo[1] = y1; co[1] = D1 * is; }
    }
}

// planar (x, c) [N,C,H,W] with batch strides  <->  pair tensor.  Used at the pyramid's entry (x0, c0 come from the caller
// as separate tensors) and by the single-layer API.
__global__ void __launch_bounds__(256) pack_pairs_kernel(const float* __restrict__ x, const float* __restrict__ c, long long bsx,
                                                         long long bsc, float2* __restrict__ out, long long per_image) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_image) return;
    const int n = blockIdx.y;
    out[(long long)n * per_image + i] = make_float2(__ldg(x + (long long)n * bsx + i), __ldg(c + (long long)n * bsc + i));
}

__global__ void __launch_bounds__(256) unpack_pairs_kernel(const float2* __restrict__ in, float* __restrict__ x,
                                                           float* __restrict__ c, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 v = __ldg(in + i);
    x[i] = v.x;
    c[i] = v.y;
}

}  // namespace adc

// K6: weight gradient of a dense 3x3 / stride-1 convolution on the tcgen05 tensor cores, plus the NHWC glue kernels of the
// tensor-core training path (the backward of confidence_depth_framework.py:464-518 and :630-669).
//
//   dW[co][ci][u][v] = sum_p  gy[p][co] * x[p + (u-1)*(W+2) + (v-1)][ci]          p = flat pixel of the zero-bordered layout
//
// Both tensors are zero-bordered NHWC bf16 ([N][H+2][W+2][C], borders of gy are zeros), so every tap is a shift of the flat
// pixel index -- the same property the forward v2 kernel uses.  GEMM view: M = Cout (128 rows, 64-channel blocks of gy),
// N = Cin (<= 160 columns), K = pixels.  The contraction index is the SLOW index of both operands in memory, i.e. both are
// "MN-major" UMMA operands: a TMA box [64 channels x KP pixels] with the 128-byte swizzle is exactly the canonical
// MN-major SWIZZLE_128B layout (8 K-rows of 128 bytes per atom, SBO = 1024 bytes between K groups, LBO = distance between
// 64-channel blocks).  A stride-2 layer is the same contraction on a zero-dilated gy (z[2i][2j] = gy[i][j]).
// Work split: CTA (g, u) owns vertical tap u and every G-th K tile; it keeps the three accumulators of taps (u, 0..2) in
// TMEM for its whole life (3 x Cin columns) and feeds the three horizontal taps from ONE x segment of KP+2 pixels (B
// descriptors start 0,1,2 rows into it).  Partial sums go to a workspace [G][9][Cout16][Cin16] and a second kernel reduces
// them in a fixed order (deterministic, no atomics) into the Parameter layout [Cout][Cin][3][3].
#include "tc_common.cuh"
#include "prof.h"

namespace tc {

struct WgradParams {
    long long F;          // flat pixels N*(H+2)*(W+2)
    int Wp;               // W+2
    int ncols, mrows;     // GEMM N = Cin of the slice rounded to 16; rows kept = Cout rounded to 16
    int mb, nb;           // 64-channel blocks of gy / of x per stage
    int kp;               // pixels per K tile (64 or 128)
    int ktiles, groups, nstages, split;
    int bseg;             // bytes reserved per 64-channel x block ((kp+2) rows of 128 bytes, rounded up to 1024)
    float* part;          // [groups][9][mrows][ncols]  (single: [groups][1][mrows][ncols])
    int single;           // 1: only the centre tap (dW[co][ci] = sum_p gy[p][co] * x[p][ci]: the weight gradient of a 1x1 convolution)
    // rowpack (split mode, Cout <= 64): the A operand is [gy_hi ; gy_lo] -- its second 64-row block is the lo tile (LBO = distance between
    // the tiles) instead of a replay of the first -- so [hi ; lo] x x_hi and [hi ; lo] x x_lo = 2 MMAs give all four products
    // (rows r: gy_hi * x, rows 64 + r: gy_lo * x, added by the reduction) where hi*hi, lo*hi, hi*lo took 3 with half of M idle.
    // stack (Cin <= 64): the three horizontal taps are ONE MMA of N = 3 x 64: the N blocks of an MN-major operand sit LBO bytes apart, and
    // LBO = 128 bytes = one pixel makes block v the x segment shifted by v pixels (accumulator columns 64 v + ci).
    int rowpack, stack;
    int prow;             // rows per partial slot: mrows, or 64 + mrows with rowpack
};

// MN-major SWIZZLE_128B operand: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout 2 <<61
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// as make_idesc, with A and B MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

__global__ void __launch_bounds__(192) conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g,
                                                               const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_gl,
                                                               const __grid_constant__ CUtensorMap map_xl, WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int a_bytes = p.mb * p.kp * 128, b_bytes = p.nb * p.bseg;
    const int half = a_bytes + b_bytes;
    const int stage_bytes = p.split ? 2 * half : half;
    const int nst = p.nstages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);    // full[8], empty[8], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = p.single ? 1 : blockIdx.x % 3, g = p.single ? blockIdx.x : blockIdx.x / 3;
    const int v0 = p.single ? 1 : 0, v1 = p.single ? 2 : 3, ntw = p.single ? 1 : 9;
    const int my_tiles = g < p.ktiles ? (p.ktiles - g + p.groups - 1) / p.groups : 0;
    const int vcols = p.stack ? 64 : p.ncols;                     // accumulator columns per horizontal tap
    uint32_t ncols_alloc = 32;
    while ((int)ncols_alloc < 3 * vcols) ncols_alloc <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[8 + s], 1);
        }
        mbar_init(&bars[16], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols_alloc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            const uint32_t tx = (uint32_t)((p.mb * p.kp + p.nb * (p.kp + 2)) * 128 * (p.split ? 2 : 1));
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                const long long p0 = (long long)(g + i * p.groups) * p.kp;
                const long long x0 = p0 + (long long)(u - 1) * p.Wp - 1;
                mbar_wait(&bars[8 + s], (use & 1) ^ 1);
                uint8_t* sa = smem + s * stage_bytes;
                mbar_expect_tx(&bars[s], tx);
                for (int b = 0; b < p.mb; ++b) tma_load_2d(sa + b * p.kp * 128, &map_g, &bars[s], b * 64, (int)p0);
                for (int b = 0; b < p.nb; ++b) tma_load_2d(sa + a_bytes + b * p.bseg, &map_x, &bars[s], b * 64, (int)x0);
                if (p.split) {
                    for (int b = 0; b < p.mb; ++b) tma_load_2d(sa + half + b * p.kp * 128, &map_gl, &bars[s], b * 64, (int)p0);
                    for (int b = 0; b < p.nb; ++b) tma_load_2d(sa + half + a_bytes + b * p.bseg, &map_xl, &bars[s], b * 64, (int)x0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            const int rowpack = p.rowpack, stack = p.stack;
            const uint32_t idesc = make_idesc_mn(128, stack ? 192 : p.ncols);
            // one block only: rows 64..127 replay rows 0..63 (never read back) -- or, with rowpack, are the lo tile
            const uint32_t lbo_a = rowpack ? (uint32_t)half : (p.mb > 1 ? (uint32_t)(p.kp * 128) : 0u);
            const uint32_t lbo_b = stack ? 128u : (uint32_t)p.bseg;
            const int ksteps = p.kp / 16;
            // descriptors by 64-bit adds from two bases (low field = shared-memory address / 16) and loop invariants in registers: the
            // MMA asm's memory clobber otherwise makes the compiler rebuild both descriptors and re-read parameters for every MMA
            const uint64_t dA0 = make_desc_mn(base, lbo_a), dB0 = make_desc_mn(base + a_bytes, lbo_b);
            const uint32_t sstep = (uint32_t)stage_bytes >> 4, hoff = (uint32_t)half >> 4;
            const int split = p.split, ncols = p.ncols;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                mbar_wait(&bars[s], use & 1);
                tc_fence_after();
                const uint32_t soff = (uint32_t)s * sstep;
                for (int v = v0; v < (stack ? v0 + 1 : v1); ++v) {
                    const uint32_t tacc = tmem_base + (uint32_t)(v * ncols);
                    const uint64_t dav = dA0 + soff, dbv = dB0 + (uint64_t)(soff + (uint32_t)(v * 8));     // + v * 128 bytes
                    for (int k = 0; k < ksteps; ++k) {       // 16 pixels = 2 swizzle atoms = 2048 bytes per K step
                        const uint64_t da = dav + (uint32_t)(k * 128), db = dbv + (uint32_t)(k * 128);
                        umma_bf16(tacc, da, db, idesc, (i | k) != 0);
                        if (rowpack) {
                            umma_bf16(tacc, da, db + hoff, idesc, 1);
                        } else if (split) {
                            umma_bf16(tacc, da + hoff, db, idesc, 1);
                            umma_bf16(tacc, da, db + hoff, idesc, 1);
                        }
                    }
                }
                umma_commit(&bars[8 + s]);
            }
            umma_commit(&bars[16]);
        }
    } else {
        // ===== epilogue warps 2..5: rows (= output channels) quarter*32 + lane =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        if (my_tiles > 0) {
            mbar_wait(&bars[16], 0);
            tc_fence_after();
        }
        const bool keep = row < p.mrows || (p.rowpack && row >= 64 && row < 64 + p.mrows);
        for (int v = v0; v < v1; ++v) {
            float* dst = p.part + (((long long)g * ntw + (p.single ? 0 : u * 3 + v)) * p.prow + row) * p.ncols;
            for (int c0 = 0; c0 < p.ncols; c0 += 16) {
                float r[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((p.stack ? (v - v0) * 64 : v * p.ncols) + c0), r);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0.f;
                }
                if (keep) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols_alloc);
}

// CTA-pair form of the flat kernel (split mode, 64 < Cin <= 128): the two CTAs of a cluster execute ONE tcgen05.mma.cta_group::2 of
// M = 256 per tap, K step and x plane.  CTA 0 holds the gy_hi tile and CTA 1 the gy_lo tile (its 128 rows of A); CTA r holds the 64-channel
// block r of x, both planes (its N/2 = 64 columns of B).  D of CTA 0 = gy_hi * (x_hi + x_lo), D of CTA 1 = gy_lo * (x_hi + x_lo): all four
// products in 2 instructions of 64 cycles (tools/micro/mma_pair_rate.cu) where the single-CTA form issues 3 of 102, and each CTA loads half
// the bytes of a K tile.  Partial slot: rows [0, mrows) from CTA 0, [mrows, 2 mrows) from CTA 1, added by the reduction.
__global__ void __launch_bounds__(192) conv3x3_wgrad_pair_tc_kernel(const __grid_constant__ CUtensorMap map_g,
                                                                    const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ CUtensorMap map_gl,
                                                                    const __grid_constant__ CUtensorMap map_xl, WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int a_bytes = p.mb * p.kp * 128;
    const int stage_bytes = a_bytes + 2 * p.bseg;                 // [gy plane of this CTA | x_hi block | x_lo block]
    const int nst = p.nstages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);    // full[8], empty[8], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pairid = blockIdx.x >> 1, u = pairid % 3, g = pairid / 3;
    const int my_tiles = g < p.ktiles ? (p.ktiles - g + p.groups - 1) / p.groups : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[8 + s], 1);
        }
        mbar_init(&bars[16], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer (both CTAs; the bytes are counted on the leader's barrier) =====
            const CUtensorMap* mg = rank ? &map_gl : &map_g;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                const long long p0 = (long long)(g + i * p.groups) * p.kp;
                const long long x0 = p0 + (long long)(u - 1) * p.Wp - 1;
                mbar_wait(&bars[8 + s], (use & 1) ^ 1);
                uint8_t* sa = smem + s * stage_bytes;
                const uint32_t fbar = mapa_u32(smem_u32(&bars[s]), 0);
                if (rank == 0) mbar_expect_tx(&bars[s], (uint32_t)(2 * (p.mb * p.kp + 2 * (p.kp + 2)) * 128));
                for (int b = 0; b < p.mb; ++b) tma_load_2d_pair(sa + b * p.kp * 128, mg, fbar, b * 64, (int)p0);
                tma_load_2d_pair(sa + a_bytes, &map_x, fbar, 64 * (int)rank, (int)x0);
                tma_load_2d_pair(sa + a_bytes + p.bseg, &map_xl, fbar, 64 * (int)rank, (int)x0);
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (leader) =====
            const uint32_t idesc = make_idesc_mn(256, 128);
            const uint32_t lbo_a = p.mb > 1 ? (uint32_t)(p.kp * 128) : 0u;
            const int ksteps = p.kp / 16;
            const uint64_t dA0 = make_desc_mn(base, lbo_a), dB0 = make_desc_mn(base + a_bytes, 0u);
            const uint32_t sstep = (uint32_t)stage_bytes >> 4, loff = (uint32_t)p.bseg >> 4;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                mbar_wait(&bars[s], use & 1);
                tc_fence_after();
                const uint32_t soff = (uint32_t)s * sstep;
                for (int v = 0; v < 3; ++v) {
                    const uint32_t tacc = tmem_base + (uint32_t)(v * 128);
                    const uint64_t dav = dA0 + soff, dbv = dB0 + (uint64_t)(soff + (uint32_t)(v * 8));     // + v pixels
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t da = dav + (uint32_t)(k * 128), db = dbv + (uint32_t)(k * 128);
                        umma_bf16_pair(tacc, da, db, idesc, (i | k) != 0);
                        umma_bf16_pair(tacc, da, db + loff, idesc, 1);
                    }
                }
                umma_commit_pair(&bars[8 + s]);
            }
            umma_commit_pair(&bars[16]);
        }
    } else {
        // ===== epilogue warps 2..5 of both CTAs: rows = channels of this CTA's gy plane =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        if (my_tiles > 0) {
            mbar_wait(&bars[16], 0);
            tc_fence_after();
        }
        for (int v = 0; v < 3; ++v) {
            float* dst = p.part + (((long long)g * 9 + u * 3 + v) * p.prow + (long long)rank * p.mrows + row) * p.ncols;
            for (int c0 = 0; c0 < p.ncols; c0 += 16) {
                float r[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(v * 128 + c0), r);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0.f;
                }
                if (row < p.mrows) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// Box variant for strided layers (GED conv1-5, stride 2): the flat-index trick needs equal geometry of gy and x, so here a K
// tile is a BH x BW block of OUTPUT pixels of one image; gy comes as a 4-D TMA box and the x operand of tap (u,v) as a 4-D box
// with element strides (stride, stride) starting at (ow0*stride + v - 1, oh0*stride + u - 1) -- out-of-bounds zero fill is
// the padding.  One tap per CTA (CTA = (group, tap), grid 9 x SMs/9), one accumulator of Cin columns; same MN-major
// descriptors, same partial layout and reduction as the flat kernel.
struct WgradBoxParams {
    int N, tiles_x, tiles_y, BW, BH;
    int stride, gy_pad, x_pad;
    int ncols, mrows, mb, nb, kp, ktiles, groups, nstages, split;
    float* part;
    int single;           // 1: the centre tap only (weight gradient of a 1x1 convolution of the same stride, no padding)
    int rowpack, prow;    // as in WgradParams
};

__global__ void __launch_bounds__(192) conv3x3_wgrad_box_tc_kernel(const __grid_constant__ CUtensorMap map_g,
                                                                   const __grid_constant__ CUtensorMap map_x,
                                                                   const __grid_constant__ CUtensorMap map_gl,
                                                                   const __grid_constant__ CUtensorMap map_xl, WgradBoxParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int blk = p.kp * 128;                                   // one 64-channel block of KP pixels
    const int a_bytes = p.mb * blk, b_bytes = p.nb * blk;
    const int half = a_bytes + b_bytes;
    const int stage_bytes = p.split ? 2 * half : half;
    const int nst = p.nstages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);    // full[8], empty[8], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap = p.single ? 4 : blockIdx.x % 9, g = p.single ? blockIdx.x : blockIdx.x / 9;
    const int u = tap / 3, v = tap % 3;
    const int my_tiles = g < p.ktiles ? (p.ktiles - g + p.groups - 1) / p.groups : 0;
    uint32_t ncols_alloc = 32;
    while ((int)ncols_alloc < p.ncols) ncols_alloc <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[8 + s], 1);
        }
        mbar_init(&bars[16], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols_alloc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t tx = (uint32_t)(half * (p.split ? 2 : 1));
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                const int t = g + i * p.groups;
                const int bx = t % p.tiles_x, by = (t / p.tiles_x) % p.tiles_y, n = t / (p.tiles_x * p.tiles_y);
                const int ow0 = bx * p.BW, oh0 = by * p.BH;
                const int gx0 = ow0 + p.gy_pad, gy0 = oh0 + p.gy_pad;
                const int xx0 = ow0 * p.stride + v - 1 + p.x_pad, xy0 = oh0 * p.stride + u - 1 + p.x_pad;
                mbar_wait(&bars[8 + s], (use & 1) ^ 1);
                uint8_t* sa = smem + s * stage_bytes;
                mbar_expect_tx(&bars[s], tx);
                for (int b = 0; b < p.mb; ++b) tma_load_4d(sa + b * blk, &map_g, &bars[s], b * 64, gx0, gy0, n);
                for (int b = 0; b < p.nb; ++b) tma_load_4d(sa + a_bytes + b * blk, &map_x, &bars[s], b * 64, xx0, xy0, n);
                if (p.split) {
                    for (int b = 0; b < p.mb; ++b) tma_load_4d(sa + half + b * blk, &map_gl, &bars[s], b * 64, gx0, gy0, n);
                    for (int b = 0; b < p.nb; ++b) tma_load_4d(sa + half + a_bytes + b * blk, &map_xl, &bars[s], b * 64, xx0, xy0, n);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(128, p.ncols);
            const int rowpack = p.rowpack;
            const uint32_t lbo_a = rowpack ? (uint32_t)half : (p.mb > 1 ? (uint32_t)blk : 0u), lbo_b = (uint32_t)blk;
            const int ksteps = p.kp / 16;
            const uint64_t dA0 = make_desc_mn(base, lbo_a), dB0 = make_desc_mn(base + a_bytes, lbo_b);      // (see the flat kernel)
            const uint32_t sstep = (uint32_t)stage_bytes >> 4, hoff = (uint32_t)half >> 4;
            const int split = p.split;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nst, use = i / nst;
                mbar_wait(&bars[s], use & 1);
                tc_fence_after();
                const uint32_t soff = (uint32_t)s * sstep;
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t da = dA0 + (uint64_t)(soff + (uint32_t)(k * 128)), db = dB0 + (uint64_t)(soff + (uint32_t)(k * 128));
                    umma_bf16(tmem_base, da, db, idesc, (i | k) != 0);
                    if (rowpack) {
                        umma_bf16(tmem_base, da, db + hoff, idesc, 1);
                    } else if (split) {
                        umma_bf16(tmem_base, da + hoff, db, idesc, 1);
                        umma_bf16(tmem_base, da, db + hoff, idesc, 1);
                    }
                }
                umma_commit(&bars[8 + s]);
            }
            umma_commit(&bars[16]);
        }
    } else {
        const int quarter = warp & 3, row = quarter * 32 + lane;
        if (my_tiles > 0) {
            mbar_wait(&bars[16], 0);
            tc_fence_after();
        }
        const bool keep = row < p.mrows || (p.rowpack && row >= 64 && row < 64 + p.mrows);
        float* dst = p.part + (p.single ? (long long)g * p.prow + row : ((long long)g * 9 + tap) * p.prow + row) * p.ncols;
        for (int c0 = 0; c0 < p.ncols; c0 += 16) {
            float r[16];
            if (my_tiles > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0.f;
            }
            if (keep) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols_alloc);
}

// dW[co][ci][tap] = sum_g part[g][tap][co][ci]   (fixed order; reads coalesced along ci)
// mrows = rows per partial slot; lo_rows = 64 when rows 64 + co hold a second term of the same element (rowpack), else 0
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int groups, int mrows,
                                                              int ncols, int cout, int cin, int ntaps, int lo_rows) {
    const int total = ntaps * cout * cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % cin, co = (i / cin) % cout, tap = i / (cin * cout);
        // four partial sums: the loads of one thread overlap (a single serial chain over 148 groups was latency-bound: 40 us)
        const long long gs = (long long)ntaps * mrows * ncols;
        const float* q = part + ((long long)tap * mrows + co) * ncols + ci;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int g = 0;
        for (; g + 3 < groups; g += 4) {
            a0 += q[(long long)g * gs];
            a1 += q[(long long)(g + 1) * gs];
            a2 += q[(long long)(g + 2) * gs];
            a3 += q[(long long)(g + 3) * gs];
        }
        for (; g < groups; ++g) a0 += q[(long long)g * gs];
        if (lo_rows) {
            const float* ql = q + (long long)lo_rows * ncols;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            for (g = 0; g + 3 < groups; g += 4) {
                b0 += ql[(long long)g * gs];
                b1 += ql[(long long)(g + 1) * gs];
                b2 += ql[(long long)(g + 2) * gs];
                b3 += ql[(long long)(g + 3) * gs];
            }
            for (; g < groups; ++g) b0 += ql[(long long)g * gs];
            a0 += b0; a1 += b1; a2 += b2; a3 += b3;
        }
        dw[((long long)co * cin + ci) * ntaps + tap] = (a0 + a1) + (a2 + a3);
    }
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 1;
    }
    return sms;
}
// ADC_WGRAD_SMEM_RESERVE=KB: shared memory the weight-gradient kernels leave free on every SM so that CTAs of the HBM-bound
// gradient-assembly pass (16 KB static each) can be resident beside them (experiment; default 0).
static int wgrad_smem_reserve() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ADC_WGRAD_SMEM_RESERVE");
        const int kb = e ? atoi(e) : 0;
        v = (kb < 0 ? 0 : (kb > 96 ? 96 : kb)) * 1024;
    }
    return v;
}
static int wgrad_groups() { return sm_count() / 3 > 0 ? sm_count() / 3 : 1; }        // flat kernel: CTA = (group, u)
static int wgrad_box_groups() { return sm_count() / 9 > 0 ? sm_count() / 9 : 1; }    // box kernel: CTA = (group, tap)

}  // namespace tc

using namespace tc;

extern "C" {

size_t adc_conv3x3_wgrad_tc_workspace_bytes(int cin_pad, int cout) {
    if (cin_pad <= 0 || cout <= 0) return 0;
    const int mrows = (cout + 15) / 16 * 16;
    const int prow = cout <= 64 ? 64 + mrows : mrows;          // rowpack: a slot holds the gy_hi rows and, 64 rows on, the gy_lo rows
    const int g = wgrad_groups() * 9 > sm_count() ? wgrad_groups() * 9 : sm_count();     // 3x3 forms: groups x 9 slots; 1x1 form: one slot per SM
    return (size_t)g * prow * cin_pad * sizeof(float);
}

// gy: zero-bordered NHWC bf16 [N][H+2][W+2][gy_stride] (channels [gy_off, gy_off+cout) are read; border pixels must be 0);
// x: zero-bordered NHWC bf16 [N][H+2][W+2][x_stride] (channels [x_off, x_off+cin_pad), cin_pad % 16 == 0, <= 160);
// dw: fp32 [cout][cin][3][3] (the first cin of the cin_pad columns); workspace: adc_conv3x3_wgrad_tc_workspace_bytes().
// Split (bf16x3) mode when gy_lo and x_lo are given: hi*hi + lo*hi + hi*lo accumulated in fp32.
static int wgrad_flat(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* x, const void* x_lo, int x_stride,
                      int x_off, float* dw, void* workspace, int N, int H, int W, int cin, int cin_pad, int cout, int single,
                      adc_stream_t stream) {
    if (!gy || !x || !dw || !workspace || N <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || cin_pad < cin) {
        adc::set_error("adc_conv3x3_wgrad_tc: bad argument");
        return ADC_E_BADARG;
    }
    if ((gy_lo == nullptr) != (x_lo == nullptr)) {
        adc::set_error("adc_conv3x3_wgrad_tc: gy_lo and x_lo must be given together (split mode)");
        return ADC_E_BADARG;
    }
    if (cin_pad % 16 != 0 || cin_pad > 160 || cout > 128 || gy_stride % 8 != 0 || gy_off % 8 != 0 || x_stride % 8 != 0 || x_off % 8 != 0 ||
        gy_off + cout > gy_stride || x_off + cin_pad > x_stride) {
        adc::set_error("adc_conv3x3_wgrad_tc: needs cin_pad %% 16 == 0, cin_pad <= 160, cout <= 128 and 16-byte aligned channel slices "
                       "(cin_pad=%d cout=%d)", cin_pad, cout);
        return ADC_E_UNSUPP;
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        adc::set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.Wp = W + 2;
    p.F = (long long)N * (H + 2) * (W + 2);
    p.ncols = cin_pad;
    p.mrows = (cout + 15) / 16 * 16;
    p.mb = cout > 64 ? 2 : 1;
    p.nb = (cin_pad + 63) / 64;
    p.split = gy_lo != nullptr;
    p.single = single;
    p.groups = single ? sm_count() : wgrad_groups();
    p.part = reinterpret_cast<float*>(workspace);
    {
        const char* e = getenv("ADC_WGRAD_PLAIN");           // (experiment: the three-MMA, one-tap-per-MMA form)
        const bool plain = e && e[0] == '1';
        p.rowpack = (p.split && p.mb == 1 && !plain) ? 1 : 0;
        p.stack = (p.nb == 1 && !single && !plain) ? 1 : 0;
        p.prow = p.rowpack ? 64 + p.mrows : p.mrows;
    }
    bool pair = false;
    {
        // Measured (gudepthcompnet18 step, B = 16): the pair form is correct but SLOWER where it applies (832 vs 496 us for the 129-channel
        // head, weight gradients 7.34 vs 7.00 ms per step): the single-CTA kernel already executes its three MMAs per product at
        // ~1.4 PFLOP/s, the kernel is fed by TMA rather than bound by MMA issue, and 24 pairs per tap row split K more coarsely than
        // 49 CTAs.  Kept behind a switch as the measured alternative.
        const char* e = getenv("ADC_WGRAD_PAIR");
        pair = p.split && p.nb == 2 && !single && sm_count() >= 6 && e && e[0] == '1';
    }
    if (pair) {
        p.rowpack = 0;
        p.prow = 2 * p.mrows;
        p.groups = sm_count() / 6;          // pairs of CTAs: (group, u)
    }
    const int budget = kSmemCap - 1024 - wgrad_smem_reserve();
    for (p.kp = 128; p.kp >= 64; p.kp /= 2) {
        p.bseg = ((p.kp + 2) * 128 + 1023) / 1024 * 1024;
        const int stage = pair ? p.mb * p.kp * 128 + 2 * p.bseg : (p.mb * p.kp * 128 + p.nb * p.bseg) * (p.split ? 2 : 1);
        p.nstages = budget / stage;
        if (p.nstages > 8) p.nstages = 8;
        if (p.nstages >= 3 || (p.kp == 64 && p.nstages >= 2)) break;
    }
    if (p.kp < 64 || p.nstages < 2) {
        adc::set_error("adc_conv3x3_wgrad_tc: stage does not fit in shared memory (cin_pad=%d cout=%d)", cin_pad, cout);
        return ADC_E_UNSUPP;
    }
    p.ktiles = (int)((p.F + p.kp - 1) / p.kp);
    const int smem = p.nstages * (pair ? p.mb * p.kp * 128 + 2 * p.bseg : (p.mb * p.kp * 128 + p.nb * p.bseg) * (p.split ? 2 : 1)) + 1024 + 256;

    auto make = [&](CUtensorMap* m, const void* base_, int stride, int off, int rows) -> int {
        const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(base_) + off;
        cuuint64_t dims[2] = {(cuuint64_t)(stride - off), (cuuint64_t)p.F};
        cuuint64_t strides[1] = {(cuuint64_t)stride * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(b), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    CUtensorMap mg, mx, mgl, mxl;
    memset(&mgl, 0, sizeof(mgl));
    memset(&mxl, 0, sizeof(mxl));
    int rc;
    if ((rc = make(&mg, gy, gy_stride, gy_off, p.kp)) || (rc = make(&mx, x, x_stride, x_off, p.kp + 2))) return rc;
    if (p.split && ((rc = make(&mgl, gy_lo, gy_stride, gy_off, p.kp)) || (rc = make(&mxl, x_lo, x_stride, x_off, p.kp + 2)))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap),
                             "cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel)");
        if (rc) return rc;
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_wgrad_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap),
                             "cudaFuncSetAttribute(conv3x3_wgrad_pair_tc_kernel)");
        if (rc) return rc;
        attr_done = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_WGRAD, 2.0 * N * H * W * (single ? 1.0 : 9.0) * cin_pad * cout);
    adc::count_launches(1);
    if (pair) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(6 * p.groups);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if ((rc = adc::check_cuda(cudaLaunchKernelEx(&cfg, conv3x3_wgrad_pair_tc_kernel, mg, mx, mgl, mxl, p), "conv3x3_wgrad_pair_tc_kernel launch")))
            return rc;
    } else {
        conv3x3_wgrad_tc_kernel<<<(single ? 1 : 3) * p.groups, 192, smem, st>>>(mg, mx, mgl, mxl, p);
        if ((rc = adc::check_cuda(cudaGetLastError(), "conv3x3_wgrad_tc_kernel launch"))) return rc;
    }
    const int total = (single ? 1 : 9) * cout * cin;
    wgrad_tc_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(p.part, dw, p.groups, p.prow, p.ncols, cout, cin, single ? 1 : 9,
                                                                pair ? p.mrows : (p.rowpack ? 64 : 0));
    return adc::check_cuda(cudaGetLastError(), "wgrad_tc_reduce_kernel launch");
}

int adc_conv3x3_wgrad_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* x, const void* x_lo, int x_stride,
                         int x_off, float* dw, void* workspace, int N, int H, int W, int cin, int cin_pad, int cout,
                         adc_stream_t stream) {
    return wgrad_flat(gy, gy_lo, gy_stride, gy_off, x, x_lo, x_stride, x_off, dw, workspace, N, H, W, cin, cin_pad, cout, 0, stream);
}

// Weight gradient of a 1x1 convolution (plain GEMM over the flat pixels): dw fp32 [cout][cin].  Same tensors, workspace query and
// determinism rules as adc_conv3x3_wgrad_tc; every SM owns a share of the pixels.
int adc_conv1x1_wgrad_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* x, const void* x_lo, int x_stride,
                         int x_off, float* dw, void* workspace, int N, int H, int W, int cin, int cin_pad, int cout,
                         adc_stream_t stream) {
    return wgrad_flat(gy, gy_lo, gy_stride, gy_off, x, x_lo, x_stride, x_off, dw, workspace, N, H, W, cin, cin_pad, cout, 1, stream);
}

// Strided / general-geometry variant (see conv3x3_wgrad_box_tc_kernel): gy [N][Ho][Wo] (+border if gy_pad), x [N][H][W]
// (+border if x_pad), Ho = (H-1)/stride + 1.  Same channel-slice, split-mode, workspace and determinism rules as above.
static int wgrad_box(const void* gy, const void* gy_lo, int gy_stride, int gy_off, int gy_pad, const void* x, const void* x_lo,
                     int x_stride, int x_off, int x_pad, float* dw, void* workspace, int N, int H, int W, int stride, int cin,
                     int cin_pad, int cout, int single, adc_stream_t stream) {
    if (!gy || !x || !dw || !workspace || N <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || cin_pad < cin || (stride != 1 && stride != 2)) {
        adc::set_error("adc_conv3x3_wgrad_box_tc: bad argument");
        return ADC_E_BADARG;
    }
    if ((gy_lo == nullptr) != (x_lo == nullptr)) {
        adc::set_error("adc_conv3x3_wgrad_box_tc: gy_lo and x_lo must be given together (split mode)");
        return ADC_E_BADARG;
    }
    if (cin_pad % 16 != 0 || cin_pad > 256 || cout > 128 || gy_stride % 8 != 0 || gy_off % 8 != 0 || x_stride % 8 != 0 || x_off % 8 != 0 ||
        gy_off + cout > gy_stride || x_off + cin_pad > x_stride) {
        adc::set_error("adc_conv3x3_wgrad_box_tc: needs cin_pad %% 16 == 0, cin_pad <= 256, cout <= 128 and 16-byte aligned channel slices "
                       "(cin_pad=%d cout=%d)", cin_pad, cout);
        return ADC_E_UNSUPP;
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        adc::set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    WgradBoxParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.stride = stride; p.gy_pad = gy_pad ? 1 : 0; p.x_pad = x_pad ? 1 : 0;
    p.ncols = cin_pad;
    p.mrows = (cout + 15) / 16 * 16;
    p.mb = cout > 64 ? 2 : 1;
    p.nb = (cin_pad + 63) / 64;
    p.split = gy_lo != nullptr;
    p.single = single ? 1 : 0;
    p.groups = single ? sm_count() : wgrad_box_groups();
    p.part = reinterpret_cast<float*>(workspace);
    {
        const char* e = getenv("ADC_WGRAD_PLAIN");
        p.rowpack = (p.split && p.mb == 1 && !(e && e[0] == '1')) ? 1 : 0;
        p.prow = p.rowpack ? 64 + p.mrows : p.mrows;
    }
    const int budget = kSmemCap - 1024 - wgrad_smem_reserve();
    for (p.kp = 128; p.kp >= 64; p.kp /= 2) {
        const int stage = (p.mb + p.nb) * p.kp * 128 * (p.split ? 2 : 1);
        p.nstages = budget / stage;
        if (p.nstages > 8) p.nstages = 8;
        if (p.nstages >= 3 || (p.kp == 64 && p.nstages >= 2)) break;
    }
    if (p.kp < 64 || p.nstages < 2) {
        adc::set_error("adc_conv3x3_wgrad_box_tc: stage does not fit in shared memory (cin_pad=%d cout=%d)", cin_pad, cout);
        return ADC_E_UNSUPP;
    }
    p.BW = Wo >= 24 ? 32 : (Wo >= 12 ? 16 : 8);
    p.BH = p.kp / p.BW;
    p.tiles_x = (Wo + p.BW - 1) / p.BW; p.tiles_y = (Ho + p.BH - 1) / p.BH;
    p.ktiles = p.tiles_x * p.tiles_y * N;
    const int smem = p.nstages * (p.mb + p.nb) * p.kp * 128 * (p.split ? 2 : 1) + 1024 + 256;

    auto make = [&](CUtensorMap* m, const void* base_, int cstride, int off, int h, int w, int pad, int st) -> int {
        const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(base_) + off;
        const int Wt = w + 2 * pad, Ht = h + 2 * pad;
        cuuint64_t dims[4] = {(cuuint64_t)(cstride - off), (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)cstride * 2 * Wt, (cuuint64_t)cstride * 2 * Wt * Ht};
        cuuint32_t box[4] = {64, (cuuint32_t)(p.BW * st), (cuuint32_t)(p.BH * st), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(b), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(wgrad box) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    CUtensorMap mg, mx, mgl, mxl;
    memset(&mgl, 0, sizeof(mgl));
    memset(&mxl, 0, sizeof(mxl));
    int rc;
    if ((rc = make(&mg, gy, gy_stride, gy_off, Ho, Wo, p.gy_pad, 1)) || (rc = make(&mx, x, x_stride, x_off, H, W, p.x_pad, stride))) return rc;
    if (p.split && ((rc = make(&mgl, gy_lo, gy_stride, gy_off, Ho, Wo, p.gy_pad, 1)) || (rc = make(&mxl, x_lo, x_stride, x_off, H, W, p.x_pad, stride))))
        return rc;
    static bool attr_done = false;
    if (!attr_done) {
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_wgrad_box_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap),
                             "cudaFuncSetAttribute(conv3x3_wgrad_box_tc_kernel)");
        if (rc) return rc;
        attr_done = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TC_WGRAD_BOX, 2.0 * N * Ho * Wo * 9.0 * cin_pad * cout);
    adc::count_launches(1);
    conv3x3_wgrad_box_tc_kernel<<<(single ? 1 : 9) * p.groups, 192, smem, st>>>(mg, mx, mgl, mxl, p);
    if ((rc = adc::check_cuda(cudaGetLastError(), "conv3x3_wgrad_box_tc_kernel launch"))) return rc;
    const int total = (single ? 1 : 9) * cout * cin;
    wgrad_tc_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(p.part, dw, p.groups, p.prow, p.ncols, cout, cin, single ? 1 : 9,
                                                                p.rowpack ? 64 : 0);
    return adc::check_cuda(cudaGetLastError(), "wgrad_tc_reduce_kernel launch");
}

int adc_conv3x3_wgrad_box_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, int gy_pad, const void* x, const void* x_lo,
                             int x_stride, int x_off, int x_pad, float* dw, void* workspace, int N, int H, int W, int stride, int cin,
                             int cin_pad, int cout, adc_stream_t stream) {
    return wgrad_box(gy, gy_lo, gy_stride, gy_off, gy_pad, x, x_lo, x_stride, x_off, x_pad, dw, workspace, N, H, W, stride, cin, cin_pad, cout, 0,
                     stream);
}

// Weight gradient dw fp32 [cout][cin] of a 1x1 convolution of stride 1 or 2 (no padding) from NHWC activations of different geometry
// (the ResNet down-sampling shortcut, s2d_u_resnet.py:99-107): the centre tap of the box kernel, every SM on a share of the pixels.
int adc_conv1x1_wgrad_box_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, int gy_pad, const void* x, const void* x_lo,
                             int x_stride, int x_off, int x_pad, float* dw, void* workspace, int N, int H, int W, int stride, int cin,
                             int cin_pad, int cout, adc_stream_t stream) {
    return wgrad_box(gy, gy_lo, gy_stride, gy_off, gy_pad, x, x_lo, x_stride, x_off, x_pad, dw, workspace, N, H, W, stride, cin, cin_pad, cout, 1,
                     stream);
}

}  // extern "C"

// K5: dense 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA), for the wide layers around the NConv pyramid (model_zoo/
// confidence_depth_framework.py:442-452,536-607).  sm_100a only.
//
//   out[n, oh, ow, co] = act( bias[co] + sum_{u,v,ci} W[co,ci,u,v] * in[n, oh*s+u-1, ow*s+v-1, ci] )
//
// Layout: activations NHWC bf16 (channel count a multiple of 16), weights bf16 [tap][Cout_pad][Cin] (tap = u*3+v).
// GEMM view: M = 128 output pixels (a BH x BW block of one image), N = Cout_pad (<= 256), K = 9 taps x Cin.
// For every (tap, 64/32/16-channel block) one TMA box copy brings the shifted pixel block (hardware zero fill outside
// the image = the convolution's padding; element strides 2 implement stride-2 convolutions) and one brings the weight
// slice, both K-major with the 128/64/32-byte hardware swizzle that the UMMA shared-memory descriptors expect.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; also owns the TMEM
// allocation), warps 2-5 = epilogue (tcgen05.ld of their TMEM lane quarter -> bias + activation -> bf16/fp32 store).
// Ring of STAGES smem slots with full/empty mbarriers; tcgen05.commit releases a slot when the MMAs that read it retire.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/adc_b200.h"

#include "tc_common.cuh"
#include "prof.h"

namespace tc {

struct ConvParams {
    int N, H, W;            // input image size
    int Ho, Wo;             // output size
    int cin, npad, cout;    // Cin (multiple of 16), Cout padded to 16, real Cout
    int stride;             // 1 or 2
    int BW, BH;             // output pixel block, BW*BH = 128
    int tiles_x, tiles_y;
    int act;                // 0 none, 1 relu, 2 leaky-relu(0.2), 3 sigmoid
    int out_f32;            // 1: fp32 output, 0: bf16
    long long out_pix_stride;   // elements between consecutive pixels in the output (>= cout: lets a layer write a channel slice)
    int out_ch_off;
    void* out;
    void* out_lo;           // split mode, bf16 output: the low halves (out = hi)
    const float* bias;      // may be null
    int split;              // 1: every operand is a (hi, lo) pair of bf16 tensors and hi*hi + lo*hi + hi*lo is accumulated
    int x_pad, out_pad;     // 1: the tensor is stored with a one-pixel zero border ([H+2][W+2] per image)
    // Tap list: input box offset (dy, dx) and weight tap index of every tap that is summed.  The plain convolution has nine
    // (dy = u-1, dx = v-1, tap = 3u+v); the data gradient of a stride-2 convolution is four sub-convolutions, one per
    // parity class of the input pixel, with 1/2/2/4 taps each.
    int ntaps;
    signed char tdy[9], tdx[9], ttap[9];
    // output pixel of tile-space pixel (oh, ow): (oh*omul + oay, ow*omul + oax) in an image of OutH x OutW
    int omul, oay, oax, OutH, OutW;
    int accum;              // fp32 output: out += result (a gradient buffer with several producers)
    int ntiles;             // tiles_x * tiles_y * N: the CTAs are persistent over them
    int nstages;            // pipeline stages that fit (2..kStages)
};

constexpr int kStages = 4;

// PAIR (split mode): clusters of two CTAs execute every MMA together (tcgen05.mma.cta_group::2, M = 256 = the two CTAs' pixel blocks).
// CTA r holds both weight planes of the output channels [r npad/2, (r+1) npad/2): half the weight bytes per CTA and step, and the three
// products of bf16x3 are three instructions of N = npad at N/2 cycles each (see conv3x3_s1_tc_v2_kernel for the barrier protocol).
template <int KBLK, bool PAIR>   // channels per K step: 64 / 32 / 16  (row bytes 128 / 64 / 32)
__global__ void __launch_bounds__(192) conv3x3_fprop_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w,
                                                               const __grid_constant__ CUtensorMap map_xl,
                                                               const __grid_constant__ CUtensorMap map_wl, ConvParams p) {
    constexpr int ROWB = KBLK * 2;
    constexpr int A_BYTES = 128 * ROWB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int brows = PAIR ? p.npad / 2 : p.npad;          // weight rows (output channels) of a tap tile held by this CTA
    const int b_bytes = brows * ROWB;
    const int b_off = A_BYTES;                                     // B_hi
    const int al_off = (A_BYTES + b_bytes + 1023) & ~1023;         // A_lo (split mode)
    const int bl_off = al_off + A_BYTES;                           // B_lo
    const int stage_bytes = p.split ? 2 * al_off : al_off;
    // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
    const int nst = p.nstages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    // Persistent CTAs (grid = SM count) with a continuous TMA pipeline and two TMEM accumulators: the first version launched one CTA per
    // 128-pixel tile, so barrier set-up, TMEM allocation, pipeline fill / drain and the epilogue of every tile ran back to back with its
    // 18-72 pipeline steps (39 waves of ~13 us for GED's conv2, ~4 us of each not overlapped with anything).
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.cin + KBLK - 1) / KBLK;
    const int ksteps = p.ntaps * kblocks;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * p.npad) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[kStages + s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars[2 * kStages + a], 1);
            mbar_init(&bars[2 * kStages + 2 + a], PAIR ? 8 : 4);
        }
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[256];
    for (int i = threadIdx.x; i < 256; i += 192) sbias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair(tmem_slot, ncols);
        else tmem_alloc(tmem_slot, ncols);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            int q = 0;
            for (int tile = blockIdx.x; tile - (int)rank < p.ntiles; tile += gridDim.x) {     // (pair: the odd tile past the end loads zeros)
                const int ow0 = (tile % p.tiles_x) * p.BW, oh0 = ((tile / p.tiles_x) % p.tiles_y) * p.BH, n = tile / tiles_per_img;
                for (int ks = 0; ks < ksteps; ++ks, ++q) {
                    const int s = q % nst, use = q / nst;
                    mbar_wait(&bars[kStages + s], (use & 1) ^ 1);
                    const int kb = ks / p.ntaps, t = ks % p.ntaps, tap = p.ttap[t];
                    const int ix = ow0 * p.stride + p.tdx[t] + p.x_pad, iy = oh0 * p.stride + p.tdy[t] + p.x_pad;
                    uint8_t* sa = smem + s * stage_bytes;
                    if (PAIR) {
                        const uint32_t fbar = mapa_u32(smem_u32(&bars[s]), 0);
                        if (rank == 0) mbar_expect_tx(&bars[s], (uint32_t)(4 * (A_BYTES + b_bytes)));
                        tma_load_4d_pair(sa, &map_x, fbar, kb * KBLK, ix, iy, n);
                        tma_load_3d_pair(sa + b_off, &map_w, fbar, kb * KBLK, (int)rank * brows, tap);
                        tma_load_4d_pair(sa + al_off, &map_xl, fbar, kb * KBLK, ix, iy, n);
                        tma_load_3d_pair(sa + bl_off, &map_wl, fbar, kb * KBLK, (int)rank * brows, tap);
                        continue;
                    }
                    mbar_expect_tx(&bars[s], (uint32_t)((A_BYTES + b_bytes) * (p.split ? 2 : 1)));
                    tma_load_4d(sa, &map_x, &bars[s], kb * KBLK, ix, iy, n);
                    tma_load_3d(sa + b_off, &map_w, &bars[s], kb * KBLK, 0, tap);
                    if (p.split) {
                        tma_load_4d(sa + al_off, &map_xl, &bars[s], kb * KBLK, ix, iy, n);
                        tma_load_3d(sa + bl_off, &map_wl, &bars[s], kb * KBLK, 0, tap);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (pair: the leader) =====
            const uint32_t idesc = make_idesc(PAIR ? 256 : 128, p.npad);
            auto mma = [](uint32_t d, uint64_t a_, uint64_t b_, uint32_t id, uint32_t acc) {
                if (PAIR) umma_bf16_pair(d, a_, b_, id, acc);
                else umma_bf16(d, a_, b_, id, acc);
            };
            auto commit = [](uint64_t* bar) {
                if (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            const int split = p.split, ntiles = p.ntiles, npad = p.npad;
            const uint64_t dA0 = make_desc<ROWB>(base), dB0 = make_desc<ROWB>(base + b_off);
            const uint32_t sstep = (uint32_t)stage_bytes >> 4, lo_a = (uint32_t)al_off >> 4, lo_b = (uint32_t)(bl_off - b_off) >> 4;
            int q = 0, it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                if (PAIR) mbar_wait_cluster(&bars[2 * kStages + 2 + a], ((it >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained it
                else mbar_wait(&bars[2 * kStages + 2 + a], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(a * npad);
                for (int ks = 0; ks < ksteps; ++ks, ++q) {
                    const int s = q % nst, use = q / nst;
                    mbar_wait(&bars[s], use & 1);
                    tc_fence_after();
                    const uint64_t da = dA0 + (uint64_t)((uint32_t)s * sstep), db = dB0 + (uint64_t)((uint32_t)s * sstep);
#pragma unroll
                    for (int k = 0; k < KBLK / 16; ++k)    // advance 16 bf16 = 32 bytes inside the swizzled row: +2 in the >>4 address field
                        mma(tacc, da + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
                    if (split) {                           // + a_lo*b_hi + a_hi*b_lo  (a_lo*b_lo ~ 2^-18 is dropped)
#pragma unroll
                        for (int k = 0; k < KBLK / 16; ++k) mma(tacc, da + lo_a + 2 * k, db + 2 * k, idesc, 1);
#pragma unroll
                        for (int k = 0; k < KBLK / 16; ++k) mma(tacc, da + 2 * k, db + lo_b + 2 * k, idesc, 1);
                    }
                    commit(&bars[kStages + s]);            // slot free once these MMAs have read it
                }
                commit(&bars[2 * kStages + a]);            // accumulator complete
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                      // GEMM row = pixel index in the block
        int it = 0;
        const uint32_t tmem_empty_remote = PAIR ? mapa_u32(smem_u32(&bars[2 * kStages + 2]), 0) : 0u;
        for (int tile = blockIdx.x; tile - (int)rank < p.ntiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const int ow0 = (tile % p.tiles_x) * p.BW, oh0 = ((tile / p.tiles_x) % p.tiles_y) * p.BH, n = tile / tiles_per_img;
            const int oh = oh0 + m / p.BW, ow = ow0 + m % p.BW;
            const bool valid = oh < p.Ho && ow < p.Wo && tile < p.ntiles;
            mbar_wait(&bars[2 * kStages + a], (it >> 1) & 1);
            tc_fence_after();
            const int qh = oh * p.omul + p.oay, qw = ow * p.omul + p.oax;
            const long long pix = p.out_pad ? ((long long)n * (p.OutH + 2) + qh + 1) * (p.OutW + 2) + qw + 1
                                            : ((long long)n * p.OutH + qh) * p.OutW + qw;
            for (int c0 = 0; c0 < p.npad; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.npad + c0), v);
                if (!valid) continue;
                epi_bias_act16(v, sbias + c0, p.act);
                if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_pix_stride + p.out_ch_off + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.cout) o[j] = p.accum ? o[j] + v[j] : v[j];
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_pix_stride + p.out_ch_off + c0;
                    __nv_bfloat16* ol = p.out_lo ? reinterpret_cast<__nv_bfloat16*>(p.out_lo) + pix * p.out_pix_stride + p.out_ch_off + c0 : nullptr;
                    uint4 h[2], l[2];
                    epi_pack_hi(v, h);
                    if (ol) epi_pack_lo(v, h, l);
                    if (c0 + 16 <= p.cout && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                        reinterpret_cast<uint4*>(o)[0] = h[0];
                        reinterpret_cast<uint4*>(o)[1] = h[1];
                        if (ol) {
                            reinterpret_cast<uint4*>(ol)[0] = l[0];
                            reinterpret_cast<uint4*>(ol)[1] = l[1];
                        }
                    } else {
                        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(h);
                        const __nv_bfloat16* ls = reinterpret_cast<const __nv_bfloat16*>(l);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < p.cout) {
                                o[j] = hs[j];
                                if (ol) ol[j] = ls[j];
                            }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                             // this warp has drained accumulator a
                if (PAIR) mbar_arrive_cluster(tmem_empty_remote + (uint32_t)(a * 8));
                else mbar_arrive(&bars[2 * kStages + 2 + a]);
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        if (PAIR) tmem_dealloc_pair(tmem_base, ncols);
        else tmem_dealloc(tmem_base, ncols);
    }
}

// ---- v2: stride-1 kernel with tap reuse in shared memory, persistent CTAs and double-buffered TMEM accumulators ------
// The 9-tap implicit GEMM above re-reads every input pixel 9 times from L2 (measured: L2-bandwidth bound, 188 TFLOP/s).
// Here activations live in a ZERO-BORDERED NHWC layout [N][H+2][W+2][C]: with the border in memory a tap (u,v) is a pure
// shift of (u-1)*(W+2)+(v-1) in the flat pixel index, so an M tile is ANY 128 consecutive flat pixels (no tile
// quantisation) and one TMA row segment of 130 pixels per vertical tap u serves the three horizontal taps: the MMA
// descriptors of v = 0,1,2 simply start 0,1,2 rows into the segment (descriptor base_offset carries the swizzle phase).
// Pipeline step = (64-channel block, u): A segment (+ low half) and the three weight tiles of taps (u,0..2) (+ low halves);
// 12 (36 in split mode) UMMAs per step.  CTAs are persistent over M tiles; the accumulator alternates between two TMEM
// buffers so the epilogue of tile i overlaps the MMAs of tile i+1.  Border pixels of the output are written as zeros.
struct ConvParamsV2 {
    long long F;             // total flat pixels N*(H+2)*(W+2)
    int Hp, Wp;              // H+2, W+2
    int cin, npad, cout;
    int act, out_f32, split;
    long long out_pix_stride;
    int out_ch_off;
    void* out;
    void* out_lo;
    const float* bias;
    int ntiles;
    int out_zero_to;         // channels [cout, out_zero_to) of the output slice are written as zeros (keeps padded inputs exact)
    int resident;            // 1: all weight tiles stay in shared memory for the CTA's lifetime
    int nstages;             // 2..8 pipeline stages
    int tma_store;           // 1: bf16 output tile staged in shared memory and written by TMA bulk stores (slice width 64); 2: fp32 output
                             // (<= 64 channels) staged as two 32-channel boxes, written by bulk stores or, with accum, bulk reduce-adds
    int dbg;                 // timing experiments: 1 = epilogue skips its global stores, 2 = epilogue skipped entirely
    int taps1;               // 1: only the centre tap exists (a 1x1 convolution / plain GEMM over the flat pixels)
    int tap_mask;            // bit 3u+v set = tap (u,v) has non-zero weights; rows of taps that are all clear are never loaded
    int accum;               // fp32 output: out += result (a gradient buffer with several producers)
    int ktail;               // 16-channel K steps of the LAST 64-channel block that hold real channels (1..4): cin = 64*(kblocks-1) + 16*ktail
    int tri;                 // CTA pair, npad >= 80: three MMAs of N = npad per product (hi*hi, hi*lo, lo*hi) into ONE accumulator; each CTA
                             // holds both weight planes of HALF the output channels (see the kernel)
    int ostage_blocks;       // bf16 bulk stores: staging slots of one 64-channel column block each (2 = used in turn with one store group per block, 1 = staged and drained in turn)
    int pf;                  // L2 prefetch distance in tiles of this CTA (0 = off): the segments of tile + pf * grid are requested early
    int hstack;              // 1: the three horizontal taps of a tap row are ONE MMA of N = 3 * (accumulator width) (see the kernel)
};

constexpr int kSegRows = 130;
constexpr int kSegBytes = ((kSegRows * 128 + 1023) / 1024) * 1024;   // 17408
constexpr int kStagesV2 = 2;

__device__ __forceinline__ uint64_t make_desc_sw128_off(uint32_t smem_addr, int boff_mode) {
    // Start address that is NOT aligned to the 1024-byte swizzle repeat (a row shift of v*128 bytes).  Measured on B200: the
    // UMMA applies the 128-byte swizzle to ABSOLUTE shared-memory address bits, so the descriptor's base_offset field must
    // stay 0 (mode 0: exact results for all nine taps; mode 1 = address bits [7,10) in base_offset: taps with v != 0 wrong).
    const uint64_t base_off = boff_mode ? ((smem_addr >> 7) & 7) : 0;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (base_off << 49) |
           (2ull << 61);
}

// PAIR: the kernel runs as clusters of two CTAs (the two SMs of a TPC) that execute every MMA together (tcgen05.mma.cta_group::2,
// M = 256 = the two CTAs' 128-pixel tiles).  Split mode only: CTA 0 holds the W_hi tile of a tap and CTA 1 the W_lo tile -- the two
// halves of the N = 2 * npad operand -- so a CTA keeps HALF the weight bytes (layers with one 64-channel block and 64 output channels
// become resident) and fetches half of B per MMA: N = 128 costs 64 cycles for 256 pixels against 102 for 128 (tc_common.cuh).  Both
// CTAs' TMA loads count on the leader's `full` barrier; the leader's tcgen05.commit multicasts to both CTAs' `empty` / `tmem_full`
// barriers; both CTAs' epilogue warps arrive on the leader's `tmem_empty`.
template <bool PAIR>
__global__ void __launch_bounds__(320) conv3x3_s1_tc_v2_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w,
                                                               const __grid_constant__ CUtensorMap map_xl,
                                                               const __grid_constant__ CUtensorMap map_wl,
                                                               const __grid_constant__ CUtensorMap map_o,
                                                               const __grid_constant__ CUtensorMap map_ol, ConvParamsV2 p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int kblocks = (p.cin + 63) / 64;
    // Weight tiles per 64-channel block: only the taps of the mask exist in shared memory (the parity blocks of a stride-2 data gradient
    // have 1, 2 or 4 live taps of 9) -- except in the stacked form, whose MMA spans the three adjacent tiles of a tap row.
    const bool compact = !p.taps1 && !p.hstack;
    const uint32_t tmask = (uint32_t)p.tap_mask;
    const int tpk = p.taps1 ? 1 : (compact ? __popc(tmask) : 9);
    const int tps = p.taps1 ? 1 : 3;                  // ... per pipeline step (at most)
    // active tap rows (the quad forms of the stride-2 data gradients have whole rows / columns of zero taps)
    int rows[3], nrows = 0;
    for (int u = 0; u < 3; ++u)
        if (p.taps1 ? u == 1 : ((p.tap_mask >> (3 * u)) & 7) != 0) rows[nrows++] = u;
    const int steps = kblocks * nrows;
    const int nst = p.nstages;
    const int b_tile = p.npad * 128;                  // one tap's weight tile (multiple of 1024: npad % 8 == 0)
    const int w_all = tpk * kblocks * b_tile;         // every weight tile of the layer (resident mode)
    // Split (bf16x3) mode: the hi and lo tiles of a tap sit back to back ([hi rows | lo rows]) so ONE B descriptor of
    // N = 2*npad rows covers both: A_hi x [W_hi | W_lo] and A_lo x [W_hi | W_lo] = 2 MMAs per (tap, 16 channels) into two
    // accumulator halves (W_hi products | W_lo products) that the epilogue adds.  Measured (tools/micro/mma_rate.cu): an
    // M = 128 MMA with shared-memory operands costs max(91, 38 + N/2) cycles whatever it computes, so 2 MMAs of N <= 128 beat
    // the 3 MMAs (hi*hi, lo*hi, hi*lo) of the first version by ~1.45x, and the lo*lo term comes for free.
    // shared memory: [resident weights: tile (kb, tap) = hi | lo] then nst stages of [A seg hi | A seg lo | B tap 0..2 (hi | lo)]
    const int mult = p.split ? 2 : 1;
    const int wmult = PAIR ? 1 : mult;                // weight tiles of a tap held by THIS CTA (pair: hi in CTA 0, lo in CTA 1)
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int res_bytes = p.resident ? w_all * wmult : 0;
    const int stage_bytes = mult * kSegBytes + (p.resident ? 0 : wmult * tps * b_tile);
    const int b_off = mult * kSegBytes;               // weights of a streamed stage
    uint8_t* stages = smem + res_bytes;
    // output staging for the TMA store: one (two in split mode) 128-pixel x 64-channel bf16 tile, 128-byte swizzled
    uint8_t* ostage = stages + nst * stage_bytes;
    const int oblk_bytes = 16384 * (p.out_lo ? 2 : 1);
    const int ostage_bytes = p.tma_store == 2 ? 32768 : (p.tma_store ? oblk_bytes * p.ostage_blocks : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + ostage_bytes);
    // bars: full[8], empty[8], tmem_full[2], tmem_empty[2], weights
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
    float* xch = reinterpret_cast<float*>(bars + 32);     // hstack: rows 0, 1 of every epilogue warp for its neighbour, [2][2][4][48]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Epilogue warps: 4 (192 threads) or 8 (320 threads).  A warp may only read the TMEM lane quarter (warp % 4); with 8 warps the two
    // warps of a quarter take alternate 16-column chunks (short-K layers were bound by the four warps' TMEM loads + conversions).
    // (12 warps -- three per quarter -- were measured for layers of five or more chunks: no change, 371 vs 363 us for conv1's GEMM.)
    const int nhalf = blockDim.x > 192 ? 2 : 1;
    const int epi_threads = 128 * nhalf;
    // Pair form, wide layers (tri): the [W_hi | W_lo] operand makes A_lo x [W_hi | W_lo] compute lo*lo too -- free while an instruction's
    // cost did not depend on N, a quarter of the tensor-pipe time now that the pair MMA costs N/2 cycles.  Here CTA r holds both planes of
    // the output channels [r npad/2, (r+1) npad/2) and the three wanted products are three MMAs of N = npad accumulating into the SAME
    // npad columns (3 x 64 cycles instead of 2 x 128 at npad = 128; half the TMEM columns; no hi/lo add in the epilogue).
    const int tri = PAIR ? p.tri : 0;
    const int acc_cols = tri ? p.npad : mult * p.npad;   // accumulator columns of one tap (split: W_hi products | W_lo products)
    // hstack (thin layers: 3 * acc_cols <= 256).  An MMA costs max(91, 38 + N/2) cycles whatever it computes, so with Cout <= 32 the
    // three horizontal taps (N = 64 each) run at a quarter of the tensor pipe's rate.  Their A operands are the SAME segment shifted
    // by v rows; instead of shifting A, one MMA of N = 3 * acc_cols multiplies rows 0..127 of the segment with [W(u,0) | W(u,1) |
    // W(u,2)] (the tap tiles are adjacent in shared memory) and the epilogue adds the three column blocks across lanes:
    // out[m0 + t] = D0[t] + D1[t + 1] + D2[t + 2].  A tile then yields 126 output pixels (tiles advance by 126 flat pixels).
    const int hst = p.hstack;
    const int acc_all = hst ? 3 * acc_cols : acc_cols;   // one accumulator buffer
    const int tile_m = hst ? 126 : 128;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * acc_all) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[8 + s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars[16 + a], 1);
            mbar_init(&bars[18 + a], (PAIR ? 8 : 4) * nhalf);
        }
        mbar_init(&bars[20], 1);
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[128];
    if (threadIdx.x < 128) sbias[threadIdx.x] = (p.bias && (int)threadIdx.x < p.cout) ? p.bias[threadIdx.x] : 0.f;
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair(tmem_slot, ncols);
        else tmem_alloc(tmem_slot, ncols);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // the peer's barriers must be initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Tiles: CTA b handles tiles b, b + grid, ...; in a pair (grid even, rank = b & 1) both CTAs run the same number of iterations (the odd
    // tile past the end is all out of range: its loads are zero-filled, its stores clipped).
    const int ntiles = p.ntiles;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            if (PAIR) {
                const CUtensorMap* mapw = rank ? &map_wl : &map_w;
                const uint32_t wbar = mapa_u32(smem_u32(&bars[20]), 0);
                if (p.resident) {
                    if (rank == 0) mbar_expect_tx(&bars[20], (uint32_t)(2 * res_bytes));
                    for (int kb = 0; kb < kblocks; ++kb)
                        for (int t = 0, slot = 0; t < (p.taps1 ? 1 : 9); ++t) {
                            if (compact && !((tmask >> t) & 1)) continue;
                            uint8_t* wt = smem + (kb * tpk + slot) * b_tile;
                            if (tri) {      // rows [rank * npad/2, +npad/2) of both planes: [hi half | lo half]
                                tma_load_3d_pair(wt, &map_w, wbar, kb * 64, (int)rank * (p.npad / 2), t);
                                tma_load_3d_pair(wt + (p.npad / 2) * 128, &map_wl, wbar, kb * 64, (int)rank * (p.npad / 2), t);
                            } else {
                                tma_load_3d_pair(wt, mapw, wbar, kb * 64, 0, p.taps1 ? 4 : t);
                            }
                            ++slot;
                        }
                }
                int q = 0;
                for (int tile = blockIdx.x; tile - (int)rank < ntiles; tile += gridDim.x) {
                    const long long m0 = (long long)tile * tile_m;
                    for (int st = 0; st < steps; ++st, ++q) {
                        const int s = q % nst, use = q / nst;
                        const int kb = st / nrows, u = rows[st % nrows];
                        mbar_wait(&bars[8 + s], (use & 1) ^ 1);
                        uint8_t* sa = stages + s * stage_bytes;
                        const uint32_t fbar = mapa_u32(smem_u32(&bars[s]), 0);
                        const uint32_t rmask = p.taps1 ? 1u : (compact ? (tmask >> (3 * u)) & 7u : 7u);     // weight tiles of this step
                        if (rank == 0) mbar_expect_tx(&bars[s], (uint32_t)(2 * (2 * kSegRows * 128 + (p.resident ? 0 : __popc(rmask) * b_tile))));
                        const long long f0 = m0 + (long long)(u - 1) * p.Wp - 1;
                        tma_load_2d_pair(sa, &map_x, fbar, kb * 64, (int)f0);
                        tma_load_2d_pair(sa + kSegBytes, &map_xl, fbar, kb * 64, (int)f0);
                        if (p.pf && tile + p.pf * (int)gridDim.x < ntiles) {
                            const long long fp = f0 + (long long)p.pf * gridDim.x * tile_m;
                            tma_prefetch_2d(&map_x, kb * 64, (int)fp);
                            tma_prefetch_2d(&map_xl, kb * 64, (int)fp);
                        }
                        if (!p.resident)
                            for (int v = 0, slot = 0; v < tps; ++v) {
                                if (!((rmask >> v) & 1)) continue;
                                uint8_t* wt = sa + b_off + slot * b_tile;
                                if (tri) {
                                    tma_load_3d_pair(wt, &map_w, fbar, kb * 64, (int)rank * (p.npad / 2), 3 * u + v);
                                    tma_load_3d_pair(wt + (p.npad / 2) * 128, &map_wl, fbar, kb * 64, (int)rank * (p.npad / 2), 3 * u + v);
                                } else {
                                    tma_load_3d_pair(wt, mapw, fbar, kb * 64, 0, p.taps1 ? 4 : 3 * u + v);
                                }
                                ++slot;
                            }
                    }
                }
            } else {
            if (p.resident) {      // all weight tiles once: they stay in shared memory for every tile of this CTA
                mbar_expect_tx(&bars[20], (uint32_t)res_bytes);
                for (int kb = 0; kb < kblocks; ++kb)
                    for (int t = 0, slot = 0; t < (p.taps1 ? 1 : 9); ++t) {
                        if (compact && !((tmask >> t) & 1)) continue;
                        uint8_t* wt = smem + (kb * tpk + slot) * (mult * b_tile);
                        const int tap = p.taps1 ? 4 : t;
                        tma_load_3d(wt, &map_w, &bars[20], kb * 64, 0, tap);
                        if (p.split) tma_load_3d(wt + b_tile, &map_wl, &bars[20], kb * 64, 0, tap);
                        ++slot;
                    }
            }
            int q = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const long long m0 = (long long)tile * tile_m;
                for (int st = 0; st < steps; ++st, ++q) {
                    const int s = q % nst, use = q / nst;
                    const int kb = st / nrows, u = rows[st % nrows];
                    mbar_wait(&bars[8 + s], (use & 1) ^ 1);
                    uint8_t* sa = stages + s * stage_bytes;
                    const uint32_t rmask = p.taps1 ? 1u : (compact ? (tmask >> (3 * u)) & 7u : 7u);         // weight tiles of this step
                    mbar_expect_tx(&bars[s], (uint32_t)((kSegRows * 128 + (p.resident ? 0 : __popc(rmask) * b_tile)) * mult));
                    const long long f0 = m0 + (long long)(u - 1) * p.Wp - 1;
                    tma_load_2d(sa, &map_x, &bars[s], kb * 64, (int)f0);
                    if (p.split) tma_load_2d(sa + kSegBytes, &map_xl, &bars[s], kb * 64, (int)f0);
                    if (p.pf && tile + p.pf * (int)gridDim.x < ntiles) {
                        const long long fp = f0 + (long long)p.pf * gridDim.x * tile_m;
                        tma_prefetch_2d(&map_x, kb * 64, (int)fp);
                        if (p.split) tma_prefetch_2d(&map_xl, kb * 64, (int)fp);
                    }
                    if (!p.resident)
                        for (int v = 0, slot = 0; v < tps; ++v) {
                            if (!((rmask >> v) & 1)) continue;
                            uint8_t* wt = sa + b_off + slot * (mult * b_tile);
                            const int tap = p.taps1 ? 4 : 3 * u + v;
                            tma_load_3d(wt, &map_w, &bars[s], kb * 64, 0, tap);
                            if (p.split) tma_load_3d(wt + b_tile, &map_wl, &bars[s], kb * 64, 0, tap);
                            ++slot;
                        }
                }
            }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc(PAIR ? 256 : 128, acc_all);
            auto mma = [](uint32_t d, uint64_t a_, uint64_t b_, uint32_t id, uint32_t acc) {
                if (PAIR) umma_bf16_pair(d, a_, b_, id, acc);
                else umma_bf16(d, a_, b_, id, acc);
            };
            auto commit = [](uint64_t* bar) {
                if (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            const uint32_t sbase = base + res_bytes;
            // Everything that does not change from MMA to MMA is computed once: the asm of an MMA clobbers "memory", so the compiler
            // re-read kernel parameters and rebuilt both 64-bit descriptors after every instruction -- ~25 scalar instructions per MMA,
            // which at two K steps per tap (K = 32 layers) made the issuing thread the bottleneck (176 cycles per MMA, ncu source view).
            // A descriptor's low field is the shared-memory address / 16, so a byte offset is one 64-bit add of offset / 16.
            const int taps1 = p.taps1, tap_mask = p.tap_mask, split = p.split, resident = p.resident, ktail = p.ktail;
            const uint64_t dA0 = make_desc_sw128_off(sbase, 0);
            const uint64_t dB0 = make_desc<128>(resident ? base : sbase + b_off);
            const uint32_t tri_lo = (uint32_t)((p.npad / 2) * 128) >> 4;
            const uint32_t wstep = (uint32_t)(wmult * b_tile) >> 4, lo_off = (uint32_t)kSegBytes >> 4, sstep = (uint32_t)stage_bytes >> 4;
            if (resident) {
                mbar_wait(&bars[20], 0);
                tc_fence_after();
            }
            int q = 0, it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {      // (leader: rank 0)
                const int a = it & 1;
                if (PAIR) mbar_wait_cluster(&bars[18 + a], ((it >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this accumulator
                else mbar_wait(&bars[18 + a], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(a * acc_all);
                uint32_t accumulate = 0;
                for (int st = 0; st < steps; ++st, ++q) {
                    const int s = q % nst, use = q / nst;
                    const int kb = st / nrows, u = rows[st % nrows];
                    mbar_wait(&bars[s], use & 1);
                    tc_fence_after();
                    const uint32_t soff = (uint32_t)s * sstep;
                    const int kn = kb == kblocks - 1 ? ktail : 4;      // a 129-channel input is 2 blocks + ONE K step, not 3 blocks
                    const int vmask = taps1 ? 2 : ((tap_mask >> (3 * u)) & 7);     // the centre tap reads the segment one row in
                    if (hst) {       // one MMA per K step over the three adjacent tap tiles (masked taps hold zero weights)
                        const uint64_t da = dA0 + (uint64_t)soff;
                        const uint32_t widx = resident ? (uint32_t)(kb * tpk + 3 * u) : 0u;
                        const uint64_t db = dB0 + (uint64_t)((resident ? 0u : soff) + widx * wstep);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < kn) mma(tacc, da + 2 * k, db + 2 * k, idesc, accumulate | (uint32_t)k);
                        accumulate = 1;
                        if (split) {
                            const uint64_t dl = da + lo_off;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < kn) mma(tacc, dl + 2 * k, db + 2 * k, idesc, 1);
                        }
                        commit(&bars[8 + s]);
                        continue;
                    }
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        if (!((vmask >> v) & 1)) continue;
                        const uint64_t da = dA0 + (uint64_t)(soff + (uint32_t)(v * 8));            // + v * 128 bytes
                        // tile slot: resident = live taps below (u, v) of the whole mask; streamed = live taps below v of this row
                        const uint32_t below = taps1 ? 0u : (compact ? (uint32_t)__popc(resident ? tmask & ((1u << (3 * u + v)) - 1u)
                                                                                                       : (uint32_t)vmask & ((1u << v) - 1u))
                                                                      : (uint32_t)(resident ? 3 * u + v : v));
                        const uint32_t widx = resident ? (uint32_t)(kb * tpk) + below : below;
                        const uint64_t db = dB0 + (uint64_t)((resident ? 0u : soff) + widx * wstep);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < kn) mma(tacc, da + 2 * k, db + 2 * k, idesc, accumulate | (uint32_t)k);
                        accumulate = 1;
                        if (tri) {          // hi*lo and lo*hi into the same columns (db = the W_hi half, + tri_lo = the W_lo half)
                            const uint64_t dl = da + lo_off, dbl = db + tri_lo;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < kn) mma(tacc, da + 2 * k, dbl + 2 * k, idesc, 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < kn) mma(tacc, dl + 2 * k, db + 2 * k, idesc, 1);
                        } else if (split) {
                            const uint64_t dl = da + lo_off;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < kn) mma(tacc, dl + 2 * k, db + 2 * k, idesc, 1);
                        }
                    }
                    commit(&bars[8 + s]);
                }
                commit(&bars[16 + a]);
            }
        }
    } else {
        // ===== epilogue warps 2..5 =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        int it = 0;
        uint32_t cc = 0;                                   // hstack: exchange-buffer parity, one per accumulator chunk
        uint32_t gblk = 0;                                 // column blocks staged so far by this CTA (staging slot = gblk & 1)
        const uint32_t tmem_empty_remote = PAIR ? mapa_u32(smem_u32(&bars[18]), 0) : 0u;
        // accumulator columns of horizontal tap v, W_hi / W_lo products: single CTA [v][hi | lo], pair [hi: v0 v1 v2 | lo: v0 v1 v2]
        const uint32_t vstr = (uint32_t)(PAIR ? p.npad : acc_cols), hstr = (uint32_t)(PAIR && hst ? 3 * p.npad : p.npad);
        for (int tile = blockIdx.x; tile - (int)rank < ntiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const long long m = (long long)tile * tile_m + quarter * 32 + lane;
            const bool inrange = m < p.F && quarter * 32 + lane < tile_m;
            // (host: F + 256 < 2^31 -- the 64-bit divisions of the first version were ~6 % of the epilogue's issue slots on thin-K layers)
            const unsigned mu = (unsigned)m, wq = mu / (unsigned)p.Wp;
            const int wp = (int)(mu - wq * (unsigned)p.Wp), hp = (int)(wq % (unsigned)p.Hp);
            const bool border = wp == 0 || wp == p.Wp - 1 || hp == 0 || hp == p.Hp - 1;
            mbar_wait(&bars[16 + a], (it >> 1) & 1);
            tc_fence_after();
            const int c_end = p.dbg == 2 ? 0 : (p.out_zero_to > p.npad ? p.out_zero_to : p.npad);
            // bf16 outputs are staged and stored one 64-channel column block at a time (an 80-channel slice = a 64- and a 16-channel box;
            // per-thread 16-byte stores at the pixel stride were measured at 3x the cost of the whole rest of a K = 48 layer)
            const int nblk = (p.tma_store && c_end > 64) ? (c_end + 63) >> 6 : 1;
            // bf16 bulk stores with two staging slots (ostage_blocks == 2): column blocks go through the slots in turn, every block is
            // its own bulk-store group, and a slot is only waited for when it comes round again -- the store of block g drains while
            // block g + 1 is staged (the single-phase form stalled all epilogue warps behind `wait_group.read 0` once per tile: 20 % of
            // the samples of conv1's K = 48 GEMM, whose epilogue is its critical path).  One slot: staged and drained in turn.
            const bool ring = p.tma_store == 1 && p.ostage_blocks >= 2;
            for (int blk = 0; blk < nblk; ++blk) {
            const int cb0 = p.tma_store ? blk * 64 : 0, cb1 = p.tma_store ? (c_end < cb0 + 64 ? c_end : cb0 + 64) : c_end;
            uint8_t* ost = ostage + (ring ? (gblk & 1u) * oblk_bytes : 0);
            if (ring) {
                if (gblk >= 2u) {
                    // the store that used this slot two blocks ago must have finished READING it (the latest group may still be in flight)
                    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");
                }
            } else if (p.tma_store && (it > 0 || blk > 0)) {
                // the previous bulk store must have finished READING the staging tile before it is overwritten
                if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");
            }
            for (int c0 = cb0 + 16 * half; c0 < cb1; c0 += 16 * nhalf) {
                float v[16];
                if (c0 < p.npad && hst) {
                    const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * acc_all + c0);
                    float s1[16], s2[16], t2[16];
                    tmem_ld16(tb, v);
                    tmem_ld16(tb + vstr, s1);
                    tmem_ld16(tb + 2 * vstr, s2);
                    if (p.split) {
                        tmem_ld16(tb + hstr, t2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += t2[j];
                        tmem_ld16(tb + vstr + hstr, t2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) s1[j] += t2[j];
                        tmem_ld16(tb + 2 * vstr + hstr, t2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) s2[j] += t2[j];
                    }
                    // rows 0 and 1 of this warp are rows 32, 33 of the warp below it
                    float* xq = xch + (((cc & 1) * 2 + half) * 4 + quarter) * 48;
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { xq[j] = s1[j]; xq[16 + j] = s2[j]; }
                    } else if (lane == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) xq[32 + j] = s2[j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        s1[j] = __shfl_down_sync(0xffffffffu, s1[j], 1);
                        s2[j] = __shfl_down_sync(0xffffffffu, s2[j], 2);
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");     // the four warps of this half
                    const float* xn = xch + (((cc & 1) * 2 + half) * 4 + ((quarter + 1) & 3)) * 48;
                    ++cc;
                    if (lane == 31) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { s1[j] = xn[j]; s2[j] = xn[32 + j]; }
                    } else if (lane == 30) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) s2[j] = xn[16 + j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += s1[j] + s2[j];
                } else if (c0 < p.npad) {
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * acc_cols + c0), v);
                    if (p.split && !tri) {      // + the W_lo products of the second accumulator half
                        float v2[16];
                        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * acc_cols + p.npad + c0), v2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += v2[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                if (!inrange || p.dbg == 1) continue;
                const int lim = p.out_zero_to > p.cout ? p.out_zero_to : p.cout;   // channels this layer owns
                if (c0 < p.npad) epi_bias_act16(v, sbias + c0, p.act);
                epi_mask16(v, c0, p.cout, border);
                if (p.tma_store == 2) {
                    // fp32 tile staged as 32-channel boxes of 128 rows x 128 bytes (16-byte chunks XOR-swizzled by row & 7): per-thread
                    // stores of a 256-byte pixel row are 32 sectors per instruction, the bulk store / reduce-add is one coalesced copy
                    const int row = quarter * 32 + lane;
                    uint8_t* rowp = ostage + ((c0 - cb0) >> 5) * 16384 + row * 128;
                    const int ch0 = (c0 & 31) >> 2;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<float4*>(rowp + (((ch0 + k) ^ (row & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    continue;
                }
                if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(p.out) + m * p.out_pix_stride + p.out_ch_off + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < lim) o[j] = p.accum ? o[j] + v[j] : v[j];
                    continue;
                }
                uint4 h[2], l[2];
                epi_pack_hi(v, h);
                if (p.out_lo) epi_pack_lo(v, h, l);
                if (p.tma_store) {
                    // stage the row in shared memory (16-byte chunks XOR-swizzled by row & 7, as the store map expects)
                    const int row = quarter * 32 + lane, ch0 = ((c0 - cb0) >> 3);        // two 16-byte chunks per 16 channels
                    uint8_t* rowp = ost + row * 128;
                    *reinterpret_cast<uint4*>(rowp + (((ch0) ^ (row & 7)) << 4)) = h[0];
                    *reinterpret_cast<uint4*>(rowp + (((ch0 + 1) ^ (row & 7)) << 4)) = h[1];
                    if (p.out_lo) {
                        *reinterpret_cast<uint4*>(rowp + 16384 + (((ch0) ^ (row & 7)) << 4)) = l[0];
                        *reinterpret_cast<uint4*>(rowp + 16384 + (((ch0 + 1) ^ (row & 7)) << 4)) = l[1];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.out_pix_stride + p.out_ch_off + c0;
                    __nv_bfloat16* ol = p.out_lo ? reinterpret_cast<__nv_bfloat16*>(p.out_lo) + m * p.out_pix_stride + p.out_ch_off + c0 : nullptr;
                    if (c0 + 16 <= lim && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                        reinterpret_cast<uint4*>(o)[0] = h[0];
                        reinterpret_cast<uint4*>(o)[1] = h[1];
                        if (ol) {
                            reinterpret_cast<uint4*>(ol)[0] = l[0];
                            reinterpret_cast<uint4*>(ol)[1] = l[1];
                        }
                    } else {
                        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(h);
                        const __nv_bfloat16* ls = reinterpret_cast<const __nv_bfloat16*>(l);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < lim) {
                                o[j] = hs[j];
                                if (ol) ol[j] = ls[j];
                            }
                    }
                }
            }
            if (blk == nblk - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                               // this warp has drained accumulator a
                    if (PAIR) mbar_arrive_cluster(tmem_empty_remote + (uint32_t)(a * 8));
                    else mbar_arrive(&bars[18 + a]);
                }
            }
            if (p.tma_store) {
                // all epilogue warps have staged their rows of this block -> one thread issues the coalesced bulk store(s)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");
                if (warp == 2 && lane == 0) {
                    const long long m0 = (long long)tile * tile_m;
                    if (p.tma_store == 2) {
                        for (int b = 0; cb0 + b * 32 < cb1 && b < 2; ++b) {      // (fp32 blocks are always staged in turn)
                            if (p.accum) tma_reduce_add_2d(&map_o, ostage + b * 16384, cb0 + b * 32, (int)m0);
                            else tma_store_2d(&map_o, ostage + b * 16384, cb0 + b * 32, (int)m0);
                        }
                    } else {
                        tma_store_2d(&map_o, ost, blk * 64, (int)m0);
                        if (p.out_lo) tma_store_2d(&map_ol, ost + 16384, blk * 64, (int)m0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            ++gblk;
            }
        }
        if (p.tma_store && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // the peer may not exit (or free TMEM) while MMAs read its shared memory
    else __syncthreads();
    if (warp == 1) {
        if (PAIR) tmem_dealloc_pair(tmem_base, ncols);
        else tmem_dealloc(tmem_base, ncols);
    }
}

// ---- layout glue ---------------------------------------------------------------------------------------------------
// NCHW fp32 (channel slice [c_off, c_off+C) of a tensor with Ctot channels) -> NHWC bf16 with Cpad channels (zero padded),
// written into channels [dst_off, dst_off+Cpad) of an NHWC tensor with dst_stride channels per pixel.
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                     __nv_bfloat16* __restrict__ dst_lo, int C, int Cpad,
                                                                     long long HW, long long src_batch_stride, int dst_stride,
                                                                     int dst_off, int W, int pad) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long pp = p0 + tx;
        tile[j][tx] = (c < C && pp < HW) ? src[(long long)n * src_batch_stride + (long long)c * HW + pp] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long long pp = p0 + j;
        const int c = c0 + tx;
        if (pp < HW && c < Cpad) {
            const long long H_ = HW / W;
            const long long dp = pad ? ((long long)n * (H_ + 2) + pp / W + 1) * (W + 2) + pp % W + 1 : (long long)n * HW + pp;
            const long long o = dp * dst_stride + dst_off + c;
            const __nv_bfloat16 hi = __float2bfloat16(tile[tx][j]);
            dst[o] = hi;
            if (dst_lo) dst_lo[o] = __float2bfloat16(tile[tx][j] - __bfloat162float(hi));
        }
    }
}

// NHWC (bf16 or fp32) channel slice -> NCHW fp32
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int C,
                                                                long long HW, int src_stride, int src_off, int W, int pad) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const long long pp = p0 + j;
        const int c = c0 + tx;
        const long long H_ = HW / W;
        const long long sp = pad ? ((long long)n * (H_ + 2) + pp / W + 1) * (W + 2) + pp % W + 1 : (long long)n * HW + pp;
        tile[j][tx] = (pp < HW && c < C) ? (float)src[sp * src_stride + src_off + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long pp = p0 + tx;
        if (c < C && pp < HW) dst[((long long)n * C + c) * HW + pp] = tile[tx][j];
    }
}

// Vectorised versions of the two conversions above for 8-channel aligned slices (the 32x32 transposes write 2-byte elements, 64 B
// per warp row, and launch 138 K CTAs at B=32, 228x304: 300-600 us per call).  CTA = 128 consecutive pixels of one image x one block
// of up to 64 channels, staged through a [64][129] fp32 tile: plane accesses are 512 B per warp, NHWC accesses 16 B per thread with
// the 8-channel groups of a pixel in adjacent lanes (one full 128-byte line per 64-channel pixel).
constexpr int kLT_P = 128, kLT_C = 64;

__device__ __forceinline__ long long nhwc_pixel(int n, long long pp, long long HW, int W, int pad) {
    if (!pad) return (long long)n * HW + pp;
    const long long H_ = HW / W;
    return ((long long)n * (H_ + 2) + pp / W + 1) * (W + 2) + pp % W + 1;
}

__global__ void __launch_bounds__(256) nchw_f32_to_nhwc_bf16_vec_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                         __nv_bfloat16* __restrict__ dst_lo, int C, int Cpad, long long HW,
                                                                         long long src_batch_stride, int dst_stride, int dst_off, int W,
                                                                         int pad) {
    __shared__ float tile[kLT_C][kLT_P + 1];
    const int n = blockIdx.z, c0 = blockIdx.y * kLT_C;
    const long long p0 = (long long)blockIdx.x * kLT_P;
    const int cb = Cpad - c0 < kLT_C ? Cpad - c0 : kLT_C;          // channels of this block (multiple of 8)
    for (int e = threadIdx.x; e < cb * kLT_P; e += 256) {
        const int c = e / kLT_P, t = e % kLT_P;
        const long long pp = p0 + t;
        tile[c][t] = (c0 + c < C && pp < HW) ? __ldg(src + (long long)n * src_batch_stride + (long long)(c0 + c) * HW + pp) : 0.f;
    }
    __syncthreads();
    const int groups = cb >> 3;
    for (int e = threadIdx.x; e < groups * kLT_P; e += 256) {
        const int g = e % groups, t = e / groups;
        const long long pp = p0 + t;
        if (pp >= HW) continue;
        uint4 h, l;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
        __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = tile[g * 8 + 2 * k][t], b = tile[g * 8 + 2 * k + 1][t];
            hh[k] = __floats2bfloat162_rn(a, b);
            ll[k] = __floats2bfloat162_rn(a - __low2float(hh[k]), b - __high2float(hh[k]));
        }
        const long long o = nhwc_pixel(n, pp, HW, W, pad) * dst_stride + dst_off + c0 + g * 8;
        *reinterpret_cast<uint4*>(dst + o) = h;
        if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + o) = l;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_vec_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                                                   int src_stride, int src_off, int W, int pad) {
    __shared__ float tile[kLT_C][kLT_P + 1];
    constexpr int VEC = 16 / sizeof(T);                             // channels per 16-byte load: 8 (bf16) or 4 (fp32)
    const int n = blockIdx.z, c0 = blockIdx.y * kLT_C;
    const long long p0 = (long long)blockIdx.x * kLT_P;
    const int cb = C - c0 < kLT_C ? C - c0 : kLT_C;                 // real channels of this block
    const int groups = (cb + VEC - 1) / VEC;                        // (reads may cover pad channels of the slice: stride permits)
    for (int e = threadIdx.x; e < groups * kLT_P; e += 256) {
        const int g = e % groups, t = e / groups;
        const long long pp = p0 + t;
        float v[VEC];
        if (pp < HW) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + nhwc_pixel(n, pp, HW, W, pad) * src_stride + src_off + c0 + g * VEC));
            if constexpr (sizeof(T) == 2) {
                const __nv_bfloat162* rr = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[2 * k] = __low2float(rr[k]); v[2 * k + 1] = __high2float(rr[k]); }
            } else {
                v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
            }
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (g * VEC + k < kLT_C) tile[g * VEC + k][t] = v[k];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cb * kLT_P; e += 256) {
        const int c = e / kLT_P, t = e % kLT_P;
        const long long pp = p0 + t;
        if (pp < HW) dst[((long long)n * C + c0 + c) * HW + pp] = tile[c][t];
    }
}

// fp32 [Cout, Cin, 3, 3] -> bf16 [9][npad][cin_pad]  (zero padded)
__global__ void prep_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                         __nv_bfloat16* __restrict__ out_lo, int cout, int cin, int npad, int cin_pad,
                                         int flip_transpose) {
    const long long total = 9LL * npad * cin_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_pad), co = (int)((i / cin_pad) % npad), tap = (int)(i / ((long long)cin_pad * npad));
        float v = 0.f;
        if (!flip_transpose) {
            if (co < cout && ci < cin) v = w[((long long)co * cin + ci) * 9 + tap];
        } else {   // data-gradient weights: out "channel" = original ci, in "channel" = original co; taps flipped (mode 1:
                   // stride-1 data gradient = a convolution of gy) or kept (mode 2: per-tap sums of the stride-2 data gradient)
            if (co < cin && ci < cout) v = w[((long long)ci * cin + co) * 9 + (flip_transpose == 1 ? 8 - tap : tap)];
        }
        const __nv_bfloat16 hi = __float2bfloat16(v);
        out[i] = hi;
        if (out_lo) out_lo[i] = __float2bfloat16(v - __bfloat162float(hi));
    }
}

// Several weight tensors in ONE launch (a training step prepares ~28 of them: every layer's forward weights, and the
// transposed / flipped data-gradient weights per output-channel chunk).  Item = channel slice [ci_off, ci_off+ci_cnt) of the
// input channels of a [cout][cin][3][3] tensor, converted like prep_conv_weights_kernel does for a contiguous tensor.
struct WPrepBatch {
    adc_wprep_item_t it[ADC_WPREP_MAX];
};
__global__ void __launch_bounds__(256) prep_conv_weights_batch_kernel(const __grid_constant__ WPrepBatch b) {
    const adc_wprep_item_t& t = b.it[blockIdx.y];
    const int oc = t.mode ? t.ci_cnt : t.cout, npad = (oc + 15) / 16 * 16;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(t.out);
    __nv_bfloat16* out_lo = reinterpret_cast<__nv_bfloat16*>(t.out_lo);
    const long long total = 9LL * npad * t.cin_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % t.cin_pad), co = (int)((i / t.cin_pad) % npad), tap = (int)(i / ((long long)t.cin_pad * npad));
        float v = 0.f;
        if (!t.mode) {
            if (co < t.cout && ci < t.ci_cnt) v = t.w[((long long)co * t.cin + t.ci_off + ci) * 9 + tap];
        } else {
            if (co < t.ci_cnt && ci < t.cout) v = t.w[((long long)ci * t.cin + t.ci_off + co) * 9 + (t.mode == 1 ? 8 - tap : tap)];
        }
        const __nv_bfloat16 hi = __float2bfloat16(v);
        out[i] = hi;
        if (out_lo) out_lo[i] = __float2bfloat16(v - __bfloat162float(hi));
    }
}

template <int KBLK>
static int launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mxl, const CUtensorMap& mwl, const ConvParams& p,
                  int smem, bool pair, cudaStream_t st) {
    static bool done = false;
    if (!done) {   // opt in once to the largest configuration of this instantiation (npad = 256)
        int max_smem = 2 * kStages * ((128 * KBLK * 2 + 256 * KBLK * 2 + 1023) & ~1023) + 1024 + 256;
        if (max_smem > kSmemCap) max_smem = kSmemCap;
        int rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_fprop_tc_kernel<KBLK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem),
                                 "cudaFuncSetAttribute(conv3x3_fprop_tc_kernel)");
        if (rc) return rc;
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_fprop_tc_kernel<KBLK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem),
                             "cudaFuncSetAttribute(conv3x3_fprop_tc_kernel<pair>)");
        if (rc) return rc;
        done = true;
    }
    if (smem > kSmemCap) {
        adc::set_error("conv3x3_fprop_tc: %d bytes of shared memory needed (Cout too large for the split mode)", smem);
        return ADC_E_UNSUPP;
    }
    adc::Prof prof(st, p.omul == 2 ? adc::KIND_TC_DGRAD_S2 : adc::KIND_TC_FPROP_V1,
                   2.0 * p.N * p.Ho * p.Wo * p.ntaps * (double)p.cin * p.cout);
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    if (pair) {
        const int pairs = (p.ntiles + 1) / 2, max_pairs = sms / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
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
        return adc::check_cuda(cudaLaunchKernelEx(&cfg, conv3x3_fprop_tc_kernel<KBLK, true>, mx, mw, mxl, mwl, p),
                               "conv3x3_fprop_tc_kernel<pair> launch");
    }
    conv3x3_fprop_tc_kernel<KBLK, false><<<p.ntiles < sms ? p.ntiles : sms, 192, smem, st>>>(mx, mw, mxl, mwl, p);
    return adc::check_cuda(cudaGetLastError(), "conv3x3_fprop_tc_kernel launch");
}

}  // namespace tc

using namespace tc;

extern "C" {

int adc_nchw_to_nhwc_bf16(const float* src, void* dst, void* dst_lo, int N, int C, int Cpad, int H, int W,
                          long long src_batch_stride, int dst_stride, int dst_off, int pad, adc_stream_t stream) {
    if (!src || !dst || N <= 0 || C <= 0 || Cpad < C || H <= 0 || W <= 0 || dst_stride < dst_off + Cpad) {
        adc::set_error("adc_nchw_to_nhwc_bf16: bad argument");
        return ADC_E_BADARG;
    }
    const long long HW = (long long)H * W;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_LAYOUT, (double)N * HW * (4.0 * C + (dst_lo ? 4.0 : 2.0) * Cpad));
    if (Cpad % 8 == 0 && dst_off % 8 == 0 && dst_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 &&
        (!dst_lo || (reinterpret_cast<uintptr_t>(dst_lo) & 15) == 0)) {
        dim3 vgrid((unsigned)((HW + kLT_P - 1) / kLT_P), (Cpad + kLT_C - 1) / kLT_C, N);
        nchw_f32_to_nhwc_bf16_vec_kernel<<<vgrid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_lo, C, Cpad, HW,
                                                                                 src_batch_stride, dst_stride, dst_off, W, pad);
        return adc::check_cuda(cudaGetLastError(), "nchw_f32_to_nhwc_bf16_vec_kernel launch");
    }
    dim3 grid((unsigned)((HW + 31) / 32), (Cpad + 31) / 32, N);
    nchw_f32_to_nhwc_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_lo, C, Cpad, HW,
                                                                         src_batch_stride, dst_stride, dst_off, W, pad);
    return adc::check_cuda(cudaGetLastError(), "nchw_f32_to_nhwc_bf16_kernel launch");
}

int adc_nhwc_to_nchw_f32(const void* src, int src_is_f32, float* dst, int N, int C, int H, int W, int src_stride, int src_off,
                         int pad, adc_stream_t stream) {
    if (!src || !dst || N <= 0 || C <= 0 || H <= 0 || W <= 0 || src_stride < src_off + C) {
        adc::set_error("adc_nhwc_to_nchw_f32: bad argument");
        return ADC_E_BADARG;
    }
    const long long HW = (long long)H * W;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_LAYOUT, (double)N * HW * C * (4.0 + (src_is_f32 ? 4.0 : 2.0)));
    {   // vector path: 16-byte aligned slice whose 16-byte groups stay inside the pixel's stride
        const int vec = src_is_f32 ? 4 : 8;
        const int cround = (C + vec - 1) / vec * vec;
        if (src_off % vec == 0 && src_stride % vec == 0 && src_off + cround <= src_stride && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            dim3 vgrid((unsigned)((HW + kLT_P - 1) / kLT_P), (C + kLT_C - 1) / kLT_C, N);
            if (src_is_f32)
                nhwc_to_nchw_f32_vec_kernel<float><<<vgrid, 256, 0, (cudaStream_t)stream>>>((const float*)src, dst, C, HW, src_stride, src_off, W, pad);
            else
                nhwc_to_nchw_f32_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, C, HW,
                                                                                                 src_stride, src_off, W, pad);
            return adc::check_cuda(cudaGetLastError(), "nhwc_to_nchw_f32_vec_kernel launch");
        }
    }
    dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N);
    if (src_is_f32)
        nhwc_to_nchw_f32_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, dst, C, HW, src_stride, src_off, W, pad);
    else
        nhwc_to_nchw_f32_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, C, HW, src_stride,
                                                                                    src_off, W, pad);
    return adc::check_cuda(cudaGetLastError(), "nhwc_to_nchw_f32_kernel launch");
}

size_t adc_conv3x3_weight_elems(int cout, int cin_pad) {
    const int npad = (cout + 15) / 16 * 16;
    return (size_t)9 * npad * cin_pad;
}

int adc_conv3x3_prep_weights(const float* w, void* out_bf16, void* out_lo, int cout, int cin, int cin_pad, int flip_transpose,
                             adc_stream_t stream) {
    const int oc = flip_transpose ? cin : cout, ic = flip_transpose ? cout : cin;
    if (!w || !out_bf16 || cout <= 0 || cin <= 0 || cin_pad < ic || cin_pad % 16 != 0) {
        adc::set_error("adc_conv3x3_prep_weights: bad argument");
        return ADC_E_BADARG;
    }
    const int npad = (oc + 15) / 16 * 16;
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_PREP, 0.0);
    prep_conv_weights_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)out_lo, cout, cin, npad,
                                                                   cin_pad, flip_transpose);
    return adc::check_cuda(cudaGetLastError(), "prep_conv_weights_kernel launch");
}

int adc_conv3x3_prep_weights_batch(const adc_wprep_item_t* items, int n, adc_stream_t stream) {
    if (!items || n <= 0 || n > ADC_WPREP_MAX) {
        adc::set_error("adc_conv3x3_prep_weights_batch: 1..%d items", ADC_WPREP_MAX);
        return ADC_E_BADARG;
    }
    WPrepBatch b;
    memset(&b, 0, sizeof(b));
    for (int k = 0; k < n; ++k) {
        const adc_wprep_item_t& t = items[k];
        const int ic = t.mode ? t.cout : t.ci_cnt;
        if (!t.w || !t.out || t.cout <= 0 || t.cin <= 0 || t.ci_off < 0 || t.ci_cnt <= 0 || t.ci_off + t.ci_cnt > t.cin || t.cin_pad < ic ||
            t.cin_pad % 16 != 0 || t.mode < 0 || t.mode > 2) {
            adc::set_error("adc_conv3x3_prep_weights_batch: bad item %d", k);
            return ADC_E_BADARG;
        }
        b.it[k] = t;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_PREP, 0.0);
    prep_conv_weights_batch_kernel<<<dim3(24, n), 256, 0, (cudaStream_t)stream>>>(b);
    return adc::check_cuda(cudaGetLastError(), "prep_conv_weights_batch_kernel launch");
}

// One launch of the v1 kernel: a sum over a tap list of [128 pixels x Cin] x [Cin x Cout] products (see ConvParams).
struct V1Launch {
    const void *x, *x_lo; int x_stride, x_off;
    const void *w, *w_lo; const float* bias;
    void *out, *out_lo; int out_f32, out_stride, out_off;
    int N, H, W;                 // stored input image (without border)
    int cin, cout, stride, act, x_pad, out_pad;
    int Ho, Wo;                  // tile-space output extent
    int ntaps; signed char tdy[9], tdx[9], ttap[9];
    int omul, oay, oax, OutH, OutW;
    int accum;
};

static int conv_v1(const V1Launch& d, const char* who, cudaStream_t st) {
    if (d.cin % 16 != 0 || d.x_stride % 8 != 0 || d.x_off % 8 != 0 || d.cout > 256) {
        adc::set_error("%s: needs cin %% 16 == 0, 16-byte aligned channel slices and cout <= 256 (cin=%d cout=%d)", who, d.cin, d.cout);
        return ADC_E_UNSUPP;
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        adc::set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    const int split = d.x_lo != nullptr;
    const int npad = (d.cout + 15) / 16 * 16;
    const int cin = d.cin, stride = d.stride;
    const int kblk = cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : (cin >= 64 ? 64 : 16));
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.N = d.N; p.H = d.H; p.W = d.W;
    p.Ho = d.Ho; p.Wo = d.Wo;
    p.cin = cin; p.npad = npad; p.cout = d.cout; p.stride = stride;
    p.BW = p.Wo >= 24 ? 32 : (p.Wo >= 12 ? 16 : 8);
    p.BH = 128 / p.BW;
    p.tiles_x = (p.Wo + p.BW - 1) / p.BW; p.tiles_y = (p.Ho + p.BH - 1) / p.BH;
    p.ntiles = p.tiles_x * p.tiles_y * d.N;
    p.act = d.act; p.out_f32 = d.out_f32; p.out_pix_stride = d.out_stride; p.out_ch_off = d.out_off; p.out = d.out; p.bias = d.bias;
    p.out_lo = d.out_f32 ? nullptr : d.out_lo;
    p.split = split;
    p.x_pad = d.x_pad ? 1 : 0;
    p.out_pad = d.out_pad ? 1 : 0;
    p.ntaps = d.ntaps;
    memcpy(p.tdy, d.tdy, 9); memcpy(p.tdx, d.tdx, 9); memcpy(p.ttap, d.ttap, 9);
    p.omul = d.omul; p.oay = d.oay; p.oax = d.oax; p.OutH = d.OutH; p.OutW = d.OutW;
    p.accum = (d.accum && d.out_f32) ? 1 : 0;
    const int Wt = d.W + 2 * p.x_pad, Ht = d.H + 2 * p.x_pad;    // stored extent of the input

    CUtensorMap mx, mw, mxl, mwl;
    memset(&mxl, 0, sizeof(mxl));
    memset(&mwl, 0, sizeof(mwl));
    auto make_x = [&](CUtensorMap* m, const void* base) -> int {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(base) + d.x_off;
        cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)d.N};
        cuuint64_t strides[3] = {(cuuint64_t)d.x_stride * 2, (cuuint64_t)d.x_stride * 2 * Wt, (cuuint64_t)d.x_stride * 2 * Wt * Ht};
        cuuint32_t box[4] = {(cuuint32_t)kblk, (cuuint32_t)(p.BW * stride), (cuuint32_t)(p.BH * stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(xb), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz(kblk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(x) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    // CTA pairs in the split mode (each CTA: both weight planes of half the output channels)
    bool pair = false;
    {
        const char* e = getenv("ADC_BOX_NO_PAIR");
        pair = split && !(e && e[0] == '1');
    }
    const int brows = pair ? npad / 2 : npad;
    auto make_w = [&](CUtensorMap* m, const void* base) -> int {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)npad, 9};
        cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * 2 * npad};
        cuuint32_t box[3] = {(cuuint32_t)kblk, (cuuint32_t)brows, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz(kblk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(w) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    int rc;
    if ((rc = make_x(&mx, d.x)) || (rc = make_w(&mw, d.w))) return rc;
    if (split && ((rc = make_x(&mxl, d.x_lo)) || (rc = make_w(&mwl, d.w_lo)))) return rc;
    const int rowb = kblk * 2;
    const int al_off = (128 * rowb + brows * rowb + 1023) & ~1023;
    const int stage_bytes = split ? 2 * al_off : al_off;
    p.nstages = (kSmemCap - 1024 - 256) / stage_bytes;
    if (p.nstages > kStages) p.nstages = kStages;
    if (p.nstages < 2) {
        adc::set_error("%s: a pipeline stage of %d bytes does not fit twice in shared memory (Cout too large)", who, stage_bytes);
        return ADC_E_UNSUPP;
    }
    const int smem = p.nstages * stage_bytes + 1024 + 256;
    if (kblk == 64) return launch<64>(mx, mw, mxl, mwl, p, smem, pair, st);
    if (kblk == 32) return launch<32>(mx, mw, mxl, mwl, p, smem, pair, st);
    return launch<16>(mx, mw, mxl, mwl, p, smem, pair, st);
}

// x: NHWC bf16 [N,H,W,x_stride] (channels [x_off, x_off+cin) are read, cin % 16 == 0); w: prepared bf16 [9][npad][cin];
// out: NHWC, out_f32 ? fp32 : bf16, channel slice [out_off, out_off+cout) of a tensor with out_stride channels per pixel.
// Split (bf16x3) mode when x_lo and w_lo are given: every operand is hi + lo with hi = bf16(v), lo = bf16(v - hi) and the
// kernel accumulates hi*hi + lo*hi + hi*lo in fp32 (relative error ~1e-5: fp32-grade parity on the tensor cores); a bf16
// output is then written as the pair (out, out_lo).
int adc_conv3x3_fprop_tc(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo,
                         const float* bias, void* out, void* out_lo, int out_f32, int out_stride, int out_off, int N, int H,
                         int W, int cin, int cout, int stride, int act, int x_pad, int out_pad, adc_stream_t stream) {
    if (!x || !w || !out || N <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || (stride != 1 && stride != 2)) {
        adc::set_error("adc_conv3x3_fprop_tc: bad argument");
        return ADC_E_BADARG;
    }
    if ((x_lo == nullptr) != (w_lo == nullptr)) {
        adc::set_error("adc_conv3x3_fprop_tc: x_lo and w_lo must be given together (split mode)");
        return ADC_E_BADARG;
    }
    V1Launch d;
    memset(&d, 0, sizeof(d));
    d.x = x; d.x_lo = x_lo; d.x_stride = x_stride; d.x_off = x_off; d.w = w; d.w_lo = w_lo; d.bias = bias;
    d.out = out; d.out_lo = out_lo; d.out_f32 = out_f32; d.out_stride = out_stride; d.out_off = out_off;
    d.N = N; d.H = H; d.W = W; d.cin = cin; d.cout = cout; d.stride = stride; d.act = act; d.x_pad = x_pad; d.out_pad = out_pad;
    d.Ho = (H + 2 - 3) / stride + 1; d.Wo = (W + 2 - 3) / stride + 1;
    d.ntaps = 9;
    for (int t = 0; t < 9; ++t) { d.tdy[t] = (signed char)(t / 3 - 1); d.tdx[t] = (signed char)(t % 3 - 1); d.ttap[t] = (signed char)t; }
    d.omul = 1; d.oay = 0; d.oax = 0; d.OutH = d.Ho; d.OutW = d.Wo;
    return conv_v1(d, "adc_conv3x3_fprop_tc", (cudaStream_t)stream);
}

// Data gradient of a 3x3 / stride-2 / pad-1 convolution (forward: x [H,W] -> y [Ho,Wo], Ho = (H-1)/2+1).
//   gx[p][q][ci] = sum over (u,v) with p+1-u, q+1-v even of  W[co][ci][u][v] * gy[(p+1-u)/2][(q+1-v)/2][co]
// Four launches of the v1 kernel, one per parity class (p%2, q%2): even rows see only u=1 (i = p/2), odd rows u=0 (i = (p+1)/2)
// and u=2 (i = (p-1)/2); likewise for columns.  w: adc_conv3x3_prep_weights(..., flip_transpose=2) = [tap][Cin_fwd][Cout_fwd]
// (transposed, taps NOT flipped).  gy: [N][Ho][Wo] (+border if gy_pad), channels [gy_off, gy_off+cgy) with cgy % 16 == 0;
// gx: channels [gx_off, gx_off+cgx) of [N][H][W] (+border if gx_pad), bf16 (pair in split mode) or fp32.
static int dgrad_s2_impl(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* w, const void* w_lo, void* gx,
                         void* gx_lo, int gx_f32, int gx_stride, int gx_off, int N, int H, int W, int cgy, int cgx, int gy_pad,
                         int gx_pad, adc_stream_t stream, int accum, int only_center) {
    if (!gy || !w || !gx || N <= 0 || H <= 0 || W <= 0 || cgy <= 0 || cgx <= 0) {
        adc::set_error("adc_conv3x3_dgrad_s2_tc: bad argument");
        return ADC_E_BADARG;
    }
    if ((gy_lo == nullptr) != (w_lo == nullptr)) {
        adc::set_error("adc_conv3x3_dgrad_s2_tc: gy_lo and w_lo must be given together (split mode)");
        return ADC_E_BADARG;
    }
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    for (int pa = 0; pa < (only_center ? 1 : 2); ++pa)       // a 1x1 / stride-2 layer reaches the even pixels only
        for (int pb = 0; pb < (only_center ? 1 : 2); ++pb) {
            V1Launch d;
            memset(&d, 0, sizeof(d));
            d.accum = accum;
            d.x = gy; d.x_lo = gy_lo; d.x_stride = gy_stride; d.x_off = gy_off; d.w = w; d.w_lo = w_lo; d.bias = nullptr;
            d.out = gx; d.out_lo = gx_lo; d.out_f32 = gx_f32; d.out_stride = gx_stride; d.out_off = gx_off;
            d.N = N; d.H = Ho; d.W = Wo; d.cin = cgy; d.cout = cgx; d.stride = 1; d.act = 0; d.x_pad = gy_pad; d.out_pad = gx_pad;
            d.Ho = (H - pa + 1) / 2; d.Wo = (W - pb + 1) / 2;      // pixels of this parity class
            if (d.Ho <= 0 || d.Wo <= 0) continue;
            const int nu = pa ? 2 : 1, nv = pb ? 2 : 1;
            const int us[2] = {pa ? 0 : 1, 2}, dys[2] = {pa ? 1 : 0, 0};
            const int vs[2] = {pb ? 0 : 1, 2}, dxs[2] = {pb ? 1 : 0, 0};
            d.ntaps = 0;
            for (int a = 0; a < nu; ++a)
                for (int b = 0; b < nv; ++b) {
                    d.tdy[d.ntaps] = (signed char)dys[a]; d.tdx[d.ntaps] = (signed char)dxs[b];
                    d.ttap[d.ntaps] = (signed char)(us[a] * 3 + vs[b]);
                    ++d.ntaps;
                }
            d.omul = 2; d.oay = pa; d.oax = pb; d.OutH = H; d.OutW = W;
            int rc = conv_v1(d, "adc_conv3x3_dgrad_s2_tc", (cudaStream_t)stream);
            if (rc) return rc;
        }
    return 0;
}

int adc_conv3x3_dgrad_s2_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* w, const void* w_lo, void* gx,
                            void* gx_lo, int gx_f32, int gx_stride, int gx_off, int N, int H, int W, int cgy, int cgx, int gy_pad,
                            int gx_pad, adc_stream_t stream) {
    return dgrad_s2_impl(gy, gy_lo, gy_stride, gy_off, w, w_lo, gx, gx_lo, gx_f32, gx_stride, gx_off, N, H, W, cgy, cgx, gy_pad, gx_pad, stream,
                         0, 0);
}


// Stride-1 3x3 convolution, v2 kernel (shared-memory tap reuse, persistent, double-buffered TMEM).  x and out are
// zero-bordered NHWC tensors ([N][H+2][W+2][channels]); needs cin % 64 == 0.  Same split-mode convention as above.
static int conv_v2(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo, const float* bias, void* out,
                   void* out_lo, int out_f32, int out_stride, int out_off, int out_zero_to, int N, int H, int W, int cin, int cout, int act,
                   int taps1, int tap_mask, adc_stream_t stream, int accum = 0) {
    if (!x || !w || !out || N <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) {
        adc::set_error("adc_conv3x3_s1_tc_v2: bad argument");
        return ADC_E_BADARG;
    }
    if ((x_lo == nullptr) != (w_lo == nullptr)) {
        adc::set_error("adc_conv3x3_s1_tc_v2: x_lo and w_lo must be given together (split mode)");
        return ADC_E_BADARG;
    }
    if (cin % 16 != 0 || x_stride % 8 != 0 || x_off % 8 != 0 || cout > 128) {
        adc::set_error("adc_conv3x3_s1_tc_v2: needs cin %% 16 == 0, 16-byte aligned channel slices and cout <= 128 (cin=%d cout=%d)", cin, cout);
        return ADC_E_UNSUPP;
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        adc::set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    const int split = x_lo != nullptr;
    const int npad = (cout + 15) / 16 * 16;
    ConvParamsV2 p;
    memset(&p, 0, sizeof(p));
    p.Hp = H + 2; p.Wp = W + 2;
    p.F = (long long)N * p.Hp * p.Wp;
    p.cin = cin; p.npad = npad; p.cout = cout; p.act = act; p.out_f32 = out_f32; p.split = split;
    p.out_pix_stride = out_stride; p.out_ch_off = out_off; p.out = out; p.out_lo = out_f32 ? nullptr : out_lo; p.bias = bias;
    {   // thin layers: the three horizontal taps as one MMA (see the kernel); a tile is then 126 output pixels
        const char* e = getenv("ADC_V2_NO_HSTACK");
        const char* m = getenv("ADC_V2_HSTACK_MAX");       // (experiment: widest stacked layer, default 32 output channels)
        p.hstack = (!taps1 && npad <= (m ? atoi(m) : 32) && !(e && e[0] == '1')) ? 1 : 0;      // (64 columns per tap: the epilogue's lane exchange costs more than the MMAs save)
    }
    { const char* e = getenv("ADC_V2_PF"); p.pf = e ? atoi(e) : 0; }      // measured: no gain (the loads are not DRAM-latency bound), kept as a switch
    const int tile_m = p.hstack ? 126 : 128;
    if (p.F + 256 >= (1LL << 31)) {
        adc::set_error("adc_conv3x3_s1_tc_v2: more than 2^31 flat pixels");
        return ADC_E_UNSUPP;
    }
    p.ntiles = (int)((p.F + tile_m - 1) / tile_m);
    p.out_zero_to = out_zero_to;
    p.taps1 = taps1;
    p.accum = (accum && out_f32) ? 1 : 0;
    p.ktail = cin % 64 ? (cin % 64) / 16 : 4;
    p.tap_mask = (tap_mask & 0x1FF) ? (tap_mask & 0x1FF) : 0x1FF;
    { const char* e = getenv("ADC_V2_DBG"); p.dbg = e ? atoi(e) : 0; }
    // CTA pairs (cta_group::2) in the split mode: each CTA holds the hi OR the lo weight tiles
    bool pair = false;
    {
        const char* e = getenv("ADC_V2_NO_PAIR");
        // measured (tools/time_rowseg.py): the pair wins wherever a tile carries enough MMAs to hide the cross-CTA handshakes; the 1x1 form
        // (one step per tile) and the stacked-tap layers (few, wide MMAs per tile: 64 -> 32 at 228x304 378 vs 443 us) are faster on single CTAs
        // ... and so are layers with little MMA work per tile (live taps x K steps < 32: the K = 32 data gradients, the tap-masked parity
        // blocks of the stride-2 data gradients: 161 vs 192 us), unless their 81..128 output channels only fit the pair's half-size weight tiles
        const int work = __builtin_popcount((unsigned)p.tap_mask) * ((cin + 15) / 16);
        pair = split && !taps1 && !p.hstack && (work >= 32 || npad > 80) && !(e && e[0] == '1');     // (a single tile: the peer's is out of range)
    }
    {
        const char* e = getenv("ADC_V2_NO_TRI");
        const char* m = getenv("ADC_V2_TRI_MIN");           // (experiment: narrowest layer in the three-product form)
        p.tri = (pair && npad >= (m ? atoi(m) : 80) && !(e && e[0] == '1')) ? 1 : 0;
    }
    CUtensorMap mx, mw, mxl, mwl;
    memset(&mxl, 0, sizeof(mxl));
    memset(&mwl, 0, sizeof(mwl));
    auto make_x = [&](CUtensorMap* m, const void* base_) -> int {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(base_) + x_off;
        cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)p.F};
        cuuint64_t strides[1] = {(cuuint64_t)x_stride * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)kSegRows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(xb), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(x, v2) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    auto make_w = [&](CUtensorMap* m, const void* base_) -> int {
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)npad, 9};
        cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * 2 * npad};
        cuuint32_t box[3] = {64, (cuuint32_t)(p.tri ? npad / 2 : npad), 1};      // one tap (tri: one CTA's half of its rows) per copy
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base_), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            adc::set_error("cuTensorMapEncodeTiled(w, v2) failed: %d", (int)r);
            return ADC_E_UNSUPP;
        }
        return 0;
    };
    int rc;
    if ((rc = make_x(&mx, x)) || (rc = make_w(&mw, w))) return rc;
    if (split && ((rc = make_x(&mxl, x_lo)) || (rc = make_w(&mwl, w_lo)))) return rc;
    // bf16 output whose slice is exactly 64 channels wide: coalesced TMA bulk stores from a swizzled staging tile
    // (per-thread 16-byte stores at a 128-byte stride were measured to cost more than the MMAs: 15.7 -> 6.4 ms for GMS)
    CUtensorMap mo, mol;
    memset(&mo, 0, sizeof(mo));
    memset(&mol, 0, sizeof(mol));
    const int width = out_zero_to > npad ? out_zero_to : npad;
    {
        const char* e = getenv("ADC_V2_NO_WIDE_TMA");     // (experiment: only 64-channel slices through the bulk store, as before)
        const bool wide = !(e && e[0] == '1');
        p.tma_store = (!out_f32 && (wide || width == 64) && out_stride % 8 == 0 && out_off % 8 == 0) ? 1 : 0;
    }
    {
        const char* e = getenv("ADC_V2_NO_F32_TMA");
        if (out_f32 && width <= 128 && out_stride % 4 == 0 && out_off % 4 == 0 && !(e && e[0] == '1')) {
            // fp32 result (the BN nets' raw convolution outputs and gradient buffers): one map of 32-channel x 128-row boxes over the
            // channels this launch owns (dims[0] clips the rest), stores or reduce-adds
            float* ob = reinterpret_cast<float*>(out) + out_off;
            const int lim = out_zero_to > cout ? out_zero_to : cout;
            cuuint64_t dims[2] = {(cuuint64_t)lim, (cuuint64_t)p.F};
            cuuint64_t strides[1] = {(cuuint64_t)out_stride * 4};
            cuuint32_t box[2] = {32, (cuuint32_t)tile_m};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ob, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                adc::set_error("cuTensorMapEncodeTiled(out f32, v2) failed: %d", (int)r);
                return ADC_E_UNSUPP;
            }
            p.tma_store = 2;
        }
    }
    if (p.tma_store == 1) {
        auto make_o = [&](CUtensorMap* m, void* base_) -> int {
            __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(base_) + out_off;
            const int lim = out_zero_to > cout ? out_zero_to : cout;      // channels this launch owns: the boxes are clipped there
            cuuint64_t dims[2] = {(cuuint64_t)lim, (cuuint64_t)p.F};
            cuuint64_t strides[1] = {(cuuint64_t)out_stride * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)tile_m};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ob, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                adc::set_error("cuTensorMapEncodeTiled(out, v2) failed: %d", (int)r);
                return ADC_E_UNSUPP;
            }
            return 0;
        };
        if ((rc = make_o(&mo, out))) return rc;
        if (p.out_lo && (rc = make_o(&mol, p.out_lo))) return rc;
    }
    // weights resident in shared memory if they fit next to >= 2 activation stages; otherwise streamed with every step
    const int kblocks = (cin + 63) / 64, mult = split ? 2 : 1;
    const int wmult = pair ? 1 : mult;
    const int w_all = (taps1 ? 1 : (p.hstack ? 9 : __builtin_popcount((unsigned)p.tap_mask))) * kblocks * npad * 128;   // live taps only
    // Shared-memory plan: weights resident if they fit next to >= 2 activation stages, otherwise streamed with every step.  A bf16
    // result gets two staging slots of one 64-channel column block each if the pipeline keeps >= 3 stages, else one slot.
    const int oblk_bytes = 16384 * (p.out_lo ? 2 : 1), oblocks = (width + 63) / 64;
    const int kTail = 256 + 3072;          // mbarriers + TMEM slot, hstack exchange rows
    int smem = 0, ostage_bytes = 0;
    auto plan = [&](int oblk) {
        p.ostage_blocks = oblk;
        ostage_bytes = p.tma_store == 2 ? 32768 : (p.tma_store ? oblk_bytes * oblk : 0);
        const int budget = kSmemCap - 1024 - kTail - ostage_bytes;
        const char* e = getenv("ADC_V2_NO_RESIDENT");
        const bool allow = !(e && e[0] == '1');
        const int a_stage = kSegBytes * mult;
        int nst_res = (budget - w_all * wmult) / a_stage;
        if (allow && nst_res >= 2) {
            p.resident = 1;
            p.nstages = nst_res > 8 ? 8 : nst_res;
            smem = w_all * wmult + p.nstages * a_stage + ostage_bytes + 1024 + kTail;
        } else {
            const int st = kSegBytes * mult + (taps1 ? 1 : 3) * npad * 128 * wmult;
            p.resident = 0;
            p.nstages = budget / st > 8 ? 8 : budget / st;
            smem = p.nstages * st + ostage_bytes + 1024 + kTail;
        }
    };
    {
        const char* e = getenv("ADC_V2_NO_RING");              // (experiment / test: one staging slot, staged and drained in turn)
        const bool ring = p.tma_store == 1 && oblocks <= 2 && !(e && e[0] == '1');
        plan(ring ? 2 : 1);                                    // two staging slots (see the epilogue) if the pipeline keeps >= 3 stages
    }
    if (p.ostage_blocks > 1 && p.nstages < 3) plan(1);
    {
        const char* e = getenv("ADC_V2_MAXST");
        if (e && atoi(e) >= 2 && atoi(e) < p.nstages) p.nstages = atoi(e);     // (experiment: the shared memory stays allocated)
    }
    if (p.nstages < 2 || smem > kSmemCap) {
        adc::set_error("adc_conv3x3_s1_tc_v2: %d bytes of shared memory needed", smem);
        return ADC_E_UNSUPP;
    }
    static bool done = false;
    static int sms = 148;
    if (!done) {
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_s1_tc_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap),
                             "cudaFuncSetAttribute(conv3x3_s1_tc_v2_kernel)");
        if (rc) return rc;
        rc = adc::check_cuda(cudaFuncSetAttribute(conv3x3_s1_tc_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap),
                             "cudaFuncSetAttribute(conv3x3_s1_tc_v2_kernel<pair>)");
        if (rc) return rc;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        done = true;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TC_FPROP_V2, 2.0 * N * H * W * (taps1 ? 1.0 : 9.0) * cin * cout);
    int threads = 320;      // 8 epilogue warps when there are at least two 16-column chunks to share
    {
        const char* e = getenv("ADC_V2_EPI4");
        if ((e && e[0] == '1') || (out_zero_to > npad ? out_zero_to : npad) < 32) threads = 192;
    }
    if (pair) {
        const int pairs = (p.ntiles + 1) / 2, max_pairs = sms / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return adc::check_cuda(cudaLaunchKernelEx(&cfg, conv3x3_s1_tc_v2_kernel<true>, mx, mw, mxl, mwl, mo, mol, p),
                               "conv3x3_s1_tc_v2_kernel<pair> launch");
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    conv3x3_s1_tc_v2_kernel<false><<<grid, threads, smem, (cudaStream_t)stream>>>(mx, mw, mxl, mwl, mo, mol, p);
    return adc::check_cuda(cudaGetLastError(), "conv3x3_s1_tc_v2_kernel launch");
}

int adc_conv3x3_s1_tc_v2(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo,
                         const float* bias, void* out, void* out_lo, int out_f32, int out_stride, int out_off, int out_zero_to,
                         int N, int H, int W, int cin, int cout, int act, int tap_mask, adc_stream_t stream) {
    return conv_v2(x, x_lo, x_stride, x_off, w, w_lo, bias, out, out_lo, out_f32, out_stride, out_off, out_zero_to, N, H, W, cin, cout, act, 0,
                   tap_mask, stream);
}

// 1x1 convolution (a plain GEMM over the flat pixels of a zero-bordered NHWC tensor) on the same persistent kernel: only the
// centre tap of the prepared [9][Cout16][Cin] weights is read (adc_conv3x3_prep_weights of a tensor whose other taps are zero).
int adc_conv1x1_tc_v2(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo, const float* bias, void* out,
                      void* out_lo, int out_f32, int out_stride, int out_off, int out_zero_to, int N, int H, int W, int cin, int cout, int act,
                      adc_stream_t stream) {
    return conv_v2(x, x_lo, x_stride, x_off, w, w_lo, bias, out, out_lo, out_f32, out_stride, out_off, out_zero_to, N, H, W, cin, cout, act, 1,
                   0, stream);
}

// Descriptor form of the convolution entry points above (one struct instead of 20 positional arguments) with the two options the BN
// nets need: `accumulate` (fp32 output += result: a gradient buffer that several layers feed) and input-channel counts that are a
// multiple of 16 but not of 64 (the row-segment kernel then issues only the K steps of the last 64-channel block that hold channels).
// kind 0: 3x3 stride 1 on the row-segment kernel (zero-bordered tensors); 1: 1x1 on the same kernel; 2: 3x3 of any stride on the box
// kernel; 3: 1x1 of stride 2 on the box kernel (ResNet down-sampling shortcut: the centre tap of prepared [9][Cout16][Cin] weights);
// 4: data gradient of a 3x3 / stride-2 / pad-1 convolution = ConvTranspose2d(3, stride 2, padding 1, output_padding 1) forward
// (H, W = the LARGER tensor; w prepared with mode 2); 5: data gradient of a 1x1 / stride-2 convolution (even pixels of the larger tensor
// only: accumulate into a buffer another layer has written, or into zeros).
int adc_conv_tc(const adc_conv_t* d, adc_stream_t stream) {
    if (!d) {
        adc::set_error("adc_conv_tc: null descriptor");
        return ADC_E_BADARG;
    }
    if (d->kind == 0 || d->kind == 1)
        return conv_v2(d->x, d->x_lo, d->x_stride, d->x_off, d->w, d->w_lo, d->bias, d->out, d->out_lo, d->out_f32, d->out_stride, d->out_off,
                       d->out_zero_to, d->N, d->H, d->W, d->cin, d->cout, d->act, d->kind == 1, d->tap_mask, stream, d->accumulate);
    if (!d->x || !d->w || !d->out || d->N <= 0 || d->H <= 0 || d->W <= 0 || d->cin <= 0 || d->cout <= 0 || (d->x_lo == nullptr) != (d->w_lo == nullptr)) {
        adc::set_error("adc_conv_tc: bad argument");
        return ADC_E_BADARG;
    }
    if (d->kind == 2 || d->kind == 3) {
        if (d->stride != 1 && d->stride != 2) {
            adc::set_error("adc_conv_tc: stride must be 1 or 2");
            return ADC_E_BADARG;
        }
        V1Launch v;
        memset(&v, 0, sizeof(v));
        v.x = d->x; v.x_lo = d->x_lo; v.x_stride = d->x_stride; v.x_off = d->x_off; v.w = d->w; v.w_lo = d->w_lo; v.bias = d->bias;
        v.out = d->out; v.out_lo = d->out_lo; v.out_f32 = d->out_f32; v.out_stride = d->out_stride; v.out_off = d->out_off;
        v.N = d->N; v.H = d->H; v.W = d->W; v.cin = d->cin; v.cout = d->cout; v.stride = d->stride; v.act = d->act;
        v.x_pad = d->x_pad; v.out_pad = d->out_pad; v.accum = d->accumulate;
        v.Ho = (d->H - 1) / d->stride + 1; v.Wo = (d->W - 1) / d->stride + 1;
        if (d->kind == 2) {
            v.ntaps = 9;
            for (int t = 0; t < 9; ++t) { v.tdy[t] = (signed char)(t / 3 - 1); v.tdx[t] = (signed char)(t % 3 - 1); v.ttap[t] = (signed char)t; }
        } else {
            v.ntaps = 1;
            v.tdy[0] = 0; v.tdx[0] = 0; v.ttap[0] = 4;
        }
        v.omul = 1; v.oay = 0; v.oax = 0; v.OutH = v.Ho; v.OutW = v.Wo;
        return conv_v1(v, "adc_conv_tc", (cudaStream_t)stream);
    }
    if (d->kind == 4 || d->kind == 5)
        return dgrad_s2_impl(d->x, d->x_lo, d->x_stride, d->x_off, d->w, d->w_lo, d->out, d->out_lo, d->out_f32, d->out_stride, d->out_off, d->N,
                             d->H, d->W, d->cin, d->cout, d->x_pad, d->out_pad, stream, d->accumulate, d->kind == 5);
    adc::set_error("adc_conv_tc: unknown kind %d", d->kind);
    return ADC_E_BADARG;
}

}  // extern "C"

// Shared device/host helpers of the tcgen05 (5th-generation tensor core) kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// UMMA shared-memory and instruction descriptors, epilogue arithmetic, and the host-side tensor-map encoder lookup.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/adc_b200.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// request a tile into L2 only (no shared-memory destination, no barrier): hides the DRAM latency of a later load of the same box
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// global[map tile] += shared tile (fp32): the accumulate form of the bulk store
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- tcgen05 wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, single CTA
// elect.sync: exactly one lane of the (converged) warp gets `true`.  Unlike `lane == 0` the compiler KNOWS the region behind it runs on
// a single thread, so the uniform-register operands of UTCHMMA / UTMALDG need no per-instruction ELECT / BRA.U.ANY "waterfall" loop:
// ncu's source view showed the issuing warp spending its time in that loop, 135 cycles per MMA (profiles/r2_ncu_rowseg_forward_b128.txt).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) execute ONE MMA of M = 256 ----------------------------
// Each CTA supplies its own 128 rows of A and N/2 rows of B from its shared memory and receives its 128 rows of D in its own TMEM;
// the leader CTA (cluster rank 0) issues.  Measured (tools/micro/mma_pair_rate.cu): max(~40, N/2) cycles per instruction -- N = 128
// costs 64 cycles for 256 rows where the single-CTA form costs 102 for 128 rows, because each SM fetches half of B.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t phase) {     // observes arrivals from the peer CTA
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAITC:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONEC;\n"
        "bra TC_WAITC;\n"
        "TC_DONEC:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the bytes are counted on the mbarrier at `bar_cluster_addr`
// (the leader's barrier: the MMA thread waits for both CTAs' operands on one barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // warp 1 of BOTH CTAs, same dst offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// 32 lanes x 16 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Epilogue arithmetic on 16 accumulator columns: + bias (staged in shared memory, zero where there is none), activation.
// The activation is selected ONCE per chunk (warp-uniform switch); the first version tested p.act, p.bias and p.cout per
// element and compiled to 415 instructions per chunk, which made the epilogue (not the MMAs or the stores) the critical
// path of the persistent kernel: 706 us per 64->64 layer against 231 us with the epilogue removed (profiles/r1_tc_epilogue.txt).
template <int ACT>
__device__ __forceinline__ void epi_bias_act16(float (&v)[16], const float* sb) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sb + j);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r = v[j + i] + bb[i];
            if (ACT == 1) r = fmaxf(r, 0.f);
            else if (ACT == 2) r = r > 0.f ? r : 0.2f * r;
            else if (ACT == 3) r = 1.f / (1.f + __expf(-r));
            v[j + i] = r;
        }
    }
}
__device__ __forceinline__ void epi_bias_act16(float (&v)[16], const float* sb, int act) {
    switch (act) {
        case 1: epi_bias_act16<1>(v, sb); break;
        case 2: epi_bias_act16<2>(v, sb); break;
        case 3: epi_bias_act16<3>(v, sb); break;
        default: epi_bias_act16<0>(v, sb); break;
    }
}
// channels >= cout of a chunk (weights and bias are zero there, so only sigmoid(0) = 0.5 needs the mask) and border pixels -> 0
__device__ __forceinline__ void epi_mask16(float (&v)[16], int c0, int cout, bool zero_all) {
    if (c0 + 16 > cout) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (c0 + j >= cout) ? 0.f : v[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = zero_all ? 0.f : v[j];
}
__device__ __forceinline__ void epi_pack_hi(const float (&v)[16], uint4 (&h)[2]) {
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(h);
#pragma unroll
    for (int j = 0; j < 8; ++j) hh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void epi_pack_lo(const float (&v)[16], const uint4 (&h)[2], uint4 (&l)[2]) {
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(h);
    __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(l);
#pragma unroll
    for (int j = 0; j < 8; ++j) ll[j] = __floats2bfloat162_rn(v[2 * j] - __low2float(hh[j]), v[2 * j + 1] - __high2float(hh[j]));
}

// Shared-memory matrix descriptor, K-major operand whose rows are ROWB bytes (= the swizzle span: 128/64/32) and whose
// 8-row swizzle atoms are packed back to back (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version 1 <<46 | layout <<61; layout 2/4/6 = SWIZZLE_128B/64B/32B).
template <int ROWB>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * ROWB) >> 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major A and B,
// N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kSmemCap = 227 * 1024 - 1024;   // dynamic shared memory limit: the opt-in maximum minus the static bias table

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static inline CUtensorMapSwizzle swz(int kblk) {
    return kblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace tc

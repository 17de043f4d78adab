// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair, K = 16, K-major SWIZZLE_128B operands in
// shared memory; each CTA supplies its 128 rows of A and N/2 rows of B) as a function of N, issued back to back by one thread of the
// leader CTA.  Compare with tools/micro/mma_rate.cu (cta_group::1: max(91, 38 + N/2) cycles per M = 128 instruction).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../aerial_depth_completion_b200/csrc -o mma_pair_rate mma_pair_rate.cu -lcuda
#include "tc_common.cuh"
namespace adc { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return (int)e; } }
using namespace tc;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192) mma_pair_rate_kernel(int N, int reps, long long* out, float* check) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t rank = cluster_rank();
    // A = 1.0 everywhere; B rows of CTA r = (r + 1): D[:, 0..N/2) = 16 * reps * 4 * 1, D[:, N/2..N) = ... * 2 if the halves come from the two CTAs
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (i * 4 >= 32768 && rank == 1) ? 0x40004000u : 0x3f803f80u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && rank == 0) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(256, N);
            const uint32_t a = base, b = base + 32768;
            const uint64_t da = make_desc<128>(a), db = make_desc<128>(b);
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, da + 2 * k, db + 2 * k, idesc, (r | k) != 0);
            }
            umma_commit_pair(&bar, 3);
            mbar_wait(&bar, 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
    } else if (warp == 1) {
        if (elect_one()) mbar_wait(&bar, 0);       // the multicast commit arrives here too
    }
    __syncwarp();
    tc_fence_after();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2 && blockIdx.x < 2) {            // lane quarter 2 of each CTA of the first pair: columns 0 and N-1 of row 64
        float v[16];
        tmem_ld16(tmem + ((uint32_t)64 << 16), v);
        float w[16];
        tmem_ld16(tmem + ((uint32_t)64 << 16) + (uint32_t)(N - 16), w);
        if ((threadIdx.x & 31) == 0) { check[rank * 2] = v[0]; check[rank * 2 + 1] = w[15]; }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d;
    float* c;
    cudaMalloc(&d, 8);
    cudaMalloc(&c, 16);
    cudaFuncSetAttribute(mma_pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 2000;
    for (int N : {32, 64, 96, 128, 160, 192, 256}) {
        cudaMemset(c, 0, 16);
        mma_pair_rate_kernel<<<148, 192, 100 * 1024>>>(N, reps, d, c);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        float hc[4] = {0, 0, 0, 0};
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
        printf("cta_group::2 M 256 N %3d: %6.1f clk/MMA   D[64][0], D[64][N-1] of CTA0: %.0f %.0f  CTA1: %.0f %.0f (expect %d and %d)  %s\n", N,
               (double)h / (reps * 4), hc[0], hc[1], hc[2], hc[3], 16 * 4 * reps, 2 * 16 * 4 * reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
    }
    return 0;
}

// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, operands in shared memory) as a function of N,
// for K-major SWIZZLE_128B and MN-major operands, issued back to back by one thread into the same / alternating accumulators.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../aerial_depth_completion_b200/csrc -o mma_rate mma_rate.cu -lcuda
#include "tc_common.cuh"
namespace adc { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return (int)e; } }
using namespace tc;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// mode 0: K-major A and B, same accumulator; 1: alternate 2 accumulators; 2: MN-major A and B; 3: K-major, A fixed / B fixed but
// 4 different k offsets (like the conv loop)
__global__ void __launch_bounds__(192) mma_rate_kernel(int N, int reps, int mode, long long* out, int elect_mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (mode == 3 || mode == 4) {   // 3: two issuing threads in different warps (accumulators 0 / 256); 4: M = 64, one thread
        const bool issuer = lane == 0 && (warp == 1 || (mode == 3 && warp == 2));
        if (issuer) {
            const uint32_t idesc = make_idesc(mode == 4 ? 64 : 128, N);
            const uint32_t a = base + (warp == 2 ? 8192 : 0), b = base + 32768 + (warp == 2 ? 8192 : 0);
            const uint32_t acc = tmem + (warp == 2 ? 256 : 0);
            const int my = mode == 3 ? reps / 2 : reps;
            long long t0 = clock64();
            for (int r = 0; r < my; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(acc, make_desc<128>(a) + 2 * k, make_desc<128>(b) + 2 * k, idesc, 1);
            }
            umma_commit(&bar);
            if (warp == 1) {
                mbar_wait(&bar, 0);
                if (mode == 3) mbar_wait(&bar, 1);
                long long t1 = clock64();
                if (blockIdx.x == 0) out[0] = t1 - t0;
            }
        }
    } else if (warp == 1 && (elect_mode ? elect_one() : lane == 0)) {
        uint32_t idesc = make_idesc(128, N);
        if (mode == 2) idesc |= (1u << 15) | (1u << 16);
        const uint32_t a = base, b = base + 32768;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = tmem + ((mode == 1) ? ((r & 1) * 256) : 0);
                if (mode == 2) umma_bf16(acc, make_desc_mn(a + k * 2048, 16384), make_desc_mn(b + k * 2048, 16384), idesc, 1);
                else umma_bf16(acc, make_desc<128>(a) + 2 * k, make_desc<128>(b) + 2 * k, idesc, 1);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 2000;
    // issuer selected by lane == 0 (the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop) vs elect.sync (modes 0..2)
    for (int elect_mode = 0; elect_mode < 2; ++elect_mode)
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {16, 32, 64, 96, 128, 192, 256}) {
            for (int grid : {148}) {
                mma_rate_kernel<<<grid, 192, 100 * 1024>>>(N, reps, mode, d, elect_mode);
                cudaError_t e = cudaDeviceSynchronize();
                long long h = 0;
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("issuer %s mode %d N %3d grid %3d: %6.1f clk/MMA  (floor %d)  %s\n", elect_mode ? "elect.sync" : "lane==0   ", mode, N, grid, (double)h / (reps * 4), 128 * N / 256,
                       e == cudaSuccess ? "" : cudaGetErrorString(e)); fflush(stdout);
            }
        }
    return 0;
}

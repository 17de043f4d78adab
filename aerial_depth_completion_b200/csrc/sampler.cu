// UniformSampling.dense_to_sparse (dataloaders/dense_to_sparse.py:29-48) on the device, BIT-EXACT with the reference's numpy stream:
//   keep = depth > 0 [& depth <= max_depth];  prob = float(num_samples) / count(keep)  (float64);
//   mask = keep & (np.random.uniform(0, 1, shape) < prob)   after np.random.seed(seed)
// numpy's legacy generator is MT19937 (Matsumoto & Nishimura): init_genrand(seed), 624-word state regenerated in place, tempered
// 32-bit outputs, one double per pixel from two outputs: ((a >> 5) * 2^26 + (b >> 6)) / 2^53.  One CTA per frame keeps the state in
// shared memory.  The in-place regeneration is sequential only across three index ranges: word i needs the OLD words i, i+1 and word
// (i+397) mod 624, which is old for i < 227 and already NEW for i >= 227 (and word 623 needs the new word 0), so the ranges
// [0,227), [227,454), [454,624) are each computed in parallel (read everything, barrier, write) -- validated against the sequential
// definition in oracle/sampling_oracle.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace sampler {

constexpr int kN = 624, kM = 397, kThreads = 256;

__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    uint32_t v = y >> 1;
    if (y & 1u) v ^= 0x9908b0dfu;
    return c ^ v;
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// one range [lo, hi) of the regeneration: every thread reads its three inputs, barrier, writes, barrier
__device__ __forceinline__ void regen_range(uint32_t* mt, int lo, int hi, bool c_is_new_low) {
    const int i = lo + (int)threadIdx.x;
    uint32_t a = 0, b = 0, c = 0;
    const bool on = i < hi;
    if (on) {
        a = mt[i];
        b = mt[(i + 1) % kN];
        c = c_is_new_low ? mt[i - (kN - kM)] : mt[i + kM];
    }
    __syncthreads();
    if (on) mt[i] = twist(a, b, c);
    __syncthreads();
}

// Where a pixel's decision goes: the uint8 mask of adc_uniform_sampling, or -- adc_sample_assemble -- straight into the network input:
// d = gt * mask and w = mask as fp32 planes (dataloader_ext.py:544-566 builds exactly these two channels from the mask on the host).
struct Sink {
    unsigned char* mask;      // [HW] or nullptr
    float* d;                 // [HW] or nullptr
    float* w;
    __device__ __forceinline__ void put(long long p, bool on, float v) const {
        if (mask) mask[p] = on ? 1 : 0;
        if (d) { d[p] = on ? v : 0.f; w[p] = on ? 1.f : 0.f; }
    }
};

__global__ void __launch_bounds__(kThreads) uniform_sampling_kernel(const float* __restrict__ depth, long long depth_frame_stride,
                                                                    const unsigned int* __restrict__ seeds,
                                                                    int num_samples, float max_depth, int use_max,
                                                                    unsigned char* __restrict__ mask, float* __restrict__ dw_out,
                                                                    long long out_frame_stride, long long HW) {
    __shared__ uint32_t mt[kN];
    __shared__ uint32_t out[kN];
    __shared__ int warp_cnt[kThreads / 32];
    __shared__ long long s_keep;
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* d = depth + (long long)n * depth_frame_stride;
    Sink m;
    m.mask = mask ? mask + (long long)n * HW : nullptr;
    m.d = dw_out ? dw_out + (long long)n * out_frame_stride : nullptr;
    m.w = dw_out ? m.d + HW : nullptr;

    // n_keep = count(depth > 0 [& depth <= max_depth])
    long long cnt = 0;
    for (long long p = tid; p < HW; p += kThreads) {
        const float v = d[p];
        cnt += (v > 0.f && (!use_max || v <= max_depth)) ? 1 : 0;
    }
    int c32 = (int)cnt;                                   // <= HW / 256 per thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c32 += __shfl_xor_sync(0xffffffffu, c32, o);
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = c32;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += warp_cnt[w];
        s_keep = t;
        // init_genrand(seed): numpy's legacy seeding of an integer seed
        uint32_t s = seeds[n];
        for (int pos = 0; pos < kN; ++pos) {
            mt[pos] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)pos + 1u;
        }
    }
    __syncthreads();
    const long long n_keep = s_keep;
    if (n_keep == 0 || num_samples == 0) {                // the reference returns mask_keep without touching the generator
        for (long long p = tid; p < HW; p += kThreads) {
            const float v = d[p];
            m.put(p, v > 0.f && (!use_max || v <= max_depth), v);
        }
        return;
    }
    const double prob = (double)num_samples / (double)n_keep;

    const long long blocks = (2 * HW + kN - 1) / kN;      // regenerations needed: 2 outputs per pixel
    for (long long r = 0; r < blocks; ++r) {
        regen_range(mt, 0, kN - kM, false);               // [0, 227): word i+397 is old
        regen_range(mt, kN - kM, 2 * (kN - kM), true);    // [227, 454): word i-227 is new
        regen_range(mt, 2 * (kN - kM), kN, true);         // [454, 624): word i-227 is new, word 623 reads the new word 0
        for (int i = tid; i < kN; i += kThreads) out[i] = temper(mt[i]);
        __syncthreads();
        for (int t = tid; t < kN / 2; t += kThreads) {
            const long long p = r * (kN / 2) + t;
            if (p < HW) {
                const uint32_t a = out[2 * t] >> 5, b = out[2 * t + 1] >> 6;
                const double u = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
                const float v = d[p];
                m.put(p, v > 0.f && (!use_max || v <= max_depth) && u < prob, v);
            }
        }
        __syncthreads();                                  // `out` is rewritten by the next round
    }
}

// rgb planes of the [N][4][HW] (r, g, b, gt) upload into channels 0..2 of the [N][5][HW] network input: 16-byte copies when HW % 4 == 0
__global__ void __launch_bounds__(256) copy_rgb_kernel(const float* __restrict__ src, float* __restrict__ dst, long long HW3, long long src_fs,
                                                       long long dst_fs, int vec) {
    const int n = blockIdx.y;
    const float* s = src + (long long)n * src_fs;
    float* o = dst + (long long)n * dst_fs;
    const long long step = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(s);
        float4* o4 = reinterpret_cast<float4*>(o);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW3 / 4; i += step) o4[i] = __ldg(s4 + i);
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW3; i += step) o[i] = __ldg(s + i);
    }
}

}  // namespace sampler

extern "C" {

int adc_uniform_sampling(const float* depth, const unsigned int* seeds, int num_samples, float max_depth, int use_max_depth,
                         unsigned char* mask, int N, long long HW, adc_stream_t stream) {
    if (!depth || !seeds || !mask || N <= 0 || HW <= 0 || num_samples < 0) {
        adc::set_error("adc_uniform_sampling: bad argument (num_samples >= 0; the random-probability modes -1 / -2 are not provided)");
        return ADC_E_BADARG;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TRAIN, (double)N * HW * 5.0);
    sampler::uniform_sampling_kernel<<<N, sampler::kThreads, 0, (cudaStream_t)stream>>>(depth, HW, seeds, num_samples, max_depth, use_max_depth,
                                                                                      mask, nullptr, 0, HW);
    return adc::check_cuda(cudaGetLastError(), "uniform_sampling_kernel launch");
}

int adc_sample_assemble(const float* rgbgt, const unsigned int* seeds, int num_samples, float max_depth, int use_max_depth, float* input5,
                        int N, long long HW, adc_stream_t stream) {
    if (!rgbgt || !seeds || !input5 || N <= 0 || HW <= 0 || num_samples < 0) {
        adc::set_error("adc_sample_assemble: bad argument (num_samples >= 0)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    adc::Prof prof(st, adc::KIND_TRAIN, (double)N * HW * (4.0 + 8.0 + 24.0));
    adc::count_launches(1);
    const int vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(rgbgt) | reinterpret_cast<uintptr_t>(input5)) % 16 == 0);
    long long blocks = (3 * HW / (vec ? 4 : 1) + 255) / 256;
    dim3 grid((unsigned)(blocks < 148 ? (blocks > 0 ? blocks : 1) : 148), (unsigned)N);
    sampler::copy_rgb_kernel<<<grid, 256, 0, st>>>(rgbgt, input5, 3 * HW, 4 * HW, 5 * HW, vec);
    int rc = adc::check_cuda(cudaGetLastError(), "copy_rgb_kernel launch");
    if (rc) return rc;
    sampler::uniform_sampling_kernel<<<N, sampler::kThreads, 0, st>>>(rgbgt + 3 * HW, 4 * HW, seeds, num_samples, max_depth, use_max_depth,
                                                                     nullptr, input5 + 3 * HW, 5 * HW, HW);
    return adc::check_cuda(cudaGetLastError(), "uniform_sampling_kernel launch");
}

}  // extern "C"

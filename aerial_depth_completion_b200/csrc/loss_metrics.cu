// Masked losses (confidence_depth_framework.py:829-914) and Result.evaluate (metrics.py:38-66) as single-pass kernels.
// The reference compacts (target - pred)[mask] and reduces with a host sync per scalar (2 per loss, ~12 per evaluate); here one
// kernel accumulates every sum over target > 0 (per-thread double accumulators, per-CTA partial rows, a one-CTA fixed-order
// finalize: deterministic) and nothing leaves the device until the caller reads the result tensor.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/adc_b200.h"
#include "prof.h"

namespace adc {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace adc

namespace lossm {

constexpr int kParts = 148 * 4;
constexpr int kNV = 10;              // values per partial row (loss: 2 used, metrics: 10)

// per-element error of the masked losses: 0 l2 (t-p)^2 | 1 l1 |t-p| | 2 il1 |1/t - 1/p| | 3 absrel |(t-p)/t|
__device__ __forceinline__ float loss_err(int kind, float p, float t) {
    switch (kind) {
        case 0: { const float d = t - p; return d * d; }
        case 1: return fabsf(t - p);
        case 2: return fabsf(1.0f / t - 1.0f / p);
        default: return fabsf((t - p) / t);
    }
}
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }   // torch.abs backward: sign(x)
__device__ __forceinline__ float loss_derr(int kind, float p, float t) {
    switch (kind) {
        case 0: return -2.0f * (t - p);
        case 1: return -sgn(t - p);
        case 2: return sgn(1.0f / t - 1.0f / p) * (1.0f / (p * p));
        default: return sgn((t - p) / t) * (-1.0f / t);
    }
}

template <int NV>
__device__ __forceinline__ void block_store(double (&v)[NV], double* __restrict__ row) {
    __shared__ double red[8][kNV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        row[threadIdx.x] = t;
    }
}

// 16-byte loads, two groups of four elements in flight per thread (the scalar grid-stride loop with its dependent, predicated pred
// load was latency-bound: 60 us for 71 MB at B = 128); scalar path for unaligned views and the tail.
template <int kind>
__global__ void __launch_bounds__(256) masked_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                  long long n, double* __restrict__ part) {
    double s0 = 0.0, s1 = 0.0;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
    const long long n4 = vec ? n >> 2 : 0;
    const float4* p4 = reinterpret_cast<const float4*>(pred);
    const float4* t4 = reinterpret_cast<const float4*>(target);
    auto acc1 = [&](float p, float t) {       // per-element fp64 accumulation, as the scalar path
        if (t > 0.f) {
            s0 += (double)loss_err(kind, p, t);
            s1 += 1.0;
        }
    };
    auto acc4 = [&](const float4& p, const float4& t) {
        acc1(p.x, t.x);
        acc1(p.y, t.y);
        acc1(p.z, t.z);
        acc1(p.w, t.w);
    };
    long long i = gtid;
    for (; i + gstride < n4; i += 2 * gstride) {
        const float4 ta = __ldg(t4 + i), tb = __ldg(t4 + i + gstride), pa = __ldg(p4 + i), pb = __ldg(p4 + i + gstride);
        acc4(pa, ta);
        acc4(pb, tb);
    }
    if (i < n4) acc4(__ldg(p4 + i), __ldg(t4 + i));
    for (long long j = n4 * 4 + gtid; j < n; j += gstride) {
        const float t = __ldg(target + j);
        if (t > 0.f) {
            s0 += (double)loss_err(kind, __ldg(pred + j), t);
            s1 += 1.0;
        }
    }
    double v[2] = {s0, s1};
    block_store<2>(v, part + (long long)blockIdx.x * kNV);
}

// warp k sums column k of the partial rows (lane-strided rows, then a fixed shuffle tree: deterministic);
// out[k] = sum_k / count for k < nv-1, out[nv-1] = count
__global__ void __launch_bounds__(32 * kNV) finalize_kernel(const double* __restrict__ part, int rows, int nv, float* __restrict__ out) {
    __shared__ double s[kNV];
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k < nv) {
        double t = 0.0;
        for (int r = lane; r < rows; r += 32) t += part[(long long)r * kNV + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s[k] = t;
    }
    __syncthreads();
    if (threadIdx.x < nv) out[threadIdx.x] = (threadIdx.x == nv - 1) ? (float)s[nv - 1] : (float)(s[threadIdx.x] / s[nv - 1]);
}

template <int kind>
__global__ void __launch_bounds__(256) masked_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                              const float* __restrict__ gloss, const float* __restrict__ count,
                                                              long long n, float* __restrict__ gpred) {
    const float scale = gloss[0] / count[0];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(gpred)) & 15) == 0;
    const long long n4 = vec ? n >> 2 : 0;
    const float4* p4 = reinterpret_cast<const float4*>(pred);
    const float4* t4 = reinterpret_cast<const float4*>(target);
    float4* g4 = reinterpret_cast<float4*>(gpred);
    auto g1 = [&](float p, float t) { return t > 0.f ? scale * loss_derr(kind, p, t) : 0.f; };
    auto gq = [&](const float4& p, const float4& t) { return make_float4(g1(p.x, t.x), g1(p.y, t.y), g1(p.z, t.z), g1(p.w, t.w)); };
    long long i = gtid;
    for (; i + gstride < n4; i += 2 * gstride) {
        const float4 ta = __ldg(t4 + i), tb = __ldg(t4 + i + gstride), pa = __ldg(p4 + i), pb = __ldg(p4 + i + gstride);
        g4[i] = gq(pa, ta);
        g4[i + gstride] = gq(pb, tb);
    }
    if (i < n4) g4[i] = gq(__ldg(p4 + i), __ldg(t4 + i));
    for (long long j = n4 * 4 + gtid; j < n; j += gstride) gpred[j] = g1(__ldg(pred + j), __ldg(target + j));
}

// sums over (target > 0 [and confidence > thr]): |d|^2, |d|, |log10 o - log10 t|, |d|/t, [r<1.25], [r<1.25^2], [r<1.25^3],
// |1/o - 1/t|^2, |1/o - 1/t|, 1   with d = o - t, r = max(o/t, t/o); per-element arithmetic in fp32 like the reference
__global__ void __launch_bounds__(256) depth_metrics_partial_kernel(const float* __restrict__ output, const float* __restrict__ target,
                                                                    const float* __restrict__ conf, float thr, long long n,
                                                                    double* __restrict__ part) {
    double v[kNV];
#pragma unroll
    for (int k = 0; k < kNV; ++k) v[k] = 0.0;
    const float inv_ln10 = 1.0f / 2.302585092994046f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float t = __ldg(target + i);
        if (!(t > 0.f) || (conf && !(__ldg(conf + i) > thr))) continue;
        const float o = __ldg(output + i);
        const float ad = fabsf(o - t);
        const float r = fmaxf(o / t, t / o);
        const float ai = fabsf(1.0f / o - 1.0f / t);
        const float lg = fabsf(logf(o) * inv_ln10 - logf(t) * inv_ln10);
        const bool rnan = (o / t) != (o / t) || (t / o) != (t / o);      // torch.max propagates NaN; NaN < x is false either way
        v[0] += (double)(ad * ad);
        v[1] += (double)ad;
        v[2] += (double)lg;
        v[3] += (double)(ad / t);
        v[4] += (!rnan && r < 1.25f) ? 1.0 : 0.0;
        v[5] += (!rnan && r < 1.25f * 1.25f) ? 1.0 : 0.0;
        v[6] += (!rnan && r < 1.25f * 1.25f * 1.25f) ? 1.0 : 0.0;
        v[7] += (double)(ai * ai);
        v[8] += (double)ai;
        v[9] += 1.0;
    }
    block_store<kNV>(v, part + (long long)blockIdx.x * kNV);
}

// ConfidencePixelwiseThrAverageMeter.evaluate (metrics.py:103-131) asks, for each of num_bins thresholds, for the mean of
// absrel = |d - t| / t * 1000 and the number of valid pixels (t > 0) whose confidence exceeds the threshold: 500 boolean-index
// passes with two host synchronisations each in the reference.  One pass here: a pixel with k = #{j : conf > thr[j]} (thresholds
// ascending, compared in fp32 like `confidence > curr_conf`) counts for exactly the thresholds j < k, so it lands in bin k of a
// (num_bins + 1)-bin histogram and the per-threshold figures are suffix sums over the bins.  Integer accumulators (pixel counts and
// absrel in 2^-16 fixed point) keep the result independent of the atomics' order.  hist: [3][num_bins + 1] unsigned 64-bit =
// count, fixed-point absrel sum, count of pixels whose absrel is not finite (their thresholds have no finite mean, as in the reference).
__global__ void __launch_bounds__(256) conf_pr_hist_kernel(const float* __restrict__ depth, const float* __restrict__ conf,
                                                           const float* __restrict__ target, long long n, const float* __restrict__ thr,
                                                           int nb, unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned long long sh[];      // [3][nb + 1]
    float* sthr = reinterpret_cast<float*>(sh + 3 * (nb + 1));
    for (int i = threadIdx.x; i < 3 * (nb + 1); i += blockDim.x) sh[i] = 0ull;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sthr[i] = thr[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float t = __ldg(target + i);
        if (!(t > 0.f)) continue;
        const float c = __ldg(conf + i);
        const float ar = (fabsf(__ldg(depth + i) - t) / t) * 1000.0f;
        int lo = 0, hi = nb;                         // first j with !(c > thr[j]); NaN confidence: 0
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c > sthr[mid]) lo = mid + 1; else hi = mid;
        }
        atomicAdd(sh + lo, 1ull);
        if (isfinite(ar)) atomicAdd(sh + (nb + 1) + lo, __double2ull_rn((double)ar * 65536.0));
        else atomicAdd(sh + 2 * (nb + 1) + lo, 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (nb + 1); i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, sh[i]);
}

static int grid_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < kParts ? (b > 0 ? b : 1) : kParts);
}

}  // namespace lossm

using namespace lossm;

extern "C" {

size_t adc_loss_metrics_workspace_bytes(void) { return (size_t)kParts * kNV * sizeof(double); }

int adc_masked_loss_fwd(const float* pred, const float* target, long long n, int kind, void* workspace, float* out, adc_stream_t stream) {
    if (!pred || !target || !workspace || !out || n <= 0 || kind < 0 || kind > 3) {
        adc::set_error("adc_masked_loss_fwd: bad argument (kind 0 l2, 1 l1, 2 il1, 3 absrel)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n);
    adc::Prof prof(st, adc::KIND_TRAIN, (double)n * 8.0);
    adc::count_launches(1);
    double* part = reinterpret_cast<double*>(workspace);
    switch (kind) {
        case 0: masked_loss_partial_kernel<0><<<grid, 256, 0, st>>>(pred, target, n, part); break;
        case 1: masked_loss_partial_kernel<1><<<grid, 256, 0, st>>>(pred, target, n, part); break;
        case 2: masked_loss_partial_kernel<2><<<grid, 256, 0, st>>>(pred, target, n, part); break;
        default: masked_loss_partial_kernel<3><<<grid, 256, 0, st>>>(pred, target, n, part); break;
    }
    int rc = adc::check_cuda(cudaGetLastError(), "masked_loss_partial_kernel launch");
    if (rc) return rc;
    finalize_kernel<<<1, 32 * kNV, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, 2, out);
    return adc::check_cuda(cudaGetLastError(), "finalize_kernel launch");
}

int adc_masked_loss_bwd(const float* pred, const float* target, const float* gloss, const float* count, long long n, int kind, float* gpred,
                        adc_stream_t stream) {
    if (!pred || !target || !gloss || !count || !gpred || n <= 0 || kind < 0 || kind > 3) {
        adc::set_error("adc_masked_loss_bwd: bad argument");
        return ADC_E_BADARG;
    }
    adc::Prof prof((cudaStream_t)stream, adc::KIND_TRAIN, (double)n * 12.0);
    long long b = (n + 255) / 256;
    const int grid = (int)(b < 148 * 16 ? b : 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case 0: masked_loss_bwd_kernel<0><<<grid, 256, 0, st>>>(pred, target, gloss, count, n, gpred); break;
        case 1: masked_loss_bwd_kernel<1><<<grid, 256, 0, st>>>(pred, target, gloss, count, n, gpred); break;
        case 2: masked_loss_bwd_kernel<2><<<grid, 256, 0, st>>>(pred, target, gloss, count, n, gpred); break;
        default: masked_loss_bwd_kernel<3><<<grid, 256, 0, st>>>(pred, target, gloss, count, n, gpred); break;
    }
    return adc::check_cuda(cudaGetLastError(), "masked_loss_bwd_kernel launch");
}

int adc_depth_metrics(const float* output, const float* target, const float* confidence, float threshold, long long n, void* workspace,
                      float* out, adc_stream_t stream) {
    if (!output || !target || !workspace || !out || n <= 0) {
        adc::set_error("adc_depth_metrics: bad argument");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n);
    adc::Prof prof(st, adc::KIND_TRAIN, (double)n * (confidence ? 12.0 : 8.0));
    adc::count_launches(1);
    depth_metrics_partial_kernel<<<grid, 256, 0, st>>>(output, target, confidence, threshold, n, reinterpret_cast<double*>(workspace));
    int rc = adc::check_cuda(cudaGetLastError(), "depth_metrics_partial_kernel launch");
    if (rc) return rc;
    finalize_kernel<<<1, 32 * kNV, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, kNV, out);
    return adc::check_cuda(cudaGetLastError(), "finalize_kernel launch");
}

int adc_conf_pr_histogram(const float* depth, const float* confidence, const float* target, long long n, const float* thresholds,
                          int num_bins, unsigned long long* hist, adc_stream_t stream) {
    if (!depth || !confidence || !target || !thresholds || !hist || n <= 0 || num_bins <= 0 || num_bins > 1500) {
        adc::set_error("adc_conf_pr_histogram: bad argument (1 <= num_bins <= 1500)");
        return ADC_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = adc::check_cuda(cudaMemsetAsync(hist, 0, 3 * (size_t)(num_bins + 1) * sizeof(unsigned long long), st), "histogram memset");
    if (rc) return rc;
    adc::Prof prof(st, adc::KIND_TRAIN, (double)n * 12.0);
    const size_t smem = 3 * (size_t)(num_bins + 1) * sizeof(unsigned long long) + (size_t)num_bins * sizeof(float);
    long long b = (n + 255) / 256;
    conf_pr_hist_kernel<<<(int)(b < 148 * 4 ? (b > 0 ? b : 1) : 148 * 4), 256, smem, st>>>(depth, confidence, target, n, thresholds, num_bins, hist);
    return adc::check_cuda(cudaGetLastError(), "conf_pr_hist_kernel launch");
}

}  // extern "C"

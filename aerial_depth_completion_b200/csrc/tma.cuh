// TMA (cp.async.bulk.tensor) + mbarrier helpers for the tile loaders.  One elected thread issues a 4-D box copy
// (W, H, channel, image) per plane; out-of-bounds box elements are zero-filled by the hardware, which IS the
// convolution's zero padding, and the copy lands asynchronously while other CTAs of the SM compute.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "adc_common.cuh"

namespace adc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Can a [N, C, H, W] fp32 tensor with the given batch stride be described to TMA?
inline bool tma_ok(const float* base, long long batch_stride, int H, int W) {
    return base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (W % 4 == 0) && (((long long)H * W) % 4 == 0) &&
           (batch_stride % 4 == 0);
}

// 4-D map (W, H, C, N) with box (box_w, box_h, 1, 1).  Returns 0 on success.
inline int make_plane_map(CUtensorMap* m, const float* base, long long batch_stride, int N, int C, int H, int W, int box_w,
                          int box_h) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)batch_stride * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d C=%d N=%d box %dx%d)", (int)r, W, H, C, N, box_w, box_h);
        return ADC_E_UNSUPP;
    }
    return 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// box at (w, h, c, n) -> dst (dense [box_h][box_w] floats), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, uint64_t* bar, int w, int h, int c, int n) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(w), "r"(h), "r"(c), "r"(n)
        : "memory");
}

}  // namespace adc

// Cross-TU instrumentation: every kernel launch of the library is counted (adc_launch_count) and, when profiling is enabled
// (adc_profile_enable), bracketed by a CUDA event pair on the launching stream (adc_profile_read).
#pragma once
#include <cuda_runtime.h>

namespace adc {
int prof_begin(cudaStream_t st, int kind, int call, double work);   // returns a record index or -1; counts one launch
void prof_end(cudaStream_t st, int idx);
void count_launches(int n);                                          // extra launches inside an already open scope

struct Prof {
    cudaStream_t st; int idx;
    Prof(cudaStream_t s, int kind, double work, int call = -1) : st(s), idx(prof_begin(s, kind, call, work)) {}
    ~Prof() { prof_end(st, idx); }
};
// record kinds of the tensor-core translation units (0..10 are the NConv kernels, adc_nconv.cu); `work` = FLOPs for the
// MMA kernels, bytes for the glue
enum { KIND_TC_FPROP_V1 = 20, KIND_TC_FPROP_V2 = 21, KIND_TC_DGRAD_S2 = 22, KIND_TC_WGRAD = 23, KIND_TC_WGRAD_BOX = 24, KIND_TC_GLUE = 25,
       KIND_TC_LAYOUT = 26, KIND_TC_PREP = 27, KIND_HEAD10 = 28, KIND_TRAIN = 29, KIND_CONF_HEAD = 30, KIND_BN = 31 };
}  // namespace adc

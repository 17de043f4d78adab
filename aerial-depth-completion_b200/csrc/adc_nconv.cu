// C-ABI entry points of the NConv path (include/adc_b200.h) + the pyramid orchestration.
// Single translation unit for every kernel that reads the constant-memory weight images.
#include <stdarg.h>
#include <string.h>

#include "adc_common.cuh"
#include "nconv_bwd.cuh"
#include "nconv_fwd.cuh"
#include "pool.cuh"
#include "prep.cuh"

namespace adc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

static float* g_stage = nullptr;  // [ADC_WSLOTS][2][ADC_PYR_CALLS][kWStride] device staging for the single-layer API

static int ensure_init() {
    if (g_stage) return 0;
    ADC_CUDA(cudaMalloc(&g_stage, sizeof(float) * ADC_WSLOTS * 2 * ADC_PYR_CALLS * kWStride));
    return 0;
}

static inline int wimg_index(int slot, int fb, int call) { return (slot * 2 + fb) * ADC_PYR_CALLS + call; }

// copy `ncalls` images (device memory) into constant memory, stream ordered
static int upload_images(const float* img, int slot, int fb, int first_call, int ncalls, cudaStream_t st) {
    return check_cuda(cudaMemcpyToSymbolAsync(c_wimg, img, sizeof(float) * kWStride * ncalls,
                                              sizeof(float) * kWStride * wimg_index(slot, fb, first_call),
                                              cudaMemcpyDeviceToDevice, st),
                      "cudaMemcpyToSymbolAsync(c_wimg)");
}

static int dispatch_fwd(int Ci, int K, const Src& s0, const Src& s1, float* y, float* cout, int N, int H, int W, int wimg,
                        cudaStream_t st) {
    if (Ci == 1 && K == 5) return launch_nconv_fwd<1, 5>(s0, s1, y, cout, N, H, W, wimg, st);
    if (Ci == 2 && K == 5) return launch_nconv_fwd<2, 5>(s0, s1, y, cout, N, H, W, wimg, st);
    if (Ci == 4 && K == 3) return launch_nconv_fwd<4, 3>(s0, s1, y, cout, N, H, W, wimg, st);
    if (Ci == 1 && K == 3) return launch_nconv_fwd<1, 3>(s0, s1, y, cout, N, H, W, wimg, st);
    if (Ci == 2 && K == 3) return launch_nconv_fwd<2, 3>(s0, s1, y, cout, N, H, W, wimg, st);
    set_error("NConv2d forward: (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}

static int dispatch_bwd_data(int Ci, int K, const Src& s0, const Src& s1, const GradIn& g, float* gx0, float* gc0, float* gx1,
                             float* gc1, int N, int H, int W, int wimg, cudaStream_t st) {
    if (Ci == 1 && K == 5) return launch_nconv_bwd_data<1, 5>(s0, s1, g, gx0, gc0, gx1, gc1, N, H, W, wimg, st);
    if (Ci == 2 && K == 5) return launch_nconv_bwd_data<2, 5>(s0, s1, g, gx0, gc0, gx1, gc1, N, H, W, wimg, st);
    if (Ci == 4 && K == 3) return launch_nconv_bwd_data<4, 3>(s0, s1, g, gx0, gc0, gx1, gc1, N, H, W, wimg, st);
    if (Ci == 1 && K == 3) return launch_nconv_bwd_data<1, 3>(s0, s1, g, gx0, gc0, gx1, gc1, N, H, W, wimg, st);
    if (Ci == 2 && K == 3) return launch_nconv_bwd_data<2, 3>(s0, s1, g, gx0, gc0, gx1, gc1, N, H, W, wimg, st);
    set_error("NConv2d backward(data): (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}

static int dispatch_bwd_weight(int Ci, int K, const Src& s0, const Src& s1, const GradIn& g, float* partial, int N, int H, int W,
                               int wimg, cudaStream_t st) {
    if (Ci == 1 && K == 5) return launch_nconv_bwd_weight<1, 5>(s0, s1, g, partial, N, H, W, wimg, st);
    if (Ci == 2 && K == 5) return launch_nconv_bwd_weight<2, 5>(s0, s1, g, partial, N, H, W, wimg, st);
    if (Ci == 4 && K == 3) return launch_nconv_bwd_weight<4, 3>(s0, s1, g, partial, N, H, W, wimg, st);
    if (Ci == 1 && K == 3) return launch_nconv_bwd_weight<1, 3>(s0, s1, g, partial, N, H, W, wimg, st);
    if (Ci == 2 && K == 3) return launch_nconv_bwd_weight<2, 3>(s0, s1, g, partial, N, H, W, wimg, st);
    set_error("NConv2d backward(weight): (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}

static inline int tiles(int H, int W) { return ceil_div(W, 64) * ceil_div(H, 32); }
static inline int blocks_1x1(long long plane) { return (int)((plane + 1023) / 1024); }

// number of per-CTA partial rows the weight-gradient kernel of a layer writes
static inline int wg_blocks(int N, int Co, int H, int W) {
    return Co == 1 ? N * blocks_1x1((long long)H * W) : N * tiles(H, W);
}

static int run_layer_fwd(const Src& s0, const Src* s1, int Ci, int Co, int K, float* y, float* cout, int N, int H, int W,
                         int wimg, cudaStream_t st) {
    Src dummy = s0;
    if (Co == 2) return dispatch_fwd(Ci, K, s0, s1 ? *s1 : dummy, y, cout, N, H, W, wimg, st);
    if (Co == 1 && Ci == 2 && K == 1 && !s1 && !s0.up) {
        const long long plane = (long long)H * W;
        dim3 grid(blocks_1x1(plane), N);
        nconv_1x1_c2_fwd_kernel<<<grid, 256, 0, st>>>(s0.x, s0.c, s0.bs, y, cout, plane, wimg);
        return check_cuda(cudaGetLastError(), "nconv_1x1_c2_fwd_kernel launch");
    }
    set_error("NConv2d forward: (Ci=%d, Co=%d, k=%d) is not instantiated", Ci, Co, K);
    return ADC_E_UNSUPP;
}

// gx/gc for an upsampled source are produced at layer resolution into `up_scratch` and then reduced.
static int run_layer_bwd(const Src& s0, const Src* s1, int Ci, int Co, int K, const GradIn& g, float* gx0, float* gc0, float* gx1,
                         float* gc1, float* partial, float* up_scratch, int N, int H, int W, int wimg_f, int wimg_b,
                         bool need_data, bool need_weight, cudaStream_t st) {
    Src dummy = s0;
    const long long plane = (long long)H * W;
    if (Co == 1) {
        if (!(Ci == 2 && K == 1 && !s1 && !s0.up)) {
            set_error("NConv2d backward: (Ci=%d, Co=1, k=%d) is not instantiated", Ci, K);
            return ADC_E_UNSUPP;
        }
        dim3 grid(blocks_1x1(plane), N);
        nconv_1x1_c2_bwd_kernel<<<grid, 256, 0, st>>>(s0.x, s0.c, s0.bs, g, need_data ? gx0 : nullptr, gc0, partial, plane,
                                                     wimg_f, wimg_b);
        return check_cuda(cudaGetLastError(), "nconv_1x1_c2_bwd_kernel launch");
    }
    if (need_weight) {
        int rc = dispatch_bwd_weight(Ci, K, s0, s1 ? *s1 : dummy, g, partial, N, H, W, wimg_f, st);
        if (rc) return rc;
    }
    if (need_data) {
        // destinations at layer resolution
        float* d0x = gx0; float* d0c = gc0; float* d1x = gx1; float* d1c = gc1;
        float* sc = up_scratch;
        if (s0.up) { d0x = sc; d0c = sc + (long long)N * s0.ch * plane; sc += 2LL * N * s0.ch * plane; }
        if (s1 && s1->up) { d1x = sc; d1c = sc + (long long)N * s1->ch * plane; }
        // the data kernel reads b and s from the image it is given: the backward image carries both
        int rc = dispatch_bwd_data(Ci, K, s0, s1 ? *s1 : dummy, g, d0x, d0c, d1x, d1c, N, H, W, wimg_b, st);
        if (rc) return rc;
        const Src* ss[2] = {&s0, s1};
        float* dx[2] = {d0x, d1x}; float* dc[2] = {d0c, d1c};
        float* ox[2] = {gx0, gx1}; float* oc[2] = {gc0, gc1};
        for (int k = 0; k < 2; ++k) {
            if (!ss[k] || !ss[k]->up) continue;
            const long long planes = (long long)N * ss[k]->ch;
            const long long total = planes * ss[k]->H * ss[k]->W;
            const int nb = (int)((total + 255) / 256);
            upsample_bwd_kernel<<<nb, 256, 0, st>>>(dx[k], ox[k], H, W, ss[k]->H, ss[k]->W, ss[k]->mh, ss[k]->mw, planes);
            upsample_bwd_kernel<<<nb, 256, 0, st>>>(dc[k], oc[k], H, W, ss[k]->H, ss[k]->W, ss[k]->mh, ss[k]->mw, planes);
            int rc2 = check_cuda(cudaGetLastError(), "upsample_bwd_kernel launch");
            if (rc2) return rc2;
        }
    }
    return 0;
}

// ---- pyramid geometry -------------------------------------------------------------------------
struct Geo {
    int N, h[4], w[4];
    long long px[4];  // N*h*w
    Geo(int N_, int H, int W) : N(N_) {
        h[0] = H; w[0] = W;
        for (int l = 1; l < 4; ++l) { h[l] = h[l - 1] / 2; w[l] = w[l - 1] / 2; }
        for (int l = 0; l < 4; ++l) px[l] = (long long)N * h[l] * w[l];
    }
};

// saved activations: each entry is an (x, c) pair of [N,2,h,w] tensors: x at off, c at off + 2*px
enum Act { A0, A1, A2, P1, A3, A4, P2, A5, P3, A6, A7, A8, A9, ACT_COUNT };
static const int kActLevel[ACT_COUNT] = {0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 2, 1, 0};
static const long long kImgFloats = (long long)ADC_PYR_CALLS * kWStride;

struct SavedLayout {
    long long off[ACT_COUNT];
    long long total;
    explicit SavedLayout(const Geo& g) {
        long long o = kImgFloats;  // forward weight images first
        for (int a = 0; a < ACT_COUNT; ++a) {
            off[a] = o;
            o += 4 * g.px[kActLevel[a]];
        }
        total = o;
    }
};

// module index / shape of each of the 11 calls (model_zoo/nconv_sd.py:38-97)
static const int kCallModule[ADC_PYR_CALLS] = {0, 1, 2, 1, 2, 1, 1, 3, 4, 5, 6};
static const int kCallLevel[ADC_PYR_CALLS] = {0, 0, 0, 1, 1, 2, 3, 2, 1, 0, 0};
static const int kModCo[ADC_PYR_MODULES] = {2, 2, 2, 2, 2, 2, 1};
static const int kModCi[ADC_PYR_MODULES] = {1, 2, 2, 4, 4, 4, 2};
static const int kModK[ADC_PYR_MODULES] = {5, 5, 5, 3, 3, 3, 1};
static inline int mod_numel(int m) { return kModCo[m] * kModCi[m] * kModK[m] * kModK[m]; }
static long long mod_offset(int m) {
    long long o = 0;
    for (int i = 0; i < m; ++i) o += mod_numel(i);
    return o;
}

static adc_src_t mk_src(const float* base, long long px, int N, int ch, int h, int w) {
    adc_src_t s;
    s.x = base;
    s.c = base + 2 * px;  // (x, c) pair layout of the saved activations (2 channels each)
    s.batch_stride = (long long)ch * h * w;
    s.channels = ch;
    s.H = h;
    s.W = w;
    (void)N;
    return s;
}

}  // namespace adc

using namespace adc;

extern "C" {

int adc_abi_version(void) { return 1; }
const char* adc_last_error(void) { return g_err; }

int adc_softplus_pow(const float* w, float* out, int numel, int times, adc_stream_t stream) {
    ADC_CHECK_ARG(w && out && numel > 0 && times >= 0, ADC_E_BADARG, "adc_softplus_pow: bad argument");
    softplus_pow_kernel<<<ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(w, out, numel, times);
    return check_cuda(cudaGetLastError(), "softplus_pow_kernel launch");
}

int adc_conf_pool_fwd(const float* x, const float* c, float* x_ds, float* c_ds, long long* idx, int N, int C, int H, int W,
                      adc_stream_t stream) {
    ADC_CHECK_ARG(x && c && x_ds && c_ds && idx && N > 0 && C > 0 && H >= 2 && W >= 2, ADC_E_BADARG,
                  "adc_conf_pool_fwd: bad argument");
    const int Ho = H / 2, Wo = W / 2;
    const long long planes = (long long)N * C, total = planes * Ho * Wo;
    conf_pool_fwd_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, c, x_ds, c_ds, idx, H, W, Ho, Wo, planes);
    return check_cuda(cudaGetLastError(), "conf_pool_fwd_kernel launch");
}

int adc_conf_pool_bwd(const float* g_xds, const float* g_cds, const long long* idx, const float* add_x, const float* add_c,
                      float* gx, float* gc, int N, int C, int H, int W, adc_stream_t stream) {
    ADC_CHECK_ARG(g_xds && g_cds && idx && gx && gc && N > 0 && C > 0 && H >= 2 && W >= 2, ADC_E_BADARG,
                  "adc_conf_pool_bwd: bad argument");
    const int Ho = H / 2, Wo = W / 2;
    const long long planes = (long long)N * C, total = planes * ((H + 1) / 2) * ((W + 1) / 2);
    conf_pool_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_xds, g_cds, idx, add_x, add_c, gx, gc, H,
                                                                                       W, Ho, Wo, planes);
    return check_cuda(cudaGetLastError(), "conf_pool_bwd_kernel launch");
}

static int check_layer_args(const char* fn, const adc_src_t* s0, const adc_src_t* s1, int N, int Co, int H, int W, int k, int slot) {
    ADC_CHECK_ARG(s0 && s0->x && s0->c && s0->channels > 0, ADC_E_BADARG, "%s: src0 missing", fn);
    ADC_CHECK_ARG(!s1 || (s1->x && s1->c && s1->channels > 0), ADC_E_BADARG, "%s: src1 incomplete", fn);
    ADC_CHECK_ARG(N > 0 && H > 0 && W > 0 && (Co == 1 || Co == 2) && k >= 1 && (k & 1), ADC_E_BADARG, "%s: bad sizes", fn);
    ADC_CHECK_ARG(slot >= 0 && slot < ADC_WSLOTS, ADC_E_BADARG, "%s: slot out of range", fn);
    const int Ci = s0->channels + (s1 ? s1->channels : 0);
    ADC_CHECK_ARG(Co * Ci * k * k <= kWMax, ADC_E_UNSUPP, "%s: Co*Ci*k*k = %d exceeds %d", fn, Co * Ci * k * k, kWMax);
    return 0;
}

int adc_nconv2d_fwd(const adc_src_t* src0, const adc_src_t* src1, const float* w, const float* b, float* y, float* cout, int N,
                    int Co, int H, int W, int k, int slot, adc_stream_t stream) {
    int rc = check_layer_args("adc_nconv2d_fwd", src0, src1, N, Co, H, W, k, slot);
    if (rc) return rc;
    ADC_CHECK_ARG(w && b && y && cout, ADC_E_BADARG, "adc_nconv2d_fwd: null pointer");
    if ((rc = ensure_init())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ci = src0->channels + (src1 ? src1->channels : 0);
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.call[0] = PrepCall{w, b, nullptr, Co, Ci, k, 0, 0, nullptr};
    pa.img = g_stage + (size_t)wimg_index(slot, 0, 0) * kWStride;
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(pa.img, slot, 0, 0, 1, st))) return rc;
    Src s0 = make_src(src0, H, W), s1;
    if (src1) s1 = make_src(src1, H, W);
    return run_layer_fwd(s0, src1 ? &s1 : nullptr, Ci, Co, k, y, cout, N, H, W, wimg_index(slot, 0, 0), st);
}

size_t adc_nconv2d_bwd_scratch_floats(int N, int Ci, int Co, int H, int W, int k) {
    (void)k;
    // per-CTA weight-gradient partials + layer-resolution gradients of (possibly) upsampled sources
    return (size_t)wg_blocks(N, Co, H, W) * kWStride + (size_t)2 * N * Ci * H * W;
}

int adc_nconv2d_bwd(const adc_src_t* src0, const adc_src_t* src1, const float* y, const float* cout, const float* gy,
                    const float* gcout, const float* w_fwd, const float* w_bwd, const float* b, float* gx0, float* gc0, float* gx1,
                    float* gc1, float* gw, float* gb, float* scratch, int N, int Co, int H, int W, int k, int slot,
                    adc_stream_t stream) {
    int rc = check_layer_args("adc_nconv2d_bwd", src0, src1, N, Co, H, W, k, slot);
    if (rc) return rc;
    ADC_CHECK_ARG(y && cout && gy && gcout && w_fwd && w_bwd && b && scratch, ADC_E_BADARG, "adc_nconv2d_bwd: null pointer");
    ADC_CHECK_ARG((gx0 == nullptr) == (gc0 == nullptr) && (gx1 == nullptr) == (gc1 == nullptr), ADC_E_BADARG,
                  "adc_nconv2d_bwd: gx/gc must be given in pairs");
    ADC_CHECK_ARG(!src1 || ((gx0 == nullptr) == (gx1 == nullptr)), ADC_E_BADARG,
                  "adc_nconv2d_bwd: input gradients are all-or-none across sources");
    if ((rc = ensure_init())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ci = src0->channels + (src1 ? src1->channels : 0);
    const bool need_data = gx0 != nullptr, need_weight = gw != nullptr;
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    float* img_f = g_stage + (size_t)wimg_index(slot, 0, 0) * kWStride;
    float* img_b = g_stage + (size_t)wimg_index(slot, 1, 0) * kWStride;
    pa.call[0] = PrepCall{w_fwd, b, nullptr, Co, Ci, k, 0, 0, nullptr};
    pa.img = img_f;
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    pa.call[0] = PrepCall{w_bwd, b, nullptr, Co, Ci, k, 0, Co == 2 ? 1 : 0, img_f};
    pa.img = img_b;
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(img_f, slot, 0, 0, 1, st))) return rc;
    if ((rc = upload_images(img_b, slot, 1, 0, 1, st))) return rc;
    Src s0 = make_src(src0, H, W), s1;
    if (src1) s1 = make_src(src1, H, W);
    GradIn g{y, cout, gy, gcout};
    float* partial = scratch;
    const int nblk = wg_blocks(N, Co, H, W);
    float* up_scratch = scratch + (size_t)nblk * kWStride;
    rc = run_layer_bwd(s0, src1 ? &s1 : nullptr, Ci, Co, k, g, gx0, gc0, gx1, gc1, partial, up_scratch, N, H, W,
                       wimg_index(slot, 0, 0), wimg_index(slot, 1, 0), need_data, need_weight || Co == 1, st);
    if (rc) return rc;
    if (need_weight) {
        WgReduceArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.call[0] = WgReduceCall{partial, nblk, 0, Co, Co * Ci * k * k};
        ra.ncalls = 1;
        ra.gw[0] = gw;
        ra.gb[0] = gb;
        ra.accumulate = 1;
        wgrad_reduce_kernel<<<1, 1024, 0, st>>>(ra);
        ADC_CUDA(cudaGetLastError());
    }
    return 0;
}

size_t adc_pyramid_weight_floats(void) { return (size_t)mod_offset(ADC_PYR_MODULES); }

size_t adc_pyramid_saved_floats(int N, int H, int W) {
    Geo g(N, H, W);
    return (size_t)SavedLayout(g).total;
}

size_t adc_pyramid_idx_count(int N, int H, int W) {
    Geo g(N, H, W);
    return (size_t)(2 * (g.px[1] + g.px[2] + g.px[3]));
}

int adc_pyramid_idx_layout(int N, int H, int W, long long offsets[3], int hs[3], int ws[3]) {
    ADC_CHECK_ARG(N > 0 && H >= 8 && W >= 8 && offsets && hs && ws, ADC_E_BADARG, "adc_pyramid_idx_layout: bad argument");
    Geo g(N, H, W);
    long long o = 0;
    for (int l = 0; l < 3; ++l) {
        offsets[l] = o;
        hs[l] = g.h[l + 1];
        ws[l] = g.w[l + 1];
        o += 2 * g.px[l + 1];
    }
    return 0;
}

int adc_pyramid_fwd(const adc_pyramid_fwd_t* a, adc_stream_t stream) {
    ADC_CHECK_ARG(a && a->x0 && a->c0 && a->saved && a->idx && a->xout && a->cout, ADC_E_BADARG, "adc_pyramid_fwd: null pointer");
    ADC_CHECK_ARG(a->N > 0 && a->H >= 8 && a->W >= 8, ADC_E_BADARG, "adc_pyramid_fwd: need N>0 and H,W >= 8 (three 2x2 poolings)");
    ADC_CHECK_ARG(a->slot >= 0 && a->slot < ADC_WSLOTS, ADC_E_BADARG, "adc_pyramid_fwd: slot out of range");
    ADC_CHECK_ARG(!a->training || a->weight_out, ADC_E_BADARG, "adc_pyramid_fwd: training needs weight_out");
    for (int m = 0; m < ADC_PYR_MODULES; ++m)
        ADC_CHECK_ARG(a->weight[m] && a->bias[m], ADC_E_BADARG, "adc_pyramid_fwd: missing parameter %d", m);
    cudaStream_t st = (cudaStream_t)stream;
    Geo g(a->N, a->H, a->W);
    SavedLayout L(g);
    float* S = a->saved;

    // 1. weights: softplus^k per call, sums, kernel layout -> constant memory
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    int seen[ADC_PYR_MODULES] = {0};
    int ncalls_of[ADC_PYR_MODULES] = {0};
    for (int k = 0; k < ADC_PYR_CALLS; ++k) ncalls_of[kCallModule[k]]++;
    for (int k = 0; k < ADC_PYR_CALLS; ++k) {
        const int m = kCallModule[k];
        seen[m]++;
        float* wout = (a->training && seen[m] == ncalls_of[m]) ? a->weight_out + mod_offset(m) : nullptr;
        pa.call[k] = PrepCall{a->weight[m], a->bias[m], wout, kModCo[m], kModCi[m], kModK[m], a->training ? seen[m] : 0, 0, nullptr};
    }
    pa.img = S;
    prep_weights_kernel<<<ADC_PYR_CALLS, 128, 0, st>>>(pa);
    ADC_CUDA(cudaGetLastError());
    int rc = upload_images(S, a->slot, 0, 0, ADC_PYR_CALLS, st);
    if (rc) return rc;

    auto act = [&](int id) { return S + L.off[id]; };
    auto src_of = [&](int id) {
        const int l = kActLevel[id];
        adc_src_t s = mk_src(act(id), g.px[l], g.N, 2, g.h[l], g.w[l]);
        return s;
    };
    auto layer = [&](int call, const adc_src_t& a0, const adc_src_t* a1, float* y, float* cout) -> int {
        const int m = kCallModule[call], l = kCallLevel[call];
        Src s0 = make_src(&a0, g.h[l], g.w[l]), s1;
        if (a1) s1 = make_src(a1, g.h[l], g.w[l]);
        return run_layer_fwd(s0, a1 ? &s1 : nullptr, kModCi[m], kModCo[m], kModK[m], y, cout, g.N, g.h[l], g.w[l],
                             wimg_index(a->slot, 0, call), st);
    };
    auto out_of = [&](int id, float*& y, float*& c) { y = act(id); c = act(id) + 2 * g.px[kActLevel[id]]; };
    auto pool = [&](int in_id, int out_id, int lvl) -> int {  // lvl = level of the input
        float *xi, *ci, *xo, *co;
        out_of(in_id, xi, ci);
        out_of(out_id, xo, co);
        long long off[3]; int hs[3], ws[3];
        adc_pyramid_idx_layout(g.N, a->H, a->W, off, hs, ws);
        return adc_conf_pool_fwd(xi, ci, xo, co, a->idx + off[lvl], g.N, 2, g.h[lvl], g.w[lvl], stream);
    };
#define RUN(expr) do { rc = (expr); if (rc) return rc; } while (0)
    adc_src_t in0;
    in0.x = a->x0; in0.c = a->c0; in0.batch_stride = a->x0_batch_stride; in0.channels = 1; in0.H = a->H; in0.W = a->W;
    ADC_CHECK_ARG(a->x0_batch_stride == a->c0_batch_stride, ADC_E_BADARG, "adc_pyramid_fwd: x0/c0 batch strides must match");
    float *y, *c;
    out_of(A0, y, c); RUN(layer(0, in0, nullptr, y, c));
    { adc_src_t s = src_of(A0); out_of(A1, y, c); RUN(layer(1, s, nullptr, y, c)); }
    { adc_src_t s = src_of(A1); out_of(A2, y, c); RUN(layer(2, s, nullptr, y, c)); }
    RUN(pool(A2, P1, 0));
    { adc_src_t s = src_of(P1); out_of(A3, y, c); RUN(layer(3, s, nullptr, y, c)); }
    { adc_src_t s = src_of(A3); out_of(A4, y, c); RUN(layer(4, s, nullptr, y, c)); }
    RUN(pool(A4, P2, 1));
    { adc_src_t s = src_of(P2); out_of(A5, y, c); RUN(layer(5, s, nullptr, y, c)); }
    RUN(pool(A5, P3, 2));
    { adc_src_t s = src_of(P3); out_of(A6, y, c); RUN(layer(6, s, nullptr, y, c)); }
    { adc_src_t s = src_of(A5), u = src_of(A6); out_of(A7, y, c); RUN(layer(7, s, &u, y, c)); }   // nconv4: cat(x3_ds, up(x4_ds))
    { adc_src_t s = src_of(A4), u = src_of(A7); out_of(A8, y, c); RUN(layer(8, s, &u, y, c)); }   // nconv5: cat(x2_ds, up(x34_ds))
    { adc_src_t u = src_of(A8), s = src_of(A2); out_of(A9, y, c); RUN(layer(9, u, &s, y, c)); }   // nconv6: cat(up(x23_ds), x1)
    { adc_src_t s = src_of(A9); RUN(layer(10, s, nullptr, a->xout, a->cout)); }
#undef RUN
    return 0;
}

// backward scratch: [bwd images][partials per call][gradient buffers per level][upsample scratch]
namespace adc {
struct BwdLayout {
    long long img, partial[ADC_PYR_CALLS], gbuf[4][4], up[3], total;
    int nblk[ADC_PYR_CALLS];
    explicit BwdLayout(const Geo& g) {
        long long o = 0;
        img = o; o += kImgFloats;
        for (int k = 0; k < ADC_PYR_CALLS; ++k) {
            const int m = kCallModule[k], l = kCallLevel[k];
            nblk[k] = wg_blocks(g.N, kModCo[m], g.h[l], g.w[l]);
            partial[k] = o; o += (long long)nblk[k] * kWStride;
        }
        for (int l = 0; l < 4; ++l)
            for (int b = 0; b < 4; ++b) { gbuf[l][b] = o; o += 4 * g.px[l]; }   // (gx, gc) pair of [N,2,h,w]
        for (int l = 0; l < 3; ++l) { up[l] = o; o += 4 * g.px[l]; }             // layer-resolution grads of an upsampled source
        total = o;
    }
};
}  // namespace adc

size_t adc_pyramid_bwd_scratch_floats(int N, int H, int W) {
    Geo g(N, H, W);
    return (size_t)BwdLayout(g).total;
}

int adc_pyramid_bwd(const adc_pyramid_bwd_t* a, adc_stream_t stream) {
    ADC_CHECK_ARG(a && a->x0 && a->c0 && a->saved && a->idx && a->xout && a->cout && a->g_xout && a->g_cout && a->scratch,
                  ADC_E_BADARG, "adc_pyramid_bwd: null pointer");
    ADC_CHECK_ARG(a->N > 0 && a->H >= 8 && a->W >= 8, ADC_E_BADARG, "adc_pyramid_bwd: bad sizes");
    ADC_CHECK_ARG(a->slot >= 0 && a->slot < ADC_WSLOTS, ADC_E_BADARG, "adc_pyramid_bwd: slot out of range");
    ADC_CHECK_ARG(!a->need_input_grad || (a->gx0 && a->gc0), ADC_E_BADARG, "adc_pyramid_bwd: gx0/gc0 missing");
    ADC_CHECK_ARG(!a->need_weight_grad || (a->gweight && a->gbias), ADC_E_BADARG, "adc_pyramid_bwd: gweight/gbias missing");
    ADC_CHECK_ARG(a->x0_batch_stride == a->c0_batch_stride, ADC_E_BADARG, "adc_pyramid_bwd: x0/c0 batch strides must match");
    for (int m = 0; m < ADC_PYR_MODULES; ++m)
        ADC_CHECK_ARG(a->weight_bwd[m] && a->bias[m], ADC_E_BADARG, "adc_pyramid_bwd: missing parameter %d", m);
    cudaStream_t st = (cudaStream_t)stream;
    Geo g(a->N, a->H, a->W);
    SavedLayout L(g);
    BwdLayout B(g);
    const float* S = a->saved;
    float* X = a->scratch;
    int rc;

    // forward images (b, s of each call) back into constant memory; backward-time weights, flipped, next to them
    if ((rc = upload_images(S, a->slot, 0, 0, ADC_PYR_CALLS, st))) return rc;
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    for (int k = 0; k < ADC_PYR_CALLS; ++k) {
        const int m = kCallModule[k];
        pa.call[k] = PrepCall{a->weight_bwd[m], a->bias[m], nullptr, kModCo[m], kModCi[m], kModK[m], 0, kModCo[m] == 2 ? 1 : 0,
                              S + (size_t)k * kWStride};
    }
    pa.img = X + B.img;
    prep_weights_kernel<<<ADC_PYR_CALLS, 128, 0, st>>>(pa);
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(X + B.img, a->slot, 1, 0, ADC_PYR_CALLS, st))) return rc;

    auto actp = [&](int id) { return S + L.off[id]; };
    auto src_of = [&](int id) {
        const int l = kActLevel[id];
        return mk_src(actp(id), g.px[l], g.N, 2, g.h[l], g.w[l]);
    };
    auto gb = [&](int l, int b) { return X + B.gbuf[l][b]; };
    const bool need_w = a->need_weight_grad != 0;

    // one layer backward.  gin: (gy, gcout) buffer of the layer output; outputs go to (gx0,gc0)/(gx1,gc1)
    auto layer_bwd = [&](int call, const adc_src_t& a0, const adc_src_t* a1, const float* yv, const float* cv, const float* gyv,
                         const float* gcv, float* gx0, float* gc0, float* gx1, float* gc1, bool need_data) -> int {
        const int m = kCallModule[call], l = kCallLevel[call];
        Src s0 = make_src(&a0, g.h[l], g.w[l]), s1;
        if (a1) s1 = make_src(a1, g.h[l], g.w[l]);
        GradIn gi{yv, cv, gyv, gcv};
        return run_layer_bwd(s0, a1 ? &s1 : nullptr, kModCi[m], kModCo[m], kModK[m], gi, gx0, gc0, gx1, gc1, X + B.partial[call],
                             l < 3 ? X + B.up[l] : nullptr, g.N, g.h[l], g.w[l], wimg_index(a->slot, 0, call),
                             wimg_index(a->slot, 1, call), need_data, need_w || kModCo[m] == 1, st);
    };
    auto yc = [&](int id, const float*& y, const float*& c) { y = actp(id); c = actp(id) + 2 * g.px[kActLevel[id]]; };
    auto gpair = [&](int l, int b, float*& gx, float*& gc) { gx = gb(l, b); gc = gb(l, b) + 2 * g.px[l]; };
    long long ioff[3]; int ihs[3], iws[3];
    adc_pyramid_idx_layout(g.N, a->H, a->W, ioff, ihs, iws);
#define RUN(expr) do { rc = (expr); if (rc) return rc; } while (0)
    const float *y, *c;
    float *gx, *gc, *hx, *hc, *kx, *kc;

    // call 10: nconv7 on A9 -> grads of A9 in gbuf[0][0]
    gpair(0, 0, gx, gc);
    { adc_src_t s = src_of(A9); RUN(layer_bwd(10, s, nullptr, a->xout, a->cout, a->g_xout, a->g_cout, gx, gc, nullptr, nullptr, true)); }
    // call 9: nconv6 on cat(up(A8), A2): grads of A8 (L1) -> gbuf[1][0], skip grads of A2 -> gbuf[0][2]
    yc(A9, y, c);
    gpair(1, 0, hx, hc); gpair(0, 2, kx, kc);
    { adc_src_t u = src_of(A8), s = src_of(A2); RUN(layer_bwd(9, u, &s, y, c, gx, gc, hx, hc, kx, kc, true)); }
    // call 8: nconv5 on cat(A4, up(A7)): skip grads of A4 -> gbuf[1][2], grads of A7 (L2) -> gbuf[2][0]
    yc(A8, y, c);
    gpair(1, 2, gx, gc); gpair(2, 0, kx, kc);
    { adc_src_t s = src_of(A4), u = src_of(A7); RUN(layer_bwd(8, s, &u, y, c, hx, hc, gx, gc, kx, kc, true)); }
    // call 7: nconv4 on cat(A5, up(A6)): skip grads of A5 -> gbuf[2][2], grads of A6 (L3) -> gbuf[3][0]
    yc(A7, y, c);
    gpair(2, 2, gx, gc); gpair(3, 0, hx, hc);
    { adc_src_t s = src_of(A5), u = src_of(A6); RUN(layer_bwd(7, s, &u, y, c, kx, kc, gx, gc, hx, hc, true)); }
    // call 6: nconv2 @L3 on P3 -> grads of P3 in gbuf[3][1]
    yc(A6, y, c);
    gpair(3, 1, gx, gc);
    { adc_src_t s = src_of(P3); RUN(layer_bwd(6, s, nullptr, y, c, hx, hc, gx, gc, nullptr, nullptr, true)); }
    // pool3 backward (+ skip grads of A5) -> grads of A5 in gbuf[2][1]
    gpair(2, 2, kx, kc); gpair(2, 1, hx, hc);
    RUN(adc_conf_pool_bwd(gx, gc, a->idx + ioff[2], kx, kc, hx, hc, g.N, 2, g.h[2], g.w[2], stream));
    // call 5: nconv2 @L2 on P2 -> grads of P2 in gbuf[2][3]
    yc(A5, y, c);
    gpair(2, 3, gx, gc);
    { adc_src_t s = src_of(P2); RUN(layer_bwd(5, s, nullptr, y, c, hx, hc, gx, gc, nullptr, nullptr, true)); }
    // pool2 backward (+ skip grads of A4 in gbuf[1][2]) -> grads of A4 in gbuf[1][1]
    gpair(1, 2, kx, kc); gpair(1, 1, hx, hc);
    RUN(adc_conf_pool_bwd(gx, gc, a->idx + ioff[1], kx, kc, hx, hc, g.N, 2, g.h[1], g.w[1], stream));
    // call 4: nconv3 @L1 on A3 -> gbuf[1][3];  call 3: nconv2 @L1 on P1 -> gbuf[1][0]
    yc(A4, y, c);
    gpair(1, 3, gx, gc);
    { adc_src_t s = src_of(A3); RUN(layer_bwd(4, s, nullptr, y, c, hx, hc, gx, gc, nullptr, nullptr, true)); }
    yc(A3, y, c);
    gpair(1, 0, hx, hc);
    { adc_src_t s = src_of(P1); RUN(layer_bwd(3, s, nullptr, y, c, gx, gc, hx, hc, nullptr, nullptr, true)); }
    // pool1 backward (+ skip grads of A2 in gbuf[0][2]) -> grads of A2 in gbuf[0][1]
    gpair(0, 2, kx, kc); gpair(0, 1, gx, gc);
    RUN(adc_conf_pool_bwd(hx, hc, a->idx + ioff[0], kx, kc, gx, gc, g.N, 2, g.h[0], g.w[0], stream));
    // call 2: nconv3 @L0 on A1 -> gbuf[0][0];  call 1: nconv2 @L0 on A0 -> gbuf[0][3];  call 0: nconv1 on (x0, c0)
    yc(A2, y, c);
    gpair(0, 0, hx, hc);
    { adc_src_t s = src_of(A1); RUN(layer_bwd(2, s, nullptr, y, c, gx, gc, hx, hc, nullptr, nullptr, true)); }
    yc(A1, y, c);
    gpair(0, 3, gx, gc);
    { adc_src_t s = src_of(A0); RUN(layer_bwd(1, s, nullptr, y, c, hx, hc, gx, gc, nullptr, nullptr, true)); }
    yc(A0, y, c);
    {
        adc_src_t in0;
        in0.x = a->x0; in0.c = a->c0; in0.batch_stride = a->x0_batch_stride; in0.channels = 1; in0.H = a->H; in0.W = a->W;
        const bool nd = a->need_input_grad != 0;
        if (nd || need_w)
            RUN(layer_bwd(0, in0, nullptr, y, c, gx, gc, nd ? a->gx0 : nullptr, nd ? a->gc0 : nullptr, nullptr, nullptr, nd));
    }
#undef RUN
    if (need_w) {
        WgReduceArgs ra;
        memset(&ra, 0, sizeof(ra));
        for (int k = 0; k < ADC_PYR_CALLS; ++k) {
            const int m = kCallModule[k];
            ra.call[k] = WgReduceCall{X + B.partial[k], B.nblk[k], m, kModCo[m], mod_numel(m)};
        }
        ra.ncalls = ADC_PYR_CALLS;
        long long bo = 0;
        for (int m = 0; m < ADC_PYR_MODULES; ++m) {
            ra.gw[m] = a->gweight + mod_offset(m);
            ra.gb[m] = a->gbias + bo;
            bo += kModCo[m];
        }
        ra.accumulate = 0;
        wgrad_reduce_kernel<<<ADC_PYR_MODULES, 1024, 0, st>>>(ra);
        ADC_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"

// C-ABI entry points of the NConv path (include/adc_b200.h) + the pyramid orchestration.
// Single translation unit for every kernel that reads the constant-memory weight images.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "adc_common.cuh"
#include "prof.h"
#include "nconv_bwd.cuh"
#include "nconv_fwd.cuh"
#include "pool.cuh"
#include "prep.cuh"
#include "tma.cuh"

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

// ---- instrumentation: launch counter (always on) and optional per-kernel CUDA-event timing -----------------------
static long long g_launches = 0;
struct ProfRec { cudaEvent_t a, b; int kind, call; double bytes; };
static const int kProfMax = 1024;
static ProfRec g_prof[kProfMax];
static int g_prof_n = 0, g_prof_on = 0, g_prof_events = 0;

int prof_begin(cudaStream_t st, int kind, int call, double work) {
    g_launches++;
    if (!g_prof_on || g_prof_n >= kProfMax) return -1;
    if (g_prof_n >= g_prof_events) {
        if (cudaEventCreate(&g_prof[g_prof_n].a) != cudaSuccess || cudaEventCreate(&g_prof[g_prof_n].b) != cudaSuccess) return -1;
        g_prof_events = g_prof_n + 1;
    }
    const int idx = g_prof_n++;
    g_prof[idx].kind = kind; g_prof[idx].call = call; g_prof[idx].bytes = work;
    cudaEventRecord(g_prof[idx].a, st);
    return idx;
}
void prof_end(cudaStream_t st, int idx) { if (idx >= 0) cudaEventRecord(g_prof[idx].b, st); }
void count_launches(int n) { g_launches += n; }

struct ProfScope {   // records an event pair around the launches made while it is alive
    cudaStream_t st; int idx;
    ProfScope(cudaStream_t s, int kind, int call, double bytes) : st(s), idx(prof_begin(s, kind, call, bytes)) {}
    ~ProfScope() { prof_end(st, idx); }
};
enum { KIND_FWD = 0, KIND_BWD_DATA = 1, KIND_BWD_WEIGHT = 2, KIND_POOL_FWD = 3, KIND_POOL_BWD = 4, KIND_UP_BWD = 5, KIND_HEAD_FWD = 6,
       KIND_HEAD_BWD = 7, KIND_PACK = 8, KIND_REDUCE = 9, KIND_PREP = 10 };
static int g_cur_call = -1;   // pyramid call index of the launches being issued (for the profile records)

static float* g_stage = nullptr;  // [ADC_WSLOTS][2][ADC_PYR_CALLS][kWStride] device staging for the single-layer API
static int* g_counters = nullptr; // tile-scheduler counters of the persistent kernels (self-resetting)
static int g_persist = 1, g_sms = 148;
static int g_use_tma = -1;
static int g_side_on = 1;        // weight-gradient kernels of adc_pyramid_bwd on a second stream (ADC_NO_SIDE_STREAM=1 disables)
static int g_side_user = 1;      // run-time switch (adc_set_side_stream): parity tests compare the two-stream schedule with the serial one
static cudaStream_t g_side = nullptr;
static cudaEvent_t g_ev_fork[ADC_PYR_CALLS], g_ev_join[ADC_PYR_CALLS];
static int g_up2 = 1;           // exact-2x upsampled sources through TMA + in-place expansion (ADC_NO_UP2=1 disables)

static int ensure_init() {
    if (g_use_tma < 0) {
        const char* e = getenv("ADC_NO_TMA");   // testing aid: force the gather loaders everywhere
        g_use_tma = (e && e[0] == '1') ? 0 : 1;
        const char* u = getenv("ADC_NO_UP2");
        g_up2 = (u && u[0] == '1') ? 0 : 1;
    }
    if (g_stage) return 0;
    const char* pe = getenv("ADC_FWD_MODE");     // 0: one tile per CTA, 2: persistent, 1 (default): persistent only for gathered sources
    g_persist = pe ? atoi(pe) : 1;
    int dev = 0;
    ADC_CUDA(cudaGetDevice(&dev));
    ADC_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    ADC_CUDA(cudaMalloc(&g_counters, sizeof(int) * 64));
    ADC_CUDA(cudaMemset(g_counters, 0, sizeof(int) * 64));
    ADC_CUDA(cudaMalloc(&g_stage, sizeof(float) * ADC_WSLOTS * 2 * ADC_PYR_CALLS * kWStride));
    const char* se = getenv("ADC_NO_SIDE_STREAM");
    g_side_on = (se && se[0] == '1') ? 0 : 1;
    if (g_side_on) {
        ADC_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
        for (int k = 0; k < ADC_PYR_CALLS; ++k) {
            ADC_CUDA(cudaEventCreateWithFlags(&g_ev_fork[k], cudaEventDisableTiming));
            ADC_CUDA(cudaEventCreateWithFlags(&g_ev_join[k], cudaEventDisableTiming));
        }
    }
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

// ---- pair-tensor sources ------------------------------------------------------------------------------------------
static SrcP make_srcp(const float2* p, long long bs, int ch, int Hs, int Ws, int H, int W) {
    SrcP r;
    r.p = p;
    r.bs = bs;
    r.ch = ch;
    r.H = Hs;
    r.W = Ws;
    r.mh = NearestMap::make(Hs, H);
    r.mw = NearestMap::make(Ws, W);
    r.up = (Hs != H || Ws != W) ? 1 : 0;
    // TMA needs a 16-byte aligned base and 16-byte multiples for every stride: even width, even batch stride
    const bool tma_able = g_use_tma == 1 && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (Ws % 2 == 0) && (bs % 2 == 0);
    r.tma = (tma_able && !r.up) ? 1 : 0;
    r.up2 = (tma_able && g_up2 && H == 2 * Hs && W == 2 * Ws) ? 1 : 0;
    return r;
}

// 4-D map over a pair tensor seen as floats: (2W, H, C, N), box (2*box_w_pairs, box_h, 1, 1)
static int make_pair_map(CUtensorMap* m, const SrcP& s, int N, int box_w_pairs, int box_h) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return ADC_E_UNSUPP;
    }
    cuuint64_t dims[4] = {(cuuint64_t)2 * s.W, (cuuint64_t)s.H, (cuuint64_t)s.ch, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)s.W * 8, (cuuint64_t)s.H * s.W * 8, (cuuint64_t)s.bs * 8};
    cuuint32_t box[4] = {(cuuint32_t)(2 * box_w_pairs), (cuuint32_t)box_h, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float2*>(s.p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d C=%d N=%d box %dx%d)", (int)r, s.W, s.H, s.ch, N,
                  box_w_pairs, box_h);
        return ADC_E_UNSUPP;
    }
    return 0;
}

template <int K, int TH_ = 32>
static int fill_maps(Maps2* maps, const SrcP& s0, const SrcP* s1, int N) {
    memset(maps, 0, sizeof(*maps));
    using C = TileCfg<K, TH_>;
    int rc = 0;
    using U = Up2Cfg<K, TH_>;
    if (s0.tma && (rc = make_pair_map(&maps->m0, s0, N, C::SWP, C::SH))) return rc;
    if (s0.up2 && (rc = make_pair_map(&maps->m0, s0, N, U::LW, U::LH))) return rc;
    if (s1 && s1->tma && (rc = make_pair_map(&maps->m1, *s1, N, C::SWP, C::SH))) return rc;
    if (s1 && s1->up2 && (rc = make_pair_map(&maps->m1, *s1, N, U::LW, U::LH))) return rc;
    return 0;
}

template <typename Kern>
static int set_smem(Kern kern, int bytes) {
    return check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute(smem)");
}

static inline dim3 tile_grid(int N, int H, int W) { return dim3(ceil_div(W, 64), ceil_div(H, 32), N); }
static inline int tiles(int N, int H, int W) { return N * ceil_div(W, 64) * ceil_div(H, 32); }
// the weight-gradient kernel works on 64 x 16 tiles (WgCfg::TH): its grid and its per-CTA partial rows
static inline dim3 wg_grid(int N, int H, int W) { return dim3(ceil_div(W, 64), ceil_div(H, 16), N); }
static inline int wg_tiles(int N, int H, int W) { return N * ceil_div(W, 64) * ceil_div(H, 16); }
static inline int blocks_1x1(long long plane) { return (int)ceil_div_ll(plane, 512); }

template <int CI, int K>
static int launch_fwd_persist(const SrcP& s0, const SrcP* s1, float2* out, int N, int H, int W, int wimg, cudaStream_t st) {
    using C = TileCfg<K>;
    static bool done = false;
    static int ctas_per_sm = 1;
    const int smem = C::smem_bytes(2);
    if (!done) {
        int rc = set_smem(nconv_fwd_persist_kernel<CI, K>, smem);
        if (rc) return rc;
        ADC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, nconv_fwd_persist_kernel<CI, K>, 128, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        done = true;
    }
    Maps2 maps;
    int rc = fill_maps<K>(&maps, s0, s1, N);
    if (rc) return rc;
    TileSched sch;
    sch.tiles_x = ceil_div(W, C::TW);
    sch.tiles_y = ceil_div(H, C::TH);
    sch.ntiles = sch.tiles_x * sch.tiles_y * N;
    sch.counters = g_counters + 2 * (wimg / (2 * ADC_PYR_CALLS));     // one counter pair per weight slot
    const int grid = sch.ntiles < g_sms * ctas_per_sm ? sch.ntiles : g_sms * ctas_per_sm;
    { ProfScope _p(st, KIND_FWD, g_cur_call, 8.0 * (CI + 2) * N * H * W);
    nconv_fwd_persist_kernel<CI, K><<<grid, 128, smem, st>>>(maps, s0, s1 ? *s1 : s0, s1 ? 1 : 0, out, H, W, wimg, sch); }
    return check_cuda(cudaGetLastError(), "nconv_fwd_persist_kernel launch");
}

template <int CI, int K>
static int launch_fwd(const SrcP& s0, const SrcP* s1, float2* out, int N, int H, int W, int wimg, cudaStream_t st) {
    using C = TileCfg<K>;
    // Measured (profiles/r1_forward_variants.txt): with every source TMA-fed the one-tile-per-CTA kernel is fastest (27 us vs
    // 32 us for nconv2@L0, B=32); when a source goes through the gather path (nearest-upsampled half of nconv4-6) the
    // persistent kernel with cp.async prefetch wins (43 us vs 61 us for nconv6@L0).
    const bool gathered = !(s0.tma || s0.up2) || (s1 && !(s1->tma || s1->up2));
    if (g_persist == 2 || (g_persist == 1 && gathered)) return launch_fwd_persist<CI, K>(s0, s1, out, N, H, W, wimg, st);
    static bool done = false;
    const int smem = C::smem_bytes(CI);
    if (!done) { int rc = set_smem(nconv_fwd_kernel<CI, K>, smem); if (rc) return rc; done = true; }
    Maps2 maps;
    int rc = fill_maps<K>(&maps, s0, s1, N);
    if (rc) return rc;
    { ProfScope _p(st, KIND_FWD, g_cur_call, 8.0 * (CI + 2) * N * H * W);
    nconv_fwd_kernel<CI, K><<<tile_grid(N, H, W), 128, smem, st>>>(maps, s0, s1 ? *s1 : s0, s1 ? 1 : 0, out, H, W, wimg); }
    return check_cuda(cudaGetLastError(), "nconv_fwd_kernel launch");
}

template <int CI, int K>
static int launch_bwd_data(const SrcP& g, const SrcP& s0, const SrcP* s1, float2* out0, float2* out1, float* bpart, int N, int H,
                           int W, int wimg, int ximg, cudaStream_t st) {
    using C = TileCfg<K>;
    static bool done = false;
    const int smem = C::smem_bytes(2);
    if (!done) { int rc = set_smem(nconv_bwd_data_kernel<CI, K>, smem); if (rc) return rc; done = true; }
    Maps2 maps;
    int rc = fill_maps<K>(&maps, g, nullptr, N);
    if (rc) return rc;
    { ProfScope _p(st, KIND_BWD_DATA, g_cur_call, 8.0 * (2 + 2 * CI) * N * H * W);
    nconv_bwd_data_kernel<CI, K><<<tile_grid(N, H, W), 128, smem, st>>>(maps, g, s0, s1 ? *s1 : s0, s1 ? 1 : 0, out0, out1, bpart,
                                                                         H, W, wimg, ximg); }
    return check_cuda(cudaGetLastError(), "nconv_bwd_data_kernel launch");
}

template <int CI, int K>
static int launch_bwd_weight(const SrcP& g, const SrcP& s0, const SrcP* s1, float* wpart, int N, int H, int W, cudaStream_t st) {
    using C = WgCfg<CI, K>;
    static bool done = false;
    const int smem = C::smem_bytes();
    if (!done) { int rc = set_smem(nconv_bwd_weight_kernel<CI, K>, smem); if (rc) return rc; done = true; }
    Maps2 maps;
    int rc = fill_maps<K, C::TH>(&maps, s0, s1, N);
    if (rc) return rc;
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    if (g.tma && (rc = make_pair_map(&gmap, g, N, C::TW, C::TH))) return rc;
    { ProfScope _p(st, KIND_BWD_WEIGHT, g_cur_call, 8.0 * (2 + CI) * N * H * W);
    nconv_bwd_weight_kernel<CI, K><<<wg_grid(N, H, W), C::NT, smem, st>>>(maps, gmap, g, s0, s1 ? *s1 : s0, s1 ? 1 : 0, wpart, H, W); }
    return check_cuda(cudaGetLastError(), "nconv_bwd_weight_kernel launch");
}

#define ADC_DISPATCH(FN, ...)                                         \
    do {                                                              \
        if (Ci == 1 && K == 5) return FN<1, 5>(__VA_ARGS__);          \
        if (Ci == 2 && K == 5) return FN<2, 5>(__VA_ARGS__);          \
        if (Ci == 4 && K == 3) return FN<4, 3>(__VA_ARGS__);          \
        if (Ci == 1 && K == 3) return FN<1, 3>(__VA_ARGS__);          \
        if (Ci == 2 && K == 3) return FN<2, 3>(__VA_ARGS__);          \
    } while (0)

static int dispatch_fwd(int Ci, int K, const SrcP& s0, const SrcP* s1, float2* out, int N, int H, int W, int wimg, cudaStream_t st) {
    ADC_DISPATCH(launch_fwd, s0, s1, out, N, H, W, wimg, st);
    set_error("NConv2d forward: (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}
static int dispatch_bwd_data(int Ci, int K, const SrcP& g, const SrcP& s0, const SrcP* s1, float2* out0, float2* out1, float* bpart,
                             int N, int H, int W, int wimg, int ximg, cudaStream_t st) {
    ADC_DISPATCH(launch_bwd_data, g, s0, s1, out0, out1, bpart, N, H, W, wimg, ximg, st);
    set_error("NConv2d backward(data): (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}
static int dispatch_bwd_weight(int Ci, int K, const SrcP& g, const SrcP& s0, const SrcP* s1, float* wpart, int N, int H, int W,
                               cudaStream_t st) {
    ADC_DISPATCH(launch_bwd_weight, g, s0, s1, wpart, N, H, W, st);
    set_error("NConv2d backward(weight): (Ci=%d, Co=2, k=%d) is not instantiated", Ci, K);
    return ADC_E_UNSUPP;
}

static int run_upsample_bwd(const float2* g_full, const SrcP& s, float2* G, float* bpart, int N, int H, int W, int ximg,
                            cudaStream_t st) {
    const long long planes = (long long)N * s.ch, total = planes * s.H * s.W;
    if (H == 2 * s.H && W == 2 * s.W && total < (1LL << 31) && (reinterpret_cast<uintptr_t>(g_full) & 15) == 0) {
        { ProfScope _p(st, KIND_UP_BWD, g_cur_call, 8.0 * s.ch * N * ((double)H * W + 2.0 * s.H * s.W));
        pair_upsample2_bwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(g_full, s.p, G, bpart, s.ch, s.H, s.W, (unsigned)total, ximg); }
        return check_cuda(cudaGetLastError(), "pair_upsample2_bwd_kernel launch");
    }
    { ProfScope _p(st, KIND_UP_BWD, g_cur_call, 8.0 * s.ch * N * ((double)H * W + 2.0 * s.H * s.W));
    pair_upsample_bwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(g_full, s.p, G, bpart, s.ch, H, W, s.H, s.W, s.mh, s.mw,
                                                                          planes, ximg); }
    return check_cuda(cudaGetLastError(), "pair_upsample_bwd_kernel launch");
}

// ---- pyramid geometry -------------------------------------------------------------------------------------------
struct Geo {
    int N, h[4], w[4];
    long long px[4];  // N*h*w
    Geo(int N_, int H, int W) : N(N_) {
        h[0] = H; w[0] = W;
        for (int l = 1; l < 4; ++l) { h[l] = h[l - 1] / 2; w[l] = w[l - 1] / 2; }
        for (int l = 0; l < 4; ++l) px[l] = (long long)N * h[l] * w[l];
    }
};

// saved activations: pair tensors [N,2,h,w] (AIN: [N,1,H,W])
enum Act { A0, A1, A2, P1, A3, A4, P2, A5, P3, A6, A7, A8, A9, AIN, ACT_COUNT };
static const int kActLevel[ACT_COUNT] = {0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 2, 1, 0, 0};
static const int kActCh[ACT_COUNT] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1};
static const long long kImgFloats = (long long)ADC_PYR_CALLS * kWStride;

struct SavedLayout {
    long long off[ACT_COUNT];
    long long total;
    explicit SavedLayout(const Geo& g) {
        long long o = kImgFloats;  // forward weight images first
        for (int a = 0; a < ACT_COUNT; ++a) {
            off[a] = o;
            o += 2LL * kActCh[a] * g.px[kActLevel[a]];
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

}  // namespace adc

using namespace adc;

extern "C" {

int adc_abi_version(void) { return 2; }

long long adc_launch_count(void) { return g_launches; }

int adc_set_side_stream(int on) {
    const int prev = g_side_user;
    g_side_user = on ? 1 : 0;
    return prev;
}

int adc_profile_enable(int on) {
    g_prof_on = on ? 1 : 0;
    g_prof_n = 0;
    return 0;
}

int adc_profile_read(int max_records, int* n_out, int* kinds, int* calls, float* ms, double* bytes) {
    ADC_CHECK_ARG(n_out && kinds && calls && ms && bytes && max_records >= 0, ADC_E_BADARG, "adc_profile_read: bad argument");
    const int n = g_prof_n < max_records ? g_prof_n : max_records;
    for (int i = 0; i < n; ++i) {
        ADC_CUDA(cudaEventSynchronize(g_prof[i].b));
        float t = 0.f;
        ADC_CUDA(cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b));
        kinds[i] = g_prof[i].kind; calls[i] = g_prof[i].call; ms[i] = t; bytes[i] = g_prof[i].bytes;
    }
    g_prof_n = 0;
    *n_out = n;
    return 0;
}
const char* adc_last_error(void) { return g_err; }

int adc_softplus_pow(const float* w, float* out, int numel, int times, adc_stream_t stream) {
    ADC_CHECK_ARG(w && out && numel > 0 && times >= 0, ADC_E_BADARG, "adc_softplus_pow: bad argument");
    { ProfScope _p((cudaStream_t)stream, KIND_PACK, g_cur_call, 0.0);
    softplus_pow_kernel<<<ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(w, out, numel, times);
    }
    return check_cuda(cudaGetLastError(), "softplus_pow_kernel launch");
}

int adc_conf_pool_fwd(const float* x, const float* c, float* x_ds, float* c_ds, long long* idx, int N, int C, int H, int W,
                      adc_stream_t stream) {
    ADC_CHECK_ARG(x && c && x_ds && c_ds && idx && N > 0 && C > 0 && H >= 2 && W >= 2, ADC_E_BADARG,
                  "adc_conf_pool_fwd: bad argument");
    const int Ho = H / 2, Wo = W / 2;
    const long long planes = (long long)N * C, total = planes * Ho * Wo;
    { ProfScope _p((cudaStream_t)stream, KIND_PACK, g_cur_call, 0.0);
    conf_pool_fwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(x, c, x_ds, c_ds, idx, H, W, Ho, Wo, planes);
    }
    return check_cuda(cudaGetLastError(), "conf_pool_fwd_kernel launch");
}

int adc_conf_pool_bwd(const float* g_xds, const float* g_cds, const long long* idx, const float* add_x, const float* add_c,
                      float* gx, float* gc, int N, int C, int H, int W, adc_stream_t stream) {
    ADC_CHECK_ARG(g_xds && g_cds && idx && gx && gc && N > 0 && C > 0 && H >= 2 && W >= 2, ADC_E_BADARG,
                  "adc_conf_pool_bwd: bad argument");
    const int Ho = H / 2, Wo = W / 2;
    const long long planes = (long long)N * C, total = planes * ((H + 1) / 2) * ((W + 1) / 2);
    { ProfScope _p((cudaStream_t)stream, KIND_PACK, g_cur_call, 0.0);
    conf_pool_bwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(g_xds, g_cds, idx, add_x, add_c, gx, gc, H,
                                                                                         W, Ho, Wo, planes);
    }
    return check_cuda(cudaGetLastError(), "conf_pool_bwd_kernel launch");
}

static int check_layer_args(const char* fn, const adc_src_t* s0, const adc_src_t* s1, int N, int Co, int H, int W, int k, int slot) {
    ADC_CHECK_ARG(s0 && s0->x && s0->c && s0->channels > 0, ADC_E_BADARG, "%s: src0 missing", fn);
    ADC_CHECK_ARG(!s1 || (s1->x && s1->c && s1->channels > 0), ADC_E_BADARG, "%s: src1 incomplete", fn);
    ADC_CHECK_ARG(N > 0 && H > 0 && W > 0 && (Co == 1 || Co == 2) && k >= 1 && (k & 1), ADC_E_BADARG, "%s: bad sizes", fn);
    ADC_CHECK_ARG(slot >= 0 && slot < ADC_WSLOTS, ADC_E_BADARG, "%s: slot out of range", fn);
    const int Ci = s0->channels + (s1 ? s1->channels : 0);
    ADC_CHECK_ARG(Co * Ci * k * k <= kWMax, ADC_E_UNSUPP, "%s: Co*Ci*k*k = %d exceeds %d", fn, Co * Ci * k * k, kWMax);
    ADC_CHECK_ARG(Co == 2 || (Ci == 2 && k == 1 && !s1 && s0->H == H && s0->W == W), ADC_E_UNSUPP,
                  "%s: Co=1 is instantiated for the 2->1 1x1 head only", fn);
    return 0;
}

static long long src_pairs(const adc_src_t* s, int N) { return s ? (long long)N * s->channels * s->H * s->W : 0; }

// planar source -> pair tensor in scratch
static int pack_source(const adc_src_t* s, int N, float2* dst, cudaStream_t st) {
    const long long per_image = (long long)s->channels * s->H * s->W;
    dim3 grid((unsigned)ceil_div_ll(per_image, 256), N);
    { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
    pack_pairs_kernel<<<grid, 256, 0, st>>>(s->x, s->c, s->batch_stride, s->batch_stride, dst, per_image);
    }
    return check_cuda(cudaGetLastError(), "pack_pairs_kernel launch");
}

size_t adc_nconv2d_fwd_scratch_floats(const adc_src_t* src0, const adc_src_t* src1, int N, int Co, int H, int W) {
    if (!src0) return 0;
    return (size_t)(2 * (src_pairs(src0, N) + src_pairs(src1, N) + (long long)N * Co * H * W));
}

int adc_nconv2d_fwd(const adc_src_t* src0, const adc_src_t* src1, const float* w, const float* b, float* y, float* cout,
                    float* scratch, int N, int Co, int H, int W, int k, int slot, adc_stream_t stream) {
    int rc = check_layer_args("adc_nconv2d_fwd", src0, src1, N, Co, H, W, k, slot);
    if (rc) return rc;
    ADC_CHECK_ARG(w && b && y && cout && scratch, ADC_E_BADARG, "adc_nconv2d_fwd: null pointer");
    if ((rc = ensure_init())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ci = src0->channels + (src1 ? src1->channels : 0);
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.call[0] = PrepCall{w, b, nullptr, Co, Ci, k, 0, 0, nullptr};
    pa.img = g_stage + (size_t)wimg_index(slot, 0, 0) * kWStride;
    { ProfScope _p(st, KIND_PREP, g_cur_call, 0.0);
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    }
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(pa.img, slot, 0, 0, 1, st))) return rc;
    float2* p0 = reinterpret_cast<float2*>(scratch);
    float2* p1 = p0 + src_pairs(src0, N);
    float2* po = p1 + src_pairs(src1, N);
    if ((rc = pack_source(src0, N, p0, st))) return rc;
    if (src1 && (rc = pack_source(src1, N, p1, st))) return rc;
    SrcP s0 = make_srcp(p0, (long long)src0->channels * src0->H * src0->W, src0->channels, src0->H, src0->W, H, W), s1;
    if (src1) s1 = make_srcp(p1, (long long)src1->channels * src1->H * src1->W, src1->channels, src1->H, src1->W, H, W);
    const int wimg = wimg_index(slot, 0, 0);
    const long long plane = (long long)H * W;
    if (Co == 1) {
        dim3 grid(blocks_1x1(plane), N);
        { ProfScope _p(st, KIND_HEAD_FWD, g_cur_call, 24.0 * grid.y * plane);
        nconv_1x1_fwd_kernel<<<grid, 256, 0, st>>>(p0, y, cout, plane, wimg);
        }
        return check_cuda(cudaGetLastError(), "nconv_1x1_fwd_kernel launch");
    }
    if ((rc = dispatch_fwd(Ci, k, s0, src1 ? &s1 : nullptr, po, N, H, W, wimg, st))) return rc;
    const long long total = (long long)N * Co * plane;
    { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
    unpack_pairs_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(po, y, cout, total);
    }
    return check_cuda(cudaGetLastError(), "unpack_pairs_kernel launch");
}

// scratch of the single-layer backward: [packed sources][G pairs][raw grads at layer res][raw low-res grads][wpart][bpart]
namespace adc {
struct LayerBwdLayout {
    long long p0, p1, G, r0, r1, l0, l1, wpart, bpart, stage, total;
    int wblocks, bblocks;
    LayerBwdLayout(const adc_src_t* s0, const adc_src_t* s1, int N, int Co, int H, int W) {
        const long long plane = (long long)H * W;
        long long o = 0;
        p0 = o; o += 2 * src_pairs(s0, N);
        p1 = o; o += 2 * src_pairs(s1, N);
        G = o; o += 2LL * N * Co * plane;
        r0 = o; o += 2LL * N * s0->channels * plane;
        r1 = o; o += s1 ? 2LL * N * s1->channels * plane : 0;
        l0 = o; o += 2 * src_pairs(s0, N);
        l1 = o; o += 2 * src_pairs(s1, N);
        wblocks = Co == 1 ? N * blocks_1x1(plane) : wg_tiles(N, H, W);
        bblocks = Co == 1 ? wblocks : N * (int)ceil_div_ll(plane, 256);
        wpart = o; o += (long long)wblocks * kWStride;
        bpart = o; o += (long long)bblocks * 4;
        stage = o; o += (long long)kRedChunks * kWStride;
        total = o;
    }
};
}  // namespace adc

size_t adc_nconv2d_bwd_scratch_floats(const adc_src_t* src0, const adc_src_t* src1, int N, int Co, int H, int W) {
    if (!src0) return 0;
    return (size_t)LayerBwdLayout(src0, src1, N, Co, H, W).total;
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
    ADC_CHECK_ARG((gw == nullptr) == (gb == nullptr), ADC_E_BADARG, "adc_nconv2d_bwd: gw/gb must be given together");
    if ((rc = ensure_init())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ci = src0->channels + (src1 ? src1->channels : 0);
    const bool need_data = gx0 != nullptr, need_weight = gw != nullptr;
    const long long plane = (long long)H * W;
    LayerBwdLayout L(src0, src1, N, Co, H, W);
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    float* img_f = g_stage + (size_t)wimg_index(slot, 0, 0) * kWStride;
    float* img_b = g_stage + (size_t)wimg_index(slot, 1, 0) * kWStride;
    pa.call[0] = PrepCall{w_fwd, b, nullptr, Co, Ci, k, 0, 0, nullptr};
    pa.img = img_f;
    { ProfScope _p(st, KIND_PREP, g_cur_call, 0.0);
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    }
    pa.call[0] = PrepCall{w_bwd, b, nullptr, Co, Ci, k, 0, 1, img_f};
    pa.img = img_b;
    { ProfScope _p(st, KIND_PREP, g_cur_call, 0.0);
    prep_weights_kernel<<<1, 128, 0, st>>>(pa);
    }
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(img_f, slot, 0, 0, 1, st))) return rc;
    if ((rc = upload_images(img_b, slot, 1, 0, 1, st))) return rc;
    const int wf = wimg_index(slot, 0, 0), wb = wimg_index(slot, 1, 0);
    float2* p0 = reinterpret_cast<float2*>(scratch + L.p0);
    float2* p1 = reinterpret_cast<float2*>(scratch + L.p1);
    float2* G = reinterpret_cast<float2*>(scratch + L.G);
    float2* r0 = reinterpret_cast<float2*>(scratch + L.r0);
    float2* r1 = reinterpret_cast<float2*>(scratch + L.r1);
    float* wpart = scratch + L.wpart;
    float* bpart = scratch + L.bpart;
    if ((rc = pack_source(src0, N, p0, st))) return rc;
    if (src1 && (rc = pack_source(src1, N, p1, st))) return rc;
    SrcP s0 = make_srcp(p0, (long long)src0->channels * src0->H * src0->W, src0->channels, src0->H, src0->W, H, W), s1;
    if (src1) s1 = make_srcp(p1, (long long)src1->channels * src1->H * src1->W, src1->channels, src1->H, src1->W, H, W);

    if (Co == 1) {
        dim3 grid(blocks_1x1(plane), N);
        { ProfScope _p(st, KIND_HEAD_BWD, g_cur_call, 48.0 * grid.y * plane);
        nconv_1x1_bwd_kernel<<<grid, 256, 0, st>>>(p0, y, cout, gy, gcout, need_data ? r0 : nullptr, wpart, bpart, nullptr, plane, wf,
                                                  wb, -1);
        }
        ADC_CUDA(cudaGetLastError());
    } else {
        dim3 grid((unsigned)ceil_div_ll(plane, 256), N);
        { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
        grad_transform_kernel<2><<<grid, 256, 0, st>>>(gy, gcout, y, cout, G, bpart, plane, wf);
        }
        ADC_CUDA(cudaGetLastError());
        SrcP g = make_srcp(G, 2 * plane, 2, H, W, H, W);
        if (need_weight && (rc = dispatch_bwd_weight(Ci, k, g, s0, src1 ? &s1 : nullptr, wpart, N, H, W, st))) return rc;
        if (need_data && (rc = dispatch_bwd_data(Ci, k, g, s0, src1 ? &s1 : nullptr, r0, r1, nullptr, N, H, W, wb, -1, st))) return rc;
    }
    if (need_data) {
        const adc_src_t* ss[2] = {src0, src1};
        const SrcP* sp[2] = {&s0, src1 ? &s1 : nullptr};
        float2* raw[2] = {r0, r1};
        float2* low[2] = {reinterpret_cast<float2*>(scratch + L.l0), reinterpret_cast<float2*>(scratch + L.l1)};
        float* ox[2] = {gx0, gx1};
        float* oc[2] = {gc0, gc1};
        for (int q = 0; q < 2; ++q) {
            if (!ss[q]) continue;
            const float2* fin = raw[q];
            if (sp[q]->up) {
                if ((rc = run_upsample_bwd(raw[q], *sp[q], low[q], nullptr, N, H, W, -1, st))) return rc;
                fin = low[q];
            }
            const long long total = src_pairs(ss[q], N);
            { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
            unpack_pairs_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(fin, ox[q], oc[q], total);
            }
            ADC_CUDA(cudaGetLastError());
        }
    }
    if (need_weight) {
        WgReduceArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.call[0] = WgReduceCall{wpart, bpart, L.wblocks, L.bblocks, 0, Co, Co * Ci * k * k};
        ra.ncalls = 1;
        ra.gw[0] = gw;
        ra.gb[0] = gb;
        ra.accumulate = 1;
        { ProfScope _p(st, KIND_REDUCE, g_cur_call, 0.0);
        wgrad_reduce1_kernel<<<dim3(1, kRedChunks), 1024, 0, st>>>(ra, scratch + L.stage);
        }
        { ProfScope _p(st, KIND_REDUCE, g_cur_call, 0.0);
        wgrad_reduce2_kernel<<<1, 128, 0, st>>>(ra, scratch + L.stage);
        }
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
    int rc = ensure_init();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Geo g(a->N, a->H, a->W);
    SavedLayout L(g);
    float* S = a->saved;

    // 1. weights: softplus^k per call, sums -> constant memory
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
    { ProfScope _p(st, KIND_PREP, g_cur_call, 0.0);
    prep_weights_kernel<<<ADC_PYR_CALLS, 128, 0, st>>>(pa);
    }
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(S, a->slot, 0, 0, ADC_PYR_CALLS, st))) return rc;

    auto act = [&](int id) { return reinterpret_cast<float2*>(S + L.off[id]); };
    auto src_of = [&](int id, int lvl) {   // activation `id` as a source of a layer at level `lvl`
        const int l = kActLevel[id], ch = kActCh[id];
        return make_srcp(act(id), (long long)ch * g.h[l] * g.w[l], ch, g.h[l], g.w[l], g.h[lvl], g.w[lvl]);
    };
    auto layer = [&](int call, int in0, int in1, int out) -> int {
        g_cur_call = call;
        const int m = kCallModule[call], l = kCallLevel[call];
        SrcP s0 = src_of(in0, l), s1;
        if (in1 >= 0) s1 = src_of(in1, l);
        return dispatch_fwd(kModCi[m], kModK[m], s0, in1 >= 0 ? &s1 : nullptr, act(out), g.N, g.h[l], g.w[l],
                            wimg_index(a->slot, 0, call), st);
    };
    long long ioff[3]; int ihs[3], iws[3];
    adc_pyramid_idx_layout(g.N, a->H, a->W, ioff, ihs, iws);
    auto pool = [&](int in_id, int out_id, int lvl) -> int {  // lvl = level of the input
        const int Ho = g.h[lvl + 1], Wo = g.w[lvl + 1];
        const long long planes = (long long)g.N * 2, total = planes * Ho * Wo;
        { ProfScope _p(st, KIND_POOL_FWD, g_cur_call, 24.0 * planes * g.h[lvl] * g.w[lvl] / 2);
        pair_pool_fwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(act(in_id), act(out_id), a->idx + ioff[lvl], g.h[lvl],
                                                                            g.w[lvl], Ho, Wo, planes);
        }
        return check_cuda(cudaGetLastError(), "pair_pool_fwd_kernel launch");
    };
#define RUN(expr) do { rc = (expr); if (rc) return rc; } while (0)
    {
        const long long per_image = (long long)a->H * a->W;
        dim3 grid((unsigned)ceil_div_ll(per_image, 256), g.N);
        { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
        pack_pairs_kernel<<<grid, 256, 0, st>>>(a->x0, a->c0, a->x0_batch_stride, a->c0_batch_stride, act(AIN), per_image);
        }
        ADC_CUDA(cudaGetLastError());
    }
    RUN(layer(0, AIN, -1, A0));
    RUN(layer(1, A0, -1, A1));
    RUN(layer(2, A1, -1, A2));
    RUN(pool(A2, P1, 0));
    RUN(layer(3, P1, -1, A3));
    RUN(layer(4, A3, -1, A4));
    RUN(pool(A4, P2, 1));
    RUN(layer(5, P2, -1, A5));
    RUN(pool(A5, P3, 2));
    RUN(layer(6, P3, -1, A6));
    RUN(layer(7, A5, A6, A7));   // nconv4: cat(x3_ds, up(x4_ds))
    RUN(layer(8, A4, A7, A8));   // nconv5: cat(x2_ds, up(x34_ds))
    RUN(layer(9, A8, A2, A9));   // nconv6: cat(up(x23_ds), x1)  -- upsampled FIRST (model_zoo/nconv_sd.py:94)
    {
        const long long plane = (long long)a->H * a->W;
        dim3 grid(blocks_1x1(plane), g.N);
        { ProfScope _p(st, KIND_HEAD_FWD, g_cur_call, 24.0 * grid.y * plane);
        nconv_1x1_fwd_kernel<<<grid, 256, 0, st>>>(act(A9), a->xout, a->cout, plane, wimg_index(a->slot, 0, 10));
        }
        ADC_CUDA(cudaGetLastError());
    }
#undef RUN
    return 0;
}

// backward scratch: [bwd images][wpart per call][bpart per call][G buffers per level][layer-res raw grads of upsampled sources]
namespace adc {
struct BwdLayout {
    long long img, wpart[ADC_PYR_CALLS], bpart[ADC_PYR_CALLS], gbuf[4][4], up[3], stage, total;
    int wblk[ADC_PYR_CALLS], bblk[ADC_PYR_CALLS];
    explicit BwdLayout(const Geo& g) {
        long long o = 0;
        img = o; o += kImgFloats;
        auto pool_blocks = [&](int l) { return (int)ceil_div_ll(2LL * g.N * ((g.h[l] + 1) / 2) * ((g.w[l] + 1) / 2), 256); };
        auto red_blocks = [&](int l) { return (int)ceil_div_ll(2 * g.px[l], 256); };
        const int b1 = g.N * blocks_1x1((long long)g.h[0] * g.w[0]);
        // who finalises the raw gradient of each call's output (see adc_pyramid_bwd)
        bblk[10] = b1; bblk[9] = b1;
        bblk[8] = red_blocks(1); bblk[7] = red_blocks(2); bblk[6] = red_blocks(3);
        bblk[5] = pool_blocks(2); bblk[4] = pool_blocks(1); bblk[2] = pool_blocks(0);
        bblk[3] = tiles(g.N, g.h[1], g.w[1]); bblk[1] = tiles(g.N, g.h[0], g.w[0]); bblk[0] = tiles(g.N, g.h[0], g.w[0]);
        for (int k = 0; k < ADC_PYR_CALLS; ++k) {
            const int l = kCallLevel[k];
            wblk[k] = k == 10 ? b1 : wg_tiles(g.N, g.h[l], g.w[l]);
            wpart[k] = o; o += (long long)wblk[k] * kWStride;
            bpart[k] = o; o += (long long)bblk[k] * 4;
        }
        for (int l = 0; l < 4; ++l)
            for (int b = 0; b < 4; ++b) { gbuf[l][b] = o; o += 4 * g.px[l]; }   // pair tensor [N,2,h,w]
        for (int l = 0; l < 3; ++l) { up[l] = o; o += 4 * g.px[l]; }
        stage = o; o += (long long)ADC_PYR_CALLS * kRedChunks * kWStride;
        total = o;
    }
};
}  // namespace adc

size_t adc_pyramid_bwd_scratch_floats(int N, int H, int W) {
    Geo g(N, H, W);
    return (size_t)BwdLayout(g).total;
}

int adc_pyramid_bwd(const adc_pyramid_bwd_t* a, adc_stream_t stream) {
    ADC_CHECK_ARG(a && a->saved && a->idx && a->xout && a->cout && a->g_xout && a->g_cout && a->scratch, ADC_E_BADARG,
                  "adc_pyramid_bwd: null pointer");
    ADC_CHECK_ARG(a->N > 0 && a->H >= 8 && a->W >= 8, ADC_E_BADARG, "adc_pyramid_bwd: bad sizes");
    ADC_CHECK_ARG(a->slot >= 0 && a->slot < ADC_WSLOTS, ADC_E_BADARG, "adc_pyramid_bwd: slot out of range");
    ADC_CHECK_ARG(!a->need_input_grad || (a->gx0 && a->gc0), ADC_E_BADARG, "adc_pyramid_bwd: gx0/gc0 missing");
    ADC_CHECK_ARG(!a->need_weight_grad || (a->gweight && a->gbias), ADC_E_BADARG, "adc_pyramid_bwd: gweight/gbias missing");
    for (int m = 0; m < ADC_PYR_MODULES; ++m)
        ADC_CHECK_ARG(a->weight_bwd[m] && a->bias[m], ADC_E_BADARG, "adc_pyramid_bwd: missing parameter %d", m);
    int rc = ensure_init();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Geo g(a->N, a->H, a->W);
    SavedLayout L(g);
    BwdLayout B(g);
    const float* S = a->saved;
    float* X = a->scratch;
    const bool need_w = a->need_weight_grad != 0;

    // forward images (b, s, 1/s of each call) back into constant memory; backward-time weights, flipped, next to them
    if ((rc = upload_images(S, a->slot, 0, 0, ADC_PYR_CALLS, st))) return rc;
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    for (int k = 0; k < ADC_PYR_CALLS; ++k) {
        const int m = kCallModule[k];
        pa.call[k] = PrepCall{a->weight_bwd[m], a->bias[m], nullptr, kModCo[m], kModCi[m], kModK[m], 0, 1, S + (size_t)k * kWStride};
    }
    pa.img = X + B.img;
    { ProfScope _p(st, KIND_PREP, g_cur_call, 0.0);
    prep_weights_kernel<<<ADC_PYR_CALLS, 128, 0, st>>>(pa);
    }
    ADC_CUDA(cudaGetLastError());
    if ((rc = upload_images(X + B.img, a->slot, 1, 0, ADC_PYR_CALLS, st))) return rc;

    auto act = [&](int id) { return reinterpret_cast<const float2*>(S + L.off[id]); };
    auto src_of = [&](int id, int lvl) {
        const int l = kActLevel[id], ch = kActCh[id];
        return make_srcp(act(id), (long long)ch * g.h[l] * g.w[l], ch, g.h[l], g.w[l], g.h[lvl], g.w[lvl]);
    };
    auto gb = [&](int l, int b) { return reinterpret_cast<float2*>(X + B.gbuf[l][b]); };
    auto upb = [&](int l) { return reinterpret_cast<float2*>(X + B.up[l]); };
    auto fimg = [&](int call) { return wimg_index(a->slot, 0, call); };
    auto bimg = [&](int call) { return wimg_index(a->slot, 1, call); };
    // profiling runs keep everything on one stream so that the per-kernel event pairs time each kernel alone
    const bool side = need_w && g_side_on && g_side_user && g_side != nullptr && !g_prof_on;
    int seq = 0, last_join = -1;
    // one layer: G (transformed gradient of the call's output) -> weight partials + gradients of its sources
    auto layer_bwd = [&](int call, const float2* G, int in0, int in1, float2* out0, float2* out1, int xcall, bool need_data) -> int {
        g_cur_call = call;
        const int m = kCallModule[call], l = kCallLevel[call];
        SrcP gsrc = make_srcp(G, 2LL * g.h[l] * g.w[l], 2, g.h[l], g.w[l], g.h[l], g.w[l]);
        SrcP s0 = src_of(in0, l), s1;
        if (in1 >= 0) s1 = src_of(in1, l);
        int r = 0;
        if (need_w) {
            // The weight gradient only feeds the final reduction: with `side` it runs on the second stream beside the data-gradient
            // chain (fork after the producer of G, join before the reduction).  G buffers are recycled two layers later at the
            // earliest (gbuf map below), and the side stream is in order, so the main stream waits for the weight gradient issued two
            // layers earlier before it launches a kernel that may overwrite that layer's G.
            cudaStream_t ws = st;
            if (side) {
                ADC_CUDA(cudaEventRecord(g_ev_fork[seq], st));
                ADC_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork[seq], 0));
                ws = g_side;
            }
            if ((r = dispatch_bwd_weight(kModCi[m], kModK[m], gsrc, s0, in1 >= 0 ? &s1 : nullptr, X + B.wpart[call], g.N, g.h[l], g.w[l], ws)))
                return r;
            if (side) {
                ADC_CUDA(cudaEventRecord(g_ev_join[seq], g_side));
                if (seq >= 2) ADC_CUDA(cudaStreamWaitEvent(st, g_ev_join[seq - 2], 0));
                last_join = seq;
            }
            ++seq;
        }
        if (need_data)
            r = dispatch_bwd_data(kModCi[m], kModK[m], gsrc, s0, in1 >= 0 ? &s1 : nullptr, out0, out1,
                                  xcall >= 0 ? X + B.bpart[xcall] : nullptr, g.N, g.h[l], g.w[l], bimg(call),
                                  xcall >= 0 ? fimg(xcall) : -1, st);
        return r;
    };
    auto reduce_up = [&](int lvl, int act_id, int xcall, float2* G) -> int {   // raw grads in up[lvl] -> G of call `xcall`
        SrcP s = src_of(act_id, lvl);
        return run_upsample_bwd(upb(lvl), s, G, X + B.bpart[xcall], g.N, g.h[lvl], g.w[lvl], fimg(xcall), st);
    };
    long long ioff[3]; int ihs[3], iws[3];
    adc_pyramid_idx_layout(g.N, a->H, a->W, ioff, ihs, iws);
    auto pool_bwd = [&](int lvl, const float2* g_ds, const float2* skip, int act_id, int xcall, float2* G) -> int {
        const int H = g.h[lvl], W = g.w[lvl];
        const long long planes = 2LL * g.N, total = planes * ((H + 1) / 2) * ((W + 1) / 2);
        { ProfScope _p(st, KIND_POOL_BWD, g_cur_call, 28.0 * planes * H * W);
        pair_pool_bwd_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(g_ds, a->idx + ioff[lvl], skip, act(act_id), G,
                                                                            X + B.bpart[xcall], 2, H, W, g.h[lvl + 1], g.w[lvl + 1],
                                                                            planes, fimg(xcall));
        }
        return check_cuda(cudaGetLastError(), "pair_pool_bwd_kernel launch");
    };
#define RUN(expr) do { rc = (expr); if (rc) return rc; } while (0)
    // call 10 (nconv7 1x1 head): raw autograd gradients -> G of call 9 in gbuf[0][0]
    {
        const long long plane = (long long)a->H * a->W;
        dim3 grid(blocks_1x1(plane), g.N);
        { ProfScope _p(st, KIND_HEAD_BWD, g_cur_call, 48.0 * grid.y * plane);
        nconv_1x1_bwd_kernel<<<grid, 256, 0, st>>>(act(A9), a->xout, a->cout, a->g_xout, a->g_cout, gb(0, 0), X + B.wpart[10],
                                                  X + B.bpart[10], X + B.bpart[9], plane, fimg(10), bimg(10), fimg(9));
        }
        ADC_CUDA(cudaGetLastError());
    }
    // call 9 (nconv6 on cat(up(A8), A2)): raw grads of up(A8) -> up[0] -> reduce -> G8 in gbuf[1][0]; skip grads of A2 -> gbuf[0][2]
    RUN(layer_bwd(9, gb(0, 0), A8, A2, upb(0), gb(0, 2), -1, true));
    RUN(reduce_up(0, A8, 8, gb(1, 0)));
    // call 8 (nconv5 on cat(A4, up(A7))): skip grads of A4 -> gbuf[1][2]; up(A7) -> up[1] -> G7 in gbuf[2][0]
    RUN(layer_bwd(8, gb(1, 0), A4, A7, gb(1, 2), upb(1), -1, true));
    RUN(reduce_up(1, A7, 7, gb(2, 0)));
    // call 7 (nconv4 on cat(A5, up(A6))): skip grads of A5 -> gbuf[2][2]; up(A6) -> up[2] -> G6 in gbuf[3][0]
    RUN(layer_bwd(7, gb(2, 0), A5, A6, gb(2, 2), upb(2), -1, true));
    RUN(reduce_up(2, A6, 6, gb(3, 0)));
    // call 6 (nconv2 @L3 on P3): raw grads of P3 -> gbuf[3][1]; pool3 backward (+ skip of A5) -> G5 in gbuf[2][1]
    RUN(layer_bwd(6, gb(3, 0), P3, -1, gb(3, 1), nullptr, -1, true));
    RUN(pool_bwd(2, gb(3, 1), gb(2, 2), A5, 5, gb(2, 1)));
    // call 5 (nconv2 @L2 on P2): raw grads of P2 -> gbuf[2][3]; pool2 backward (+ skip of A4) -> G4 in gbuf[1][1]
    RUN(layer_bwd(5, gb(2, 1), P2, -1, gb(2, 3), nullptr, -1, true));
    RUN(pool_bwd(1, gb(2, 3), gb(1, 2), A4, 4, gb(1, 1)));
    // call 4 (nconv3 @L1 on A3): G3 in gbuf[1][3];  call 3 (nconv2 @L1 on P1): raw grads of P1 -> gbuf[1][0]
    RUN(layer_bwd(4, gb(1, 1), A3, -1, gb(1, 3), nullptr, 3, true));
    RUN(layer_bwd(3, gb(1, 3), P1, -1, gb(1, 0), nullptr, -1, true));
    // pool1 backward (+ skip of A2) -> G2 in gbuf[0][1]
    RUN(pool_bwd(0, gb(1, 0), gb(0, 2), A2, 2, gb(0, 1)));
    // call 2 (nconv3 @L0 on A1): G1 in gbuf[0][0];  call 1 (nconv2 @L0 on A0): G0 in gbuf[0][3];  call 0 (nconv1 on the input)
    RUN(layer_bwd(2, gb(0, 1), A1, -1, gb(0, 0), nullptr, 1, true));
    RUN(layer_bwd(1, gb(0, 0), A0, -1, gb(0, 3), nullptr, 0, true));
    {
        const bool nd = a->need_input_grad != 0;
        if (nd || need_w) RUN(layer_bwd(0, gb(0, 3), AIN, -1, gb(0, 1), nullptr, -1, nd));
        if (nd) {
            const long long total = g.px[0];
            { ProfScope _p(st, KIND_PACK, g_cur_call, 0.0);
            unpack_pairs_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, st>>>(gb(0, 1), a->gx0, a->gc0, total);
            }
            ADC_CUDA(cudaGetLastError());
        }
    }
#undef RUN
    if (side && last_join >= 0) ADC_CUDA(cudaStreamWaitEvent(st, g_ev_join[last_join], 0));
    if (need_w) {
        WgReduceArgs ra;
        memset(&ra, 0, sizeof(ra));
        for (int k = 0; k < ADC_PYR_CALLS; ++k) {
            const int m = kCallModule[k];
            ra.call[k] = WgReduceCall{X + B.wpart[k], X + B.bpart[k], B.wblk[k], B.bblk[k], m, kModCo[m], mod_numel(m)};
        }
        ra.ncalls = ADC_PYR_CALLS;
        long long bo = 0;
        for (int m = 0; m < ADC_PYR_MODULES; ++m) {
            ra.gw[m] = a->gweight + mod_offset(m);
            ra.gb[m] = a->gbias + bo;
            bo += kModCo[m];
        }
        ra.accumulate = 0;
        { ProfScope _p(st, KIND_REDUCE, g_cur_call, 0.0);
        wgrad_reduce1_kernel<<<dim3(ADC_PYR_CALLS, kRedChunks), 1024, 0, st>>>(ra, X + B.stage);
        }
        { ProfScope _p(st, KIND_REDUCE, g_cur_call, 0.0);
        wgrad_reduce2_kernel<<<ADC_PYR_MODULES, 128, 0, st>>>(ra, X + B.stage);
        }
        ADC_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"

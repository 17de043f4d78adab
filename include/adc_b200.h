/*
 * adc_b200.h -- C ABI of the B200-native (sm_100a) hot path of aerial-depth-completion.
 *
 * Plain pointers and sizes only; no torch types.  All tensor pointers are DEVICE pointers to fp32
 * NCHW data (channel planes contiguous: element (n,ch,h,w) lives at n*batch_stride + ch*H*W + h*W + w;
 * batch_stride is C*H*W unless stated), indices are int64.  Every call only ENQUEUES work on `stream`
 * (a cudaStream_t) and never synchronises, so calls are CUDA-graph capturable.
 *
 * Return value: 0 on success; a positive cudaError_t value if the CUDA runtime failed; a negative
 * ADC_E_* for argument errors.  adc_last_error() returns a thread-local message for the last failure.
 * There is no CPU fallback: without a CUDA device the calls fail with the CUDA error.
 *
 * Each entry point cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef ADC_B200_H
#define ADC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* adc_stream_t; /* cudaStream_t */

#define ADC_E_BADARG   (-1) /* null pointer / non-positive size                                   */
#define ADC_E_UNSUPP   (-2) /* (Ci,Co,k) or layer shape outside the instantiated set              */
#define ADC_E_ALIGN    (-3) /* pointer not 4-byte aligned                                         */

#define ADC_PYR_CALLS   11  /* NConv2d calls in one CNN.forward (model_zoo/nconv_sd.py:38-97)     */
#define ADC_PYR_MODULES 7   /* nconv1..nconv7 (model_zoo/nconv_sd.py:21-32)                       */
#define ADC_WSLOTS      4   /* independent constant-memory weight slots (one per CNN instance)    */

int adc_abi_version(void);
const char* adc_last_error(void);

/* Instrumentation (bench.py): number of kernels this library has launched so far, and optional per-kernel timing with
 * CUDA events recorded on the launching stream around every kernel (kind: 0 nconv fwd, 1 nconv bwd data, 2 nconv bwd
 * weight, 3 pool fwd, 4 pool bwd, 5 upsample bwd, 6 head fwd, 7 head bwd, 8 pack/unpack, 9 reduce, 10 prep; call = pyramid
 * call index; bytes = the kernel's algorithmic HBM bytes).  adc_profile_read synchronises on the recorded events. */
long long adc_launch_count(void);
int adc_profile_enable(int on);
/* adc_pyramid_bwd issues its weight-gradient kernels on a second stream beside the data-gradient chain (fork / join inside the call,
 * CUDA-graph capturable).  0 keeps every kernel on the caller's stream (what the parity tests compare the two-stream schedule with);
 * returns the previous setting. */
int adc_set_side_stream(int on);
int adc_profile_read(int max_records, int* n_out, int* kinds, int* calls, float* ms, double* bytes);

/* ---- a source of input channels for one NConv2d call ------------------------------------------
 * `channels` planes of data `x` and confidence `c` of size H x W.  If (H,W) differs from the layer's
 * output size the source is read through ATen's nearest-neighbour index map, i.e. it stands for
 * F.interpolate(src, out_size, mode='nearest') followed by torch.cat -- neither tensor is ever
 * materialised (model_zoo/nconv_sd.py:81-94). */
typedef struct {
    const float* x;
    const float* c;
    long long batch_stride; /* in floats */
    int channels, H, W;
} adc_src_t;

/* NConv2d.forward (model_zoo/nconv.py:37-67): y = conv(x*c,w)/(conv(c,w)+1e-20)+b, cout = conv(c,w)/sum(w).
 * stride 1, dilation 1, groups 1, padding k/2 -- the only configuration the nets use
 * (model_zoo/nconv_sd.py:21-32).  w is [Co,Ci,k,k], b is [Co], Ci = src0.channels (+ src1.channels).
 * src1 may be NULL.  Channel order of the virtual concatenation is src0 then src1.
 * `scratch` must hold adc_nconv2d_fwd_scratch_floats() floats (the kernels work on interleaved {x,c} pair tensors;
 * this single-layer entry packs/unpacks the planar tensors, the pyramid entry below stays packed throughout).
 * slot selects the constant-memory weight slot (0..ADC_WSLOTS-1); calls that may run concurrently on
 * different streams must use different slots. */
size_t adc_nconv2d_fwd_scratch_floats(const adc_src_t* src0, const adc_src_t* src1, int N, int Co, int H, int W);
int adc_nconv2d_fwd(const adc_src_t* src0, const adc_src_t* src1, const float* w, const float* b,
                    float* y, float* cout, float* scratch, int N, int Co, int H, int W, int k, int slot,
                    adc_stream_t stream);

/* Backward of the above as the reference's autograd graph computes it:
 * w_fwd = the weight the forward call used (fixes s = sum(w)); w_bwd = the Parameter's contents at
 * backward time (used by the conv data-gradients; model_zoo/nconv.py:114-117 swaps weight.data).
 * gx0/gc0 (and gx1/gc1 for src1) receive the input gradients -- for an upsampled source they are the
 * gradients w.r.t. the LOW-resolution tensors.  They may be NULL (no input gradient needed; all or none).
 * gw [Co,Ci,k,k] and gb [Co] are ACCUMULATED into (+=) unless NULL (both or none).  `scratch` must hold
 * adc_nconv2d_bwd_scratch_floats() floats. */
size_t adc_nconv2d_bwd_scratch_floats(const adc_src_t* src0, const adc_src_t* src1, int N, int Co, int H, int W);
int adc_nconv2d_bwd(const adc_src_t* src0, const adc_src_t* src1, const float* y, const float* cout,
                    const float* gy, const float* gcout, const float* w_fwd, const float* w_bwd, const float* b,
                    float* gx0, float* gc0, float* gx1, float* gc1, float* gw, float* gb, float* scratch,
                    int N, int Co, int H, int W, int k, int slot, adc_stream_t stream);

/* Confidence-argmax pooling (model_zoo/nconv_sd.py:43-49): c_ds,idx = max_pool2d(c,2,2,return_indices);
 * x_ds = x.view(-1)[idx]; c_ds /= 4.  idx is int64 ih*W+iw per (n,ch) plane, bit-exact with ATen (ties ->
 * first in row-major window order, NaN propagates, odd H/W floor).  Outputs are [N,C,H/2,W/2]. */
int adc_conf_pool_fwd(const float* x, const float* c, float* x_ds, float* c_ds, long long* idx,
                      int N, int C, int H, int W, adc_stream_t stream);
/* Backward: gx = scatter(g_xds) (+ add_x), gc = scatter(g_cds/4) (+ add_c).  add_* may be NULL; they are the
 * gradients arriving over the skip connection (same [N,C,H,W] shape) so the sum needs no extra pass. */
int adc_conf_pool_bwd(const float* g_xds, const float* g_cds, const long long* idx, const float* add_x,
                      const float* add_c, float* gx, float* gc, int N, int C, int H, int W, adc_stream_t stream);

/* softplus^n(w, beta=10, threshold=20) element-wise (EnforcePos, model_zoo/nconv.py:114-131). */
int adc_softplus_pow(const float* w, float* out, int numel, int times, adc_stream_t stream);

/* ---- the whole unguided pyramid (CNN.forward, model_zoo/nconv_sd.py:35-101) in one call ---------- */
typedef struct {
    int N, H, W;                 /* level-0 size; levels are floor-halved three times                */
    int training;                /* 1: every call first applies softplus to its module's weight      */
    int slot;                    /* constant-memory weight slot                                      */
    const float* x0;             /* [N,1,H,W] sparse depth                                           */
    const float* c0;             /* [N,1,H,W] input confidence                                       */
    long long x0_batch_stride, c0_batch_stride;
    const float* weight[ADC_PYR_MODULES]; /* nconv1..7 .weight as stored in the Parameters           */
    const float* bias[ADC_PYR_MODULES];
    float* weight_out;           /* training: adc_pyramid_weight_floats() floats, the weights the      *
                                  * Parameters hold after the forward (softplus^calls), module-major  */
    float* saved;                /* adc_pyramid_saved_floats(N,H,W) floats: activations kept for bwd  */
    long long* idx;              /* adc_pyramid_idx_count(N,H,W) int64: the three pooling index maps  */
    float* xout;                 /* [N,1,H,W]                                                        */
    float* cout;                 /* [N,1,H,W]                                                        */
} adc_pyramid_fwd_t;

size_t adc_pyramid_weight_floats(void);
size_t adc_pyramid_saved_floats(int N, int H, int W);
size_t adc_pyramid_idx_count(int N, int H, int W);
/* offsets (in elements) of the three index maps inside `idx`, and their [N,2,h,w] sizes */
int adc_pyramid_idx_layout(int N, int H, int W, long long offsets[3], int hs[3], int ws[3]);
int adc_pyramid_fwd(const adc_pyramid_fwd_t* a, adc_stream_t stream);

typedef struct {
    int N, H, W;
    int slot;
    int need_input_grad;         /* 0: x0/c0 need no gradient (first layer dgrad skipped)            */
    int need_weight_grad;        /* 0: frozen net, only data gradients                               */
    const float* x0;
    const float* c0;
    long long x0_batch_stride, c0_batch_stride;
    const float* weight_bwd[ADC_PYR_MODULES]; /* Parameter contents at backward time                 */
    const float* bias[ADC_PYR_MODULES];
    const float* saved;          /* from the forward                                                 */
    const long long* idx;
    const float* xout;           /* forward outputs                                                  */
    const float* cout;
    const float* g_xout;         /* [N,1,H,W] incoming gradients                                     */
    const float* g_cout;
    float* scratch;              /* adc_pyramid_bwd_scratch_floats(N,H,W) floats                      */
    float* gx0;                  /* [N,1,H,W] or NULL                                                */
    float* gc0;
    float* gweight;              /* adc_pyramid_weight_floats() floats, module-major, OVERWRITTEN     */
    float* gbias;                /* 13 floats (2,2,2,2,2,2,1), OVERWRITTEN                            */
} adc_pyramid_bwd_t;

size_t adc_pyramid_bwd_scratch_floats(int N, int H, int W);
int adc_pyramid_bwd(const adc_pyramid_bwd_t* a, adc_stream_t stream);

/* ---- dense 3x3 convolution on the tensor cores (tcgen05 / TMEM / TMA implicit GEMM) -------------------------------
 * Replaces the nn.Conv2d 3x3 layers of the wide encoder-decoders (model_zoo/confidence_depth_framework.py:442-452,
 * 536-607; F.conv2d + bias + ReLU/LeakyReLU) in the bf16 tensor-core mode (parity 1e-2, BASELINE.json north_star).
 * Activations are NHWC bf16 with a channel count that is a multiple of 16; a layer may read / write a channel slice of a
 * wider tensor (x_off / out_off, strides in channels per pixel), which is how torch.cat is folded away. */
/* pad = 1: the NHWC tensor carries a one-pixel zero border, [N][H+2][W+2][channels] (the layout of the v2 kernel) */
int adc_nchw_to_nhwc_bf16(const float* src, void* dst, void* dst_lo /* NULL or the low halves */, int N, int C, int Cpad,
                          int H, int W, long long src_batch_stride, int dst_stride, int dst_off, int pad, adc_stream_t stream);
int adc_nhwc_to_nchw_f32(const void* src, int src_is_f32, float* dst, int N, int C, int H, int W, int src_stride, int src_off,
                         int pad, adc_stream_t stream);
/* weights fp32 [Cout,Cin,3,3] -> bf16 [9][round16(Cout)][cin_pad] (+ low halves), input channels zero padded to cin_pad
 * (a multiple of 16, >= Cin); flip_transpose=1 = data-gradient form (transposed, taps flipped), 2 = transposed only */
size_t adc_conv3x3_weight_elems(int cout, int cin_pad);
int adc_conv3x3_prep_weights(const float* w, void* out_bf16, void* out_lo, int cout, int cin, int cin_pad, int flip_transpose,
                             adc_stream_t stream);
/* Batched form: up to ADC_WPREP_MAX conversions in one launch.  Item k converts input channels [ci_off, ci_off+ci_cnt) of the
 * fp32 tensor w [cout][cin][3][3]; mode 0: out [9][round16(cout)][cin_pad] (forward weights); mode 1 / 2: transposed
 * [9][round16(ci_cnt)][cin_pad] with flipped / unflipped taps (data-gradient weights, contraction over cout <= cin_pad). */
#define ADC_WPREP_MAX 40
typedef struct {
    const float* w;
    void* out;
    void* out_lo; /* NULL: plain bf16 */
    int cout, cin, ci_off, ci_cnt, cin_pad, mode;
} adc_wprep_item_t;
int adc_conv3x3_prep_weights_batch(const adc_wprep_item_t* items, int n, adc_stream_t stream);

/* act: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 sigmoid.  stride 1 or 2, padding 1.  x_lo/w_lo non-NULL selects the bf16x3 split
 * mode: operands are (hi, lo) bf16 pairs, hi*hi + lo*hi + hi*lo accumulated in fp32 (fp32-grade parity, 3 MMAs per product). */
int adc_conv3x3_fprop_tc(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo,
                         const float* bias, void* out, void* out_lo, int out_f32, int out_stride, int out_off, int N, int H,
                         int W, int cin, int cout, int stride, int act, int x_pad, int out_pad, adc_stream_t stream);
/* Stride-1 variant with shared-memory tap reuse, persistent CTAs and double-buffered TMEM accumulators; x and out are
 * zero-bordered NHWC tensors, cin % 64 == 0, cout <= 128; channels [cout, out_zero_to) of the output slice are zeroed. */
int adc_conv3x3_s1_tc_v2(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo,
                         const float* bias, void* out, void* out_lo, int out_f32, int out_stride, int out_off, int out_zero_to,
                         int N, int H, int W, int cin, int cout, int act, int tap_mask, adc_stream_t stream);
/* tap_mask of adc_conv3x3_s1_tc_v2: bit 3u+v set = tap (u,v) may be non-zero (0 = all nine); rows of clear taps are never loaded.
 * 1x1 convolution (plain GEMM over the flat pixels) on the same kernel: only the centre tap of the prepared weights is read. */
int adc_conv1x1_tc_v2(const void* x, const void* x_lo, int x_stride, int x_off, const void* w, const void* w_lo, const float* bias, void* out,
                      void* out_lo, int out_f32, int out_stride, int out_off, int out_zero_to, int N, int H, int W, int cin, int cout, int act,
                      adc_stream_t stream);
/* Data gradient of a 3x3 / stride-2 / pad-1 convolution (backward of the encoder layers conv1-5 of GEDNet,
 * confidence_depth_framework.py:442-446): four parity-class launches of the tcgen05 kernel.  gy: NHWC bf16 [N][Ho][Wo]
 * (+ zero border if gy_pad), channels [gy_off, gy_off+cgy), cgy % 16 == 0; w from adc_conv3x3_prep_weights(flip_transpose=2);
 * gx: channels [gx_off, gx_off+cgx) of NHWC [N][H][W] (+ border if gx_pad; the border is not written), bf16/pair or fp32. */
int adc_conv3x3_dgrad_s2_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* w, const void* w_lo, void* gx,
                            void* gx_lo, int gx_f32, int gx_stride, int gx_off, int N, int H, int W, int cgy, int cgx, int gy_pad,
                            int gx_pad, adc_stream_t stream);
/* Weight gradient of a 3x3 / stride-1 / pad-1 convolution (the autograd backward of nn.Conv2d as used in
 * confidence_depth_framework.py:442-452,536-607): dw[co][ci][u][v] = sum over pixels of gy[.][co] * x[. + tap shift][ci].
 * gy and x are zero-bordered NHWC bf16 tensors of the same spatial size (gy's border pixels must be zeros); x channels
 * [x_off, x_off+cin_pad) are read (cin_pad % 16 == 0, <= 160), the first cin of them written to dw (fp32 [cout][cin][3][3]);
 * cout <= 128.  A stride-2 layer passes gy zero-dilated to the input resolution.  gy_lo/x_lo non-NULL = bf16x3 split mode.
 * workspace: adc_conv3x3_wgrad_tc_workspace_bytes(cin_pad, cout) bytes; the reduction order is fixed (deterministic). */
size_t adc_conv3x3_wgrad_tc_workspace_bytes(int cin_pad, int cout);
int adc_conv3x3_wgrad_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* x, const void* x_lo, int x_stride,
                         int x_off, float* dw, void* workspace, int N, int H, int W, int cin, int cin_pad, int cout,
                         adc_stream_t stream);
/* Weight gradient of a 1x1 convolution: dw fp32 [cout][cin] = sum over pixels of gy[p][co] * x[p][ci]; same workspace query. */
int adc_conv1x1_wgrad_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, const void* x, const void* x_lo, int x_stride,
                         int x_off, float* dw, void* workspace, int N, int H, int W, int cin, int cin_pad, int cout,
                         adc_stream_t stream);
/* General-geometry variant (stride 1 or 2; gy [N][Ho][Wo], x [N][H][W], each with or without the zero border): one tap per
 * CTA, 4-D TMA boxes with element strides.  cin_pad <= 256.  Same workspace query. */
int adc_conv3x3_wgrad_box_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, int gy_pad, const void* x, const void* x_lo,
                             int x_stride, int x_off, int x_pad, float* dw, void* workspace, int N, int H, int W, int stride, int cin,
                             int cin_pad, int cout, adc_stream_t stream);
/* dw fp32 [cout][cin] of a 1x1 convolution of stride 1 or 2 (no padding): the centre tap of the kernel above, every SM on a share of
 * the pixels (ResNet down-sampling shortcut).  Same tensors, workspace and determinism rules. */
int adc_conv1x1_wgrad_box_tc(const void* gy, const void* gy_lo, int gy_stride, int gy_off, int gy_pad, const void* x, const void* x_lo,
                             int x_stride, int x_off, int x_pad, float* dw, void* workspace, int N, int H, int W, int stride, int cin,
                             int cin_pad, int cout, adc_stream_t stream);

/* ---- NHWC glue of the tensor-core training path (zero-bordered NHWC bf16 tensors, optional low halves) --------------------
 * Channel slices [off, off+C) with C % 8 == 0 and 16-byte alignment.
 * adc_nhwc_upsample_nearest: F.interpolate(mode='nearest') of a slice into a slice of a larger tensor (the decoder's
 *   up-sample + torch.cat, confidence_depth_framework.py:488-504).
 * adc_nhwc_grad_assemble: out = d(act) * ( sum over the up-sampling children of A (at Hd x Wd, identity if equal) + B ),
 *   d(act) = 1 where act > 0 else slope: the backward of up-sample / cat / ReLU / LeakyReLU in one pass; a, b, act optional.
 *   a_quad: `a` is the quad form of a stride-2 data gradient instead: [N][ceil(H/2)+2][ceil(W/2)+2][4*C], pixel (y,x) reads pixel
 *   (y/2, x/2), channel block (2*(y&1) + (x&1)) * C (Hd, Wd = its size).
 *   colsum_part (optional): [adc_nhwc_grad_assemble_parts()][C] per-block sums of the output over its pixels; their sum over
 *   dim 0 in index order is the bias gradient of the layer whose activation is `act` (saves adc_nhwc_colsum's extra read).
 * adc_nhwc_colsum: per-channel sum over all pixels (bias gradient), deterministic two-stage reduction. */
/* out[c] = sum over r of part[r][c], fixed order: second stage of the deterministic two-stage reductions (partial rows written by
 * adc_nhwc_grad_assemble, adc_quad_head_*). */
int adc_sum_rows(const float* part, int rows, int C, float* out, adc_stream_t stream);
int adc_nhwc_upsample_nearest(const void* src, const void* src_lo, int src_stride, int src_off, void* dst, void* dst_lo, int dst_stride,
                              int dst_off, int N, int C, int Hs, int Ws, int Hd, int Wd, adc_stream_t stream);
int adc_nhwc_grad_assemble(const void* a, const void* a_lo, int a_stride, int a_off, int Hd, int Wd, const void* b, const void* b_lo,
                           int b_stride, int b_off, const void* act, int act_stride, int act_off, float slope, void* out, void* out_lo,
                           int out_stride, int out_off, int N, int C, int H, int W, float* colsum_part, int a_quad, adc_stream_t stream);
int adc_nhwc_grad_assemble_parts(void);
size_t adc_nhwc_colsum_workspace_bytes(int C);
int adc_nhwc_colsum(const void* g, const void* g_lo, int g_stride, int g_off, int N, int C, int H, int W, float* out, void* workspace,
                    adc_stream_t stream);

/* ---- thin last layer of GEDNet (conv10: cat(up(x9) 32 ch, x0_d) -> oc in {1,2}, LeakyReLU 0.2; confidence_depth_framework.py:
 * 504-506) as direct kernels: the nearest up-sampling is folded into the addressing of x9 (zero-bordered NHWC bf16 pair at its own
 * resolution H1 x W1), x0_d / output / their gradients stay NCHW fp32.  The backward writes G9 = leaky'(x9) * (data gradient
 * summed over the up-sampling children [+ the feature gradient glli]) for the tensor-core kernels of conv9, optionally the
 * gradient of x0_d, and dw [oc][33][3][3], db [oc] (deterministic two-stage reduction). */
int adc_ged_head_fwd(const void* x9, const void* x9_lo, int x9_stride, int H1, int W1, const float* x0d, const float* w, const float* bias,
                     float* out, int N, int oc, int H, int W, adc_stream_t stream);
size_t adc_ged_head_bwd_workspace_bytes(int oc);
int adc_ged_head_bwd(const void* x9, const void* x9_lo, int x9_stride, int H1, int W1, const float* x0d, const float* w, const float* gy,
                     const float* x10, const float* glli, int glli_ch, void* g9, void* g9_lo, int g9_stride, float* gx0d, float* dw,
                     float* db, void* workspace, int N, int oc, int H, int W, adc_stream_t stream);
int adc_ged_head_bwd_x0d(const float* gy, const float* x10, const float* glli, int glli_ch, const float* w, float* gx0d, int N, int oc,
                         int H, int W, adc_stream_t stream);

/* ---- "quad" (pixel-shuffle) form of GEDNet's two full-resolution thin layers (confidence_depth_framework.py:442,452,469-506).
 * With the exact 2x nearest up-sampling of x9, conv10(up(x9)) at pixel (2i+py, 2j+px) only touches a 2x2 block of x9: it is a
 * stride-1 3x3 convolution of x9 at ITS resolution with 4 parity channels per output channel (weights = taps of conv10 that fall
 * on the same x9 pixel, summed) and runs on adc_conv3x3_s1_tc_v2 / adc_conv3x3_wgrad_tc on a quarter of the pixels.  Read
 * backwards, the same identity turns the data gradient of conv1 (3x3, stride 2) into a stride-1 convolution of its output
 * gradient with 4 parity channels per input channel.  These entry points are the elementwise remainder:
 * adc_ged_pack_input:     (xout_d [N][1][H][W], rgb [N][3][H][W] with batch stride, cout_d) -> zero-bordered NHWC bf16 (hi, lo)
 *                         channels {xout_d, r, g, b, cout_d, 0 x 11} of a tensor with dst_stride >= 16 channels per pixel.
 * adc_quad_head_finish:   out[n][o][2i+py][2j+px] = LeakyReLU_0.2( q[n][i+1][j+1][(2py+px)*oc + o] + bias[o]
 *                         + sum_uv w[o][32][u][v] * x0d[n][2i+py+u-1][2j+px+v-1] ); q fp32 zero-bordered NHWC, w [oc][33][3][3].
 * adc_quad_head_bwd_prep: gpre = gy * LeakyReLU'(x10) -> gq (fp32 zero-bordered [N][H/2+2][W/2+2][4*oc], channel (2py+px)*oc + o) and
 *                         part[adc_quad_head_bwd_prep_parts()][oc][10]: [0,9) weight gradient of the x0_d channel, [9] bias
 *                         gradient (sum over dim 0 in index order = deterministic).
 * adc_quad_unshuffle:     out[k][n][2i+py][2j+px] = q[n][i+1][j+1][(2py+px)*K + k], K in {1,2}; any H, W (H1 = ceil(H/2)). */
int adc_ged_pack_input(const float* xout, const float* cout, const float* rgb, long long rgb_batch_stride, void* dst, void* dst_lo,
                       int dst_stride, int N, int H, int W, adc_stream_t stream);
/* conv1 (3x3, stride 2) as a GEMM: im2col at the output resolution, x45[n][i+1][j+1][(3u+v)*5 + c] = in[n][c][2i+u-1][2j+v-1],
 * channels [0,48) of a zero-bordered NHWC bf16 (hi, lo) tensor [N][H1+2][W1+2][dst_stride], H1 = ceil(H/2). */
int adc_ged_pack_im2col(const float* xout, const float* cout, const float* rgb, long long rgb_batch_stride, void* dst, void* dst_lo,
                        int dst_stride, int N, int H, int W, adc_stream_t stream);
int adc_quad_head_finish(const float* q, int q_stride, const float* x0d, const float* w, const float* bias, float* out, int N, int oc, int H,
                         int W, adc_stream_t stream);
int adc_quad_head_bwd_prep_parts(void);
int adc_quad_head_bwd_prep(const float* gy, const float* x10, const float* x0d, float* gq, float* part, int N, int oc, int H, int W,
                           adc_stream_t stream);
/* Backward contractions of the quad conv10 as direct kernels (4*oc parity channels are too thin for the tensor pipe); every
 * output comes with adc_quad_head_bwd_prep_parts() per-block partial rows to be summed over dim 0 in index order.
 * adc_quad_head_wgrad: part[.][4*oc][32][9]: dWq[qo][c][3a+b] = sum_p gq[p][qo] * x9[p + (a-1, b-1)][c].
 * adc_quad_head_dgrad: G9[p][c] = LeakyReLU'(x9[p][c]) * ( sum Wq[qo][c][a][b] * gq[p - (a-1, b-1)][qo] + extra[p][c] ) as bf16
 *   (hi, lo), part[.][32] = per-channel sums of G9 (bias gradient of conv9); gq fp32 zero-bordered [N][H1+2][W1+2][4*oc]. */
int adc_quad_head_wgrad(const void* x9, const void* x9_lo, int x9_stride, const float* gq, float* part, int N, int oc, int H1, int W1,
                        adc_stream_t stream);
int adc_quad_head_dgrad(const void* x9, int x9_stride, const float* gq, const float* wq, const void* extra, const void* extra_lo,
                        int extra_stride, void* g9, void* g9_lo, int g9_stride, float* part, int N, int oc, int H1, int W1,
                        adc_stream_t stream);
int adc_quad_unshuffle(const float* q, int q_stride, float* out, int N, int K, int H, int W, adc_stream_t stream);
/* Quad (parity-block) forms of a 3x3 weight tensor w [A][B][9] (fp32), replacing the per-step broadcast product + reduction over the
 * constant 0/1 tap maps: out [4][A][B][9] (or [4][B][A][9] when `transposed`), parity q = 2*py + px.
 *   which 0 ('up', conv10 behind the 2x nearest up-sampling, confidence_depth_framework.py:504-506): low-resolution tap
 *           (((py+u-1)>>1)+1, ((px+v-1)>>1)+1) accumulates the original tap (u, v);
 *   which 1 ('s2', data gradient of a 3x3 / stride-2 / pad-1 convolution, :442-452, and ConvTranspose2d(3, 2, 1, 1),
 *           model_zoo/s2d_u_resnet.py:45-61): low-resolution tap (a, b) holds the original tap (py+3-2a, px+3-2b) where it exists.
 * adc_quad_tap_mix_adjoint: the transpose of the 'up' form, out[l][t] = sum_q in[q][l][slot(q, t)] (weight gradient of conv10). */
int adc_quad_tap_mix(const float* w, float* out, int A, int B, int which, int transposed, adc_stream_t stream);
int adc_quad_tap_mix_adjoint(const float* in, float* out, int L, adc_stream_t stream);

/* ---- cbr3-c1 confidence head behind its 3x3 convolution (confidence_depth_framework.py:299-329: conv1 -> bn1 -> LeakyReLU(0.2) ->
 * conv2 1x1 + bias -> sigmoid).  conv1, its data gradient and its weight gradient run on the tensor-core kernels above; these
 * entry points replace nn.BatchNorm2d + nn.LeakyReLU + the 1x1 nn.Conv2d + torch.sigmoid and their autograd nodes.
 * y: conv1's output, fp32 zero-bordered NHWC [N][H+2][W+2][y_stride], channels [0,C), C <= 128, pad channels zero.
 * coef: adc_conf_head_coef_floats() floats, rows of 128: a = gamma*rstd | b = beta - mean*a | mean | rstd | k1 | k2 (backward).
 * adc_conf_bn_stats: training != 0: batch statistics (biased variance, like F.batch_norm) and, if running_mean/var are non-NULL,
 *   their momentum update with the unbiased variance; training == 0: the running statistics.  Fills coef rows 0..3.
 * adc_conf_head_apply: conf [N][1][H][W] = sigmoid(b2 + sum_c w2[c] * lrelu_0.2(a[c]*y[c] + b[c])).
 * adc_conf_head_bwd: from gconf = dL/dconf: gw2 [C], gb2 [1], ggamma [C], gbeta [C] (deterministic two-stage reductions) and
 *   gy = dL/dy as zero-bordered NHWC bf16 (hi, lo; lo may be NULL) [N][H+2][W+2][gy_stride] (interior pixels, ALL gy_stride
 *   channels written, zeros past C; the border is not touched).  workspace: adc_conf_head_workspace_bytes(C) bytes. */
size_t adc_conf_head_workspace_bytes(int C);
int adc_conf_head_coef_floats(void);
int adc_conf_bn_stats(const float* y, int y_stride, int N, int C, int H, int W, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, int training, float* coef, void* workspace,
                      adc_stream_t stream);
int adc_conf_head_apply(const float* y, int y_stride, const float* coef, const float* w2, const float* b2, float* conf, int N, int C, int H,
                        int W, adc_stream_t stream);
int adc_conf_head_bwd(const float* y, int y_stride, float* coef, const float* w2, const float* gconf, const float* conf, int training,
                      void* gy, void* gy_lo, int gy_stride, float* gw2, float* gb2, float* ggamma, float* gbeta, void* workspace, int N, int C,
                      int H, int W, adc_stream_t stream);

/* ---- masked losses and depth metrics (confidence_depth_framework.py:829-914 MaskedMSELoss / MaskedL1Loss / InvertedMaskedL1Loss /
 * MaskedAbsRelLoss; metrics.py:38-66 Result.evaluate).  pred / target / output / confidence: contiguous fp32 tensors of n elements.
 * kind: 0 l2 (t-p)^2, 1 l1 |t-p|, 2 il1 |1/t - 1/p|, 3 absrel |(t-p)/t|; the mask is target > 0.
 * adc_masked_loss_fwd: out[0] = mean of the error over the mask, out[1] = number of valid pixels (as float).
 * adc_masked_loss_bwd: gpred[i] = gloss[0] / count[0] * d err / d pred on the mask, 0 elsewhere (gloss, count: device scalars).
 * adc_depth_metrics: out[0..9] = mse, mae, lg10, absrel, delta1, delta2, delta3, imse, imae (means over target > 0 [and
 *   confidence > threshold if confidence != NULL]) and the number of valid pixels.
 * workspace: adc_loss_metrics_workspace_bytes() bytes.  Deterministic (fixed-order two-stage reduction in double). */
size_t adc_loss_metrics_workspace_bytes(void);
int adc_masked_loss_fwd(const float* pred, const float* target, long long n, int kind, void* workspace, float* out, adc_stream_t stream);
int adc_masked_loss_bwd(const float* pred, const float* target, const float* gloss, const float* count, long long n, int kind, float* gpred,
                        adc_stream_t stream);
int adc_depth_metrics(const float* output, const float* target, const float* confidence, float threshold, long long n, void* workspace,
                      float* out, adc_stream_t stream);
/* Precision/recall table of the confidence (metrics.py:103-131 ConfidencePixelwiseThrAverageMeter.evaluate): one pass instead of
 * num_bins boolean-index passes.  thresholds: device fp32 [num_bins], ascending.  hist: device uint64 [3][num_bins + 1] (zeroed by the
 * call): bin k collects the valid pixels (target > 0) with exactly k thresholds below their confidence -- [0] pixel count, [1] sum of
 * absrel = |d - t| / t * 1000 in 2^-16 fixed point, [2] pixels whose absrel is not finite.  Threshold j's figures are the sums over
 * bins k > j.  Integer accumulation: deterministic. */
int adc_conf_pr_histogram(const float* depth, const float* confidence, const float* target, long long n, const float* thresholds,
                          int num_bins, unsigned long long* hist, adc_stream_t stream);

/* ---- descriptor form of the tensor-core convolutions (csrc/conv_tc.cu), used by the BN nets (model_zoo/s2d_u_resnet.py:64-190)
 * kind 0: 3x3 / stride 1 (row-segment kernel, zero-bordered tensors, cin % 16 == 0, cout <= 128 -- in the split mode up to 128 through the CTA-pair form, 80 with ADC_V2_NO_PAIR);
 * kind 1: 1x1 on the same kernel; kind 2: 3x3 of stride 1 or 2 (box kernel, tensors with or without border, cout <= 256);
 * kind 3: 1x1 / stride 2 (the ResNet shortcut, s2d_u_resnet.py:99-107); kind 4: ConvTranspose2d(3, stride 2, padding 1,
 * output_padding 1) forward = data gradient of a stride-2 convolution (H, W = the larger tensor, w prepared with mode 2);
 * kind 5: data gradient of a 1x1 / stride-2 convolution (touches the even pixels of the larger tensor only).
 * accumulate: fp32 output += result.  Split (bf16x3) mode when x_lo and w_lo are set. */
typedef struct {
    int kind;
    const void *x, *x_lo; int x_stride, x_off, x_pad;
    const void *w, *w_lo; const float* bias;
    void *out, *out_lo; int out_f32, out_stride, out_off, out_pad, out_zero_to, accumulate;
    int N, H, W, cin, cout, stride, act, tap_mask;
} adc_conv_t;
int adc_conv_tc(const adc_conv_t* d, adc_stream_t stream);

/* ---- BatchNorm2d (+ residual) + ReLU / LeakyReLU(0.2) behind a convolution, forward and backward (csrc/bn_act.cu), replacing
 * nn.BatchNorm2d + activation modules of model_zoo/s2d_u_resnet.py:26-61, torchvision BasicBlock (:99-107) and the bias + ReLU layers of
 * GMSNet (confidence_depth_framework.py:536-607; gamma == NULL = identity affine).  y: raw convolution output, fp32 zero-bordered NHWC;
 * activations / gradient operands: zero-bordered NHWC bf16 (hi, lo) pairs (lo may be NULL); raw gradients: fp32 zero-bordered NHWC.
 * Every tensor is a channel slice [off, off + C) of `stride` channels per pixel.  coef: adc_bn_coef_floats(C) floats written by
 * adc_bn_stats, read by the others (adc_bn_bwd adds its two means); workspace: adc_bn_workspace_bytes(C).  act: 0 none, 1 ReLU,
 * 2 LeakyReLU(0.2).  adc_bn_bwd: g = gradient of the activation output -> dy (operand of the gradient GEMMs), ggamma / gbeta (may be
 * NULL), and gres (may be NULL) = gradient of the residual input, written or accumulated. */
size_t adc_bn_workspace_bytes(int C);
int adc_bn_coef_floats(int C);
int adc_bn_stats(const float* y, int y_stride, int y_off, int N, int C, int H, int W, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps, int training, float* coef, void* workspace,
                 adc_stream_t stream);
int adc_bn_apply(const float* y, int y_stride, int y_off, const float* coef, const void* res_hi, const void* res_lo, int res_stride, int res_off,
                 int act, void* out_hi, void* out_lo, int out_stride, int out_off, int N, int C, int H, int W, adc_stream_t stream);
int adc_bn_bwd(const float* g, int g_stride, int g_off, const float* y, int y_stride, int y_off, const void* out_hi, int out_stride, int out_off,
               float* coef, int act, int training, void* dy_hi, void* dy_lo, int dy_stride, int dy_off, float* gres, int gres_stride,
               int gres_off, int gres_accum, float* ggamma, float* gbeta, void* workspace, int N, int C, int H, int W, adc_stream_t stream);
/* 1x1 convolution to ONE channel over a (hi, lo) NHWC slice and its backward (s2d_u_resnet.py:128 convtf): y fp32 [N][H][W] planar;
 * y_hi / y_lo (may be NULL): the raw result also as channel y_off of a zero-bordered NHWC (hi, lo) tensor -- cat(last_layer_input, y)
 * of s2d_u_resnet.py:179 without a copy.  gx: fp32 zero-bordered NHWC gradient slice (written or accumulated), gw [C], gb [1];
 * gy_extra (may be NULL): a gradient of y that another consumer left at gy_extra[pixel * gy_extra_stride] of an NHWC gradient tensor
 * (added to gy); workspace: adc_bn_workspace_bytes(C). */
int adc_dot_head_fwd(const void* x_hi, const void* x_lo, int x_stride, int x_off, const float* w, const float* bias, float* y, int relu, int N,
                     int C, int H, int W, void* y_hi, void* y_lo, int y_stride, int y_off, adc_stream_t stream);
int adc_dot_head_bwd(const void* x_hi, const void* x_lo, int x_stride, int x_off, const float* w, const float* gy, float* gx, int gx_stride,
                     int gx_off, int gx_accum, float* gw, float* gb, void* workspace, int N, int C, int H, int W, const float* gy_extra,
                     int gy_extra_stride, adc_stream_t stream);
/* im2col of a 3x3 / pad-1 window (stride 1 or 2) over the C = c0 + c1 + c2 <= 7 fp32 planes of up to three NCHW sources (s1, s2 may be
 * NULL; bs = batch stride in floats) at the output resolution: channel (3u+v)*C + c of a 64-channel (hi, lo) zero-bordered NHWC pixel.
 * The thin stems of s2d_u_resnet.py:91-97 (3 -> 48, 2 -> 16) become one GEMM launch. */
int adc_im2col3x3_pack(const float* s0, int c0, long long bs0, const float* s1, int c1, long long bs1, const float* s2, int c2, long long bs2,
                       void* dst, void* dst_lo, int dst_stride, int N, int H, int W, int stride, adc_stream_t stream);
/* fp32 NCHW planes -> channels [dst_off, dst_off + C) of a fp32 zero-bordered NHWC tensor (written or accumulated). */
int adc_nchw_to_nhwc_f32(const float* src, float* dst, int dst_stride, int dst_off, int accumulate, int N, int C, int H, int W,
                         adc_stream_t stream);

/* ---- optimizer step of the trainer (trainer.py:134-146, torch.optim.Adam over the trainable parameters) on flat buffers: every
 * parameter / gradient / moment is a view into one fp32 vector of n elements (16-byte aligned).  `step` is a device scalar holding the
 * number of updates done so far; the call increments it first (so a replayed CUDA graph advances the bias correction) and applies
 * torch.optim.Adam's update (amsgrad=False): g' = g + wd*p; m += (g'-m)(1-b1); v = b2 v + (1-b2) g'^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps). */
int adc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step, long long n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, adc_stream_t stream);

/* Same update with the hyper-parameters resident on the device: hyper = double[5] {lr, beta1, beta2, eps, weight_decay}, state =
 * float[4] {t, step_size, 1/sqrt(1-beta2^t), apply} (only t is an input).  A captured CUDA graph follows a learning-rate schedule
 * (trainer.py:236-241 multiplies lr by 0.1 every 5 epochs) because the host rewrites `hyper` between replays.  guard (may be NULL):
 * device scalar, normally the step's loss; when it is not finite the call changes nothing, t included -- the reference trainer skips
 * such a batch on the host (trainer.py:206-208). */
int adc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* state, const double* hyper,
                      const float* guard, long long n, adc_stream_t stream);

/* ---- S2DResNet (model_zoo/s2d_resnet.py:159-242, resnet18) around the tensor-core convolutions: csrc/s2d_ops.cu ------------------------
 * Activations: zero-bordered NHWC bf16 (hi, lo) pairs (lo NULL in the plain bf16 mode); raw results / gradients: zero-bordered NHWC
 * fp32; channel slices [off, off + C) of `stride` channels per pixel.
 * adc_im2col_kxk_pack:   im2col of a K x K window (stride 1 / 2, zero padding `pad`) over the C planes of an fp32 NCHW source at the
 *                        OUTPUT resolution: channel (u*K + v)*C + c, zeros up to dst_channels (multiple of 8).  The 7x7 / stride-2 stem
 *                        (s2d_resnet.py:175) becomes one GEMM.
 * adc_maxpool3x3s2_fwd:  nn.MaxPool2d(3, 2, 1) (:181) with ATen's scan order and tie rule; idx: uint8 [N][Ho][Wo][C] window position
 *                        3u + v of the maximum.  _bwd: gin (written) = scatter of g by idx, computed as a deterministic gather.
 * adc_pixel_shuffle_f32 / adc_pixel_unshuffle_pair: Y[n][2i+py][2j+px][c] <-> Q[n][i][j][(2py+px)*C + c], the parity-block form in
 *                        which UpProj's 5x5 convolutions over the zero-stuffed Unpool output (:11-24, :105-143), the stride-2
 *                        ConvTranspose2d layers of s2d_u_resnet.py:45-61 and the data gradients of stride-2 convolutions are 3x3 stride-1
 *                        convolutions of the low-resolution tensor.  The shuffle writes or accumulates into an H x W tensor with
 *                        H in {2h-1, 2h} (odd sizes of a stride-2 layer's larger tensor).
 * adc_bilinear_ac_fwd:   F.interpolate(mode='bilinear', align_corners=True) (:230-234) of an NHWC slice (fp32, or a bf16 pair when
 *                        src_is_pair) to fp32 NCHW [N][C][Hd][Wd], ATen's index / weight arithmetic.  _bwd: gsrc (written or
 *                        accumulated) = its adjoint, computed as a deterministic gather. */
int adc_im2col_kxk_pack(const float* src, long long batch_stride, int C, int K, int stride, int pad, void* dst, void* dst_lo, int dst_stride,
                        int dst_channels, int N, int H, int W, adc_stream_t stream);
/* Adjoint of adc_im2col_kxk_pack: gcol fp32 zero-bordered NHWC [N][Ho+2][Wo+2][gcol_stride] (channel (u*K + v)*C + c) -> gx fp32 NCHW
 * [N][C][H][W]: the data gradient of the im2col'ed stem convolution (deterministic gather). */
int adc_col2im_kxk(const float* gcol, int gcol_stride, int C, int K, int stride, int pad, float* gx, int N, int H, int W, adc_stream_t stream);
int adc_maxpool3x3s2_fwd(const void* x, const void* x_lo, int x_stride, int x_off, void* out, void* out_lo, int out_stride, int out_off,
                         unsigned char* idx, int N, int C, int H, int W, adc_stream_t stream);
int adc_maxpool3x3s2_bwd(const float* g, int g_stride, int g_off, const unsigned char* idx, float* gin, int gin_stride, int gin_off, int N, int C,
                         int H, int W, adc_stream_t stream);
int adc_pixel_shuffle_f32(const float* q, int q_stride, int q_off, float* y, int y_stride, int y_off, int N, int C, int h, int w, int H, int W,
                          int accumulate, adc_stream_t stream);
int adc_pixel_unshuffle_pair(const void* d, const void* d_lo, int d_stride, int d_off, void* q, void* q_lo, int q_stride, int q_off, int N, int C,
                             int h, int w, adc_stream_t stream);
int adc_bilinear_ac_fwd(const void* src, const void* src_lo, int src_is_pair, int src_stride, int src_off, int N, int C, int hs, int ws, float* dst,
                        int Hd, int Wd, adc_stream_t stream);
int adc_bilinear_ac_bwd(const float* g, int N, int C, int Hd, int Wd, float* gsrc, int gsrc_stride, int gsrc_off, int hs, int ws, int accumulate,
                        adc_stream_t stream);

/* ---- sparse-depth sampling (dataloaders/dense_to_sparse.py:29-48, UniformSampling.dense_to_sparse) on the device, bit-exact with
 * the reference's numpy stream: frame n draws np.random.uniform(0, 1, HW) after np.random.seed(seeds[n]) (MT19937, 53-bit doubles)
 * and keeps mask = (depth > 0 [& depth <= max_depth]) & (u < num_samples / count(keep)).  depth: fp32 [N][HW]; seeds: device
 * uint32 [N]; mask: uint8 [N][HW] (0 / 1).  num_samples = 0 or no valid pixel: mask = keep, like the reference. */
int adc_uniform_sampling(const float* depth, const unsigned int* seeds, int num_samples, float max_depth, int use_max_depth,
                         unsigned char* mask, int N, long long HW, adc_stream_t stream);
/* The loader's modality assembly on the device (dataloaders/dataloader_ext.py:544-566 after dense_to_sparse.py:29-48): from the
 * uploaded frames rgbgt [N][4][HW] = (r, g, b, ground-truth depth) builds the network input input5 [N][5][HW] = (r, g, b, d, w) with
 * d = gt * mask, w = mask (1.0 / 0.0), mask = the bit-exact UniformSampling mask of frame n under seeds[n].  Only 4 of the 6 planes a
 * host-side loader ships (5 input + 1 target) cross PCIe; the target is the gt plane of rgbgt. */
int adc_sample_assemble(const float* rgbgt, const unsigned int* seeds, int num_samples, float max_depth, int use_max_depth, float* input5,
                        int N, long long HW, adc_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* ADC_B200_H */

"""Tensor-core training / inference path of the BatchNorm nets: the full-resolution ResNet-18 U-Net of
``model_zoo/s2d_u_resnet.py:64-190`` (udepthcompnet18 / gudepthcompnet18) and the general cbr3-c1 confidence head
(``confidence_depth_framework.py:299-329``, any channel count), forward AND backward on the hand-written kernels.

Layout (as in tc_train.py): every activation is a zero-bordered NHWC bf16 (hi, lo) pair ``[N][H+2][W+2][C]`` living in the channel slice
of the buffer that consumes it, so the five ``torch.cat`` of the decoder are slice writes and their backward slice reads.  Per layer:

  conv (tcgen05 implicit GEMM, fp32 result Y)  ->  bn_stats(Y)  ->  bn_apply: act(a*Y + b [+ residual]) as the next (hi, lo) operand
  backward: raw gradients are fp32 NHWC buffers that several consumers ACCUMULATE into (convolution epilogues with accumulate = 1);
  bn_bwd turns one into the (hi, lo) operand dY of the two gradient GEMMs (+ dgamma, dbeta, + the residual branch's gradient).

Convolution forms (csrc/conv_tc.cu, adc_conv_tc):
  3x3 s1            row-segment kernel, output channels in chunks of 64 (split mode) / 128
  3x3 s2, 1x1 s2    box kernel (strided TMA boxes); the shortcut's 1x1 reads the centre tap of an embedded 3x3 weight tensor
  ConvT 3x3 s2 op1  = the data gradient of a stride-2 convolution whose weight tensor is the ConvTranspose2d weight itself ([in][out]
                      is [Cout_conv][Cin_conv]): four parity-class launches writing the 2x larger tensor directly
  ConvT 3x3 s1      = a stride-1 convolution with transposed + flipped weights (prep mode 1)
  stems 3->48, 2->16  one GEMM over an im2col tensor (45 of 64 channels), block-diagonal weights
  1x1 128 -> 1      CUDA-core dot-product kernel over the (hi, lo) slice (HBM-bound, AI 0.5)
Gradients of every form are the same kernels with roles swapped (see the helpers below); weight gradients beyond the kernels' channel
limits are computed per (128 x 128) channel block through channel-slice offsets.

Buffers are allocated once per (module, batch shape, precision) and zero-filled once; kernels write interiors only, so borders and
pad channels stay zero and a captured CUDA graph sees fixed addresses.
"""
import ctypes

import torch

from . import dense
from ._lib import lib, check, ConvDesc

NONE, RELU, LEAKY = 0, 1, 2
K_S1, K_1X1, K_BOX, K_1X1S2, K_CONVT2, K_1X1S2_DGRAD = 0, 1, 2, 3, 4, 5


def _chunk():
    """Output channels per launch (dense.out_chunk: 128, or 64 for the single-CTA split form)."""
    return dense.out_chunk()


def _chunks(c, step):
    return [(a, min(a + step, c)) for a in range(0, c, step)]


def _r8(c):
    return (c + 7) // 8 * 8


def _ptr(t):
    return t.data_ptr() if t is not None else None


def f32_act(n, h, w, c, dev):
    return dense.new_act(n, h, w, c, dev, pad=True, zero=True, dtype=torch.float32)


def pair_act(n, h, w, c, dev):
    return dense.new_act(n, h, w, c, dev, pad=True, zero=True)


def conv(kind, x, x_off, cin, wp, out, out_off, cout, H, W, stride=1, bias=None, act=NONE, accumulate=False, out_zero_to=0, tap_mask=0):
    """One launch of adc_conv_tc.  x / out: dense.Act (out fp32 or a bf16 pair); H, W: input size (kinds 0-3) or the size of the LARGER
    tensor (kinds 4, 5)."""
    d = ConvDesc()
    d.kind = kind
    d.x, d.x_lo, d.x_stride, d.x_off, d.x_pad = x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off, int(x.pad)
    d.w, d.w_lo = wp[0].data_ptr(), _ptr(wp[1])
    d.bias = bias.detach().data_ptr() if bias is not None else None
    f32 = out.hi.dtype == torch.float32
    d.out, d.out_lo, d.out_f32 = out.hi.data_ptr(), (None if f32 else _ptr(out.lo)), int(f32)
    d.out_stride, d.out_off, d.out_pad, d.out_zero_to, d.accumulate = out.cs, out_off, int(out.pad), out_zero_to, int(bool(accumulate))
    d.N, d.H, d.W, d.cin, d.cout, d.stride, d.act, d.tap_mask = x.n, H, W, cin, cout, stride, act, tap_mask
    check(lib.adc_conv_tc(ctypes.byref(d), dense._stream()), "adc_conv_tc")
    return out


def embed_1x1(w):
    """[Cout, Cin, 1, 1] -> [Cout, Cin, 3, 3] with the weights in the centre tap (what the 1x1 kernels read)."""
    full = w.new_zeros(w.shape[0], w.shape[1], 3, 3)
    full[:, :, 1, 1] = w.detach()[:, :, 0, 0]
    return full


class WPrep:
    """Collects weight-preparation requests of one pass and runs them in as few launches as possible."""

    def __init__(self):
        self.reqs, self.keys, self.out = [], [], {}

    def add(self, key, weight, mode, cin_pad, ci_off, ci_cnt):
        self.keys.append(key)
        self.reqs.append((weight, mode, cin_pad, ci_off, ci_cnt))

    def fwd(self, key, weight, cin_pad, step=None):
        """Forward weights of a [Cout, Cin, 3, 3] convolution, one item per output-channel chunk."""
        w = weight.detach()
        for a, b in _chunks(w.shape[0], step or _chunk()):
            self.add((key, a), w[a:b], 0, cin_pad, 0, w.shape[1])

    def transposed(self, key, weight, cin_pad, flip, step=None):
        """[Cout, Cin, 3, 3] read as Cin output channels contracting Cout (taps flipped or kept), per chunk of Cin."""
        w = weight.detach()
        for a, b in _chunks(w.shape[1], step or _chunk()):
            self.add((key, a), w, 1 if flip else 2, cin_pad, a, b - a)

    def quad_s2(self, key, w_io, cin_pad):
        """Parity-block weights of a stride-2 transposed convolution / stride-2 data gradient (``quad_s2_weights``), one item per chunk."""
        wq = quad_s2_weights(w_io)
        self.fwd(key, wq, cin_pad, step=min(_chunk(), w_io.shape[1]))

    def run(self):
        if self.reqs:
            for k, v in zip(self.keys, dense.prep_conv3x3_weights_batch(self.reqs)):
                self.out[k] = v
        self.reqs, self.keys = [], []
        return self.out


# ---- BatchNorm + activation passes (csrc/bn_act.cu) ---------------------------------------------------------------------------------
class BNState:
    __slots__ = ("coef", "ws", "training")


def bn_forward(Y, y_off, C, bn, training, out, out_off, act, res=None, res_off=0):
    """Y[:, y_off:y_off+C] (fp32) -> out[:, out_off:out_off+C] = act(BN(Y) [+ res]); returns the state bn_backward needs.
    bn None = identity affine (a plain bias + activation layer whose bias the convolution already added)."""
    dev = Y.hi.device
    s = BNState()
    s.coef = torch.empty(lib.adc_bn_coef_floats(C), device=dev, dtype=torch.float32)
    s.ws = torch.empty(lib.adc_bn_workspace_bytes(C) // 8, device=dev, dtype=torch.float64)
    s.training = bool(training and bn is not None)
    if bn is not None:
        use_batch = bool(training or bn.running_mean is None)
        s.training = use_batch
        momentum = bn.momentum if bn.momentum is not None else 0.1
        check(lib.adc_bn_stats(Y.hi.data_ptr(), Y.cs, y_off, Y.n, C, Y.h, Y.w, bn.weight.detach().data_ptr(), bn.bias.detach().data_ptr(),
                               _ptr(bn.running_mean), _ptr(bn.running_var), float(momentum), float(bn.eps), int(use_batch), s.coef.data_ptr(),
                               s.ws.data_ptr(), dense._stream()), "adc_bn_stats")
        if use_batch and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
    else:
        check(lib.adc_bn_stats(Y.hi.data_ptr(), Y.cs, y_off, Y.n, C, Y.h, Y.w, None, None, None, None, 0.0, 0.0, 0, s.coef.data_ptr(),
                               s.ws.data_ptr(), dense._stream()), "adc_bn_stats")
    rh, rl, rs = (res.hi.data_ptr(), _ptr(res.lo), res.cs) if res is not None else (None, None, 0)
    check(lib.adc_bn_apply(Y.hi.data_ptr(), Y.cs, y_off, s.coef.data_ptr(), rh, rl, rs, res_off, act, out.hi.data_ptr(), _ptr(out.lo), out.cs,
                           out_off, Y.n, C, Y.h, Y.w, dense._stream()), "adc_bn_apply")
    return s


def bn_backward(G, g_off, Y, y_off, out, out_off, state, act, dY, dy_off, C, gres=None, gres_off=0, gres_accum=False, affine=True):
    """dL/d(out) (fp32 slice of G) -> dY (bf16 pair slice) [+ the residual input's gradient]; returns (dgamma, dbeta) fp32 [C]."""
    dev = Y.hi.device
    gg = torch.empty(C, device=dev, dtype=torch.float32) if affine else None
    gb = torch.empty(C, device=dev, dtype=torch.float32)
    check(lib.adc_bn_bwd(G.hi.data_ptr(), G.cs, g_off, Y.hi.data_ptr(), Y.cs, y_off, out.hi.data_ptr() if act else None, out.cs if act else 0, out_off,
                         state.coef.data_ptr(), act, int(state.training), dY.hi.data_ptr(), _ptr(dY.lo), dY.cs, dy_off,
                         gres.hi.data_ptr() if gres is not None else None, gres.cs if gres is not None else 0, gres_off, int(bool(gres_accum)),
                         _ptr(gg), gb.data_ptr(), state.ws.data_ptr(), Y.n, C, Y.h, Y.w, dense._stream()), "adc_bn_bwd")
    return gg, gb


def nchw_into(src, G, off, accumulate):
    src = src.contiguous()
    n, c, h, w = src.shape
    check(lib.adc_nchw_to_nhwc_f32(src.data_ptr(), G.hi.data_ptr(), G.cs, off, int(bool(accumulate)), n, c, h, w, dense._stream()),
          "adc_nchw_to_nhwc_f32")


# ---- forward convolutions into an fp32 result buffer --------------------------------------------------------------------------------
def fwd_s1(x, x_off, cin, wp, key, Y, y_off, cout, bias=None, act=NONE, step=None):
    for a, b in _chunks(cout, step or _chunk()):
        conv(K_S1, x, x_off, cin, wp[(key, a)], Y, y_off + a, b - a, x.h, x.w, bias=bias[a:b] if bias is not None else None, act=act)


def fwd_box(x, x_off, cin, wp, key, Y, y_off, cout, stride, kind=K_BOX):
    for a, b in _chunks(cout, _chunk()):
        conv(kind, x, x_off, cin, wp[(key, a)], Y, y_off + a, b - a, x.h, x.w, stride=stride)


# ---- stride-2 transposed convolutions and stride-2 data gradients in parity-block form -------------------------------------------------
# out[2m+py][2n+px][o] = sum over (di, dj), c of x[m+di][n+dj][c] * w_io[c][o][u][v]  with  u = py + 1 - 2*di, v = px + 1 - 2*dj  in 0..2
# (ConvTranspose2d(3, 2, 1, 1) with weight [in][out]; equally the data gradient of a 3x3 / stride-2 / pad-1 convolution with weight
# [Cout][Cin] = [in][out]).  Per parity (py, px) that is a stride-1 3x3 convolution of the LOW-resolution tensor using 1, 2, 2 or 4 of
# its nine taps (dense.quad_maps 's2'): the row-segment kernel with a tap mask per parity block of output channels writes a
# [N][h][w][4*out] tensor, adc_pixel_shuffle_f32 scatters (or accumulates) it into the 2x larger tensor.  Measured on the config-3 step
# (B = 16, 240x320): the per-tap box kernel this replaces ran at 49 TFLOP/s, the row-segment kernel runs these layers at 190-250.
_SCRATCH = {}


def _scratch_f32(n, h, w, c, dev):
    key = (n, h, w, c, str(dev))
    if key not in _SCRATCH:
        _SCRATCH[key] = f32_act(n, h, w, c, dev)
    return _SCRATCH[key]


def quad_s2_weights(w_io):
    """[in, out, 3, 3] (contracted over ``in``) -> [4*out, in, 3, 3]: row (2py+px)*out + o holds, at low-resolution tap (di+1, dj+1), the
    original tap (py+1-2di, px+1-2dj) where it exists."""
    cin, cout = w_io.shape[:2]
    return dense.quad_tap_mix(w_io, "s2", transposed=True).reshape(4 * cout, cin, 3, 3)


def quad_s2(x, x_off, cin, wp, key, Y, y_off, cout, accumulate=False):
    """x (low resolution) -> Y[:, y_off:y_off+cout] (the 2x larger tensor, fp32; odd sizes allowed), written or accumulated."""
    masks, _ = dense.quad_s2_tap_masks()
    step = min(_chunk(), cout)
    Yq = _scratch_f32(x.n, x.h, x.w, dense._r64(4 * cout), x.hi.device)
    for a, b in _chunks(4 * cout, step):
        m = 0
        for q in range(a // cout, (b - 1) // cout + 1):
            m |= masks[q]
        conv(K_S1, x, x_off, cin, wp[(key, a)], Yq, a, b - a, x.h, x.w, tap_mask=m)
    check(lib.adc_pixel_shuffle_f32(Yq.hi.data_ptr(), Yq.cs, 0, Y.hi.data_ptr(), Y.cs, y_off, x.n, cout, x.h, x.w, Y.h, Y.w, int(bool(accumulate)),
                                    dense._stream()), "adc_pixel_shuffle_f32")


def fwd_convt2(x, x_off, cin, wp, key, Y, y_off, cout):
    """ConvTranspose2d(3, stride 2, padding 1, output_padding 1): Y is the 2x larger tensor."""
    quad_s2(x, x_off, cin, wp, key, Y, y_off, cout, False)


# ---- data gradients into an fp32 gradient buffer (first producer writes, later ones accumulate) --------------------------------------
def dgrad_s1(dY, cy, wp, key, G, g_off, cin, accumulate, step=None):
    for a, b in _chunks(cin, step or _chunk()):
        conv(K_S1, dY, 0, cy, wp[(key, a)], G, g_off + a, b - a, dY.h, dY.w, accumulate=accumulate)


def dgrad_s2(dY, cy, wp, key, G, g_off, cin, accumulate, kind=K_CONVT2):
    """Data gradient of a stride-2 3x3 convolution (parity-block form, weights from WPrep.quad_s2) or of a 1x1 / stride-2 convolution
    (kind 5: box kernel on the even pixels, weights from WPrep.transposed): G is the 2x larger tensor."""
    if kind == K_CONVT2:
        return quad_s2(dY, 0, cy, wp, key, G, g_off, cin, accumulate)
    for a, b in _chunks(cin, _chunk()):
        conv(kind, dY, 0, cy, wp[(key, a)], G, g_off + a, b - a, G.h, G.w, accumulate=accumulate)


def dgrad_convt2(dY, cy, wp, key, G, g_off, cin, accumulate):
    """Data gradient of the stride-2 ConvTranspose2d = a stride-2 convolution of dY (the larger tensor)."""
    for a, b in _chunks(cin, _chunk()):
        conv(K_BOX, dY, 0, cy, wp[(key, a)], G, g_off + a, b - a, dY.h, dY.w, stride=2, accumulate=accumulate)


# ---- weight gradients per channel block -----------------------------------------------------------------------------------------------
def _blocked(cout, cin, co_step, ci_step, fn, taps=True):
    if dense._r16(cin) <= 160:
        ci_step = cin                          # the kernels take up to 160 input channels in one launch (the 129-channel head: 144)
    dev_out = None
    blocks = [(a, b, c, d) for a, b in _chunks(cout, co_step) for c, d in _chunks(cin, ci_step)]
    if len(blocks) == 1:
        return fn(0, cout, 0, cin)
    for a, b, c, d in blocks:
        part = fn(a, b, c, d)
        if dev_out is None:
            dev_out = part.new_empty((cout, cin) + tuple(part.shape[2:]))
        dev_out[a:b, c:d] = part
    return dev_out


def wgrad_s1(dY, dy_off, cout, x, x_off, cin):
    """[cout, cin, 3, 3] of a stride-1 3x3 convolution (flat kernel: <= 128 output x 128 input channels per launch)."""
    return _blocked(cout, cin, 128, 128, lambda a, b, c, d: dense.conv3x3_wgrad_tc(dY, x, d - c, b - a, x_off=x_off + c, gy_off=dy_off + a,
                                                                                   cin_pad=dense._r16(d - c)))


def wgrad_box(dY, dy_off, cout, x, x_off, cin, stride):
    return _blocked(cout, cin, 128, 128, lambda a, b, c, d: dense.conv3x3_wgrad_box_tc(dY, x, d - c, b - a, stride, x_off=x_off + c,
                                                                                       gy_off=dy_off + a, cin_pad=dense._r16(d - c)))


def wgrad_1x1_box(dY, dy_off, cout, x, x_off, cin, stride):
    def one(a, b, c, d):
        ci, co = d - c, b - a
        cin_pad = dense._r16(ci)
        dw = torch.empty((co, ci), device=x.hi.device, dtype=torch.float32)
        ws = torch.empty(lib.adc_conv3x3_wgrad_tc_workspace_bytes(cin_pad, co) // 4, device=x.hi.device, dtype=torch.float32)
        check(lib.adc_conv1x1_wgrad_box_tc(dY.hi.data_ptr(), _ptr(dY.lo), dY.cs, dy_off + a, int(dY.pad), x.hi.data_ptr(), _ptr(x.lo), x.cs,
                                           x_off + c, int(x.pad), dw.data_ptr(), ws.data_ptr(), x.n, x.h, x.w, stride, ci, cin_pad, co,
                                           dense._stream()), "adc_conv1x1_wgrad_box_tc")
        return dw
    return _blocked(cout, cin, 128, 128, one).reshape(cout, cin, 1, 1)


# ---- 1x1 head to one / two channels ---------------------------------------------------------------------------------------------------
def dot_head(x, x_off, C, weight, bias, relu=False, y_act=None, y_off=0):
    """[oc, C, 1, 1] convolution over a (hi, lo) slice -> fp32 NCHW [n, oc, h, w]; with ``y_act`` (oc == 1) the result is also written
    as channel ``y_off`` of that NHWC activation."""
    oc = weight.shape[0]
    y = torch.empty((x.n, oc, x.h, x.w), device=x.hi.device, dtype=torch.float32)
    w2 = weight.detach().reshape(oc, C).contiguous()
    b = bias.detach() if bias is not None else None
    if oc == 1:
        yh, yl, ys = (y_act.hi.data_ptr(), _ptr(y_act.lo), y_act.cs) if y_act is not None else (None, None, 0)
        check(lib.adc_dot_head_fwd(x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off, w2.data_ptr(), _ptr(b), y.data_ptr(), int(relu), x.n, C, x.h, x.w,
                                   yh, yl, ys, y_off, dense._stream()), "adc_dot_head_fwd")
        return y
    planes = []
    for o in range(oc):
        yo = torch.empty((x.n, 1, x.h, x.w), device=x.hi.device, dtype=torch.float32)
        check(lib.adc_dot_head_fwd(x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off, w2[o].data_ptr(), b[o:o + 1].data_ptr() if b is not None else None,
                                   yo.data_ptr(), int(relu), x.n, C, x.h, x.w, None, None, 0, 0, dense._stream()), "adc_dot_head_fwd")
        planes.append(yo)
    return torch.cat(planes, 1)


def dot_head_backward(x, x_off, C, weight, gy, G, g_off, accumulate, gy_extra=None):
    """Backward of dot_head: G[:, g_off:g_off+C] (+)= gy * w; returns (dweight [oc, C, 1, 1], dbias [oc]).  gy_extra = (fp32 NHWC
    activation, channel): a second gradient of y sitting in that channel (added to gy; oc == 1)."""
    oc = weight.shape[0]
    dev = x.hi.device
    w2 = weight.detach().reshape(oc, C).contiguous()
    gy = gy.contiguous()
    gw = torch.empty((oc, C), device=dev, dtype=torch.float32)
    gb = torch.empty(oc, device=dev, dtype=torch.float32)
    ws = torch.empty(lib.adc_bn_workspace_bytes(C) // 8, device=dev, dtype=torch.float64)
    for o in range(oc):
        g_o = gy[:, o].contiguous() if oc > 1 else gy
        ex, exs = (gy_extra[0].hi.data_ptr() + 4 * gy_extra[1], gy_extra[0].cs) if (gy_extra is not None and oc == 1) else (None, 0)
        check(lib.adc_dot_head_bwd(x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off, w2[o].data_ptr(), g_o.data_ptr(), G.hi.data_ptr(), G.cs, g_off,
                                   int(bool(accumulate) or o > 0), gw[o].data_ptr(), gb[o:o + 1].data_ptr(), ws.data_ptr(), x.n, C, x.h, x.w,
                                   ex, exs, dense._stream()), "adc_dot_head_bwd")
    return gw.reshape(oc, C, 1, 1), gb


# ---- a ResNet stage of two BasicBlocks (torchvision names conv1 / bn1 / conv2 / bn2 / downsample.{0,1}) ----------------------------------
# Shared by the U-Net body below and by tc_s2d (S2DResNet).  ``P`` is a plan: P.y / P.x / P.g / P.d hand out the fp32 result, activation,
# fp32 gradient and gradient-operand buffers of a (key, level); P.S keeps the BatchNorm states.
def stage_prep_fwd(wp, st, name, cin, cout):
    for bi in (0, 1):
        blk = st[bi]
        wp.fwd((name, bi, 1), blk.conv1.weight, cin if bi == 0 else cout)
        wp.fwd((name, bi, 2), blk.conv2.weight, cout)
        if blk.downsample is not None:
            wp.fwd((name, bi, "d"), embed_1x1(blk.downsample[0].weight), cin)


def stage_forward(P, st, name, lvl, cin, cout, stride, xin, xoff, dst, doff, W, training):
    """xin[:, xoff:xoff+cin] -> dst[:, doff:doff+cout] (level ``lvl`` = the stage's OUTPUT level)."""
    for bi in (0, 1):
        blk = st[bi]
        ci = cin if bi == 0 else cout
        s_ = stride if bi == 0 else 1
        Ya = P.y((name, bi, 1), lvl, cout)
        if s_ == 1:
            fwd_s1(xin, xoff, ci, W, (name, bi, 1), Ya, 0, cout)
        else:
            fwd_box(xin, xoff, ci, W, (name, bi, 1), Ya, 0, cout, 2)
        hmid = P.x((name, bi, "h"), lvl, cout)
        P.S[(name, bi, 1)] = bn_forward(Ya, 0, cout, blk.bn1, training, hmid, 0, RELU)
        Yb = P.y((name, bi, 2), lvl, cout)
        fwd_s1(hmid, 0, cout, W, (name, bi, 2), Yb, 0, cout)
        if blk.downsample is not None:
            Yd = P.y((name, bi, "d"), lvl, cout)
            fwd_box(xin, xoff, ci, W, (name, bi, "d"), Yd, 0, cout, s_, kind=K_1X1S2 if s_ == 2 else K_BOX)
            sc = P.x((name, bi, "sc"), lvl, cout)
            P.S[(name, bi, "d")] = bn_forward(Yd, 0, cout, blk.downsample[1], training, sc, 0, NONE)
            res, roff = sc, 0
        else:
            res, roff = xin, xoff
        if bi == 0:
            o, ooff = P.x((name, "b0"), lvl, cout), 0
        else:
            o, ooff = dst, doff
        P.S[(name, bi, 2)] = bn_forward(Yb, 0, cout, blk.bn2, training, o, ooff, RELU, res=res, res_off=roff)
        xin, xoff = o, ooff


def stage_prep_bwd(wp, st, name, cin, cout, stride):
    for bi in (0, 1):
        blk = st[bi]
        s_ = stride if bi == 0 else 1
        if s_ == 1:
            wp.transposed((name, bi, 1), blk.conv1.weight, dense._r64(cout), flip=True)
        else:
            wp.quad_s2((name, bi, 1), blk.conv1.weight, dense._r64(cout))
        wp.transposed((name, bi, 2), blk.conv2.weight, dense._r64(cout), flip=True)
        if blk.downsample is not None:
            wp.transposed((name, bi, "d"), embed_1x1(blk.downsample[0].weight), dense._r64(cout), flip=False)


def stage_backward(P, st, name, lvl, cin, cout, stride, xin0, xoff0, Gin0, goff0, gin0_written, out, ooff_last, Gout, gooff_last, W, grads,
                   prefix):
    """Backward of stage_forward.  Gout[:, gooff_last:...] holds dL/d(stage output); the input gradient is written (or, when
    ``gin0_written``, accumulated) into Gin0[:, goff0:goff0+cin]; parameter gradients land in ``grads`` under ``prefix``."""
    for bi in (1, 0):
        blk = st[bi]
        ci = cin if bi == 0 else cout
        s_ = stride if bi == 0 else 1
        if bi == 1:
            o, ooff, Go, gooff = out, ooff_last, Gout, gooff_last
            xin, xoff = P.X[(name, "b0")], 0
            Gin, goff, gin_written = P.g((name, "b0"), lvl, cout), 0, False
        else:
            o, ooff, Go, gooff = P.X[(name, "b0")], 0, P.G[(name, "b0")], 0
            xin, xoff, Gin, goff, gin_written = xin0, xoff0, Gin0, goff0, gin0_written
        pre = prefix + "%d." % bi
        # bn2 (+ residual) + ReLU
        Db = P.d((name, bi, 2), lvl, cout)
        if blk.downsample is not None:
            Gsc = P.g((name, bi, "sc"), lvl, cout)
            gg, gbeta = bn_backward(Go, gooff, P.Y[(name, bi, 2)], 0, o, ooff, P.S[(name, bi, 2)], RELU, Db, 0, cout, gres=Gsc, gres_off=0,
                                    gres_accum=False)
        else:
            gg, gbeta = bn_backward(Go, gooff, P.Y[(name, bi, 2)], 0, o, ooff, P.S[(name, bi, 2)], RELU, Db, 0, cout, gres=Gin, gres_off=goff,
                                    gres_accum=gin_written)
            gin_written = True
        grads[pre + "bn2.weight"], grads[pre + "bn2.bias"] = gg, gbeta
        hmid = P.X[(name, bi, "h")]
        grads[pre + "conv2.weight"] = wgrad_s1(Db, 0, cout, hmid, 0, cout)
        Gh = P.g((name, bi, "h"), lvl, cout)
        dgrad_s1(Db, dense._r64(cout), W, (name, bi, 2), Gh, 0, cout, False)
        # bn1 + ReLU
        Da = P.d((name, bi, 1), lvl, cout)
        gg, gbeta = bn_backward(Gh, 0, P.Y[(name, bi, 1)], 0, hmid, 0, P.S[(name, bi, 1)], RELU, Da, 0, cout)
        grads[pre + "bn1.weight"], grads[pre + "bn1.bias"] = gg, gbeta
        if s_ == 1:
            grads[pre + "conv1.weight"] = wgrad_s1(Da, 0, cout, xin, xoff, ci)
            dgrad_s1(Da, dense._r64(cout), W, (name, bi, 1), Gin, goff, ci, gin_written)
        else:
            grads[pre + "conv1.weight"] = wgrad_box(Da, 0, cout, xin, xoff, ci, 2)
            dgrad_s2(Da, dense._r64(cout), W, (name, bi, 1), Gin, goff, ci, gin_written)
        gin_written = True
        if blk.downsample is not None:
            Dd = P.d((name, bi, "d"), lvl, cout)
            gg, gbeta = bn_backward(P.G[(name, bi, "sc")], 0, P.Y[(name, bi, "d")], 0, None, 0, P.S[(name, bi, "d")], NONE, Dd, 0, cout)
            grads[pre + "downsample.1.weight"], grads[pre + "downsample.1.bias"] = gg, gbeta
            grads[pre + "downsample.0.weight"] = wgrad_1x1_box(Dd, 0, cout, xin, xoff, ci, s_)
            dgrad_s2(Dd, dense._r64(cout), W, (name, bi, "d"), Gin, goff, ci, True, kind=K_1X1S2_DGRAD)


# ---- the ResNet-18 U-Net body ---------------------------------------------------------------------------------------------------------
class _Plan:
    """Every buffer of one (batch, height, width, precision) configuration of an S2DUResNet."""

    def __init__(self, net, n, h, w, dev):
        assert h % 16 == 0 and w % 16 == 0, "S2DUResNet needs H, W = 0 mod 16 (reference limitation, s2d_u_resnet.py)"
        self.n, self.h, self.w = n, h, w
        hs = [h >> i for i in range(5)]
        ws = [w >> i for i in range(5)]
        self.hs, self.ws = hs, ws
        P = lambda lvl, c: pair_act(n, hs[lvl], ws[lvl], c, dev)
        F = lambda lvl, c: f32_act(n, hs[lvl], ws[lvl], c, dev)
        self.P, self.F = P, F
        self.x45 = P(0, 64)                       # im2col of the stems' inputs
        self.cat_last = P(0, 192)                 # [convt1 out 64 | conv1 64 | y 1 | zeros]: the 129 feature channels of the confidence head
        self.cat = {2: P(0, 128), 3: P(1, 192), 4: P(2, 384), 5: P(3, 768)}   # cat_k = [convt_k out | conv_k (encoder stage output)]
        self.c6 = P(4, 512)
        self.Y, self.X, self.S = {}, {}, {}       # fp32 conv results, block-internal activations, BN states
        self.G = {}                               # fp32 gradient buffers (allocated at the first backward)
        self.D = {}                               # (hi, lo) dY operands of the gradient GEMMs
        self.forward_id = 0
        self.head_grad_id = -1                    # forward_id whose confidence-head data gradient is in G["cat_last"]

    def y(self, key, lvl, c):
        if key not in self.Y:
            self.Y[key] = self.F(lvl, dense._r64(c))
        return self.Y[key]

    def x(self, key, lvl, c):
        if key not in self.X:
            self.X[key] = self.P(lvl, dense._r64(c))
        return self.X[key]

    def g(self, key, lvl, c):
        if key not in self.G:
            self.G[key] = self.F(lvl, c)
        return self.G[key]

    def d(self, key, lvl, c):
        if key not in self.D:
            self.D[key] = self.P(lvl, dense._r64(c))
        return self.D[key]


def _plan(net, n, h, w, dev):
    plans = net.__dict__.setdefault("_tc_plans", {})
    key = (n, h, w, dense.PRECISION, str(dev))
    if key not in plans:
        plans.clear()                              # one live configuration per module: these buffers are GBs at batch 16
        plans[key] = _Plan(net, n, h, w, dev)
    return plans[key]


# (stage name in the module, level of its output, channels in, channels out, stride of its first block)
_STAGES = (("conv2", 0, 64, 64, 1), ("conv3", 1, 64, 128, 2), ("conv4", 2, 128, 256, 2), ("conv5", 3, 256, 512, 2))
# (name, level of its OUTPUT, channels in, channels out) of the stride-2 transposed convolutions, deepest first
_CONVT = (("convt5", 3, 512, 256), ("convt4", 2, 768, 128), ("convt3", 1, 384, 64), ("convt2", 0, 192, 64))


def _stems(net, x):
    """[(conv module, bn module, [source planes tensors])] of the first layer (s2d_u_resnet.py:133-143), depth stem first."""
    rgb = x[:, :3]
    if net.in_channels == 3:
        return [(net.conv1_img[0], net.conv1_img[1], [rgb])], None
    if net.unguided:
        c0 = (x[:, 3:4] > 0).float() if net.in_channels == 4 else x[:, 4:5]
        densex, coutx = net.unguidednet(x[:, 3:4], c0)
        d_src = [densex, coutx]
    else:
        d_src = [x[:, 3:4] if net.in_channels == 4 else x[:, 3:5]]
    return [(net.conv1_d[0], net.conv1_d[1], d_src), (net.conv1_img[0], net.conv1_img[1], [rgb])], (d_src if net.unguided else None)


def _planes(t):
    n, c, h, w = t.shape
    if t.stride(3) != 1 or t.stride(2) != w or (c > 1 and t.stride(1) != h * w):
        t = t.contiguous()
    return t, (t.stride(0) if n > 1 else c * h * w)


class S2DUBodyTC(torch.autograd.Function):
    """(rgb-d input planes..., parameters) -> (y [n, oc, h, w], last_layer_input [n, 128, h, w] | None).  Inputs: the source plane
    tensors of the stems (only those of an unguided depth stem receive a gradient), then every parameter of the body in
    ``_param_list`` order."""

    @staticmethod
    def forward(ctx, net, want_lli, nsrc, *args):
        srcs, params = args[:nsrc], args[nsrc:]
        stems = net._tc_stems                       # [(conv, bn, number of its sources)] set by s2du_forward
        n, _, h, w = srcs[0].shape
        dev = srcs[0].device
        P = _plan(net, n, h, w, dev)
        training = net.training
        hs, ws = P.hs, P.ws
        # ---- weights of the whole forward pass: two launches
        wp = WPrep()
        C5 = sum(s.shape[1] for s in srcs)
        w1 = torch.zeros((64, 9, C5), device=dev, dtype=torch.float32)
        ro = co = si = 0
        for conv_m, bn_m, k in stems:
            cs = sum(s.shape[1] for s in srcs[si:si + k])
            w1[ro:ro + conv_m.out_channels, :, co:co + cs] = conv_m.weight.detach().permute(0, 2, 3, 1).reshape(conv_m.out_channels, 9, cs)
            ro, co, si = ro + conv_m.out_channels, co + cs, si + k
        w1e = torch.zeros((64, 64, 3, 3), device=dev, dtype=torch.float32)
        w1e[:, :9 * C5, 1, 1] = w1.reshape(64, 9 * C5)
        wp.fwd("stem", w1e, 64)
        for name, lvl, cin, cout, stride in _STAGES:
            stage_prep_fwd(wp, getattr(net, name), name, cin, cout)
        wp.fwd("conv6", net.conv6[0].weight, 512)
        for name, lvl, cin, cout in _CONVT:
            # the ConvTranspose2d weight [in][out][3][3] IS the weight of the stride-2 convolution this layer is the data gradient of
            # ([Cout_conv = in][Cin_conv = out]): prep mode 2 (transposed, taps kept) gives rows = out chunk, columns = in
            wp.quad_s2(name, getattr(net, name)[0].weight, cin)
        wp.transposed("convt1", net.convt1[0].weight, 128, flip=True)            # Wc[o][i][tap] = Wt[i][o][8-tap]: prep mode 1
        W = wp.run()
        # ---- stems: im2col + ONE GEMM, then BatchNorm + LeakyReLU per stem into cat_last[64:128] (= conv1 of the reference)
        pl = [_planes(s) for s in srcs] + [(None, 0)] * (3 - len(srcs))
        ch = [s.shape[1] for s in srcs] + [0] * (3 - len(srcs))
        check(lib.adc_im2col3x3_pack(pl[0][0].data_ptr(), ch[0], pl[0][1], _ptr(pl[1][0]), ch[1], pl[1][1], _ptr(pl[2][0]), ch[2], pl[2][1],
                                     P.x45.hi.data_ptr(), _ptr(P.x45.lo), P.x45.cs, n, h, w, 1, dense._stream()), "adc_im2col3x3_pack")
        Y1 = P.y("stem", 0, 64)
        conv(K_1X1, P.x45, 0, 64, W[("stem", 0)], Y1, 0, 64, h, w)
        ro = 0
        for i, (conv_m, bn_m, k) in enumerate(stems):
            P.S[("stem", i)] = bn_forward(Y1, ro, conv_m.out_channels, bn_m, training, P.cat_last, 64 + ro, LEAKY)
            ro += conv_m.out_channels
        # ---- encoder: four stages of two BasicBlocks; stage output lands in the decoder's concat buffer
        xin, xoff = P.cat_last, 64
        for name, lvl, cin, cout, stride in _STAGES:
            dst, doff = P.cat[int(name[-1])], {2: 64, 3: 64, 4: 128, 5: 256}[int(name[-1])]
            stage_forward(P, getattr(net, name), name, lvl, cin, cout, stride, xin, xoff, dst, doff, W, training)
            xin, xoff = dst, doff
        # ---- bottleneck and decoder
        Y6 = P.y("conv6", 4, 512)
        fwd_box(P.cat[5], 256, 512, W, "conv6", Y6, 0, 512, 2)
        P.S["conv6"] = bn_forward(Y6, 0, 512, net.conv6[1], training, P.c6, 0, LEAKY)
        xin, cin_prev = P.c6, 512
        for name, lvl, cin, cout in _CONVT:
            Yt = P.y(name, lvl, cout)
            fwd_convt2(xin, 0, cin, W, name, Yt, 0, cout)
            dst = P.cat[int(name[-1])]
            P.S[name] = bn_forward(Yt, 0, cout, getattr(net, name)[1], training, dst, 0, LEAKY)
            xin = dst
        Yt1 = P.y("convt1", 0, 64)
        fwd_s1(P.cat[2], 0, 128, W, "convt1", Yt1, 0, 64)
        P.S["convt1"] = bn_forward(Yt1, 0, 64, net.convt1[1], training, P.cat_last, 0, LEAKY)
        oc = net.convtf[0].weight.shape[0]
        y = dot_head(P.cat_last, 0, 128, net.convtf[0].weight, net.convtf[0].bias, y_act=P.cat_last if oc == 1 else None, y_off=128)
        # want_lli: 1 = materialise last_layer_input as an NCHW tensor; 2 = hand the NHWC buffer to the confidence head through a token
        # (no NCHW round trip: the head reads cat_last[0:129] in place and leaves its input gradient in this plan's gradient buffer)
        lli = dense.to_nchw_f32(P.cat_last, 128) if want_lli == 1 else (torch.zeros(1, device=dev) if want_lli == 2 else None)
        P.forward_id += 1
        ctx.net, ctx.plan, ctx.fid, ctx.nsrc, ctx.training, ctx.want_lli = net, P, P.forward_id, nsrc, training, want_lli
        ctx.src_ch = [s.shape[1] for s in srcs]
        ctx.save_for_backward(*params)
        return y, lli

    @staticmethod
    def backward(ctx, g_y, g_lli):
        net, P = ctx.net, ctx.plan
        if P.forward_id != ctx.fid:
            raise RuntimeError("S2DUBodyTC: the module ran another forward before this backward; its saved activations are gone")
        params = ctx.saved_tensors
        names = _param_names(net)
        grads = {}
        stems = net._tc_stems
        n, h, w = P.n, P.h, P.w
        dev = g_y.device
        # ---- weights of every data gradient: transposed views of the same Parameters
        wp = WPrep()
        for name, lvl, cin, cout, stride in _STAGES:
            stage_prep_bwd(wp, getattr(net, name), name, cin, cout, stride)
        wp.quad_s2("conv6", net.conv6[0].weight, 512)
        for name, lvl, cin, cout in _CONVT:
            wp.fwd(name, getattr(net, name)[0].weight, dense._r64(cout))       # ConvT weight [in][out] read as a convolution [Cout = in][Cin = out]
        wp.fwd("convt1", net.convt1[0].weight, 64)
        W = wp.run()
        # ---- head: 1x1 convolution + the features' gradient
        Gl = P.g("cat_last", 0, 192)
        head_wrote = ctx.want_lli == 2 and P.head_grad_id == ctx.fid     # the confidence head's data gradient already sits in Gl[0:129]
        if g_lli is not None and ctx.want_lli == 1:
            nchw_into(g_lli, Gl, 0, False)
        if g_y is None:
            g_y = torch.zeros((n, net.convtf[0].weight.shape[0], h, w), device=Gl.hi.device)
        have = (g_lli is not None and ctx.want_lli == 1) or head_wrote
        gw, gb = dot_head_backward(P.cat_last, 0, 128, net.convtf[0].weight, g_y, Gl, 0, have, gy_extra=(Gl, 128) if head_wrote else None)
        grads["convtf.0.weight"], grads["convtf.0.bias"] = gw, gb
        # ---- convt1 (stride-1 transposed convolution = convolution with flipped weights)
        D = P.d("convt1", 0, 64)
        gg, gbeta = bn_backward(Gl, 0, P.Y["convt1"], 0, P.cat_last, 0, P.S["convt1"], LEAKY, D, 0, 64)
        grads["convt1.1.weight"], grads["convt1.1.bias"] = gg, gbeta
        dwc = wgrad_s1(D, 0, 64, P.cat[2], 0, 128)                                # gradient of Wc[o][i][u][v] = Wt[i][o][2-u][2-v]
        grads["convt1.0.weight"] = dwc.flip(2, 3).transpose(0, 1).contiguous()
        G2 = P.g(("cat", 2), 0, 128)
        dgrad_s1(D, 64, W, "convt1", G2, 0, 128, False)
        # ---- stride-2 transposed convolutions, shallowest first (reverse of the forward)
        Gk = {2: G2}
        for name, lvl, cin, cout in reversed(_CONVT):
            k = int(name[-1])
            src = P.c6 if name == "convt5" else P.cat[k + 1]
            D = P.d(name, lvl, cout)
            gg, gbeta = bn_backward(Gk[k], 0, P.Y[name], 0, P.cat[k], 0, P.S[name], LEAKY, D, 0, cout)
            grads[name + ".1.weight"], grads[name + ".1.bias"] = gg, gbeta
            # dWt[ci][co][u][v] = sum x_in[i,j][ci] * dY[2i+u-1, 2j+v-1][co]: the weight gradient of a stride-2 convolution whose "output
            # gradient" is the layer's INPUT and whose "input" is dY
            grads[name + ".0.weight"] = wgrad_box(src, 0, cin, D, 0, cout, 2)
            if name == "convt5":
                Gsrc = P.g("c6", 4, 512)
            else:
                Gsrc = P.g(("cat", k + 1), lvl + 1, cin)
                Gk[k + 1] = Gsrc
            dgrad_convt2(D, dense._r64(cout), W, name, Gsrc, 0, cin, False)
        # ---- conv6
        D = P.d("conv6", 4, 512)
        gg, gbeta = bn_backward(P.G["c6"], 0, P.Y["conv6"], 0, P.c6, 0, P.S["conv6"], LEAKY, D, 0, 512)
        grads["conv6.1.weight"], grads["conv6.1.bias"] = gg, gbeta
        grads["conv6.0.weight"] = wgrad_box(D, 0, 512, P.cat[5], 256, 512, 2)
        dgrad_s2(D, 512, W, "conv6", Gk[5], 256, 512, True)
        # ---- encoder stages, deepest first
        for name, lvl, cin, cout, stride in reversed(_STAGES):
            k = int(name[-1])
            doff = {2: 64, 3: 64, 4: 128, 5: 256}[k]
            if k == 2:
                xin0, xoff0, Gin0, goff0 = P.cat_last, 64, Gl, 64
            else:
                xin0, xoff0, Gin0, goff0 = P.cat[k - 1], {3: 64, 4: 64, 5: 128}[k], Gk[k - 1], {3: 64, 4: 64, 5: 128}[k]
            stage_backward(P, getattr(net, name), name, lvl, cin, cout, stride, xin0, xoff0, Gin0, goff0, True, P.cat[k], doff, Gk[k], doff, W,
                           grads, name + ".")
        # ---- stems
        D1 = P.d("stem", 0, 64)
        ro = 0
        for i, (conv_m, bn_m, k) in enumerate(stems):
            c = conv_m.out_channels
            gg, gbeta = bn_backward(Gl, 64 + ro, P.Y["stem"], ro, P.cat_last, 64 + ro, P.S[("stem", i)], LEAKY, D1, ro, c)
            pre = "conv1_d" if (len(stems) == 2 and i == 0) else "conv1_img"
            grads[pre + ".1.weight"], grads[pre + ".1.bias"] = gg, gbeta
            ro += c
        C5 = sum(ctx.src_ch)
        dw1 = dense.conv1x1_wgrad_tc(D1, P.x45, 9 * C5, 64, cin_pad=dense._r16(9 * C5)).reshape(64, 9, C5)
        ro = co = si = 0
        src_grads = [None] * ctx.nsrc
        for i, (conv_m, bn_m, k) in enumerate(stems):
            cs = sum(ctx.src_ch[si:si + k])
            c = conv_m.out_channels
            pre = "conv1_d" if (len(stems) == 2 and i == 0) else "conv1_img"
            grads[pre + ".0.weight"] = dw1[ro:ro + c, :, co:co + cs].permute(0, 2, 1).reshape(c, cs, 3, 3).contiguous()
            if any(ctx.needs_input_grad[3 + si + j] for j in range(k)):
                # data gradient of this stem: a stride-1 convolution of its slice of dY with the flipped weights
                wfull = torch.zeros((64, cs, 3, 3), device=dev, dtype=torch.float32)
                wfull[ro:ro + c] = conv_m.weight.detach()
                wq = dense.prep_conv3x3_weights_batch([(wfull, 1, 64, 0, cs)])[0]
                gin = f32_act(n, h, w, 16, dev)
                conv(K_S1, D1, 0, 64, wq, gin, 0, cs, h, w)
                planes = dense.to_nchw_f32(gin, cs)
                o = 0
                for j in range(k):
                    src_grads[si + j] = planes[:, o:o + ctx.src_ch[si + j]]
                    o += ctx.src_ch[si + j]
            ro, co, si = ro + c, co + cs, si + k
        return (None, None, None, *src_grads, *[grads.get(nm) for nm in names])


def _param_names(net):
    """Names (relative to the S2DUResNet module) of the parameters the body function takes, in order."""
    return [k for k, _ in net.named_parameters() if not k.startswith("unguidednet.")]


def s2du_forward(net, x, build_conf_input):
    """S2DUResNet.forward (s2d_u_resnet.py:132-190) on the tensor-core path."""
    stems, _ = _stems(net, x)
    srcs = []
    net._tc_stems = []
    for conv_m, bn_m, src in stems:
        net._tc_stems.append((conv_m, bn_m, len(src)))
        srcs += src
    names = _param_names(net)
    pd = dict(net.named_parameters())
    want = bool(build_conf_input) and net.out_channels == 1
    lazy = want and getattr(net, "_adc_lazy_features", False)
    y, lli = S2DUBodyTC.apply(net, (2 if lazy else 1) if want else 0, len(srcs), *srcs, *[pd[k] for k in names])
    features = None
    if build_conf_input:
        if lazy:
            features = LazyFeatures(lli, net, 129)
        else:
            features = torch.cat((lli, y), 1) if want else torch.sigmoid(y[:, 1:2, :, :])
    depth = y[:, 0:1, :, :] if net.training else torch.relu(y[:, 0:1, :, :])
    return depth, features


# ---- cbr3-c1 confidence head, any channel count -----------------------------------------------------------------------------------------
class LazyFeatures:
    """The feature tensor of a depth-completion net (s2d_u_resnet.py:179: cat(last_layer_input, y), 129 channels) left where it was
    computed: channels [0, 129) of the net's NHWC (hi, lo) concat buffer.  ``CBR3C1Confidence`` consumes it in place; autograd sees the
    1-element ``token``.  Handed out only when the framework model asks for it (``net._adc_lazy_features``), so a direct caller of
    ``S2DUResNet.forward`` still receives the NCHW tensor of the reference."""

    def __init__(self, token, net, channels):
        self.token, self.net, self.channels = token, net, channels

    def size(self):
        P = _current_plan(self.net)
        return torch.Size((P.n, self.channels, P.h, P.w))


def _current_plan(net):
    return next(iter(net.__dict__["_tc_plans"].values()))


def _head_forward(X, c, head, w1, w2, b2):
    """Shared core: X = (hi, lo) activation holding the c feature channels at offset 0 (contraction width round16(c))."""
    n, h, w, dev = X.n, X.h, X.w, X.hi.device
    bn = head.bn1
    cp, cs = dense._r16(c), dense._r64(c)
    step = 80 if dense._split() else 128               # the row-segment kernel takes 80 output channels in split mode: 129 = 80 + 49
    wp = WPrep()
    wp.fwd("c1", w1, cp, step)
    W = wp.run()
    Y = f32_act(n, h, w, cs, dev)
    fwd_s1(X, 0, cp, W, "c1", Y, 0, c, step=step)
    Z = pair_act(n, h, w, cs, dev)
    state = bn_forward(Y, 0, c, bn, bn.training, Z, 0, LEAKY)
    conf = torch.sigmoid(dot_head(Z, 0, c, w2, b2))
    return conf, (Y, Z, state)


def _head_backward(X, c, saved, conf, gconf, w1, w2, Gx):
    """Returns (gw1, dgamma, dbeta, gw2, gb2); with ``Gx`` (fp32 NHWC) the input gradient is WRITTEN into its channels [0, c)."""
    Y, Z, state = saved
    n, h, w, dev = X.n, X.h, X.w, X.hi.device
    cs, cp = dense._r64(c), dense._r16(c)
    gs = (gconf * conf * (1.0 - conf)).contiguous()
    Gz = f32_act(n, h, w, cs, dev)
    gw2, gb2 = dot_head_backward(Z, 0, c, w2, gs, Gz, 0, False)
    D = pair_act(n, h, w, cs, dev)
    gg, gbeta = bn_backward(Gz, 0, Y, 0, Z, 0, state, LEAKY, D, 0, c)
    gw1 = wgrad_s1(D, 0, c, X, 0, c)
    if Gx is not None:
        step = 80 if dense._split() else 128
        wp = WPrep()
        wp.transposed("c1", w1, cp, flip=True, step=step)
        W = wp.run()
        dgrad_s1(D, cp, W, "c1", Gx, 0, c, False, step=step)
    return gw1, gg, gbeta, gw2, gb2


class ConfHeadBN(torch.autograd.Function):
    """conv3x3 (no bias) -> BatchNorm -> LeakyReLU(0.2) -> conv1x1 -> sigmoid (confidence_depth_framework.py:299-329) for channel counts
    beyond tc_train.ConfHeadTC's fused 80-channel path (129 for the S2DU features): the 3x3 convolution in output-channel chunks with a
    K tail (129 channels = 2 blocks of 64 + ONE 16-channel step), general BatchNorm passes, the 1x1 as the dot-product kernel."""

    @staticmethod
    def forward(ctx, feat, w1, gamma, beta, w2, b2, head):
        n, c, h, w = feat.shape
        X = dense.to_nhwc_bf16(feat, cpad=dense._r64(c), pad=True)
        conf, saved = _head_forward(X, c, head, w1, w2, b2)
        ctx.X, ctx.saved, ctx.c = X, saved, c
        ctx.save_for_backward(conf, w1, w2)
        return conf

    @staticmethod
    def backward(ctx, gconf):
        conf, w1, w2 = ctx.saved_tensors
        X, c = ctx.X, ctx.c
        Gx = f32_act(X.n, X.h, X.w, X.cs, conf.device) if ctx.needs_input_grad[0] else None
        gw1, gg, gbeta, gw2, gb2 = _head_backward(X, c, ctx.saved, conf, gconf, w1, w2, Gx)
        gfeat = dense.to_nchw_f32(Gx, c) if Gx is not None else None
        ctx.X = ctx.saved = None
        return gfeat, gw1, gg, gbeta, gw2, gb2, None


class ConfHeadLazy(torch.autograd.Function):
    """The same head reading its features in place from the depth-completion net's NHWC buffer (LazyFeatures); its input gradient goes
    straight into that net's fp32 gradient buffer, where the net's own backward (which autograd runs afterwards) accumulates onto it."""

    @staticmethod
    def forward(ctx, token, w1, gamma, beta, w2, b2, head, feats):
        P = _current_plan(feats.net)
        conf, saved = _head_forward(P.cat_last, feats.channels, head, w1, w2, b2)
        ctx.plan, ctx.fid, ctx.saved, ctx.c = P, P.forward_id, saved, feats.channels
        ctx.save_for_backward(conf, w1, w2)
        return conf

    @staticmethod
    def backward(ctx, gconf):
        conf, w1, w2 = ctx.saved_tensors
        P = ctx.plan
        if P.forward_id != ctx.fid:
            raise RuntimeError("ConfHeadLazy: the depth-completion net ran another forward before this backward; its features are gone")
        Gx = None
        if ctx.needs_input_grad[0]:
            Gx = P.g("cat_last", 0, 192)
            P.head_grad_id = ctx.fid
        gw1, gg, gbeta, gw2, gb2 = _head_backward(P.cat_last, ctx.c, ctx.saved, conf, gconf, w1, w2, Gx)
        ctx.saved = None
        return (torch.zeros_like(conf[:1, 0, 0, :1]).reshape(1) if Gx is not None else None), gw1, gg, gbeta, gw2, gb2, None, None


def conf_head(head, feat):
    p = (head.conv1.weight, head.bn1.weight, head.bn1.bias, head.conv2.weight, head.conv2.bias)
    if isinstance(feat, LazyFeatures):
        return ConfHeadLazy.apply(feat.token, *p, head, feat)
    return ConfHeadBN.apply(feat, *p, head)

"""Host side of the tensor-core path of the dense (wide-channel) layers around the NConv pyramid: the arithmetic-mode switch, the
NHWC activation container and thin typed wrappers of the C-ABI entry points (layout conversions, weight preparation, the tcgen05
convolution / weight-gradient / glue kernels).  The nets' modules are plain ``nn.Conv2d`` / ``nn.BatchNorm2d`` parameter containers
(identical ``state_dict`` keys); in the modes 'bf16x3' / 'bf16' their parameters are consumed by tc_train / tc_gms / tc_bn / tc_s2d, in
'fp32' / 'tf32' the modules' own ATen forward runs (``BACKEND``: the measurement baseline).  Nothing here runs on the CPU.
"""
import os

import torch
import torch.nn.functional as F

BACKEND = "cudnn"
PRECISION = "bf16x3"


def set_precision(mode):
    """'bf16x3' (DEFAULT) / 'bf16': the dense layers run on the hand-written tcgen05 kernels (NHWC activations, fp32 accumulation
    in TMEM) with hi+lo bf16 operand pairs (fp32-grade results, the 1e-4 parity mode) or plain bf16 operands (the 1e-2 mode).
    'fp32' / 'tf32' route the dense layers to the ATen/cuDNN library instead -- the baseline the kernels are measured against, kept
    as a debugging aid; the NConv pyramid, losses, metrics and optimizer are this library's kernels in every mode."""
    global PRECISION
    if mode not in ("fp32", "tf32", "bf16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'tf32', 'bf16' or 'bf16x3'")
    PRECISION = mode
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"


set_precision("bf16x3")      # the B200 path is the default: hand-written tcgen05 kernels, fp32-grade results (hi/lo split operands)


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("dense layer input must be a CUDA tensor: the B200 path has no CPU fallback")


# ---- tensor-core path (tcgen05 implicit GEMM, NHWC bf16) ---------------------------------------------------------------
# An activation (class Act) is a pair (hi, lo) of NHWC bf16 tensors; lo is None in the plain 'bf16' mode.  In the 'bf16x3'
# mode every value is hi + lo (hi = bf16(v), lo = bf16(v - hi), 16 mantissa bits together) and the kernel accumulates
# hi*hi + lo*hi + hi*lo in fp32: fp32-grade results from the bf16 tensor cores at three MMAs per product.
# pad=True: the tensor carries a one-pixel zero border ([N][H+2][W+2][C]), the layout of the v2 (tap-reuse) kernel.
USE_V2 = True


def _r16(c):
    return (c + 15) // 16 * 16


# Output channels per row-segment launch.  In the split mode a CTA pair holds one weight plane per CTA, so 128-channel chunks fit the way
# they do in the plain bf16 mode; ADC_CHUNK128=0 restores the 64-channel chunks of the single-CTA form.
CHUNK128 = os.environ.get("ADC_CHUNK128", "1") != "0"


def out_chunk():
    return 128 if (PRECISION != "bf16x3" or CHUNK128) else 64


def _stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _split():
    return PRECISION == "bf16x3"


def _ptr(t):
    return t.data_ptr() if t is not None else None


class Act:
    """NHWC activation: hi (+ lo) tensors of logical size [n, h, w, cs]; stored [n, h+2, w+2, cs] when pad."""
    __slots__ = ("hi", "lo", "n", "h", "w", "cs", "pad")

    def __init__(self, hi, lo, n, h, w, cs, pad):
        self.hi, self.lo, self.n, self.h, self.w, self.cs, self.pad = hi, lo, n, h, w, cs, pad


def _r64(c):
    return (c + 63) // 64 * 64


def new_act(n, h, w, c, device, pad=False, zero=False, dtype=torch.bfloat16, split=None, border_only=False):
    """zero=True clears everything; border_only clears just the one-pixel border of a padded tensor (the v1 kernel and the
    layout conversion write the interior, the v2 kernel writes interior AND border itself)."""
    split = _split() if split is None else split
    mk = torch.zeros if zero else torch.empty
    shape = (n, h + 2, w + 2, c) if pad else (n, h, w, c)
    ts = [mk(shape, device=device, dtype=dtype)]
    if split and dtype == torch.bfloat16:
        ts.append(mk(shape, device=device, dtype=dtype))
    if pad and border_only and not zero:
        for t in ts:
            t[:, 0].zero_()
            t[:, -1].zero_()
            t[:, :, 0].zero_()
            t[:, :, -1].zero_()
    return Act(ts[0], ts[1] if len(ts) > 1 else None, n, h, w, c, pad)


def to_nhwc_bf16(x, cpad=None, dst=None, dst_off=0, pad=False):
    """NCHW fp32 (any batch stride, contiguous planes) -> NHWC bf16 activation with the channel count padded to a multiple
    of 16; optionally into channels [dst_off, dst_off+cpad) of an existing activation (folds torch.cat)."""
    from ._lib import lib, check
    _need_cuda(x)
    n, c, h, w = x.shape
    if x.stride(3) != 1 or x.stride(2) != w or (c > 1 and x.stride(1) != h * w):
        x = x.contiguous()
    cpad = _r16(c) if cpad is None else cpad
    if dst is None:
        dst = new_act(n, h, w, cpad, x.device, pad=pad, border_only=True)
    check(lib.adc_nchw_to_nhwc_bf16(x.data_ptr(), dst.hi.data_ptr(), _ptr(dst.lo), n, c, cpad, h, w,
                                    x.stride(0) if n > 1 else c * h * w, dst.cs, dst_off, int(dst.pad), _stream()),
          "adc_nchw_to_nhwc_bf16")
    return dst


def to_nchw_f32(act, c, off=0):
    """Channel slice [off, off+c) of an NHWC activation -> NCHW fp32 (hi + lo in the split mode)."""
    from ._lib import lib, check
    out = torch.empty((act.n, c, act.h, act.w), device=act.hi.device, dtype=torch.float32)
    check(lib.adc_nhwc_to_nchw_f32(act.hi.data_ptr(), int(act.hi.dtype == torch.float32), out.data_ptr(), act.n, c, act.h, act.w,
                                   act.cs, off, int(act.pad), _stream()), "adc_nhwc_to_nchw_f32")
    if act.lo is not None:
        low = torch.empty_like(out)
        check(lib.adc_nhwc_to_nchw_f32(act.lo.data_ptr(), 0, low.data_ptr(), act.n, c, act.h, act.w, act.cs, off, int(act.pad),
                                       _stream()), "adc_nhwc_to_nchw_f32")
        out += low
    return out


def prep_conv3x3_weights(weight, flip_transpose=False, cin_pad=None):
    """fp32 [Cout,Cin,3,3] -> (hi, lo) bf16 [9][round16(Cout)][cin_pad] as the tensor-core kernels read them (input channels
    zero padded to ``cin_pad``, default round16(Cin))."""
    from ._lib import lib, check
    co, ci = weight.shape[:2]
    oc, ic = (ci, co) if flip_transpose else (co, ci)
    cin_pad = _r16(ic) if cin_pad is None else cin_pad
    hi = torch.empty((9, _r16(oc), cin_pad), device=weight.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if _split() else None
    check(lib.adc_conv3x3_prep_weights(weight.detach().contiguous().data_ptr(), hi.data_ptr(), _ptr(lo), co, ci, cin_pad,
                                       int(flip_transpose), _stream()), "adc_conv3x3_prep_weights")
    return (hi, lo)


def prep_conv3x3_weights_batch(reqs):
    """Several prep_conv3x3_weights in ONE launch.  reqs: (weight [Cout,Cin,3,3], mode, cin_pad, ci_off, ci_cnt) with mode 0 =
    forward weights of input channels [ci_off, ci_off+ci_cnt), 1 / 2 = transposed data-gradient weights (taps flipped / kept)
    producing those channels.  Returns [(hi, lo)] in order."""
    import ctypes
    from ._lib import lib, check, WPrepItem, ADC_WPREP_MAX
    outs = []
    for k0 in range(0, len(reqs), ADC_WPREP_MAX):
        chunk = reqs[k0:k0 + ADC_WPREP_MAX]
        items = (WPrepItem * len(chunk))()
        for it, (weight, mode, cin_pad, ci_off, ci_cnt) in zip(items, chunk):
            w = weight.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            co, ci = w.shape[:2]
            oc = ci_cnt if mode else co
            hi = torch.empty((9, _r16(oc), cin_pad), device=w.device, dtype=torch.bfloat16)
            lo = torch.empty_like(hi) if _split() else None
            it.w, it.out, it.out_lo = w.data_ptr(), hi.data_ptr(), _ptr(lo)
            it.cout, it.cin, it.ci_off, it.ci_cnt, it.cin_pad, it.mode = co, ci, ci_off, ci_cnt, cin_pad, mode
            outs.append((hi, lo, w))            # keep the (possibly copied) source alive until the launch is enqueued
        check(lib.adc_conv3x3_prep_weights_batch(ctypes.cast(items, ctypes.c_void_p), len(chunk), _stream()),
              "adc_conv3x3_prep_weights_batch")
    return [(h, l) for h, l, _ in outs]


def conv3x3_tc(act, cin, w_prep, bias, cout, stride=1, act_fn=0, x_off=0, out=None, out_off=0, out_f32=False, out_pad=None,
               out_zero_to=0, tap_mask=0):
    """3x3 / pad 1 convolution on the tensor cores.  Channels [x_off, x_off+cin) of ``act`` are read (cin a multiple of 16,
    zero padded); the result goes to channels [out_off, out_off+cout) of ``out`` (allocated if None)."""
    from ._lib import lib, check
    n, h, w = act.n, act.h, act.w
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    # v2 streams 3 taps x Cout16 weight rows per pipeline stage: in the split mode two stages fit only up to 80 output channels
    v2 = USE_V2 and stride == 1 and act.pad and cin % 16 == 0 and cout <= (80 if (_split() and not CHUNK128) else 128) and (out is None or out.pad)
    assert v2 or not tap_mask, "tap masks are a feature of the row-segment kernel"
    zero_to = out_zero_to if v2 else 0                      # slice width the layer owns (channels past cout written as 0)
    if out is None:
        out_pad = act.pad if out_pad is None else out_pad
        v2 = v2 and out_pad
        cs = _r64(cout) if v2 else _r16(cout)               # v2 chains want 64-channel blocks
        out = new_act(n, ho, wo, cs, act.hi.device, pad=out_pad, zero=(not v2) and cs != cout, border_only=not v2,
                      dtype=torch.float32 if out_f32 else torch.bfloat16)
        zero_to = cs                                         # v2 writes the pad channels (and the border) as zeros itself
    b = bias.detach().contiguous().data_ptr() if bias is not None else None
    is_f32 = int(out.hi.dtype == torch.float32)
    if v2:
        check(lib.adc_conv3x3_s1_tc_v2(act.hi.data_ptr(), _ptr(act.lo), act.cs, x_off, w_prep[0].data_ptr(), _ptr(w_prep[1]), b,
                                       out.hi.data_ptr(), _ptr(out.lo), is_f32, out.cs, out_off, zero_to, n, h, w, cin, cout, act_fn,
                                       tap_mask, _stream()), "adc_conv3x3_s1_tc_v2")
    else:
        check(lib.adc_conv3x3_fprop_tc(act.hi.data_ptr(), _ptr(act.lo), act.cs, x_off, w_prep[0].data_ptr(), _ptr(w_prep[1]), b,
                                       out.hi.data_ptr(), _ptr(out.lo), is_f32, out.cs, out_off, n, h, w, cin, cout, stride, act_fn,
                                       int(act.pad), int(out.pad), _stream()), "adc_conv3x3_fprop_tc")
    return out


def conv3x3_wgrad_tc(gy, x, cin, cout, x_off=0, gy_off=0, cin_pad=None):
    """Weight gradient [cout, cin, 3, 3] (fp32) of a 3x3 / stride-1 convolution from zero-bordered NHWC activations: ``x``
    (channels [x_off, x_off+cin_pad)) and the output gradient ``gy`` (channels [gy_off, gy_off+cout), zero border)."""
    from ._lib import lib, check
    assert x.pad and gy.pad and (x.n, x.h, x.w) == (gy.n, gy.h, gy.w)
    cin_pad = _r16(cin) if cin_pad is None else cin_pad
    dw = torch.empty((cout, cin, 3, 3), device=x.hi.device, dtype=torch.float32)
    ws = torch.empty(lib.adc_conv3x3_wgrad_tc_workspace_bytes(cin_pad, cout) // 4, device=x.hi.device, dtype=torch.float32)
    check(lib.adc_conv3x3_wgrad_tc(gy.hi.data_ptr(), _ptr(gy.lo), gy.cs, gy_off, x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off,
                                   dw.data_ptr(), ws.data_ptr(), x.n, x.h, x.w, cin, cin_pad, cout, _stream()), "adc_conv3x3_wgrad_tc")
    return dw


def conv3x3_wgrad_box_tc(gy, x, cin, cout, stride, x_off=0, gy_off=0, cin_pad=None):
    """Weight gradient of a 3x3 / pad-1 convolution of any stride (1 or 2) from NHWC activations with or without border."""
    from ._lib import lib, check
    assert x.n == gy.n and gy.h == (x.h - 1) // stride + 1 and gy.w == (x.w - 1) // stride + 1
    cin_pad = _r16(cin) if cin_pad is None else cin_pad
    dw = torch.empty((cout, cin, 3, 3), device=x.hi.device, dtype=torch.float32)
    ws = torch.empty(lib.adc_conv3x3_wgrad_tc_workspace_bytes(cin_pad, cout) // 4, device=x.hi.device, dtype=torch.float32)
    check(lib.adc_conv3x3_wgrad_box_tc(gy.hi.data_ptr(), _ptr(gy.lo), gy.cs, gy_off, int(gy.pad), x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off,
                                       int(x.pad), dw.data_ptr(), ws.data_ptr(), x.n, x.h, x.w, stride, cin, cin_pad, cout, _stream()),
          "adc_conv3x3_wgrad_box_tc")
    return dw


def conv3x3_dgrad_s2_tc(gy, cgy, w_prep, cgx, h, w, gy_off=0, out=None, out_off=0, out_f32=False, out_pad=True):
    """Data gradient of a 3x3 / stride-2 / pad-1 convolution whose input was [n, h, w]: ``gy`` channels [gy_off, gy_off+cgy)
    -> channels [out_off, out_off+cgx) of ``out``; ``w_prep`` = prep_conv3x3_weights(weight, flip_transpose=2, cin_pad=cgy)."""
    from ._lib import lib, check
    if out is None:
        out = new_act(gy.n, h, w, _r16(cgx), gy.hi.device, pad=out_pad, zero=True, dtype=torch.float32 if out_f32 else torch.bfloat16)
    is_f32 = int(out.hi.dtype == torch.float32)
    check(lib.adc_conv3x3_dgrad_s2_tc(gy.hi.data_ptr(), _ptr(gy.lo), gy.cs, gy_off, w_prep[0].data_ptr(), _ptr(w_prep[1]),
                                      out.hi.data_ptr(), _ptr(out.lo), is_f32, out.cs, out_off, gy.n, h, w, cgy, cgx, int(gy.pad),
                                      int(out.pad), _stream()), "adc_conv3x3_dgrad_s2_tc")
    return out


def upsample_into(src, c, dst, dst_off, src_off=0):
    """Nearest up-sampling of channels [src_off, src_off+c) of ``src`` into channels [dst_off, dst_off+c) of the (larger)
    zero-bordered activation ``dst``: F.interpolate + torch.cat of the decoder in one pass."""
    from ._lib import lib, check
    assert src.pad and dst.pad and src.n == dst.n
    check(lib.adc_nhwc_upsample_nearest(src.hi.data_ptr(), _ptr(src.lo), src.cs, src_off, dst.hi.data_ptr(), _ptr(dst.lo), dst.cs, dst_off,
                                        src.n, c, src.h, src.w, dst.h, dst.w, _stream()), "adc_nhwc_upsample_nearest")
    return dst


def sum_rows(part):
    """Sum over dim 0 of a contiguous fp32 tensor [rows, ...] in a fixed order (second stage of the two-stage reductions)."""
    from ._lib import lib, check
    rows = part.shape[0]
    out = torch.empty(part.shape[1:], device=part.device, dtype=torch.float32)
    check(lib.adc_sum_rows(part.data_ptr(), rows, out.numel(), out.data_ptr(), _stream()), "adc_sum_rows")
    return out


def grad_assemble(out, c, out_off=0, a=None, a_off=0, b=None, b_off=0, act=None, act_off=0, slope=0.0, colsum=False, a_quad=False,
                  reduce=None):
    """out[:, out_off:out_off+c] = d(act) * (sum over the up-sampling children of ``a`` + ``b``); see adc_nhwc_grad_assemble.
    a_quad: ``a`` is the quad form of a stride-2 data gradient (4 parity blocks of c channels at half resolution).
    colsum=True also returns the per-channel sum of the result over all pixels (fp32 [c]: a bias gradient) from the same pass;
    ``reduce``: callable that sums the per-CTA partial rows (default ``sum_rows`` on the current stream; a caller may move this
    second stage off its critical path)."""
    from ._lib import lib, check
    z = lambda t: (t.hi.data_ptr(), _ptr(t.lo), t.cs) if t is not None else (None, None, 0)
    ah, al, ast = z(a)
    bh, bl, bst = z(b)
    part = torch.empty((lib.adc_nhwc_grad_assemble_parts(), c), device=out.hi.device, dtype=torch.float32) if colsum else None
    check(lib.adc_nhwc_grad_assemble(ah, al, ast, a_off, a.h if a is not None else 0, a.w if a is not None else 0, bh, bl, bst, b_off,
                                     act.hi.data_ptr() if act is not None else None, act.cs if act is not None else 0, act_off,
                                     float(slope), out.hi.data_ptr(), _ptr(out.lo), out.cs, out_off, out.n, c, out.h, out.w, _ptr(part),
                                     int(a_quad), _stream()), "adc_nhwc_grad_assemble")
    return (out, (reduce or sum_rows)(part)) if colsum else out


def colsum(g, c, off=0):
    """Per-channel sum over all pixels of channels [off, off+c) of a zero-bordered activation (fp32 [c]): the bias gradient."""
    from ._lib import lib, check
    out = torch.empty(c, device=g.hi.device, dtype=torch.float32)
    ws = torch.empty(lib.adc_nhwc_colsum_workspace_bytes(c) // 4, device=g.hi.device, dtype=torch.float32)
    check(lib.adc_nhwc_colsum(g.hi.data_ptr(), _ptr(g.lo), g.cs, off, g.n, c, g.h, g.w, out.data_ptr(), ws.data_ptr(), _stream()),
          "adc_nhwc_colsum")
    return out


def ged_head_fwd(x9, x0d, weight, bias, h, w):
    """conv10 of GEDNet + LeakyReLU(0.2) from x9 (zero-bordered NHWC activation, 32 channels, its own resolution) and x0_d
    (NCHW fp32 [n,1,h,w]): [n, oc, h, w] fp32."""
    from ._lib import lib, check
    oc = weight.shape[0]
    out = torch.empty((x9.n, oc, h, w), device=x0d.device, dtype=torch.float32)
    check(lib.adc_ged_head_fwd(x9.hi.data_ptr(), _ptr(x9.lo), x9.cs, x9.h, x9.w, x0d.data_ptr(), weight.detach().contiguous().data_ptr(),
                               bias.detach().contiguous().data_ptr(), out.data_ptr(), x9.n, oc, h, w, _stream()), "adc_ged_head_fwd")
    return out


def ged_head_bwd(x9, x0d, weight, gy, x10, glli, g9, need_x0d):
    """Backward of ged_head_fwd: fills channels [0,32) of ``g9`` with the pre-activation gradient of conv9's output and returns
    (gx0d | None, dweight, dbias)."""
    from ._lib import lib, check
    n, oc, h, w = gy.shape
    dev = gy.device
    dw = torch.empty_like(weight, dtype=torch.float32)
    db = torch.empty(oc, device=dev, dtype=torch.float32)
    gx0d = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32) if need_x0d else None
    ws = torch.empty(lib.adc_ged_head_bwd_workspace_bytes(oc) // 4, device=dev, dtype=torch.float32)
    gy = gy.contiguous()
    glli = glli.contiguous() if glli is not None else None
    check(lib.adc_ged_head_bwd(x9.hi.data_ptr(), _ptr(x9.lo), x9.cs, x9.h, x9.w, x0d.data_ptr(), weight.detach().contiguous().data_ptr(),
                               gy.data_ptr(), x10.data_ptr(), _ptr(glli), glli.shape[1] if glli is not None else 0, g9.hi.data_ptr(),
                               _ptr(g9.lo), g9.cs, _ptr(gx0d), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, oc, h, w, _stream()),
          "adc_ged_head_bwd")
    return gx0d, dw, db


# ---- "quad" (pixel-shuffle) form of the full-resolution thin layers (csrc/ged_quad.cu) ---------------------------------------
_QUAD_MAPS = {}


def quad_maps(device):
    """Constant 0/1 tensors [4 parities, 9 low-resolution taps, 9 original taps] of the two quad identities:
    'up': conv3x3(up2(x))[2i+py] tap u falls on low-resolution row i + (py+u-1)//2                     (conv10 after the up-sampling)
    's2': d/dx of conv3x3-stride-2 at x row 2i+py receives output row i + (py+1-u)/2 when py+1-u is even  (conv1 data gradient)"""
    key = str(device)
    if key not in _QUAD_MAPS:
        up = torch.zeros(4, 9, 9)
        s2 = torch.zeros(4, 9, 9)
        for py in range(2):
            for px in range(2):
                for u in range(3):
                    for v in range(3):
                        a, b = (py + u - 1) // 2 + 1, (px + v - 1) // 2 + 1
                        up[py * 2 + px, a * 3 + b, u * 3 + v] = 1.0
                        if (py + 1 - u) % 2 == 0 and (px + 1 - v) % 2 == 0:
                            a, b = (py + 1 - u) // 2 + 1, (px + 1 - v) // 2 + 1
                            s2[py * 2 + px, a * 3 + b, u * 3 + v] = 1.0
        _QUAD_MAPS[key] = (up.to(device), s2.to(device))
    return _QUAD_MAPS[key]


def quad_tap_mix(w, which, transposed=False):
    """Quad form of a 3x3 weight tensor ``w`` [A, B, 3, 3] (or [A, B, 9]) in one launch: [4, A, B, 9], or [4, B, A, 9] when
    ``transposed``; ``which`` = 'up' | 's2' (the two identities of ``quad_maps``)."""
    from ._lib import lib, check
    _need_cuda(w)
    a, b = w.shape[:2]
    w = w.detach().reshape(a, b, 9).contiguous().float()
    out = torch.empty((4, b, a, 9) if transposed else (4, a, b, 9), device=w.device, dtype=torch.float32)
    check(lib.adc_quad_tap_mix(w.data_ptr(), out.data_ptr(), a, b, {"up": 0, "s2": 1}[which], int(bool(transposed)), _stream()),
          "adc_quad_tap_mix")
    return out


def quad_tap_mix_adjoint(g):
    """Transpose of ``quad_tap_mix(., 'up')``: ``g`` [4, L, 9] (gradient of the quad weights) -> [L, 9]."""
    from ._lib import lib, check
    _need_cuda(g)
    g = g.contiguous().float()
    out = torch.empty((g.shape[1], 9), device=g.device, dtype=torch.float32)
    check(lib.adc_quad_tap_mix_adjoint(g.data_ptr(), out.data_ptr(), g.shape[1], _stream()), "adc_quad_tap_mix_adjoint")
    return out


def quad_s2_tap_masks():
    """tap_mask (bit 3a+b) of the low-resolution taps each parity block of a quad stride-2 data gradient uses, and their union."""
    masks = []
    for py in range(2):
        for px in range(2):
            m = 0
            for u in range(3):
                for v in range(3):
                    if (py + 1 - u) % 2 == 0 and (px + 1 - v) % 2 == 0:
                        m |= 1 << (((py + 1 - u) // 2 + 1) * 3 + (px + 1 - v) // 2 + 1)
            masks.append(m)
    return masks, masks[0] | masks[1] | masks[2] | masks[3]


def ged_pack_input(xout_d, cout_d, rgb, dst):
    """torch.cat((xout_d, rgb, cout_d), 1) + NCHW->NHWC + bf16 (hi, lo) split in one pass, into channels [0,16) of ``dst``."""
    from ._lib import lib, check
    _need_cuda(xout_d)
    n, _, h, w = xout_d.shape
    xout_d, cout_d = xout_d.contiguous(), cout_d.contiguous()
    if rgb.stride(3) != 1 or rgb.stride(2) != w or rgb.stride(1) != h * w:
        rgb = rgb.contiguous()
    check(lib.adc_ged_pack_input(xout_d.data_ptr(), cout_d.data_ptr(), rgb.data_ptr(), rgb.stride(0) if n > 1 else 3 * h * w,
                                 dst.hi.data_ptr(), _ptr(dst.lo), dst.cs, n, h, w, _stream()), "adc_ged_pack_input")
    return dst


def ged_pack_im2col(xout_d, cout_d, rgb, dst):
    """im2col of conv1's input (cat(xout_d, rgb, cout_d), 3x3 window, stride 2) at the output resolution into channels [0,48) of
    ``dst`` (channel (3u+v)*5 + c): conv1 becomes a 1x1 convolution / GEMM at a quarter of the pixels."""
    from ._lib import lib, check
    _need_cuda(xout_d)
    n, _, h, w = xout_d.shape
    xout_d, cout_d = xout_d.contiguous(), cout_d.contiguous()
    if rgb.stride(3) != 1 or rgb.stride(2) != w or rgb.stride(1) != h * w:
        rgb = rgb.contiguous()
    check(lib.adc_ged_pack_im2col(xout_d.data_ptr(), cout_d.data_ptr(), rgb.data_ptr(), rgb.stride(0) if n > 1 else 3 * h * w,
                                  dst.hi.data_ptr(), _ptr(dst.lo), dst.cs, n, h, w, _stream()), "adc_ged_pack_im2col")
    return dst


def conv1x1_tc(act, cin, w_prep, bias, cout, act_fn=0, x_off=0, out=None, out_off=0, out_zero_to=0):
    """1x1 convolution of channels [x_off, x_off+cin) of a zero-bordered activation (cin % 16 == 0) into a channel slice of
    ``out``; ``w_prep`` = prepared weights of a [cout, cin, 3, 3] tensor whose only non-zero tap is the centre."""
    from ._lib import lib, check
    assert act.pad and out.pad and (act.n, act.h, act.w) == (out.n, out.h, out.w)
    b = bias.detach().contiguous().data_ptr() if bias is not None else None
    check(lib.adc_conv1x1_tc_v2(act.hi.data_ptr(), _ptr(act.lo), act.cs, x_off, w_prep[0].data_ptr(), _ptr(w_prep[1]), b,
                                out.hi.data_ptr(), _ptr(out.lo), int(out.hi.dtype == torch.float32), out.cs, out_off, out_zero_to,
                                act.n, act.h, act.w, cin, cout, act_fn, _stream()), "adc_conv1x1_tc_v2")
    return out


def conv1x1_wgrad_tc(gy, x, cin, cout, x_off=0, gy_off=0, cin_pad=None):
    """dW [cout, cin] (fp32) = sum over pixels of gy[p][co] * x[p][ci] from zero-bordered NHWC activations."""
    from ._lib import lib, check
    assert x.pad and gy.pad and (x.n, x.h, x.w) == (gy.n, gy.h, gy.w)
    cin_pad = _r16(cin) if cin_pad is None else cin_pad
    dw = torch.empty((cout, cin), device=x.hi.device, dtype=torch.float32)
    ws = torch.empty(lib.adc_conv3x3_wgrad_tc_workspace_bytes(cin_pad, cout) // 4, device=x.hi.device, dtype=torch.float32)
    check(lib.adc_conv1x1_wgrad_tc(gy.hi.data_ptr(), _ptr(gy.lo), gy.cs, gy_off, x.hi.data_ptr(), _ptr(x.lo), x.cs, x_off,
                                   dw.data_ptr(), ws.data_ptr(), x.n, x.h, x.w, cin, cin_pad, cout, _stream()), "adc_conv1x1_wgrad_tc")
    return dw


def quad_head_finish(q, x0d, weight, bias, h, w):
    from ._lib import lib, check
    oc = weight.shape[0]
    out = torch.empty((q.n, oc, h, w), device=x0d.device, dtype=torch.float32)
    check(lib.adc_quad_head_finish(q.hi.data_ptr(), q.cs, x0d.data_ptr(), weight.detach().contiguous().data_ptr(),
                                   bias.detach().contiguous().data_ptr(), out.data_ptr(), q.n, oc, h, w, _stream()), "adc_quad_head_finish")
    return out


def quad_head_bwd_prep(gy, x10, x0d, gq):
    """Fills ``gq`` (fp32 [n, h/2+2, w/2+2, 4*oc], zero border kept) with the pre-activation gradient of conv10 in quad layout and
    returns the summed [oc, 10] x0_d-channel weight / bias gradient."""
    from ._lib import lib, check
    n, oc, h, w = gy.shape
    parts = torch.empty((lib.adc_quad_head_bwd_prep_parts(), oc, 10), device=gy.device, dtype=torch.float32)
    check(lib.adc_quad_head_bwd_prep(gy.contiguous().data_ptr(), x10.data_ptr(), x0d.data_ptr(), gq.data_ptr(), parts.data_ptr(), n, oc, h, w,
                                     _stream()), "adc_quad_head_bwd_prep")
    return sum_rows(parts)


def quad_head_wgrad(x9, gq, oc):
    """dWq [4*oc, 32, 3, 3] = weight gradient of the quad conv10 with respect to its parity weights."""
    from ._lib import lib, check
    parts = torch.empty((lib.adc_quad_head_bwd_prep_parts(), 4 * oc, 32, 9), device=gq.device, dtype=torch.float32)
    check(lib.adc_quad_head_wgrad(x9.hi.data_ptr(), _ptr(x9.lo), x9.cs, gq.data_ptr(), parts.data_ptr(), x9.n, oc, x9.h, x9.w, _stream()),
          "adc_quad_head_wgrad")
    return sum_rows(parts).reshape(4 * oc, 32, 3, 3)


def quad_head_dgrad(x9, gq, wq, g9, extra=None):
    """G9 = LeakyReLU'(x9) * (data gradient of the quad conv10 [+ extra]) into channels [0,32) of ``g9``; returns its per-channel
    sums (the bias gradient of conv9)."""
    from ._lib import lib, check
    oc = wq.shape[0] // 4
    parts = torch.empty((lib.adc_quad_head_bwd_prep_parts(), 32), device=gq.device, dtype=torch.float32)
    eh, el, es = (extra.hi.data_ptr(), _ptr(extra.lo), extra.cs) if extra is not None else (None, None, 0)
    check(lib.adc_quad_head_dgrad(x9.hi.data_ptr(), x9.cs, gq.data_ptr(), wq.contiguous().data_ptr(), eh, el, es, g9.hi.data_ptr(), _ptr(g9.lo),
                                  g9.cs, parts.data_ptr(), x9.n, oc, x9.h, x9.w, _stream()), "adc_quad_head_dgrad")
    return sum_rows(parts)


def quad_unshuffle(q, k, h, w):
    """[N][H1+2][W1+2][4k] fp32 parity channels -> [k, N, H, W] fp32."""
    from ._lib import lib, check
    out = torch.empty((k, q.n, h, w), device=q.hi.device, dtype=torch.float32)
    check(lib.adc_quad_unshuffle(q.hi.data_ptr(), q.cs, out.data_ptr(), q.n, k, h, w, _stream()), "adc_quad_unshuffle")
    return out


def ged_head_bwd_x0d(gy, x10, glli, weight):
    from ._lib import lib, check
    n, oc, h, w = gy.shape
    gx0d = torch.empty((n, 1, h, w), device=gy.device, dtype=torch.float32)
    glli = glli.contiguous() if glli is not None else None
    check(lib.adc_ged_head_bwd_x0d(gy.contiguous().data_ptr(), x10.data_ptr(), _ptr(glli), glli.shape[1] if glli is not None else 0,
                                   weight.detach().contiguous().data_ptr(), gx0d.data_ptr(), n, oc, h, w, _stream()), "adc_ged_head_bwd_x0d")
    return gx0d


# ---- cbr3-c1 confidence head behind its 3x3 convolution (csrc/conf_head.cu) -----------------------------------------------------
def conf_bn_stats(y, c, gamma, beta, running_mean, running_var, momentum, eps, training):
    """Batch (training) or running (eval) statistics of conv1's fp32 NHWC output -> the coefficient block the other conf_head calls
    read; in training mode the running statistics are updated in place like nn.BatchNorm2d does."""
    from ._lib import lib, check
    dev = y.hi.device
    coef = torch.empty(lib.adc_conf_head_coef_floats(), device=dev, dtype=torch.float32)
    ws = torch.empty(lib.adc_conf_head_workspace_bytes(c) // 8, device=dev, dtype=torch.float64)
    check(lib.adc_conf_bn_stats(y.hi.data_ptr(), y.cs, y.n, c, y.h, y.w, gamma.detach().data_ptr(), beta.detach().data_ptr(),
                                _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), int(training), coef.data_ptr(),
                                ws.data_ptr(), _stream()), "adc_conf_bn_stats")
    return coef, ws


def conf_head_apply(y, c, coef, w2, b2):
    """sigmoid(conv1x1(LeakyReLU_0.2(BN(y)))) in one pass: [n, 1, h, w] fp32."""
    from ._lib import lib, check
    conf = torch.empty((y.n, 1, y.h, y.w), device=y.hi.device, dtype=torch.float32)
    check(lib.adc_conf_head_apply(y.hi.data_ptr(), y.cs, coef.data_ptr(), w2.detach().contiguous().data_ptr(), b2.detach().data_ptr(),
                                  conf.data_ptr(), y.n, c, y.h, y.w, _stream()), "adc_conf_head_apply")
    return conf


def conf_head_bwd(y, c, coef, ws, w2, gconf, conf, training, gy):
    """Backward of conf_head_apply + BatchNorm: fills ``gy`` (zero-bordered NHWC bf16 activation) with dL/dy and returns
    (gw2 [c], gb2 [1], ggamma [c], gbeta [c])."""
    from ._lib import lib, check
    dev = y.hi.device
    gw2, gg, gb = (torch.empty(c, device=dev, dtype=torch.float32) for _ in range(3))
    gb2 = torch.empty(1, device=dev, dtype=torch.float32)
    check(lib.adc_conf_head_bwd(y.hi.data_ptr(), y.cs, coef.data_ptr(), w2.detach().contiguous().data_ptr(), gconf.data_ptr(), conf.data_ptr(),
                                int(training), gy.hi.data_ptr(), _ptr(gy.lo), gy.cs, gw2.data_ptr(), gb2.data_ptr(), gg.data_ptr(),
                                gb.data_ptr(), ws.data_ptr(), y.n, c, y.h, y.w, _stream()), "adc_conf_head_bwd")
    return gw2, gb2, gg, gb


def cached_weights(conv, embed_1x1=False, cin_pad=None):
    """Prepared bf16 weights of a Conv2d, kept ON the module and re-prepared only when the Parameter changes (storage,
    in-place version counter) or the precision mode / channel padding does."""
    w = conv.weight
    tag = (w.data_ptr(), w._version, PRECISION, cin_pad)
    hit = getattr(conv, "_adc_wprep", None)
    if hit is not None and hit[0] == tag:
        return hit[1]
    wt = w.detach()
    if embed_1x1:      # a 1x1 convolution = a 3x3 one whose only non-zero tap is the centre
        full = torch.zeros(wt.shape[0], wt.shape[1], 3, 3, device=wt.device, dtype=wt.dtype)
        full[:, :, 1, 1] = wt[:, :, 0, 0]
        wt = full
    prep = prep_conv3x3_weights(wt, cin_pad=cin_pad)
    conv._adc_wprep = (tag, prep)
    return prep


def tc_enabled():
    """True in the tensor-core precision modes (training and inference paths of the nets that have one)."""
    return PRECISION in ("bf16", "bf16x3")


def tc_inference_enabled():
    return PRECISION in ("bf16", "bf16x3") and not torch.is_grad_enabled()

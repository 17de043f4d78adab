"""torch.autograd bindings of the C-ABI kernels (``_lib``).  Device memory, streams and autograd bookkeeping are
torch's; every number is computed by the sm_100a kernels.  No CPU path: non-CUDA tensors raise."""
import ctypes

import torch

from . import _lib
from ._lib import lib, check

PYR_MODULES = ("nconv1", "nconv2", "nconv3", "nconv4", "nconv5", "nconv6", "nconv7")
_PYR_SHAPES = ((2, 1, 5, 5), (2, 2, 5, 5), (2, 2, 5, 5), (2, 4, 3, 3), (2, 4, 3, 3), (2, 4, 3, 3), (1, 2, 1, 1))
_PYR_WNUMEL = tuple(a * b * c * d for a, b, c, d in _PYR_SHAPES)
_PYR_BNUMEL = tuple(s[0] for s in _PYR_SHAPES)

_slot_counter = [0]


def new_weight_slot():
    """Constant-memory weight slot for one pyramid instance (see include/adc_b200.h: ADC_WSLOTS)."""
    s = _slot_counter[0] % (_lib.ADC_WSLOTS - 1)
    _slot_counter[0] += 1
    return s


LAYER_SLOT = _lib.ADC_WSLOTS - 1     # stand-alone NConv2d layers share the last slot


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (got %s)" % (name, t.dtype))


def _planes(name, t):
    """Return ``t`` with contiguous (H, W) planes and channel stride H*W (copying only if needed) and its batch stride."""
    _require_cuda_f32(name, t)
    n, c, h, w = t.shape
    if t.stride(3) != 1 or t.stride(2) != w or (c > 1 and t.stride(1) != h * w) or (n > 1 and t.stride(0) < c * h * w):
        t = t.contiguous()
    bs = t.stride(0) if n > 1 else c * h * w
    return t, bs


def _src(x, c, bs):
    n, ch, h, w = x.shape
    return _lib.Src(x.data_ptr(), c.data_ptr(), bs, ch, h, w)


def softplus_pow(w, times):
    """softplus(beta=10) applied ``times`` times (EnforcePos, model_zoo/nconv.py:114-131)."""
    _require_cuda_f32("weight", w)
    w = w.contiguous()
    out = torch.empty_like(w)
    check(lib.adc_softplus_pow(w.data_ptr(), out.data_ptr(), w.numel(), int(times), _stream()), "adc_softplus_pow")
    return out


class NConv2dFunction(torch.autograd.Function):
    """One NConv2d call (model_zoo/nconv.py:37-67).  ``module.weight`` is read again at backward time: its
    contents then feed the data gradients, like the reference's aliased autograd graph."""

    @staticmethod
    def forward(ctx, x, c, weight, bias, module):
        x, bsx = _planes("data", x)
        c, bsc = _planes("conf", c)
        if bsx != bsc:
            x, c = x.contiguous(), c.contiguous()
            bsx = bsc = x.shape[1] * x.shape[2] * x.shape[3]
        _require_cuda_f32("weight", weight)
        n, ci, h, w = x.shape
        co, ci_w, k, k2 = weight.shape
        if ci_w != ci or k != k2 or c.shape != x.shape:
            raise RuntimeError("NConv2d: shape mismatch data %s conf %s weight %s" % (tuple(x.shape), tuple(c.shape), tuple(weight.shape)))
        wv = weight.detach().clone()     # forward-time values (s = sum of w): a later call's EnforcePos hook rewrites the Parameter in place
        bv = bias.detach().contiguous()
        y = torch.empty((n, co, h, w), device=x.device, dtype=torch.float32)
        cout = torch.empty_like(y)
        s0 = _src(x, c, bsx)
        scratch = torch.empty(lib.adc_nconv2d_fwd_scratch_floats(ctypes.byref(s0), None, n, co, h, w), device=x.device,
                              dtype=torch.float32)
        check(lib.adc_nconv2d_fwd(ctypes.byref(s0), None, wv.data_ptr(), bv.data_ptr(), y.data_ptr(), cout.data_ptr(),
                                  scratch.data_ptr(), n, co, h, w, k, LAYER_SLOT, _stream()), "adc_nconv2d_fwd")
        ctx.save_for_backward(x, c, y, cout, wv, bv)
        ctx.bs = bsx
        ctx.module = module
        return y, cout

    @staticmethod
    def backward(ctx, gy, gcout):
        x, c, y, cout, w_fwd, b = ctx.saved_tensors
        n, ci, h, w = x.shape
        co, _, k, _ = w_fwd.shape
        gy = torch.zeros_like(y) if gy is None else gy.contiguous()
        gcout = torch.zeros_like(cout) if gcout is None else gcout.contiguous()
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        w_bwd = (ctx.module.weight if ctx.module is not None else w_fwd).detach().contiguous()
        gx = torch.empty((n, ci, h, w), device=x.device, dtype=torch.float32) if need_in else None
        gc = torch.empty_like(gx) if need_in else None
        gw = torch.zeros_like(w_fwd) if need_w else None
        gb = torch.zeros_like(b) if need_w else None
        s0 = _src(x, c, ctx.bs)
        scratch = torch.empty(lib.adc_nconv2d_bwd_scratch_floats(ctypes.byref(s0), None, n, co, h, w), device=x.device,
                              dtype=torch.float32)
        ptr = lambda t: t.data_ptr() if t is not None else None
        check(lib.adc_nconv2d_bwd(ctypes.byref(s0), None, y.data_ptr(), cout.data_ptr(), gy.data_ptr(), gcout.data_ptr(),
                                  w_fwd.data_ptr(), w_bwd.data_ptr(), b.data_ptr(), ptr(gx), ptr(gc), None, None, ptr(gw), ptr(gb),
                                  scratch.data_ptr(), n, co, h, w, k, LAYER_SLOT, _stream()), "adc_nconv2d_bwd")
        return gx, gc, gw, gb, None


def conf_pool(x, c):
    """(x_ds, c_ds, idx) of model_zoo/nconv_sd.py:43-49; forward only (the pyramid op owns the fused backward)."""
    x, c = x.contiguous(), c.contiguous()
    _require_cuda_f32("x", x)
    _require_cuda_f32("c", c)
    n, ch, h, w = x.shape
    x_ds = torch.empty((n, ch, h // 2, w // 2), device=x.device, dtype=torch.float32)
    c_ds = torch.empty_like(x_ds)
    idx = torch.empty(x_ds.shape, device=x.device, dtype=torch.int64)
    check(lib.adc_conf_pool_fwd(x.data_ptr(), c.data_ptr(), x_ds.data_ptr(), c_ds.data_ptr(), idx.data_ptr(), n, ch, h, w,
                                _stream()), "adc_conf_pool_fwd")
    return x_ds, c_ds, idx


def conf_pool_backward(g_xds, g_cds, idx, in_hw, add_x=None, add_c=None):
    n, ch = idx.shape[:2]
    h, w = in_hw
    gx = torch.empty((n, ch, h, w), device=idx.device, dtype=torch.float32)
    gc = torch.empty_like(gx)
    ptr = lambda t: t.contiguous().data_ptr() if t is not None else None
    check(lib.adc_conf_pool_bwd(g_xds.contiguous().data_ptr(), g_cds.contiguous().data_ptr(), idx.data_ptr(), ptr(add_x),
                                ptr(add_c), gx.data_ptr(), gc.data_ptr(), n, ch, h, w, _stream()), "adc_conf_pool_bwd")
    return gx, gc


def pyramid_idx_views(idx, n, h, w):
    off = (ctypes.c_longlong * 3)()
    hs = (ctypes.c_int * 3)()
    ws = (ctypes.c_int * 3)()
    check(lib.adc_pyramid_idx_layout(n, h, w, off, hs, ws), "adc_pyramid_idx_layout")
    return [idx[off[i]: off[i] + n * 2 * hs[i] * ws[i]].view(n, 2, hs[i], ws[i]) for i in range(3)]


class PyramidFunction(torch.autograd.Function):
    """The whole unguided pyramid (CNN.forward, model_zoo/nconv_sd.py:35-101) as one op: 11 fused NConv2d kernels,
    3 argmax-pool kernels, upsample+concat folded into the loaders; one C call forward, one backward."""

    @staticmethod
    def forward(ctx, x0, c0, module, training, *params):
        x0, bsx = _planes("x0", x0)
        c0, bsc = _planes("c0", c0)
        if x0.shape[1] != 1 or c0.shape != x0.shape:
            raise RuntimeError("unguided pyramid expects x0, c0 of shape [N,1,H,W]; got %s, %s" % (tuple(x0.shape), tuple(c0.shape)))
        if bsx != bsc:
            x0, c0 = x0.contiguous(), c0.contiguous()
            bsx = bsc = x0.shape[2] * x0.shape[3]
        n, _, h, w = x0.shape
        dev = x0.device
        ws = [p.detach().contiguous() for p in params[0::2]]
        bs = [p.detach().contiguous() for p in params[1::2]]
        for i, (wt, sh) in enumerate(zip(ws, _PYR_SHAPES)):
            _require_cuda_f32(PYR_MODULES[i] + ".weight", wt)
            if tuple(wt.shape) != sh:
                raise RuntimeError("%s.weight has shape %s, expected %s" % (PYR_MODULES[i], tuple(wt.shape), sh))
        a = _lib.PyramidFwd()
        a.N, a.H, a.W, a.training, a.slot = n, h, w, int(bool(training)), module._adc_slot
        a.x0, a.c0, a.x0_batch_stride, a.c0_batch_stride = x0.data_ptr(), c0.data_ptr(), bsx, bsc
        for i in range(7):
            a.weight[i] = ws[i].data_ptr()
            a.bias[i] = bs[i].data_ptr()
        weight_out = torch.empty(lib.adc_pyramid_weight_floats(), device=dev, dtype=torch.float32) if training else None
        saved = torch.empty(lib.adc_pyramid_saved_floats(n, h, w), device=dev, dtype=torch.float32)
        idx = torch.empty(lib.adc_pyramid_idx_count(n, h, w), device=dev, dtype=torch.int64)
        xout = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32)
        cout = torch.empty_like(xout)
        a.weight_out = weight_out.data_ptr() if training else None
        a.saved, a.idx, a.xout, a.cout = saved.data_ptr(), idx.data_ptr(), xout.data_ptr(), cout.data_ptr()
        check(lib.adc_pyramid_fwd(ctypes.byref(a), _stream()), "adc_pyramid_fwd")
        if training:
            # EnforcePos (model_zoo/nconv.py:114-117): the Parameters now hold softplus^calls of what they held.  The values
            # are copied INTO the parameter storage (the reference rebinds .data; the observable contents are the same) so
            # parameter addresses stay fixed -- required for CUDA-graph replay and for flat-bucket data parallelism.
            o = 0
            for i, name in enumerate(PYR_MODULES):
                getattr(module, name).weight.data.copy_(weight_out[o:o + _PYR_WNUMEL[i]].view(_PYR_SHAPES[i]))
                o += _PYR_WNUMEL[i]
        ctx.save_for_backward(x0, c0, saved, idx, xout, cout, *bs)
        ctx.module, ctx.bs, ctx.dims = module, bsx, (n, h, w)
        module._adc_last_idx = (idx, n, h, w)
        return xout, cout

    @staticmethod
    def backward(ctx, g_xout, g_cout):
        x0, c0, saved, idx, xout, cout = ctx.saved_tensors[:6]
        bs = ctx.saved_tensors[6:]
        n, h, w = ctx.dims
        dev = x0.device
        module = ctx.module
        g_xout = torch.zeros_like(xout) if g_xout is None else g_xout.contiguous()
        g_cout = torch.zeros_like(cout) if g_cout is None else g_cout.contiguous()
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = any(ctx.needs_input_grad[4:])
        w_bwd = [getattr(module, name).weight.detach().contiguous() for name in PYR_MODULES]   # contents NOW
        a = _lib.PyramidBwd()
        a.N, a.H, a.W, a.slot = n, h, w, module._adc_slot
        a.need_input_grad, a.need_weight_grad = int(need_in), int(need_w)
        a.x0, a.c0, a.x0_batch_stride, a.c0_batch_stride = x0.data_ptr(), c0.data_ptr(), ctx.bs, ctx.bs
        for i in range(7):
            a.weight_bwd[i] = w_bwd[i].data_ptr()
            a.bias[i] = bs[i].data_ptr()
        a.saved, a.idx, a.xout, a.cout = saved.data_ptr(), idx.data_ptr(), xout.data_ptr(), cout.data_ptr()
        a.g_xout, a.g_cout = g_xout.data_ptr(), g_cout.data_ptr()
        scratch = torch.empty(lib.adc_pyramid_bwd_scratch_floats(n, h, w), device=dev, dtype=torch.float32)
        a.scratch = scratch.data_ptr()
        gx0 = torch.empty_like(xout) if need_in else None
        gc0 = torch.empty_like(xout) if need_in else None
        gweight = torch.empty(sum(_PYR_WNUMEL), device=dev, dtype=torch.float32) if need_w else None
        gbias = torch.empty(sum(_PYR_BNUMEL), device=dev, dtype=torch.float32) if need_w else None
        a.gx0 = gx0.data_ptr() if need_in else None
        a.gc0 = gc0.data_ptr() if need_in else None
        a.gweight = gweight.data_ptr() if need_w else None
        a.gbias = gbias.data_ptr() if need_w else None
        check(lib.adc_pyramid_bwd(ctypes.byref(a), _stream()), "adc_pyramid_bwd")
        grads = []
        ow = ob = 0
        for i in range(7):
            if need_w:
                grads.append(gweight[ow:ow + _PYR_WNUMEL[i]].view(_PYR_SHAPES[i]))
                grads.append(gbias[ob:ob + _PYR_BNUMEL[i]])
            else:
                grads += [None, None]
            ow += _PYR_WNUMEL[i]
            ob += _PYR_BNUMEL[i]
        return (gx0, gc0, None, None, *grads)


# ---- masked losses / depth metrics (csrc/loss_metrics.cu) -------------------------------------------------------------------------
LOSS_KINDS = {"l2": 0, "l1": 1, "il1": 2, "absrel": 3}


def _lm_workspace(dev):
    return torch.empty(lib.adc_loss_metrics_workspace_bytes() // 8, device=dev, dtype=torch.float64)


class MaskedLossFunction(torch.autograd.Function):
    """(pred, target) -> [mean error over target > 0, number of valid pixels] (one 2-element tensor); one reduction kernel
    forward, one elementwise kernel backward (confidence_depth_framework.py:829-914)."""

    @staticmethod
    def forward(ctx, pred, target, kind):
        _require_cuda_f32("depth_prediction", pred)
        _require_cuda_f32("depth_target", target)
        if pred.shape != target.shape:
            raise RuntimeError("masked loss: prediction %s and target %s differ in shape" % (tuple(pred.shape), tuple(target.shape)))
        p, t = pred.contiguous(), target.contiguous()
        out = torch.empty(2, device=p.device, dtype=torch.float32)
        check(lib.adc_masked_loss_fwd(p.data_ptr(), t.data_ptr(), p.numel(), kind, _lm_workspace(p.device).data_ptr(), out.data_ptr(),
                                      _stream()), "adc_masked_loss_fwd")
        ctx.kind = kind
        ctx.save_for_backward(p, t, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        p, t, out = ctx.saved_tensors
        g = torch.empty_like(p)
        gl = gout.contiguous()          # element 0 is dL/dloss; the count output carries no gradient
        check(lib.adc_masked_loss_bwd(p.data_ptr(), t.data_ptr(), gl.data_ptr(), out[1:].data_ptr(), p.numel(), ctx.kind, g.data_ptr(),
                                      _stream()), "adc_masked_loss_bwd")
        return g, None, None


def masked_loss(pred, target, kind):
    """(loss, number of valid pixels): 0-dim tensors; only the loss is differentiable."""
    out = MaskedLossFunction.apply(pred, target, LOSS_KINDS[kind])
    return out[0], out[1].detach()


def depth_metrics(output, target, confidence=None, threshold=0.0):
    """[mse, mae, lg10, absrel, delta1, delta2, delta3, imse, imae, count] (fp32 device tensor) over target > 0
    [& confidence > threshold] in one pass (metrics.py:38-66)."""
    _require_cuda_f32("output", output)
    _require_cuda_f32("target", target)
    o, t = output.contiguous(), target.contiguous()
    c = None
    if confidence is not None:
        _require_cuda_f32("confidence", confidence)
        c = confidence.contiguous()
    if o.shape != t.shape or (c is not None and c.shape != t.shape):
        raise RuntimeError("depth_metrics: tensors differ in shape")
    out = torch.empty(10, device=o.device, dtype=torch.float32)
    check(lib.adc_depth_metrics(o.data_ptr(), t.data_ptr(), c.data_ptr() if c is not None else None, float(threshold), o.numel(),
                                _lm_workspace(o.device).data_ptr(), out.data_ptr(), _stream()), "adc_depth_metrics")
    return out


def conf_pr_histogram(depth, confidence, target, thresholds):
    """Histogram behind ConfidencePixelwiseThrAverageMeter.evaluate (metrics.py:103-131): int64 tensor [3][bins + 1] on the device
    (pixel count, absrel sum in 2^-16 fixed point, count of non-finite absrel) -- see adc_conf_pr_histogram.  thresholds: ascending
    fp32 CUDA tensor."""
    for name, t in (("depth", depth), ("confidence", confidence), ("target", target), ("thresholds", thresholds)):
        _require_cuda_f32(name, t)
    d, c, t = depth.contiguous(), confidence.contiguous(), target.contiguous()
    if d.shape != t.shape or c.shape != t.shape:
        raise RuntimeError("conf_pr_histogram: tensors differ in shape")
    thr = thresholds.contiguous()
    hist = torch.empty(3, thr.numel() + 1, device=d.device, dtype=torch.int64)
    check(lib.adc_conf_pr_histogram(d.data_ptr(), c.data_ptr(), t.data_ptr(), d.numel(), thr.data_ptr(), thr.numel(), hist.data_ptr(),
                                    _stream()), "adc_conf_pr_histogram")
    return hist

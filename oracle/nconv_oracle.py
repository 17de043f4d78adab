"""Oracle (TEST INFRASTRUCTURE): normalized convolution + the unguided pyramid.

Functional, dtype-generic restatement in torch CPU ops, with an *explicit*
backward (no autograd inside), because the reference's backward is not what
autograd of a clean re-implementation would give (see ``pyramid_backward``).

Reference lines followed (relative to /root/reference):
  model_zoo/nconv.py:37-67     NConv2d.forward
  model_zoo/nconv.py:114-131   EnforcePos pre-hook (softplus, beta=10)
  model_zoo/nconv_sd.py:35-101 CNN.forward (11 NConv calls, 3 pools, 3 upsample+cat)
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS = 1e-20  # nconv.py:25

# (module name, pyramid level) in call order -- nconv_sd.py:38-97
PYRAMID_CALLS = (
    ("nconv1", 0), ("nconv2", 0), ("nconv3", 0),
    ("nconv2", 1), ("nconv3", 1),
    ("nconv2", 2),
    ("nconv2", 3),
    ("nconv4", 2), ("nconv5", 1), ("nconv6", 0), ("nconv7", 0),
)
PYRAMID_MODULES = ("nconv1", "nconv2", "nconv3", "nconv4", "nconv5", "nconv6", "nconv7")


def softplus_pos(w):
    """nconv.py:130-131 -- F.softplus(p, beta=10) (threshold 20 => identity for 10*w > 20)."""
    return F.softplus(w, beta=10.0)


def nconv2d_forward(x, c, w, b, eps=EPS):
    """nconv.py:37-67.  Returns (y, cout, D) with D the un-normalised confidence sum.

    Only the configuration every net uses is restated: stride 1, dilation 1,
    groups 1, padding k//2 (nconv_sd.py:21-32)."""
    p = w.shape[-1] // 2
    D = F.conv2d(c, w, None, 1, p)                 # nconv.py:40-41
    Nm = F.conv2d(x * c, w, None, 1, p)            # nconv.py:42-43
    y = Nm / (D + eps) + b.view(1, -1, 1, 1)       # nconv.py:44-51
    s = w.reshape(w.shape[0], -1).sum(-1)          # nconv.py:58-61
    cout = D / s.view(1, -1, 1, 1)                 # nconv.py:63-64
    return y, cout, D


def nconv2d_backward(gy, gcout, x, c, y, D, w_fwd, w_bwd, b, eps=EPS, need_input_grad=True):
    """Backward of nconv2d_forward as the reference's autograd graph computes it.

    ``w_fwd`` is the weight the forward of this call used (it fixes ``s``);
    ``w_bwd`` is what the Parameter holds when backward runs -- the conv
    data-gradients use it because ``EnforcePos`` swaps ``weight.data`` after
    autograd saved the Parameter object (nconv.py:114-117).
    Returns (gx, gc, gw, gb)."""
    p = w_fwd.shape[-1] // 2
    Co = w_fwd.shape[0]
    s = w_fwd.reshape(Co, -1).sum(-1).view(1, -1, 1, 1)
    inv = 1.0 / (D + eps)
    gNm = gy * inv
    gD = -gy * (y - b.view(1, -1, 1, 1)) * inv + gcout / s
    gb = gy.sum((0, 2, 3))
    gs = -(gcout * (D / s) / s).sum((0, 2, 3))
    xc = x * c
    gw = (torch.nn.grad.conv2d_weight(xc, w_fwd.shape, gNm, padding=p)
          + torch.nn.grad.conv2d_weight(c, w_fwd.shape, gD, padding=p)
          + gs.view(-1, 1, 1, 1))
    if not need_input_grad:
        return None, None, gw, gb
    g_xc = F.conv_transpose2d(gNm, w_bwd, None, 1, p)
    g_c = F.conv_transpose2d(gD, w_bwd, None, 1, p)
    return g_xc * c, g_xc * x + g_c, gw, gb


def conf_pool_forward(x, c):
    """nconv_sd.py:43-49 -- 2x2 max-pool of the confidence, gather x at the argmax, c/4.

    idx is the int64 flat index ih*W+iw inside each (n, ch) plane; ties pick the
    first element in row-major window order; odd H/W floor."""
    c_ds, idx = F.max_pool2d(c, 2, 2, return_indices=True)
    x_ds = x.flatten(2).gather(2, idx.flatten(2)).view_as(c_ds)
    return x_ds, c_ds / 4, idx


def conf_pool_backward(g_xds, g_cds, idx, in_hw):
    """Backward of conf_pool_forward: both grads scatter to the argmax position."""
    N, C = idx.shape[:2]
    H, W = in_hw
    gx = torch.zeros(N, C, H * W, dtype=g_xds.dtype)
    gc = torch.zeros(N, C, H * W, dtype=g_xds.dtype)
    flat = idx.flatten(2)
    gx.scatter_(2, flat, g_xds.flatten(2))
    gc.scatter_(2, flat, (g_cds / 4).flatten(2))
    return gx.view(N, C, H, W), gc.view(N, C, H, W)


def nearest_src_index(out_size, in_size):
    """ATen 'nearest' index rule used by F.interpolate(mode='nearest') (nconv_sd.py:81-93):
    identity if equal, dst>>1 if out == 2*in, else min(floor(dst * float32(in/out)), in-1)."""
    dst = torch.arange(out_size)
    if out_size == in_size:
        return dst
    if out_size == 2 * in_size:
        return dst >> 1
    scale = torch.tensor(float(in_size), dtype=torch.float32) / torch.tensor(float(out_size), dtype=torch.float32)
    src = torch.floor(dst.to(torch.float32) * scale).to(torch.int64)
    return torch.clamp(src, max=in_size - 1)


def upsample_nearest(t, out_hw):
    ih = nearest_src_index(out_hw[0], t.shape[2])
    iw = nearest_src_index(out_hw[1], t.shape[3])
    return t[:, :, ih][:, :, :, iw]


def upsample_nearest_backward(g, in_hw):
    ih = nearest_src_index(g.shape[2], in_hw[0])
    iw = nearest_src_index(g.shape[3], in_hw[1])
    N, C = g.shape[:2]
    tmp = torch.zeros(N, C, in_hw[0], g.shape[3], dtype=g.dtype).index_add_(2, ih, g)
    return torch.zeros(N, C, in_hw[0], in_hw[1], dtype=g.dtype).index_add_(3, iw, tmp)


def pyramid_forward(P, x0, c0, training):
    """nconv_sd.py:35-101.  ``P`` maps 'nconv{1..7}.{weight,bias}' to tensors.

    In training mode every *call* of a module first replaces its weight by
    softplus(weight) (nconv.py:114-117), so call k of a module uses
    softplus^k of the stored weight; ``cache['final']`` holds the weights the
    Parameters contain after the forward.  Returns (xout, cout, cache)."""
    W = {m: P[m + ".weight"].detach() for m in PYRAMID_MODULES}
    B = {m: P[m + ".bias"].detach() for m in PYRAMID_MODULES}
    calls = []

    def call(m, x, c):
        if training:
            W[m] = softplus_pos(W[m])
        y, co, D = nconv2d_forward(x, c, W[m], B[m])
        calls.append(dict(m=m, x=x, c=c, y=y, D=D, w=W[m]))
        return y, co

    x1, c1 = call("nconv1", x0, c0)
    x1, c1 = call("nconv2", x1, c1)
    x1, c1 = call("nconv3", x1, c1)
    x1_ds, c1_ds, idx1 = conf_pool_forward(x1, c1)              # :43-49
    x2_ds, c2_ds = call("nconv2", x1_ds, c1_ds)
    x2_ds, c2_ds = call("nconv3", x2_ds, c2_ds)
    x2_dss, c2_dss, idx2 = conf_pool_forward(x2_ds, c2_ds)      # :56-63
    x3_ds, c3_ds = call("nconv2", x2_dss, c2_dss)
    x3_dss, c3_dss, idx3 = conf_pool_forward(x3_ds, c3_ds)      # :69-76
    x4_ds, c4_ds = call("nconv2", x3_dss, c3_dss)
    hw2, hw1, hw0 = c3_ds.shape[2:], c2_ds.shape[2:], c0.shape[2:]
    x4, c4 = upsample_nearest(x4_ds, hw2), upsample_nearest(c4_ds, hw2)       # :81-82
    x34_ds, c34_ds = call("nconv4", torch.cat((x3_ds, x4), 1), torch.cat((c3_ds, c4), 1))
    x34, c34 = upsample_nearest(x34_ds, hw1), upsample_nearest(c34_ds, hw1)   # :86-87
    x23_ds, c23_ds = call("nconv5", torch.cat((x2_ds, x34), 1), torch.cat((c2_ds, c34), 1))
    x23, c23 = upsample_nearest(x23_ds, hw0), upsample_nearest(c23_ds, hw0)   # :92-93
    xo, co = call("nconv6", torch.cat((x23, x1), 1), torch.cat((c23, c1), 1))  # :94 (upsampled FIRST here)
    xout, cout = call("nconv7", xo, co)
    cache = dict(calls=calls, idx=(idx1, idx2, idx3), bias=B, final=W,
                 hw=(tuple(hw0), tuple(hw1), tuple(hw2), tuple(c4_ds.shape[2:])))
    return xout, cout, cache


def pyramid_backward(cache, g_xout, g_cout, w_bwd=None, need_input_grad=True):
    """Explicit backward of pyramid_forward (what the reference's autograd does).

    ``w_bwd`` (module -> weight) are the weights the Parameters hold at backward
    time; default = ``cache['final']`` (no optimizer step in between).
    Returns (g_x0, g_c0, grads) with grads keyed like the state_dict."""
    if w_bwd is None:
        w_bwd = cache["final"]
    calls, B = cache["calls"], cache["bias"]
    hw0, hw1, hw2, hw3 = cache["hw"]
    idx1, idx2, idx3 = cache["idx"]
    grads = OrderedDict()

    def back(k, gy, gc, need_in=True):
        r = calls[k]
        m = r["m"]
        gx, gci, gw, gb = nconv2d_backward(gy, gc, r["x"], r["c"], r["y"], r["D"], r["w"], w_bwd[m], B[m],
                                           need_input_grad=need_in)
        grads[m + ".weight"] = grads.get(m + ".weight", 0) + gw
        grads[m + ".bias"] = grads.get(m + ".bias", 0) + gb
        return gx, gci

    gx, gc = back(10, g_xout, g_cout)                    # nconv7
    gx, gc = back(9, gx, gc)                             # nconv6: in = cat(up(x23), x1)
    g_x23, g_x1 = gx[:, :2], gx[:, 2:]
    g_c23, g_c1 = gc[:, :2], gc[:, 2:]
    gx, gc = back(8, upsample_nearest_backward(g_x23, hw1), upsample_nearest_backward(g_c23, hw1))  # nconv5
    g_x2ds, g_x34 = gx[:, :2], gx[:, 2:]
    g_c2ds, g_c34 = gc[:, :2], gc[:, 2:]
    gx, gc = back(7, upsample_nearest_backward(g_x34, hw2), upsample_nearest_backward(g_c34, hw2))  # nconv4
    g_x3ds, g_x4 = gx[:, :2], gx[:, 2:]
    g_c3ds, g_c4 = gc[:, :2], gc[:, 2:]
    gx, gc = back(6, upsample_nearest_backward(g_x4, hw3), upsample_nearest_backward(g_c4, hw3))    # nconv2 @L3
    gx, gc = conf_pool_backward(gx, gc, idx3, hw2)
    gx, gc = back(5, gx + g_x3ds, gc + g_c3ds)           # nconv2 @L2
    gx, gc = conf_pool_backward(gx, gc, idx2, hw1)
    gx, gc = back(4, gx + g_x2ds, gc + g_c2ds)           # nconv3 @L1
    gx, gc = back(3, gx, gc)                             # nconv2 @L1
    gx, gc = conf_pool_backward(gx, gc, idx1, hw0)
    gx, gc = back(2, gx + g_x1, gc + g_c1)               # nconv3 @L0
    gx, gc = back(1, gx, gc)                             # nconv2 @L0
    gx, gc = back(0, gx, gc, need_in=need_input_grad)    # nconv1
    return gx, gc, grads


class PyramidFn(torch.autograd.Function):
    """Autograd wrapper so the net oracles (nets_oracle.py) can use plain autograd
    around the pyramid.  ``holder`` is the dict of leaf tensors; in training mode
    the forward overwrites ``holder[...].data`` like the reference's hook does, and
    the backward reads the weights from ``holder`` *at backward time*."""

    @staticmethod
    def forward(ctx, x0, c0, holder, prefix, training, *leaves):
        P = {k: holder[prefix + k] for k in (m + s for m in PYRAMID_MODULES for s in (".weight", ".bias"))}
        xout, cout, cache = pyramid_forward(P, x0.detach(), c0.detach(), training)
        if training:
            for m in PYRAMID_MODULES:
                holder[prefix + m + ".weight"].data = cache["final"][m]
        ctx.cache, ctx.holder, ctx.prefix = cache, holder, prefix
        return xout, cout

    @staticmethod
    def backward(ctx, g_xout, g_cout):
        w_bwd = {m: ctx.holder[ctx.prefix + m + ".weight"].detach() for m in PYRAMID_MODULES}
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gx, gc, grads = pyramid_backward(ctx.cache, g_xout, g_cout, w_bwd, need_in)
        out = [grads[m + s] for m in PYRAMID_MODULES for s in (".weight", ".bias")]
        return (gx, gc, None, None, None, *out)


def pyramid_apply(holder, prefix, x0, c0, training):
    leaves = [holder[prefix + m + s] for m in PYRAMID_MODULES for s in (".weight", ".bias")]
    return PyramidFn.apply(x0, c0, holder, prefix, training, *leaves)

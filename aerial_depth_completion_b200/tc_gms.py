"""Tensor-core TRAINING path of GMSNet's three full-resolution 3x3 stacks (confidence_depth_framework.py:521-669): ``d`` (1 -> 16 ->
... -> 16), ``rgb`` (4 -> 64 -> ... -> 64), ``fuse`` (80 -> 64 -> 64 -> 64 -> 32 -> 32 -> 32) and the 1x1 ``last_layer``, forward and
backward as one autograd node on the kernels of tc_bn.py / conv_tc.cu.

Forward: every layer is the row-segment kernel with bias + ReLU fused in its epilogue, writing the (hi, lo) operand of the next layer;
the depth stream lives in 16-channel buffers and contracts ONE 16-channel K step per tap (the kernel's K tail) instead of a padded
64-channel block, the first fuse layer reads [rgb stream 64 | depth stream 16] as 64 + 16 channels.  Backward per layer: ReLU' and the
bias gradient in one pass over the fp32 gradient (bn_bwd with the identity affine), then the weight-gradient GEMM and the data-gradient
convolution (transposed, flipped weights) accumulating into the producer's fp32 gradient buffer.
"""
import torch

from . import dense
from . import tc_bn as T
from ._lib import lib, check

_STACKS = (("d", 1, (16,) * 6), ("rgb", 4, (64,) * 6), ("fuse", 80, (64, 64, 64, 32, 32, 32)))


def _convs(seq):
    return [m for m in seq if isinstance(m, torch.nn.Conv2d)]


class _Plan:
    def __init__(self, n, h, w, dev):
        self.n, self.h, self.w = n, h, w
        P = lambda c: T.pair_act(n, h, w, c, dev)
        self.d_in, self.r_in = P(16), P(16)                  # channel 0 = xout_d ; channels 0..3 = rgb, cout_d
        self.fuse_in = P(128)                                # [rgb stream 64 | depth stream 16 | zeros]
        self.act = {"d": [P(16) for _ in range(5)] + [self.fuse_in], "rgb": [P(64) for _ in range(5)] + [self.fuse_in],
                    "fuse": [P(64) for _ in range(6)]}
        self.off = {"d": [0] * 5 + [64], "rgb": [0] * 6, "fuse": [0] * 6}
        self.G, self.D = {}, {}
        self.ident = {}
        self.forward_id = 0
        self.dev = dev

    def g(self, key, c):
        if key not in self.G:
            self.G[key] = T.f32_act(self.n, self.h, self.w, c, self.dev)
        return self.G[key]

    def dy(self, key, c):
        if key not in self.D:
            self.D[key] = T.pair_act(self.n, self.h, self.w, c, self.dev)
        return self.D[key]

    def identity(self, C):
        """BN state of the identity affine (a = 1, no statistics): bn_bwd then computes g * act'(out) and the bias gradient."""
        if C not in self.ident:
            s = T.BNState()
            s.coef = torch.empty(lib.adc_bn_coef_floats(C), device=self.dev, dtype=torch.float32)
            s.ws = torch.empty(lib.adc_bn_workspace_bytes(C) // 8, device=self.dev, dtype=torch.float64)
            s.training = False
            dummy = self.g(("ident", C), dense._r16(C))
            check(lib.adc_bn_stats(dummy.hi.data_ptr(), dummy.cs, 0, self.n, C, self.h, self.w, None, None, None, None, 0.0, 0.0, 0,
                                   s.coef.data_ptr(), s.ws.data_ptr(), dense._stream()), "adc_bn_stats")
            self.ident[C] = s
        return self.ident[C]


def _plan(net, n, h, w, dev):
    plans = net.__dict__.setdefault("_tc_train_plans", {})
    key = (n, h, w, dense.PRECISION, str(dev))
    if key not in plans:
        plans.clear()
        plans[key] = _Plan(n, h, w, dev)
    return plans[key]


def _inputs(P, name):
    """(input activation, channel offset, contraction width) of every layer of a stack."""
    first = {"d": (P.d_in, 0, 16), "rgb": (P.r_in, 0, 16), "fuse": (P.fuse_in, 0, 80)}[name]
    outs, offs = P.act[name], P.off[name]
    widths = dict(_STACKS_W)[name]
    res = [first]
    for i in range(5):
        res.append((outs[i], offs[i], widths[i]))
    return res


_STACKS_W = tuple((name, widths) for name, _, widths in _STACKS)


class GMSDenseTC(torch.autograd.Function):
    """(xout_d, cout_d, rgb, parameters...) -> (xout [n, oc, h, w], last_layer_input [n, 32, h, w] | None, depth stream [n, 16, h, w] | None)."""

    @staticmethod
    def forward(ctx, net, want_feat, xout_d, cout_d, rgb, *params):
        n, _, h, w = xout_d.shape
        dev = xout_d.device
        P = _plan(net, n, h, w, dev)
        mods = {name: _convs(getattr(net, name)) for name, _, _ in _STACKS}
        wp = T.WPrep()
        for name, cin0, widths in _STACKS:
            for i, m in enumerate(mods[name]):
                wp.fwd((name, i), m.weight, dense._r16(m.in_channels))
        W = wp.run()
        dense.to_nhwc_bf16(xout_d, cpad=16, dst=P.d_in, dst_off=0)
        dense.to_nhwc_bf16(torch.cat((rgb, cout_d), 1), cpad=16, dst=P.r_in, dst_off=0)
        for name, cin0, widths in _STACKS:
            ins = _inputs(P, name)
            for i, m in enumerate(mods[name]):
                x, xoff, cin = ins[i]
                out, ooff = P.act[name][i], P.off[name][i]
                zero_to = 64 if (out.cs >= ooff + 64 and widths[i] >= 32) else 0      # 32-channel layers own a 64-channel slice (pads written as 0)
                T.conv(T.K_S1, x, xoff, cin, W[((name, i), 0)], out, ooff, widths[i], h, w, bias=m.bias, act=T.RELU, out_zero_to=zero_to)
        lli = P.act["fuse"][5]
        xout = T.dot_head(lli, 0, 32, net.last_layer.weight, net.last_layer.bias)
        f_lli = dense.to_nchw_f32(lli, 32) if want_feat else None
        f_d = dense.to_nchw_f32(P.fuse_in, 16, 64) if want_feat else None
        P.forward_id += 1
        ctx.net, ctx.plan, ctx.fid = net, P, P.forward_id
        ctx.save_for_backward(*params)
        return xout, f_lli, f_d

    @staticmethod
    def backward(ctx, g_xout, g_lli, g_d):
        net, P = ctx.net, ctx.plan
        if P.forward_id != ctx.fid:
            raise RuntimeError("GMSDenseTC: the module ran another forward before this backward; its saved activations are gone")
        n, h, w, dev = P.n, P.h, P.w, P.dev
        mods = {name: _convs(getattr(net, name)) for name, _, _ in _STACKS}
        wp = T.WPrep()
        for name, cin0, widths in _STACKS:
            for i, m in enumerate(mods[name]):
                if not (i == 0 and name == "fuse") and not (i == 0 and name == "rgb" and not ctx.needs_input_grad[3]) and \
                        not (i == 0 and name == "d" and not ctx.needs_input_grad[2]):
                    wp.transposed((name, i), m.weight, dense._r16(widths[i]), flip=True)
        wp.transposed(("fuse", 0), mods["fuse"][0].weight, 64, flip=True)
        W = wp.run()
        grads = {}
        # last_layer (1x1) and the features' gradients enter the fp32 gradient buffers
        G = P.g(("fuse", 5), 32)
        if g_lli is not None:
            T.nchw_into(g_lli, G, 0, False)
        gw, gb = T.dot_head_backward(P.act["fuse"][5], 0, 32, net.last_layer.weight, g_xout, G, 0, g_lli is not None)
        grads["last_layer.weight"], grads["last_layer.bias"] = gw, gb
        Gfi = P.g("fuse_in", 80)
        src_grad = {}
        for name in ("fuse", "rgb", "d"):
            ins = _inputs(P, name)
            widths = dict(_STACKS_W)[name]
            for i in (5, 4, 3, 2, 1, 0):
                m = mods[name][i]
                c = widths[i]
                out, ooff = P.act[name][i], P.off[name][i]
                if i == 5 and name != "fuse":
                    Gout, goff = Gfi, ooff                              # the stack's output is a slice of the fuse input
                else:
                    Gout, goff = P.g((name, i), dense._r16(c)), 0
                D = P.dy((name, i), dense._r16(c))
                _, gbias = T.bn_backward(Gout, goff, Gout, goff, out, ooff, P.identity(c), T.RELU, D, 0, c, affine=False)
                x, xoff, cin = ins[i]
                creal = m.in_channels
                grads["%s.%d.bias" % (name, 2 * i)] = gbias
                grads["%s.%d.weight" % (name, 2 * i)] = T.wgrad_s1(D, 0, c, x, xoff, creal)
                if i > 0:
                    Gin = P.g((name, i - 1), dense._r16(widths[i - 1]))
                    T.dgrad_s1(D, dense._r16(c), W, (name, i), Gin, 0, widths[i - 1], False)
                elif name == "fuse":
                    T.dgrad_s1(D, 64, W, (name, 0), Gfi, 0, 80, False)
                    if g_d is not None:
                        T.nchw_into(g_d, Gfi, 64, True)
                elif (name == "rgb" and ctx.needs_input_grad[3]) or (name == "d" and ctx.needs_input_grad[2]):
                    Gin = P.g((name, "in"), 16)
                    T.dgrad_s1(D, dense._r16(c), W, (name, 0), Gin, 0, creal, False)
                    src_grad[name] = Gin
        g_xd = dense.to_nchw_f32(src_grad["d"], 1) if "d" in src_grad else None
        g_cd = dense.to_nchw_f32(src_grad["rgb"], 1, 3) if "rgb" in src_grad else None
        names = [k for k, _ in net.named_parameters() if not k.startswith("d_net.")]
        return (None, None, g_xd, g_cd, None, *[grads.get(k) for k in names])


def gms_dense(net, xout_d, cout_d, rgb, want_feat):
    names = [k for k, _ in net.named_parameters() if not k.startswith("d_net.")]
    pd = dict(net.named_parameters())
    return GMSDenseTC.apply(net, want_feat, xout_d.contiguous(), cout_d.contiguous(), rgb, *[pd[k] for k in names])

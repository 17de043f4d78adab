"""Tensor-core training / inference path of ``S2DResNet`` (``model_zoo/s2d_resnet.py:159-242``; BASELINE config 5's ``resnet18``): every
layer forward AND backward on this library's kernels, built from the pieces of ``tc_bn`` (tcgen05 convolutions, BatchNorm passes, the
shared BasicBlock stage) plus ``csrc/s2d_ops.cu`` (im2col of the 7x7 stem, max-pooling, parity-block shuffles, bilinear resize).

  conv1 7x7 s2 (+bn1, ReLU)   im2col at the output resolution (49*Cin of 208 channels) -> ONE GEMM launch (1x1 mode of the row-segment kernel)
  maxpool 3x3 s2              adc_maxpool3x3s2_fwd / _bwd (window position per element, deterministic gather backward)
  layer1..layer4              tc_bn.stage_forward / stage_backward (odd sizes: 57x76 -> 29x38 -> 15x19 -> 8x10 at 228x304)
  conv2 1x1 512 -> 256 (+bn2) 1x1 mode; data gradient = 1x1 with transposed weights; weight gradient = the flat single-tap kernel
  UpProj x4 (:105-143)        Unpool (:11-24) puts x at the even pixels of a 2x larger zero tensor, so a 5x5 convolution of it at output pixel
                              (2i+py, 2j+px) only touches taps u = py (mod 2), v = px (mod 2):
                                  out[2i+py][2j+px] = sum_{a,b} W[2a-py][2b-px] * x[i+a-1][j+b-1]      (a, b in 0..2; taps outside 0..4 absent)
                              = a 3x3 stride-1 convolution of the LOW-resolution x with 4 parity blocks of output channels.  Both 5x5 convolutions
                              of a module (upper conv1, bottom conv) share x: ONE convolution with 8*Cout output channels on a quarter of the
                              pixels, 25 of 36 taps live, the zero-stuffed tensor never formed.  Its data gradient is one stride-1 convolution
                              of the parity-block gradient, its weight gradient one flat 3x3 weight-gradient call; the 5x5 tensors are
                              gathered back from the (8*Cout, Cin, 3, 3) result.
  conv3 3x3 16 -> out         row-segment kernel (output channels padded to 16), no BatchNorm
  bilinear align_corners      adc_bilinear_ac_fwd / _bwd to and from the fp32 NCHW tensors the caller sees
"""
import ctypes

import torch

from . import dense
from ._lib import lib, check
from . import tc_bn as T
from .tc_bn import NONE, RELU, K_1X1, WPrep, bn_forward, bn_backward, conv, fwd_s1, dgrad_s1, wgrad_s1, _ptr, _chunks, _chunk, _blocked

_STAGES = (("layer1", 1, 64, 64, 1), ("layer2", 2, 64, 128, 2), ("layer3", 3, 128, 256, 2), ("layer4", 4, 256, 512, 2))


def _half(v):
    return (v - 1) // 2 + 1


class _Plan:
    """Buffers of one (batch, height, width, precision) configuration.  Levels: 0 conv1 output, 1 max-pool / layer1, 2..4 layer2..4,
    5..8 the four UpProj outputs."""

    def __init__(self, net, n, h, w, dev):
        self.n, self.h, self.w, self.dev = n, h, w, dev
        hs, ws = [_half(h)], [_half(w)]
        for _ in range(4):
            hs.append(_half(hs[-1]))
            ws.append(_half(ws[-1]))
        for _ in range(4):
            hs.append(2 * hs[-1])
            ws.append(2 * ws[-1])
        self.hs, self.ws = hs, ws
        self.Y, self.X, self.S, self.G, self.D = {}, {}, {}, {}, {}
        self.forward_id = 0
        self.kc = dense._r16(49 * net.in_channels)                    # contraction width of the stem GEMM
        self.x49 = T.pair_act(n, hs[0], ws[0], dense._r64(self.kc), dev)
        self.pool_idx = torch.empty((n, hs[1], ws[1], 64), device=dev, dtype=torch.uint8)

    def _get(self, store, key, lvl, c, f32):
        if key not in store:
            mk = T.f32_act if f32 else T.pair_act
            store[key] = mk(self.n, self.hs[lvl], self.ws[lvl], dense._r64(c), self.dev)
        return store[key]

    def y(self, key, lvl, c):
        return self._get(self.Y, key, lvl, c, True)

    def x(self, key, lvl, c):
        return self._get(self.X, key, lvl, c, False)

    def g(self, key, lvl, c):
        return self._get(self.G, key, lvl, c, True)

    def d(self, key, lvl, c):
        return self._get(self.D, key, lvl, c, False)


def _plan(net, n, h, w, dev):
    plans = net.__dict__.setdefault("_tc_plans", {})
    key = (n, h, w, dense.PRECISION, str(dev))
    if key not in plans:
        plans.clear()
        plans[key] = _Plan(net, n, h, w, dev)
    return plans[key]


# ---- parity-block form of a 5x5 convolution over a zero-stuffed tensor -------------------------------------------------------------------
_QUAD5 = {}


def _quad5_index(dev):
    """idx [4, 3, 3] into the 25 taps (+ a 26th zero tap): parity block p = 2*py + px, low-resolution tap (a, b) <- 5x5 tap (2a-py, 2b-px);
    inv [25] = position in the flattened [4, 3, 3] block table of each 5x5 tap (every tap appears exactly once)."""
    if dev not in _QUAD5:
        idx = torch.full((4, 3, 3), 25, dtype=torch.long)
        inv = torch.zeros(25, dtype=torch.long)
        for py in range(2):
            for px in range(2):
                for a in range(3):
                    for b in range(3):
                        u, v = 2 * a - py, 2 * b - px
                        if 0 <= u <= 4 and 0 <= v <= 4:
                            idx[2 * py + px, a, b] = u * 5 + v
                            inv[u * 5 + v] = ((2 * py + px) * 3 + a) * 3 + b
        _QUAD5[dev] = (idx.to(dev), inv.to(dev))
    return _QUAD5[dev]


def quad5_weights(w5):
    """[Cout, Cin, 5, 5] -> [4*Cout, Cin, 3, 3] (row = parity * Cout + co)."""
    co, ci = w5.shape[:2]
    idx, _ = _quad5_index(w5.device)
    wp = torch.cat((w5.detach().reshape(co, ci, 25), w5.new_zeros(co, ci, 1)), 2)
    return wp[:, :, idx.reshape(-1)].reshape(co, ci, 4, 3, 3).permute(2, 0, 1, 3, 4).reshape(4 * co, ci, 3, 3)


def quad5_weight_grad(dwq, co):
    """Inverse gather: [4*Cout, Cin, 3, 3] gradient of the parity-block weights -> [Cout, Cin, 5, 5]."""
    ci = dwq.shape[1]
    _, inv = _quad5_index(dwq.device)
    t = dwq.reshape(4, co, ci, 9).permute(1, 2, 0, 3).reshape(co, ci, 36)
    return t[:, :, inv].reshape(co, ci, 5, 5).contiguous()


def _shuffle(Yq, q_off, Y, C):
    check(lib.adc_pixel_shuffle_f32(Yq.hi.data_ptr(), Yq.cs, q_off, Y.hi.data_ptr(), Y.cs, 0, Yq.n, C, Yq.h, Yq.w, Y.h, Y.w, 0, dense._stream()),
          "adc_pixel_shuffle_f32")


def _unshuffle(D, Dq, q_off, C):
    check(lib.adc_pixel_unshuffle_pair(D.hi.data_ptr(), _ptr(D.lo), D.cs, 0, Dq.hi.data_ptr(), _ptr(Dq.lo), Dq.cs, q_off, Dq.n, C, Dq.h, Dq.w,
                                       dense._stream()), "adc_pixel_unshuffle_pair")


def _bilinear(src, C, H, W, pair):
    out = torch.empty((src.n, C, H, W), device=src.hi.device, dtype=torch.float32)
    check(lib.adc_bilinear_ac_fwd(src.hi.data_ptr(), _ptr(src.lo) if pair else None, int(pair), src.cs, 0, src.n, C, src.h, src.w, out.data_ptr(),
                                  H, W, dense._stream()), "adc_bilinear_ac_fwd")
    return out


def _bilinear_bwd(g, G, C, accumulate):
    g = g.contiguous()
    n, c, H, W = g.shape
    assert c == C
    check(lib.adc_bilinear_ac_bwd(g.data_ptr(), n, C, H, W, G.hi.data_ptr(), G.cs, 0, G.h, G.w, int(bool(accumulate)), dense._stream()),
          "adc_bilinear_ac_bwd")


def _identity_state(Y, C):
    """BatchNorm state of a layer without BatchNorm (conv3): identity affine, so bn_backward just converts the fp32 gradient into the
    (hi, lo) operand of the gradient GEMMs."""
    s = T.BNState()
    s.coef = torch.empty(lib.adc_bn_coef_floats(C), device=Y.hi.device, dtype=torch.float32)
    s.ws = torch.empty(lib.adc_bn_workspace_bytes(C) // 8, device=Y.hi.device, dtype=torch.float64)
    s.training = False
    check(lib.adc_bn_stats(Y.hi.data_ptr(), Y.cs, 0, Y.n, C, Y.h, Y.w, None, None, None, None, 0.0, 0.0, 0, s.coef.data_ptr(), s.ws.data_ptr(),
                           dense._stream()), "adc_bn_stats")
    return s


def _wgrad_1x1(D, cout, x, x_off, cin):
    """[cout, cin] of a 1x1 convolution, in (128 x 128) channel blocks of the flat single-tap weight-gradient kernel."""
    return _blocked(cout, cin, 128, 128, lambda a, b, c, d: dense.conv1x1_wgrad_tc(D, x, d - c, b - a, x_off=x_off + c, gy_off=a,
                                                                                   cin_pad=dense._r16(d - c)))


def _cw(c):
    """Contraction width of a gradient GEMM over c output channels."""
    return dense._r16(c) if c < 64 else dense._r64(c)


class S2DBodyTC(torch.autograd.Function):
    """(x [n, in_channels, h, w], parameters in ``_param_names`` order) -> (y [n, oc, h, w], features [n, 16, h, w] | None)."""

    @staticmethod
    def forward(ctx, net, want_feat, x, *params):
        n, cin0, h, w = x.shape
        dev = x.device
        P = _plan(net, n, h, w, dev)
        training = net.training
        hs, ws = P.hs, P.ws
        dec = [getattr(net.decoder, "layer%d" % k) for k in range(1, 5)]
        # ---- weights of the whole forward pass
        wp = WPrep()
        w1e = torch.zeros((64, P.kc, 3, 3), device=dev, dtype=torch.float32)
        w1e[:, :49 * cin0, 1, 1] = net.conv1.weight.detach().permute(0, 2, 3, 1).reshape(64, 49 * cin0)      # channel (7u+v)*Cin + c
        wp.fwd("stem", w1e, P.kc)
        for name, lvl, ci, co, stride in _STAGES:
            T.stage_prep_fwd(wp, getattr(net, name), name, ci, co)
        wp.fwd("conv2", T.embed_1x1(net.conv2.weight), 512)
        for k, m in enumerate(dec):
            ci = 256 >> k
            wp.fwd(("up", k, "q"), torch.cat((quad5_weights(m.upper_branch.conv1.weight), quad5_weights(m.bottom_branch.conv.weight)), 0), ci)
            wp.fwd(("up", k, 2), m.upper_branch.conv2.weight, dense._r16(ci // 2))
        wp.fwd("conv3", net.conv3.weight, 16)
        W = wp.run()
        # ---- stem: im2col -> GEMM -> bn1 + ReLU -> max-pool
        xc = x.contiguous()
        check(lib.adc_im2col_kxk_pack(xc.data_ptr(), xc.stride(0), cin0, 7, 2, 3, P.x49.hi.data_ptr(), _ptr(P.x49.lo), P.x49.cs, P.x49.cs, n, h, w,
                                      dense._stream()), "adc_im2col_kxk_pack")
        Y1 = P.y("stem", 0, 64)
        conv(K_1X1, P.x49, 0, P.kc, W[("stem", 0)], Y1, 0, 64, hs[0], ws[0])
        A = P.x("stem", 0, 64)
        P.S["stem"] = bn_forward(Y1, 0, 64, net.bn1, training, A, 0, RELU)
        M = P.x("pool", 1, 64)
        check(lib.adc_maxpool3x3s2_fwd(A.hi.data_ptr(), _ptr(A.lo), A.cs, 0, M.hi.data_ptr(), _ptr(M.lo), M.cs, 0, P.pool_idx.data_ptr(), n, 64,
                                       hs[0], ws[0], dense._stream()), "adc_maxpool3x3s2_fwd")
        # ---- ResNet stages
        xin = M
        for name, lvl, ci, co, stride in _STAGES:
            dst = P.x((name, "out"), lvl, co)
            T.stage_forward(P, getattr(net, name), name, lvl, ci, co, stride, xin, 0, dst, 0, W, training)
            xin = dst
        # ---- conv2 (1x1) + bn2
        Yc = P.y("conv2", 4, 256)
        for a, b in _chunks(256, _chunk()):
            conv(K_1X1, xin, 0, 512, W[("conv2", a)], Yc, a, b - a, hs[4], ws[4])
        xin = P.x("dec_in", 4, 256)
        P.S["conv2"] = bn_forward(Yc, 0, 256, net.bn2, training, xin, 0, NONE)
        # ---- UpProj decoder
        for k, m in enumerate(dec):
            ci, co, lo_, hi_ = 256 >> k, 128 >> k, 4 + k, 5 + k
            Yq = P.y(("up", k, "q"), lo_, 8 * co)
            fwd_s1(xin, 0, ci, W, ("up", k, "q"), Yq, 0, 8 * co)
            Yu, Yb = P.y(("up", k, "u"), hi_, co), P.y(("up", k, "b"), hi_, co)
            _shuffle(Yq, 0, Yu, co)
            _shuffle(Yq, 4 * co, Yb, co)
            A1 = P.x(("up", k, "a1"), hi_, co)
            P.S[("up", k, 1)] = bn_forward(Yu, 0, co, m.upper_branch.batchnorm1, training, A1, 0, RELU)
            Y2 = P.y(("up", k, 2), hi_, co)
            fwd_s1(A1, 0, dense._r16(co), W, ("up", k, 2), Y2, 0, co)
            Rb = P.x(("up", k, "rb"), hi_, co)
            P.S[("up", k, "b")] = bn_forward(Yb, 0, co, m.bottom_branch.batchnorm, training, Rb, 0, NONE)
            Xo = P.x(("up", k, "out"), hi_, co)
            P.S[("up", k, 2)] = bn_forward(Y2, 0, co, m.upper_branch.batchnorm2, training, Xo, 0, RELU, res=Rb, res_off=0)
            xin = Xo
        # ---- conv3 + bilinear resize to the input resolution
        oc = net.conv3.weight.shape[0]
        Y3 = P.y("conv3", 8, oc)
        fwd_s1(xin, 0, 16, W, "conv3", Y3, 0, oc)
        y = _bilinear(Y3, oc, h, w, pair=False)
        feat = _bilinear(xin, 16, h, w, pair=True) if want_feat else None
        P.forward_id += 1
        ctx.net, ctx.plan, ctx.fid, ctx.want_feat = net, P, P.forward_id, want_feat
        ctx.save_for_backward(*params)
        return y, feat

    @staticmethod
    def backward(ctx, g_y, g_feat):
        net, P = ctx.net, ctx.plan
        if P.forward_id != ctx.fid:
            raise RuntimeError("S2DBodyTC: the module ran another forward before this backward; its saved activations are gone")
        names = _param_names(net)
        grads = {}
        n, h, w = P.n, P.h, P.w
        hs, ws = P.hs, P.ws
        dec = [getattr(net.decoder, "layer%d" % k) for k in range(1, 5)]
        oc = net.conv3.weight.shape[0]
        # ---- weights of every data gradient
        wp = WPrep()
        wp.transposed("conv3", net.conv3.weight, _cw(oc), flip=True)
        for k, m in enumerate(dec):
            ci, co = 256 >> k, 128 >> k
            wp.transposed(("up", k, "q"), torch.cat((quad5_weights(m.upper_branch.conv1.weight), quad5_weights(m.bottom_branch.conv.weight)), 0),
                          dense._r64(8 * co), flip=True)
            wp.transposed(("up", k, 2), m.upper_branch.conv2.weight, _cw(co), flip=True)
        wp.transposed("conv2", T.embed_1x1(net.conv2.weight), 256, flip=False)
        for name, lvl, ci, co, stride in _STAGES:
            T.stage_prep_bwd(wp, getattr(net, name), name, ci, co, stride)
        cin0 = net.in_channels
        if ctx.needs_input_grad[2]:      # S2DResNet as a loss net: rgb | depth1 | conf1 -> the last two need their gradient
            w1e = torch.zeros((64, P.kc, 3, 3), device=P.dev, dtype=torch.float32)
            w1e[:, :49 * cin0, 1, 1] = net.conv1.weight.detach().permute(0, 2, 3, 1).reshape(64, 49 * cin0)
            wp.transposed("stem", w1e, 64, flip=False)
        W = wp.run()
        # ---- bilinear + conv3
        F_ = P.X[("up", 3, "out")]
        G3 = P.g("conv3", 8, oc)
        if g_y is None:
            g_y = torch.zeros((n, oc, h, w), device=P.dev)
        _bilinear_bwd(g_y, G3, oc, False)
        D3 = P.d("conv3", 8, oc)
        bn_backward(G3, 0, P.Y["conv3"], 0, None, 0, _identity_state(P.Y["conv3"], oc), NONE, D3, 0, oc, affine=False)
        grads["conv3.weight"] = wgrad_s1(D3, 0, oc, F_, 0, 16)
        Gout = P.g(("up", 3, "out"), 8, 16)
        dgrad_s1(D3, _cw(oc), W, "conv3", Gout, 0, 16, False)
        if g_feat is not None and ctx.want_feat:
            _bilinear_bwd(g_feat, Gout, 16, True)
        # ---- UpProj modules, last first
        for k in (3, 2, 1, 0):
            m = dec[k]
            ci, co, lo_, hi_ = 256 >> k, 128 >> k, 4 + k, 5 + k
            pre = "decoder.layer%d." % (k + 1)
            xin = P.X[("up", k - 1, "out")] if k > 0 else P.X["dec_in"]
            D2, Gb = P.d(("up", k, 2), hi_, co), P.g(("up", k, "rb"), hi_, co)
            gg, gbeta = bn_backward(Gout, 0, P.Y[("up", k, 2)], 0, P.X[("up", k, "out")], 0, P.S[("up", k, 2)], RELU, D2, 0, co, gres=Gb, gres_off=0,
                                    gres_accum=False)
            grads[pre + "upper_branch.batchnorm2.weight"], grads[pre + "upper_branch.batchnorm2.bias"] = gg, gbeta
            A1 = P.X[("up", k, "a1")]
            grads[pre + "upper_branch.conv2.weight"] = wgrad_s1(D2, 0, co, A1, 0, co)
            Ga1 = P.g(("up", k, "a1"), hi_, co)
            dgrad_s1(D2, _cw(co), W, ("up", k, 2), Ga1, 0, co, False)
            D1 = P.d(("up", k, 1), hi_, co)
            gg, gbeta = bn_backward(Ga1, 0, P.Y[("up", k, "u")], 0, A1, 0, P.S[("up", k, 1)], RELU, D1, 0, co)
            grads[pre + "upper_branch.batchnorm1.weight"], grads[pre + "upper_branch.batchnorm1.bias"] = gg, gbeta
            Db = P.d(("up", k, "b"), hi_, co)
            gg, gbeta = bn_backward(Gb, 0, P.Y[("up", k, "b")], 0, None, 0, P.S[("up", k, "b")], NONE, Db, 0, co)
            grads[pre + "bottom_branch.batchnorm.weight"], grads[pre + "bottom_branch.batchnorm.bias"] = gg, gbeta
            Dq = P.d(("up", k, "q"), lo_, 8 * co)
            _unshuffle(D1, Dq, 0, co)
            _unshuffle(Db, Dq, 4 * co, co)
            dwq = wgrad_s1(Dq, 0, 8 * co, xin, 0, ci)
            grads[pre + "upper_branch.conv1.weight"] = quad5_weight_grad(dwq[:4 * co], co)
            grads[pre + "bottom_branch.conv.weight"] = quad5_weight_grad(dwq[4 * co:], co)
            Gout = P.g(("up", k - 1, "out"), lo_, ci) if k > 0 else P.g("dec_in", 4, 256)
            dgrad_s1(Dq, dense._r64(8 * co), W, ("up", k, "q"), Gout, 0, ci, False)
        # ---- bn2 + conv2 (1x1)
        x4 = P.X[("layer4", "out")]
        Dc = P.d("conv2", 4, 256)
        gg, gbeta = bn_backward(Gout, 0, P.Y["conv2"], 0, None, 0, P.S["conv2"], NONE, Dc, 0, 256)
        grads["bn2.weight"], grads["bn2.bias"] = gg, gbeta
        grads["conv2.weight"] = _wgrad_1x1(Dc, 256, x4, 0, 512).reshape(256, 512, 1, 1)
        G4 = P.g(("layer4", "out"), 4, 512)
        for a, b in _chunks(512, _chunk()):
            conv(K_1X1, Dc, 0, 256, W[("conv2", a)], G4, a, b - a, hs[4], ws[4])
        # ---- ResNet stages, deepest first
        Gs = G4
        for name, lvl, ci, co, stride in reversed(_STAGES):
            if name == "layer1":
                xin0, Gin0 = P.X["pool"], P.g("pool", 1, 64)
            else:
                prev = _STAGES[lvl - 2][0]
                xin0, Gin0 = P.X[(prev, "out")], P.g((prev, "out"), lvl - 1, ci)
            T.stage_backward(P, getattr(net, name), name, lvl, ci, co, stride, xin0, 0, Gin0, 0, False, P.X[(name, "out")], 0, Gs, 0, W, grads,
                             name + ".")
            Gs = Gin0
        # ---- max-pool, bn1, conv1 (weight gradient only: the input frames need no gradient)
        A = P.X["stem"]
        Ga = P.g("stem", 0, 64)
        check(lib.adc_maxpool3x3s2_bwd(Gs.hi.data_ptr(), Gs.cs, 0, P.pool_idx.data_ptr(), Ga.hi.data_ptr(), Ga.cs, 0, n, 64, hs[0], ws[0],
                                       dense._stream()), "adc_maxpool3x3s2_bwd")
        D1 = P.d("stem", 0, 64)
        gg, gbeta = bn_backward(Ga, 0, P.Y["stem"], 0, A, 0, P.S["stem"], RELU, D1, 0, 64)
        grads["bn1.weight"], grads["bn1.bias"] = gg, gbeta
        dw1 = _wgrad_1x1(D1, 64, P.x49, 0, 49 * cin0)
        grads["conv1.weight"] = dw1.reshape(64, 7, 7, cin0).permute(0, 3, 1, 2).contiguous()
        gx = None
        if ctx.needs_input_grad[2]:
            # data gradient of the stem: the GEMM's transpose gives the gradient of the im2col tensor, col2im gathers it back
            Gcol = P.g("col", 0, P.kc)
            for a, b in _chunks(P.kc, _chunk()):
                conv(K_1X1, D1, 0, 64, W[("stem", a)], Gcol, a, b - a, hs[0], ws[0])
            gx = torch.empty((n, cin0, h, w), device=P.dev, dtype=torch.float32)
            check(lib.adc_col2im_kxk(Gcol.hi.data_ptr(), Gcol.cs, cin0, 7, 2, 3, gx.data_ptr(), n, h, w, dense._stream()), "adc_col2im_kxk")
        return (None, None, gx, *[grads.get(nm) for nm in names])


def _param_names(net):
    return [k for k, _ in net.named_parameters()]


def s2d_forward(net, x, build_conf_input):
    """S2DResNet.forward (s2d_resnet.py:212-242) on the tensor-core path."""
    names = _param_names(net)
    pd = dict(net.named_parameters())
    want = bool(build_conf_input) and net.out_channels == 1
    y, feat = S2DBodyTC.apply(net, want, x, *[pd[k] for k in names])
    features = None
    if build_conf_input:
        features = feat if want else torch.sigmoid(y[:, 1:2, :, :])
    return y[:, 0:1, :, :], features

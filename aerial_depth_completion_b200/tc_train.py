"""Tensor-core training path of the dense layers of GEDNet (confidence_depth_framework.py:426-518): forward AND backward of
conv1..conv10 on the hand-written tcgen05 kernels, one autograd node.

Everything between the NConv pyramid and the loss lives in zero-bordered NHWC bf16 buffers (hi/lo pairs in the bf16x3 mode):

* every encoder activation x_k is written by its convolution straight into the channel slice of the decoder's concat buffer
  that will consume it (``cat9 = [up(x8) | x1]`` ...), the next encoder layer reads it from there: no torch.cat, no copies;
* forward: stride-2 layers = box kernel (adc_conv3x3_fprop_tc), stride-1 layers = row-segment kernel (adc_conv3x3_s1_tc_v2),
  ReLU / LeakyReLU(0.2) fused, nearest up-sampling = adc_nhwc_upsample_nearest into the slice;
* backward: data gradients = the forward kernel on the output gradient with transposed+flipped weights (stride 1) or four
  parity-class launches (stride 2, adc_conv3x3_dgrad_s2_tc); weight gradients = adc_conv3x3_wgrad_tc / _wgrad_box_tc straight
  from the NHWC buffers (MN-major UMMA operands); up-sampling backward + skip-gradient add + activation derivative =
  adc_nhwc_grad_assemble, whose per-channel sums are the bias gradients.

Buffers are allocated once per (module, batch shape, precision) and zero-filled once: kernels write interiors (or interiors
plus zero borders) only, so borders and pad channels stay zero for the lifetime of the plan, and a captured CUDA graph sees
fixed addresses.  A second forward of the same module before the backward of the first would overwrite saved activations; the
node checks for that and raises.
"""
import os

import torch

from . import dense

_RELU, _LEAKY = 1, 2
_SIDE_STREAM = os.environ.get("ADC_TC_SIDE_STREAM", "1") != "0"
# weight gradient of a layer enqueued AFTER its data gradient (the side stream then starts it when the data gradient has finished, beside
# the HBM-bound gradient-assembly pass of the next layer, instead of racing the data gradient for the same SMs)
_WGRAD_AFTER = os.environ.get("ADC_WGRAD_AFTER", "1") != "0"
_SUM_OFFPATH = os.environ.get("ADC_SUM_OFFPATH", "1") != "0"      # second stage of the bias-gradient reductions on the second stream


def _half(v):
    return (v - 1) // 2 + 1


class _Plan:
    def __init__(self, n, h, w, out_ch, device):
        self.n, self.h, self.w, self.out_ch = n, h, w, out_ch
        hs, ws = [h], [w]
        for _ in range(5):
            hs.append(_half(hs[-1]))
            ws.append(_half(ws[-1]))
        self.hs, self.ws = hs, ws
        A = lambda lvl, c, **kw: dense.new_act(n, hs[lvl], ws[lvl], c, device, pad=True, zero=True, **kw)
        # forward
        self.x45 = A(1, 64)            # im2col of conv1's input at conv1's output resolution: channel (3u+v)*5 + c, 45 real
        self.cat10 = None              # [up(x9) 32 | x0_d 1 | 0...]: only materialised when the confidence head wants it
        self.cat9 = A(1, 128)          # [up(x8) 32 | x1 80 | 0...]
        self.cat8 = A(2, 192)          # [up(x7) 64 | x2 80 | 0...]
        self.cat7 = A(3, 192)          # [up(x6) 64 | x3 80 | 0...]
        self.cat6 = A(4, 192)          # [up(x5) 80 | x4 80 | 0...]
        self.x5 = A(5, 128)
        self.x6, self.x7, self.x8, self.x9 = A(4, 64), A(3, 64), A(2, 64), A(1, 64)
        # conv10 in its quad form (exact 2x up-sampling only): 4 parity channels per output channel at x9's resolution, fp32
        self.quad = h % 2 == 0 and w % 2 == 0
        self.q10 = A(1, 16, dtype=torch.float32) if self.quad else None
        self.bwd = None
        self.forward_id = 0
        self.side = None               # second stream: weight gradients run beside the data-gradient chain

    def backward_buffers(self):
        if self.bwd is None:
            n, hs, ws = self.n, self.hs, self.ws
            dev = self.x45.hi.device
            A = lambda lvl, c, **kw: dense.new_act(n, hs[lvl], ws[lvl], c, dev, pad=True, zero=True, **kw)
            b = {}
            b["G9"], b["gcat9"] = A(1, 64), A(1, 128)
            b["G8"], b["gcat8"] = A(2, 64), A(2, 192)
            b["G7"], b["gcat7"] = A(3, 64), A(3, 192)
            b["G6"], b["gcat6"] = A(4, 64), A(4, 192)
            b["G5"] = A(5, 128)
            for k, lvl in (("4", 4), ("3", 3), ("2", 2), ("1", 1)):
                # gx_k: data gradient of the stride-2 layer that reads x_k, in quad form (4 parity blocks of 80 channels at the
                # layer's OUTPUT resolution); G_k: pre-activation gradient of x_k
                b["gx" + k], b["G" + k] = A(lvl + 1, 320), A(lvl, 128)
            b["qin0"] = A(1, 16, dtype=torch.float32)      # conv1's data gradient as 4 parity channels x {xout_d, cout_d}
            if self.quad:      # conv10's pre-activation gradient in quad layout, fp32, zero border
                b["gq"] = torch.zeros((n, hs[1] + 2, ws[1] + 2, 4 * self.out_ch), device=dev, dtype=torch.float32)
            self.bwd = b
        return self.bwd


def _plan(net, n, h, w, device):
    plans = net.__dict__.setdefault("_tc_plans", {})
    key = (n, h, w, dense.PRECISION, str(device))
    if key not in plans:
        plans[key] = _Plan(n, h, w, net.out_channels, device)
    return plans[key]


def _prep(weight, cin_pad, mode=0):
    return dense.prep_conv3x3_weights(weight, flip_transpose=mode, cin_pad=cin_pad)


def _dgrad_chunks(lo, hi):
    """Output-channel chunks of a stride-1 data gradient: the row-segment kernel's shared-memory budget takes 64 output channels
    in the split mode, 128 otherwise."""
    step = dense.out_chunk()
    return [(a, min(a + step, hi)) for a in range(lo, hi, step)]


def _dgrad_reqs(weight, cin_pad, lo, hi):
    """Weight-preparation requests (dense.prep_conv3x3_weights_batch) of _dgrad_s1 for the same arguments."""
    return [(weight, 1, cin_pad, a, b - a) for a, b in _dgrad_chunks(lo, hi)]


def _dgrad_s1(G, weight, out, lo, hi, prepped=None, cy=None):
    """Channels [lo, hi) of the input gradient of a stride-1 layer with Parameter ``weight`` [Cout, Cin, 3, 3] from its output
    gradient ``G`` (all of G's channels are contracted; pad channels are zeros).  ``prepped``: the prepared weights of
    _dgrad_reqs(weight, G.cs, lo, hi) if the caller batched them."""
    cy = G.cs if cy is None else cy          # contraction width: the layer's real output channels rounded up to 16
    if prepped is None:
        prepped = dense.prep_conv3x3_weights_batch(_dgrad_reqs(weight, cy, lo, hi))
    for (a, b), wp in zip(_dgrad_chunks(lo, hi), prepped):
        dense.conv3x3_tc(G, cy, wp, None, b - a, 1, 0, out=out, out_off=a, out_zero_to=64 if a + 64 <= out.cs else 0)


def _quad_head_weights(w10):
    """[oc, 33, 3, 3] -> [4*oc, 32, 3, 3]: per parity (py, px) of the output pixel, the taps of conv10 that read the same pixel of
    x9 through the 2x nearest up-sampling, summed (row (2py+px)*oc + o); one launch (dense.quad_tap_mix)."""
    oc = w10.shape[0]
    return dense.quad_tap_mix(w10[:, :32], "up").reshape(4 * oc, 32, 3, 3)


def _quad_head_backward(P, b, w10, g_x10, x10, x0_d, g_lli, need_x0d):
    """Backward of the quad conv10: G9 (pre-activation gradient of conv9's output) into b['G9']; returns (g_x0d | None, dW, db,
    bias gradient of conv9).  With 4*oc parity channels both contractions are direct CUDA-core kernels (csrc/ged_quad.cu)."""
    oc = w10.shape[0]
    gq = b["gq"]
    s = dense.quad_head_bwd_prep(g_x10, x10, x0_d, gq)                    # [oc, 10]: x0_d-channel weight gradient | bias gradient
    extra = None
    if g_lli is not None:      # gradient of the feature tensor cat(up(x9), x0_d): its first 32 channels sum over the 2x2 children
        import torch.nn.functional as F
        extra = dense.to_nhwc_bf16(F.avg_pool2d(g_lli[:, :32], 2) * 4.0, cpad=32, pad=True)
    gb9 = dense.quad_head_dgrad(P.x9, gq, _quad_head_weights(w10), b["G9"], extra)
    dwq = dense.quad_head_wgrad(P.x9, gq, oc)                             # [4*oc, 32, 3, 3]
    dw = dense.quad_tap_mix_adjoint(dwq.reshape(4, oc * 32, 9)).reshape(oc, 32, 3, 3)                      # adjoint of quad_tap_mix 'up'
    dw = torch.cat((dw, s[:, :9].reshape(oc, 1, 3, 3)), 1)
    g_x0d = dense.ged_head_bwd_x0d(g_x10, x10, g_lli, w10) if need_x0d else None
    return g_x0d, dw, s[:, 9].contiguous(), gb9


class GEDDenseTC(torch.autograd.Function):
    """(xout_d, cout_d, rgb, x0_d) -> (x10, last_layer_input | None); parameters conv1..conv10 (weight, bias) as inputs."""

    @staticmethod
    def forward(ctx, xout_d, cout_d, rgb, x0_d, want_lli, net, *params):
        n, _, h, w = xout_d.shape
        P = _plan(net, n, h, w, xout_d.device)
        W = params[0::2]
        Bz = params[1::2]
        c3 = dense.conv3x3_tc
        dense.ged_pack_im2col(xout_d, cout_d, rgb, P.x45)
        oc = W[9].shape[0]
        # every layer's weights -> bf16 (hi, lo) [9][Cout16][Cin_pad] in ONE launch
        # conv1 as a GEMM over the im2col tensor: [80, 45] weights in the centre tap of a 3x3 tensor (the 1x1 kernel reads only it)
        w1 = torch.zeros((80, 45, 3, 3), device=W[0].device, dtype=torch.float32)
        w1[:, :, 1, 1] = W[0].permute(0, 2, 3, 1).reshape(80, 45)
        # contraction widths = the REAL channel counts rounded to the MMA's K = 16 (the row-segment kernel takes a K tail): 48 of the
        # 64-channel im2col pixel, 160 / 144 / 144 / 112 of the 192 / 192 / 192 / 128-channel concat buffers, 32 of x9's 64
        reqs = [(w1, 0, 48, 0, 45)] + [(W[i], 0, cp, 0, W[i].shape[1]) for i, cp in ((1, 80), (2, 80), (3, 80), (4, 80), (5, 160), (6, 144),
                                                                                   (7, 144), (8, 112))]
        if P.quad:
            reqs.append((_quad_head_weights(W[9]), 0, 32, 0, 32))
        wp = dense.prep_conv3x3_weights_batch(reqs)
        # encoder: 3x3 stride 2 + ReLU, each output lands in the concat buffer of the decoder layer that reuses it
        dense.conv1x1_tc(P.x45, 48, wp[0], Bz[0], 80, _RELU, out=P.cat9, out_off=32)
        c3(P.cat9, 80, wp[1], Bz[1], 80, 2, _RELU, x_off=32, out=P.cat8, out_off=64)
        c3(P.cat8, 80, wp[2], Bz[2], 80, 2, _RELU, x_off=64, out=P.cat7, out_off=64)
        c3(P.cat7, 80, wp[3], Bz[3], 80, 2, _RELU, x_off=64, out=P.cat6, out_off=80)
        c3(P.cat6, 80, wp[4], Bz[4], 80, 2, _RELU, x_off=80, out=P.x5, out_off=0)
        # decoder: nearest up-sampling into the slice, 3x3 stride 1 + LeakyReLU(0.2)
        dense.upsample_into(P.x5, 80, P.cat6, 0)
        c3(P.cat6, 160, wp[5], Bz[5], 64, 1, _LEAKY, out=P.x6, out_zero_to=64)
        dense.upsample_into(P.x6, 64, P.cat7, 0)
        c3(P.cat7, 144, wp[6], Bz[6], 64, 1, _LEAKY, out=P.x7, out_zero_to=64)
        dense.upsample_into(P.x7, 64, P.cat8, 0)
        c3(P.cat8, 144, wp[7], Bz[7], 32, 1, _LEAKY, out=P.x8, out_zero_to=64)
        dense.upsample_into(P.x8, 32, P.cat9, 0)
        c3(P.cat9, 112, wp[8], Bz[8], 32, 1, _LEAKY, out=P.x9, out_zero_to=64)
        # conv10 (33 -> 1 or 2 channels at full resolution): direct kernel, up-sampling folded into the addressing of x9
        x0_d = x0_d.contiguous()
        if P.quad:
            # quad form: a stride-1 convolution of x9 at its own resolution on the tensor cores (4 parity channels per output
            # channel), then un-shuffle + x0_d taps + bias + LeakyReLU at full resolution
            c3(P.x9, 32, wp[9], None, 4 * oc, 1, 0, out=P.q10)
            x10 = dense.quad_head_finish(P.q10, x0_d, W[9], Bz[9], h, w)
        else:
            x10 = dense.ged_head_fwd(P.x9, x0_d, W[9], Bz[9], h, w)
        lli = None
        if want_lli:       # the confidence head wants cat(up(x9), x0_d) as features: materialise it only then
            if P.cat10 is None:
                P.cat10 = dense.new_act(n, h, w, 64, xout_d.device, pad=True, zero=True)
            dense.upsample_into(P.x9, 32, P.cat10, 0)
            dense.to_nhwc_bf16(x0_d, cpad=8, dst=P.cat10, dst_off=32)
            lli = dense.to_nchw_f32(P.cat10, 33)
        P.forward_id += 1
        ctx.plan, ctx.fid, ctx.oc, ctx.want_lli = P, P.forward_id, oc, want_lli
        ctx.save_for_backward(x10, x0_d, *W)
        return x10, lli

    @staticmethod
    def backward(ctx, g_x10, g_lli):
        P = ctx.plan
        if P.forward_id != ctx.fid:
            raise RuntimeError("GEDDenseTC: the module ran another forward before this backward; its saved activations are gone")
        x10, x0_d = ctx.saved_tensors[0], ctx.saved_tensors[1]
        W = ctx.saved_tensors[2:]
        b = P.backward_buffers()
        hs, ws, oc = P.hs, P.ws, ctx.oc
        asm, cs = dense.grad_assemble, dense.colsum
        gW, gB = [None] * 10, [None] * 10
        dev = x10.device
        # Weight gradients leave the critical path: each is enqueued on a second stream as soon as its output gradient exists and
        # overlaps the data-gradient chain of the layers below (the small levels launch far fewer CTAs than there are SMs).
        main = torch.cuda.current_stream()
        if P.side is None:
            P.side = torch.cuda.Stream()
        side = P.side if _SIDE_STREAM else main

        def off_path(fn, *a, **kw):
            if side is main:
                return fn(*a, **kw)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = fn(*a, **kw)
            out.record_stream(main)
            return out
        def off_sum(part):
            """Second stage of a bias-gradient reduction on the second stream (nothing on the data-gradient chain waits for it)."""
            if side is main or not _SUM_OFFPATH:
                return dense.sum_rows(part)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = dense.sum_rows(part)
            part.record_stream(side)
            out.record_stream(main)
            return out
        wg = lambda *a, **kw: off_path(dense.conv3x3_wgrad_tc, *a, **kw)
        wgb = lambda *a, **kw: off_path(dense.conv3x3_wgrad_box_tc, *a, **kw)
        # all data-gradient weights of the step (transposed / flipped, per output-channel chunk) in ONE launch
        dec_cin = {8: 112, 7: 144, 6: 144, 5: 160}
        reqs, where = [], {}
        def add(key, rs):
            where[key] = (len(reqs), len(reqs) + len(rs))
            reqs.extend(rs)
        dec_cy = {8: 32, 7: 32, 6: 64, 5: 64}        # contraction widths of the decoder data gradients = the layers' output channels
        for i in (8, 7, 6, 5):
            add(i, _dgrad_reqs(W[i], dec_cy[i], 0, dec_cin[i]))
        # stride-2 layers conv5..conv2: quad form of the data gradient (dense.quad_maps 's2'): per parity block q a stride-1
        # convolution of the output gradient with the taps of W that reach that parity, [q][ci][co][3x3]
        masks, mask_all = dense.quad_s2_tap_masks()
        for i in (4, 3, 2, 1):
            wq = dense.quad_tap_mix(W[i], "s2", transposed=True).reshape(4, 80, 80, 3, 3)
            add(i, [(wq[q], 0, 80, 0, 80) for q in range(4)])
        # conv1's data gradient is needed for xout_d and cout_d only (channels 0 and 4 of its input): quad form = a stride-1
        # convolution of G1 at ITS resolution with 4 parity x 2 output channels (dense.quad_maps 's2')
        Wd = dense.quad_tap_mix(torch.stack((W[0][:, 0], W[0][:, 4]), 1), "s2", transposed=True).reshape(8, 80, 3, 3)
        add(0, [(Wd, 0, 80, 0, 80)])
        prepped = dense.prep_conv3x3_weights_batch(reqs)
        wp = {k: prepped[a:b_] for k, (a, b_) in where.items()}
        # conv10 (LeakyReLU on the output too, confidence_depth_framework.py:506); produces G9 for conv9.
        # Channel 32 of its input is x0_d itself: raw data for the depth-completion net, but the PREDICTED depth when this net is
        # the loss net (confidence_depth_framework.py:230) -- then its gradient is needed too
        if P.quad:
            g_x0d, gW[9], gB[9], gb9 = _quad_head_backward(P, b, W[9], g_x10, x10, x0_d, g_lli, ctx.needs_input_grad[3])
        else:
            g_x0d, gW[9], gB[9] = dense.ged_head_bwd(P.x9, x0_d, W[9], g_x10, x10, g_lli, b["G9"], ctx.needs_input_grad[3])
            gb9 = cs(b["G9"], 32)
        # decoder layers 9..6: G_k = leaky'(x_k) * sum_children(grad of up(x_k)); the bias gradient (per-channel sum of G_k) comes
        # out of the same pass
        dec = ((8, "G9", None, P.x9, 32, P.cat9, 112, "gcat9"), (7, "G8", "gcat9", P.x8, 32, P.cat8, 144, "gcat8"),
               (6, "G7", "gcat8", P.x7, 64, P.cat7, 144, "gcat7"), (5, "G6", "gcat7", P.x6, 64, P.cat6, 160, "gcat6"))
        for i, gname, src, xk, co, catk, cin, gcat in dec:
            if src is not None:
                G, gB[i] = asm(b[gname], co, a=b[src], act=xk, slope=0.2, colsum=True, reduce=off_sum)
            else:
                G, gB[i] = b[gname], gb9
            if not _WGRAD_AFTER:
                gW[i] = wg(G, catk, cin, co, cin_pad=cin)
            _dgrad_s1(G, W[i], b[gcat], 0, cin, prepped=wp[i], cy=dec_cy[i])
            if _WGRAD_AFTER:
                gW[i] = wg(G, catk, cin, co, cin_pad=cin)
        # encoder layers 5..1 (stride 2, ReLU): G5 from the up-sampling children only, G4..G1 = relu' * (dgrad + skip gradient)
        G, gb = asm(b["G5"], 80, a=b["gcat6"], act=P.x5, slope=0.0, colsum=True, reduce=off_sum)
        enc = ((4, P.cat6, 80, 4, "gx4", "G4", "gcat6", 80), (3, P.cat7, 64, 3, "gx3", "G3", "gcat7", 64),
               (2, P.cat8, 64, 2, "gx2", "G2", "gcat8", 64), (1, P.cat9, 32, 1, "gx1", "G1", "gcat9", 32))
        for i, xin, xoff, lvl, gxn, gn, gcat, goff in enc:
            gB[i] = gb
            if not _WGRAD_AFTER:
                gW[i] = wgb(G, xin, 80, 80, 2, x_off=xoff, cin_pad=80)
            for q in range(4):
                dense.conv3x3_tc(G, 80, wp[i][q], None, 80, 1, 0, out=b[gxn], out_off=80 * q, tap_mask=masks[q])
            if _WGRAD_AFTER:
                gW[i] = wgb(G, xin, 80, 80, 2, x_off=xoff, cin_pad=80)
            G, gb = asm(b[gn], 80, a=b[gxn], b=b[gcat], b_off=goff, act=xin, act_off=xoff, slope=0.0, colsum=True, a_quad=True, reduce=off_sum)
        gB[0] = gb
        gw1 = off_path(dense.conv1x1_wgrad_tc, G, P.x45, 45, 80, cin_pad=48)       # [80, 45], channel (3u+v)*5 + c
        dense.conv3x3_tc(G, 80, wp[0][0], None, 8, 1, 0, out=b["qin0"], tap_mask=mask_all)
        gin = dense.quad_unshuffle(b["qin0"], 2, hs[0], ws[0])
        if side is not main:
            main.wait_stream(side)
        gW[0] = gw1.reshape(80, 3, 3, 5).permute(0, 3, 1, 2).contiguous()
        grads = []
        for i in range(10):
            grads += [gW[i], gB[i]]
        return (gin[0].unsqueeze(1), gin[1].unsqueeze(1), None, g_x0d, None, None, *grads)


def ged_dense(net, xout_d, cout_d, rgb, x0_d, want_lli):
    params = []
    for i in range(1, 11):
        m = getattr(net, "conv%d" % i)
        params += [m.weight, m.bias]
    return GEDDenseTC.apply(xout_d, cout_d, rgb, x0_d, want_lli, net, *params)


class ConfHeadTC(torch.autograd.Function):
    """cbr3-c1 confidence head (confidence_depth_framework.py:299-329) on the B200 path: features [N,C,H,W] fp32 -> conf [N,1,H,W].
    conv1 (3x3, no bias), its data gradient and its weight gradient = tcgen05 kernels on zero-bordered NHWC bf16 (hi, lo) tensors;
    BatchNorm (batch statistics + running-statistics update in train mode) + LeakyReLU(0.2) + conv2 (1x1) + sigmoid and their
    backward = the four fused passes of csrc/conf_head.cu over conv1's fp32 output."""

    @staticmethod
    def forward(ctx, feat, w1, gamma, beta, w2, b2, bn):
        n, c, h, w = feat.shape
        training = bool(bn.training or bn.running_mean is None)
        cpad = dense._r64(c)
        X = dense.to_nhwc_bf16(feat, cpad=cpad, pad=True)
        ck = dense._r16(c)                     # contraction width: 48 of the 64-channel pixel for the 36 GED features
        Y = dense.conv3x3_tc(X, ck, dense.prep_conv3x3_weights(w1, cin_pad=ck), None, c, 1, 0, out_f32=True)
        momentum = bn.momentum if bn.momentum is not None else 0.1
        coef, ws = dense.conf_bn_stats(Y, c, gamma, beta, bn.running_mean, bn.running_var, momentum, bn.eps, training)
        conf = dense.conf_head_apply(Y, c, coef, w2, b2)
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        ctx.X, ctx.Y, ctx.coef, ctx.ws, ctx.training = X, Y, coef, ws, training
        ctx.save_for_backward(conf, w1, w2)
        return conf

    @staticmethod
    def backward(ctx, gconf):
        conf, w1, w2 = ctx.saved_tensors
        X, Y = ctx.X, ctx.Y
        n, c, h, w = X.n, w1.shape[0], X.h, X.w
        G = dense.new_act(n, h, w, X.cs, conf.device, pad=True, border_only=True)
        gw2, gb2, gg, gb = dense.conf_head_bwd(Y, c, ctx.coef, ctx.ws, w2, gconf.contiguous(), conf, ctx.training, G)
        gw1 = dense.conv3x3_wgrad_tc(G, X, c, c, cin_pad=X.cs)
        gfeat = None
        if ctx.needs_input_grad[0]:
            out = dense.new_act(n, h, w, dense._r64(c), conf.device, pad=True, border_only=True, dtype=torch.float32)
            _dgrad_s1(G, w1, out, 0, c, cy=dense._r16(c))
            gfeat = dense.to_nchw_f32(out, c)
        ctx.X = ctx.Y = None
        return gfeat, gw1, gg, gb, gw2.view_as(w2), gb2, None


def conf_head(head, feat):
    """CBR3C1Confidence.forward on the tensor-core path."""
    return ConfHeadTC.apply(feat, head.conv1.weight, head.bn1.weight, head.bn1.bias, head.conv2.weight, head.conv2.bias, head.bn1)

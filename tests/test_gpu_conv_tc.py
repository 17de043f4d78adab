"""GPU tests of the tcgen05 implicit-GEMM 3x3 convolution (bf16 operands, fp32 accumulation in TMEM) against a plain
PyTorch reference of the same op.
  * 'bf16' mode, fp32 output, reference evaluated on the bf16-rounded operands: only the accumulation order differs -> 2e-5.
  * 'bf16' mode vs the un-rounded fp32 convolution: inside the 1e-2 bar of the bf16/TF32 mode (BASELINE.json north_star).
  * 'bf16x3' split mode (hi+lo operands, three MMAs per product) vs the un-rounded fp64 convolution: 2e-5, i.e. inside the
    STRICT fp32 bar of 1e-4 with margin."""
import pytest
import torch
import torch.nn.functional as F

gpu = pytest.mark.gpu


def _act(y, act):
    if act == 1:
        return F.relu(y)
    if act == 2:
        return F.leaky_relu(y, 0.2)
    if act == 3:
        return torch.sigmoid(y)
    return y


def _ref(x, w, b, stride, act, rounded=True):
    xr = x.bfloat16().double() if rounded else x.double()
    wr = w.bfloat16().double() if rounded else w.double()
    return _act(F.conv2d(xr, wr, b.double() if b is not None else None, stride, 1), act).float()


CASES = [
    (2, 64, 64, 40, 56, 1, 1), (1, 80, 80, 57, 76, 2, 1), (2, 16, 16, 33, 47, 1, 1), (1, 160, 64, 15, 19, 1, 2),
    (1, 144, 32, 57, 76, 1, 2), (2, 112, 32, 30, 44, 1, 2), (1, 32, 32, 64, 80, 1, 1), (1, 48, 1, 24, 40, 1, 0),
    (1, 64, 64, 228, 304, 1, 1), (3, 80, 80, 29, 38, 2, 1), (1, 64, 32, 16, 8, 1, 3),
]


@gpu
@pytest.mark.parametrize("cfg", CASES)
def test_conv3x3_tc_matches_reference(cfg):
    from aerial_depth_completion_b200 import dense
    n, ci, co, h, w, stride, act = cfg
    g = torch.Generator().manual_seed(ci * 1000 + co + h)
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    ref = _ref(x, wt, b, stride, act)
    full = _ref(x, wt, b, stride, act, rounded=False)
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    try:
        dense.set_precision("bf16")
        xn = dense.to_nhwc_bf16(x.cuda())
        wp = dense.prep_conv3x3_weights(wt.cuda())
        out32 = dense.conv3x3_tc(xn, xn.cs, wp, b.cuda(), co, stride, act, out_f32=True)
        got = dense.to_nchw_f32(out32, co).cpu()
        assert rel(got, ref) < 2e-5, rel(got, ref)
        out16 = dense.conv3x3_tc(xn, xn.cs, wp, b.cuda(), co, stride, act)
        assert rel(dense.to_nchw_f32(out16, co).cpu(), ref) < 1e-2
        assert rel(got, full) < 1e-2
        dense.set_precision("bf16x3")
        xs = dense.to_nhwc_bf16(x.cuda())
        ws = dense.prep_conv3x3_weights(wt.cuda())
        o32 = dense.conv3x3_tc(xs, xs.cs, ws, b.cuda(), co, stride, act, out_f32=True)
        e3 = rel(dense.to_nchw_f32(o32, co).cpu(), full)
        assert e3 < 2e-5, e3
        opair = dense.conv3x3_tc(xs, xs.cs, ws, b.cuda(), co, stride, act)      # (hi, lo) bf16 output
        e3p = rel(dense.to_nchw_f32(opair, co).cpu(), full)
        assert e3p < 5e-5, e3p
    finally:
        dense.set_precision("bf16x3")


@gpu
@pytest.mark.parametrize("cfg", [(2, 64, 64, 40, 56, 1), (1, 64, 32, 57, 76, 2), (1, 128, 64, 33, 47, 1), (1, 64, 64, 228, 304, 1),
                                 (3, 64, 16, 9, 11, 0), (1, 64, 80, 24, 40, 3)])
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_conv3x3_v2_tap_reuse_kernel(cfg, prec):
    """v2 kernel (zero-bordered layout, 3 row segments per channel block, persistent, double-buffered TMEM)."""
    from aerial_depth_completion_b200 import dense
    n, ci, co, h, w, act = cfg
    g = torch.Generator().manual_seed(ci * 100 + co + h)
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    try:
        dense.set_precision(prec)
        ref = _ref(x, wt, b, 1, act, rounded=(prec == "bf16"))
        xa = dense.to_nhwc_bf16(x.cuda(), pad=True)
        wp = dense.prep_conv3x3_weights(wt.cuda())
        o32 = dense.conv3x3_tc(xa, xa.cs, wp, b.cuda(), co, 1, act, out_f32=True)
        assert o32.pad
        e = rel(dense.to_nchw_f32(o32, co).cpu(), ref)
        assert e < 2e-5, e
        assert float(o32.hi[:, 0].abs().max()) == 0.0 and float(o32.hi[:, :, 0].abs().max()) == 0.0      # zero border kept
        o16 = dense.conv3x3_tc(xa, xa.cs, wp, b.cuda(), co, 1, act)
        e16 = rel(dense.to_nchw_f32(o16, co).cpu(), ref)
        assert e16 < (5e-5 if prec == "bf16x3" else 1e-2), e16
        # chained second layer straight from the kernel's own padded output
        w2 = torch.randn(32, co, 3, 3, generator=g) * (2.0 / (co * 9)) ** 0.5
        if co % 64 == 0:
            o2 = dense.conv3x3_tc(o16, o16.cs, dense.prep_conv3x3_weights(w2.cuda()), None, 32, 1, 0, out_f32=True)
            mid = dense.to_nchw_f32(o16, co).cpu()
            ref2 = _ref(mid, w2, None, 1, 0, rounded=(prec == "bf16"))
            assert rel(dense.to_nchw_f32(o2, 32).cpu(), ref2) < 5e-5
    finally:
        dense.set_precision("bf16x3")


@gpu
def test_channel_slices_fold_concat():
    """Two layers writing into channel slices of one NHWC tensor == torch.cat of their outputs; reading a slice == slicing."""
    from aerial_depth_completion_b200 import dense
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 32, 20, 28, generator=g)
    w1 = torch.randn(16, 32, 3, 3, generator=g) * 0.1
    w2 = torch.randn(32, 32, 3, 3, generator=g) * 0.1
    w3 = torch.randn(16, 32, 3, 3, generator=g) * 0.1
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    try:
        dense.set_precision("bf16")
        xn = dense.to_nhwc_bf16(x.cuda())
        cat = dense.new_act(1, 20, 28, 48, "cuda", zero=True)
        dense.conv3x3_tc(xn, 32, dense.prep_conv3x3_weights(w1.cuda()), None, 16, out=cat, out_off=0)
        dense.conv3x3_tc(xn, 32, dense.prep_conv3x3_weights(w2.cuda()), None, 32, out=cat, out_off=16)
        ref = torch.cat((_ref(x, w1, None, 1, 0), _ref(x, w2, None, 1, 0)), 1)
        got = dense.to_nchw_f32(cat, 48).cpu()
        assert rel(got, ref) < 1e-2
        o = dense.conv3x3_tc(cat, 32, dense.prep_conv3x3_weights(w3.cuda()), None, 16, x_off=16, out_f32=True)
        ref3 = _ref(got[:, 16:48], w3, None, 1, 0)
        assert rel(dense.to_nchw_f32(o, 16).cpu(), ref3) < 2e-5
    finally:
        dense.set_precision("bf16x3")


@gpu
@pytest.mark.parametrize("cfg", [(2, 112, 32, 57, 76, 2, "bf16x3"), (1, 32, 8, 64, 80, 0, "bf16x3"), (3, 80, 8, 29, 39, 0, "bf16x3"),
                                 (1, 144, 32, 15, 19, 2, "bf16x3"), (2, 16, 16, 33, 47, 1, "bf16x3"), (1, 64, 64, 57, 76, 1, "bf16"),
                                 (2, 48, 80, 24, 41, 3, "bf16"), (1, 64, 16, 130, 126, 0, "bf16x3")])
def test_row_segment_kernel_stacked_horizontal_taps(cfg):
    """Thin layers (3 * accumulator width <= 256 columns) run the three horizontal taps of a tap row as ONE MMA and add the column
    blocks across lanes in the epilogue (tiles of 126 pixels).  Same results as the one-MMA-per-tap form (ADC_V2_NO_HSTACK=1) and as
    the fp64 convolution, for fp32 / bf16-pair outputs, channel slices, accumulation into a gradient buffer and masked taps."""
    import os
    from aerial_depth_completion_b200 import dense
    n, ci, co, h, w, act, prec = cfg
    g = torch.Generator().manual_seed(ci * 7 + co + h)
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    try:
        dense.set_precision(prec)
        bar = 2e-5
        ref = _ref(x, wt, b, 1, act, rounded=(prec == "bf16"))
        xa = dense.to_nhwc_bf16(x.cuda(), pad=True)
        wp = dense.prep_conv3x3_weights(wt.cuda(), cin_pad=ci)
        outs = {}
        for flag in ("0", "1"):
            os.environ["ADC_V2_NO_HSTACK"] = flag
            o32 = dense.conv3x3_tc(xa, ci, wp, b.cuda(), co, 1, act, out_f32=True)
            o16 = dense.conv3x3_tc(xa, ci, wp, b.cuda(), co, 1, act)
            outs[flag] = (dense.to_nchw_f32(o32, co).cpu(), dense.to_nchw_f32(o16, co).cpu(), o32)
        os.environ["ADC_V2_NO_HSTACK"] = "0"
        # the CTA-pair form (cta_group::2, split mode) against the single-CTA form of the same kernel
        os.environ["ADC_V2_NO_PAIR"] = "1"
        single = dense.to_nchw_f32(dense.conv3x3_tc(xa, ci, wp, b.cuda(), co, 1, act, out_f32=True), co).cpu()
        os.environ["ADC_V2_NO_PAIR"] = "0"
        assert rel(outs["0"][0], single) < 5e-6
        e = rel(outs["0"][0], ref)
        assert e < bar, e
        assert rel(outs["0"][0], outs["1"][0]) < 5e-6
        assert rel(outs["0"][1], ref) < (5e-5 if prec == "bf16x3" else 1e-2)
        o32 = outs["0"][2]
        assert float(o32.hi[:, 0].abs().max()) == 0.0 and float(o32.hi[:, :, -1].abs().max()) == 0.0 and float(o32.hi[:, -1].abs().max()) == 0.0
        # accumulate into an fp32 gradient buffer that already holds a term
        acc = dense.new_act(n, h, w, dense._r64(co), "cuda", pad=True, zero=True, dtype=torch.float32)
        from aerial_depth_completion_b200 import tc_bn
        for rep in range(2):
            tc_bn.conv(0, xa, 0, ci, wp, acc, 0, co, h, w, accumulate=True)
        ref0 = _ref(x, wt, None, 1, 0, rounded=(prec == "bf16"))
        assert rel(dense.to_nchw_f32(acc, co).cpu(), 2 * ref0) < bar
        # masked taps (zero weights, whole tap row 0 and tap (1, 2) absent)
        wm = wt.clone()
        wm[:, :, 0] = 0
        wm[:, :, 1, 2] = 0
        mask = 0
        for u in range(3):
            for v in range(3):
                if float(wm[:, :, u, v].abs().max()) > 0:
                    mask |= 1 << (3 * u + v)
        om = dense.conv3x3_tc(xa, ci, dense.prep_conv3x3_weights(wm.cuda(), cin_pad=ci), None, co, 1, 0, out_f32=True, tap_mask=mask)
        assert rel(dense.to_nchw_f32(om, co).cpu(), _ref(x, wm, None, 1, 0, rounded=(prec == "bf16"))) < bar
    finally:
        os.environ.pop("ADC_V2_NO_HSTACK", None)
        os.environ.pop("ADC_V2_NO_PAIR", None)
        dense.set_precision("bf16x3")


@gpu
@pytest.mark.parametrize("cfg", [(2, 48, 80, 57, 76, 32, 128), (1, 32, 112, 64, 80, 0, 128), (3, 64, 64, 29, 39, 0, 64), (2, 80, 32, 33, 47, 64, 128),
                                 (1, 160, 128, 30, 40, 0, 128)])
def test_row_segment_kernel_store_staging_forms(cfg):
    """bf16 (hi, lo) outputs leave the row-segment kernel through shared-memory staging slots and TMA bulk stores: two slots used in
    turn with one store group per 64-channel column block (default) against one slot staged and drained in turn (ADC_V2_NO_RING=1).
    Bit-identical tensors, many tiles per CTA (the slot ring wraps), one and two column blocks, channel-slice outputs."""
    import os
    from aerial_depth_completion_b200 import dense
    n, ci, co, h, w, off, cs = cfg
    g = torch.Generator().manual_seed(ci + 3 * co + h)
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    try:
        xa = dense.to_nhwc_bf16(x.cuda(), pad=True)
        wp = dense.prep_conv3x3_weights(wt.cuda(), cin_pad=ci)
        outs = []
        for flag in ("0", "1"):
            os.environ["ADC_V2_NO_RING"] = flag
            out = dense.new_act(n, h, w, cs, "cuda", pad=True, zero=True)
            dense.conv3x3_tc(xa, ci, wp, b.cuda(), co, 1, 1, out=out, out_off=off)
            outs.append((out.hi.clone(), out.lo.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        got = (outs[0][0].float() + outs[0][1].float())[:, 1:-1, 1:-1, off:off + co].permute(0, 3, 1, 2).cpu()
        ref = _ref(x, wt, b, 1, 1, rounded=False)
        assert float((got - ref).abs().max() / ref.abs().max()) < 5e-5
        if off > 0:
            assert float(outs[0][0][..., :off].abs().max()) == 0.0          # the neighbouring slice is untouched
    finally:
        os.environ.pop("ADC_V2_NO_RING", None)

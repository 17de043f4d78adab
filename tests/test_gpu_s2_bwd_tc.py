"""Backward of the stride-2 encoder convolutions on the tensor cores: parity-class data gradient and box-mode weight gradient
(csrc/conv_tc.cu adc_conv3x3_dgrad_s2_tc, csrc/conv_tc_bwd.cu adc_conv3x3_wgrad_box_tc) against fp64 autograd."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [(2, 12, 20, 16, 80), (1, 19, 27, 80, 80), (2, 15, 19, 80, 80), (1, 57, 76, 80, 80), (2, 8, 10, 64, 32)]


def _setup(case, prec, stride):
    from aerial_depth_completion_b200 import dense
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(h * 100 + w + cin)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    gy = torch.randn(n, cout, ho, wo, generator=g)
    return dense, n, h, w, cin, cout, x, wt, gy, ho, wo


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_ci%d_co%d" % (c[1], c[2], c[3], c[4]))
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
@pytest.mark.parametrize("pad", [True, False])
def test_dgrad_s2(case, prec, pad):
    dense, n, h, w, cin, cout, x, wt, gy, ho, wo = _setup(case, prec, 2)
    prev = dense.PRECISION
    dense.set_precision(prec)
    try:
        cg = dense._r16(cout)
        ga = dense.new_act(n, ho, wo, cg, "cuda", pad=pad, zero=True)
        dense.to_nhwc_bf16(gy.cuda(), cpad=cout, dst=ga, dst_off=0)
        wp = dense.prep_conv3x3_weights(wt.cuda(), flip_transpose=2, cin_pad=cg)
        gx = dense.conv3x3_dgrad_s2_tc(ga, cg, wp, cin, h, w, out_f32=True, out_pad=pad)
        got = dense.to_nchw_f32(gx, cin).cpu()
    finally:
        dense.set_precision(prev)
    xx = torch.zeros(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    rnd = (lambda t: t.bfloat16().double()) if prec == "bf16" else (lambda t: t.double())
    (ref,) = torch.autograd.grad(F.conv2d(xx, rnd(wt), None, 2, 1), xx, rnd(gy))
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-5, "relative max error %.3e" % err


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_ci%d_co%d" % (c[1], c[2], c[3], c[4]))
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
@pytest.mark.parametrize("stride,pad", [(2, True), (2, False), (1, True)])
def test_wgrad_box(case, prec, stride, pad):
    dense, n, h, w, cin, cout, x, wt, gy, ho, wo = _setup(case, prec, stride)
    prev = dense.PRECISION
    dense.set_precision(prec)
    try:
        xa = dense.new_act(n, h, w, dense._r64(cin), "cuda", pad=pad, zero=True)
        dense.to_nhwc_bf16(x.cuda(), cpad=cin, dst=xa, dst_off=0)
        ga = dense.new_act(n, ho, wo, dense._r64(cout), "cuda", pad=pad, zero=True)
        dense.to_nhwc_bf16(gy.cuda(), cpad=cout, dst=ga, dst_off=0)
        dw = dense.conv3x3_wgrad_box_tc(ga, xa, cin, cout, stride).cpu()
    finally:
        dense.set_precision(prev)
    ww = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    rnd = (lambda t: t.bfloat16().double()) if prec == "bf16" else (lambda t: t.double())
    (ref,) = torch.autograd.grad(F.conv2d(rnd(x), ww, None, stride, 1), ww, rnd(gy))
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-5, "relative max error %.3e" % err

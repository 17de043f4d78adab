"""Data gradient of a 3x3 / stride-1 convolution = the forward tcgen05 kernel on the output gradient with transposed, flipped
weights (adc_conv3x3_prep_weights flip_transpose=1), against the fp64 autograd input gradient of F.conv2d."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(2, 12, 20, 64, 64), (1, 19, 27, 80, 32), (2, 9, 14, 128, 64), (1, 16, 16, 33, 1)],
                         ids=lambda c: "ci%d_co%d" % (c[3], c[4]))
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_dgrad_is_fprop_with_flipped_weights(case, prec):
    from aerial_depth_completion_b200 import dense
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(cin + 7 * cout)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    gy = torch.randn(n, cout, h, w, generator=g)
    prev = dense.PRECISION
    dense.set_precision(prec)
    try:
        ga = dense.new_act(n, h, w, dense._r64(cout), "cuda", pad=True, zero=True)
        dense.to_nhwc_bf16(gy.cuda(), cpad=cout, dst=ga, dst_off=0)
        # the training path's helper: output channels in chunks the kernel's shared-memory budget takes (64 in the split mode)
        from aerial_depth_completion_b200.tc_train import _dgrad_s1
        gx = dense.new_act(n, h, w, dense._r64(cin), "cuda", pad=True, zero=True)
        _dgrad_s1(ga, wt.cuda(), gx, 0, cin)
        got = dense.to_nchw_f32(gx, cin).cpu()
    finally:
        dense.set_precision(prev)
    x = torch.zeros(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    if prec == "bf16":
        y = F.conv2d(x, wt.bfloat16().double(), None, 1, 1)
        (ref,) = torch.autograd.grad(y, x, gy.bfloat16().double())
        tol = 1e-2          # the output itself is rounded to bf16
    else:
        y = F.conv2d(x, wt.double(), None, 1, 1)
        (ref,) = torch.autograd.grad(y, x, gy.double())
        tol = 5e-5
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, "relative max error %.3e" % err

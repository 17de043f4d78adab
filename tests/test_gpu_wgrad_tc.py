"""tcgen05 weight-gradient kernel (csrc/conv_tc_bwd.cu) against the fp64 autograd weight gradient of F.conv2d."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # n, h, w, cin, cin_pad(act channels), cout, gy channels
    (2, 12, 20, 64, 64, 64, 64),
    (1, 9, 13, 16, 16, 32, 64),
    (2, 17, 23, 160, 192, 80, 128),
    (3, 30, 41, 4, 64, 64, 64),
    (1, 57, 76, 112, 128, 32, 64),
    (2, 8, 8, 128, 128, 128, 128),
    (2, 33, 47, 128, 128, 80, 128),
    (1, 60, 80, 96, 128, 128, 128),
]


def _ref(x, gy, wshape):
    w = torch.zeros(wshape, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), w, None, 1, 1)
    (g,) = torch.autograd.grad(y, w, gy.double())
    return g


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dx%d_ci%d_co%d" % (c[0], c[1], c[2], c[3], c[5]))
@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_wgrad_tc_matches_autograd(case, prec):
    from aerial_depth_completion_b200 import dense
    n, h, w, cin, cpad, cout, gch = case
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    gy = torch.randn(n, cout, h, w, generator=g)
    prev = dense.PRECISION
    dense.set_precision(prec)
    try:
        xa = dense.new_act(n, h, w, cpad, "cuda", pad=True, zero=True)
        dense.to_nhwc_bf16(x.cuda(), cpad=cin, dst=xa, dst_off=0)
        ga = dense.new_act(n, h, w, gch, "cuda", pad=True, zero=True)
        dense.to_nhwc_bf16(gy.cuda(), cpad=cout, dst=ga, dst_off=0)
        dw = dense.conv3x3_wgrad_tc(ga, xa, cin, cout, cin_pad=min(dense._r16(cin), cpad)).cpu()
        dw2 = dense.conv3x3_wgrad_tc(ga, xa, cin, cout, cin_pad=min(dense._r16(cin), cpad)).cpu()
    finally:
        dense.set_precision(prev)
    assert torch.equal(dw, dw2), "the reduction must be deterministic"
    if prec == "bf16":     # exact operands: compare with the gradient of the bf16-rounded tensors
        ref = _ref(x.bfloat16().float(), gy.bfloat16().float(), (cout, cin, 3, 3))
        tol = 2e-5
    else:
        ref = _ref(x, gy, (cout, cin, 3, 3))
        tol = 5e-5
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, "relative max error %.3e" % err


@pytest.mark.parametrize("switch", ["ADC_WGRAD_PAIR", "ADC_WGRAD_PLAIN"])
@pytest.mark.parametrize("case", [(2, 33, 47, 128, 128, 80, 128), (1, 57, 76, 112, 128, 32, 64), (2, 12, 20, 64, 64, 64, 64)],
                         ids=lambda c: "ci%d_co%d" % (c[3], c[5]))
def test_wgrad_tc_alternative_forms(case, switch):
    """The measured alternatives of the weight-gradient kernel stay correct behind their switches: the CTA-pair form
    (cta_group::2, gy_hi / gy_lo rows in the two CTAs) and the plain form (three MMAs per product, one tap per MMA)."""
    import os
    from aerial_depth_completion_b200 import dense
    n, h, w, cin, cpad, cout, gch = case
    g = torch.Generator().manual_seed(cin * 17 + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    gy = torch.randn(n, cout, h, w, generator=g)
    xa = dense.new_act(n, h, w, cpad, "cuda", pad=True, zero=True)
    dense.to_nhwc_bf16(x.cuda(), cpad=cin, dst=xa, dst_off=0)
    ga = dense.new_act(n, h, w, gch, "cuda", pad=True, zero=True)
    dense.to_nhwc_bf16(gy.cuda(), cpad=cout, dst=ga, dst_off=0)
    base = dense.conv3x3_wgrad_tc(ga, xa, cin, cout, cin_pad=min(dense._r16(cin), cpad)).cpu()
    os.environ[switch] = "1"
    try:
        alt = dense.conv3x3_wgrad_tc(ga, xa, cin, cout, cin_pad=min(dense._r16(cin), cpad)).cpu()
    finally:
        os.environ.pop(switch, None)
    ref = _ref(x, gy, (cout, cin, 3, 3))
    rel = lambda a: float((a.double() - ref).abs().max() / ref.abs().max())
    assert rel(alt) < 5e-5 and rel(base) < 5e-5, (rel(alt), rel(base))

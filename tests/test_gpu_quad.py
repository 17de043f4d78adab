"""GPU: the quad (pixel-shuffle) forms of GEDNet's full-resolution thin layers (csrc/ged_quad.cu + tc_train.py) against the
fp64 CPU oracle of the same net (oracle.nets_oracle.ged_forward): conv10 after the exact 2x up-sampling as a low-resolution
tensor-core convolution with 4 parity channels, conv1's data gradient likewise, the fused input packing.  Even, odd and headline
(228x304) sizes, 1 and 2 output channels, with and without the confidence-head features.  Per-tensor max-norm bars: outputs 1e-4,
gradients 5e-3 (the absolute term of the bf16x3 mode, golden_util.GRAD_ABS: forward activations exact to ~2e-5 flip ReLU masks of
near-zero activations ~100x more often than an fp32 run does)."""
import pytest
import torch

from golden_util import relmax
from oracle import synth

pytestmark = pytest.mark.gpu


def _ged(out_channels, seed):
    from aerial_depth_completion_b200.model_zoo.confidence_depth_framework import GEDNet
    m = GEDNet(out_channels=out_channels)
    # well-conditioned seeded weights (positive NConv weights): the constructor's Kaiming-initialised 1x1 NConv head can divide
    # by a near-zero weight sum, which turns every gradient into a measurement of rounding noise (values ~1e7)
    m.load_state_dict(synth.seeded_state({k: v.shape for k, v in m.state_dict().items()}, seed=seed))
    return m.cuda().train()


def _run(m, x, build, wts):
    for p in m.parameters():
        p.grad = None
    x = x.clone().requires_grad_(True)
    d, feat = m(x, build)
    loss = (d * wts[0]).sum()
    if feat is not None:
        loss = loss + (feat * wts[1][:, :feat.shape[1]]).sum()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    return d.detach(), None if feat is None else feat.detach(), x.grad.clone(), grads


def _oracle(state, x, build, wts):
    from oracle import nets_oracle as nets
    S = {k: v.detach().double().cpu().clone().requires_grad_(True) for k, v in state.items()}
    x = x.detach().double().cpu().clone().requires_grad_(True)
    d, feat = nets.ged_forward(S, "", x, build, True)
    loss = (d * wts[0].double().cpu()).sum()
    if feat is not None:
        loss = loss + (feat * wts[1][:, :feat.shape[1]].double().cpu()).sum()
    loss.backward()
    return d.detach(), None if feat is None else feat.detach(), x.grad, {k: v.grad for k, v in S.items() if v.grad is not None}


@pytest.mark.parametrize("shape", [(2, 48, 64), (1, 228, 304), (1, 37, 52)], ids=lambda s: "%dx%dx%d" % s)
@pytest.mark.parametrize("oc,build", [(1, False), (1, True), (2, True)])
def test_ged_tc_training_matches_fp64_oracle(shape, oc, build):
    n, h, w = shape
    m = _ged(oc, 5)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    inp, _ = synth.synth_frames(n, h, w, seed=31, num_samples=max(60, h * w // 140))
    x = inp.cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    # positive loss weights: a random-sign loss makes every parameter gradient a cancelling sum, which measures rounding noise
    wts = (torch.rand(n, 1, h, w, device="cuda", generator=g) + 0.5, torch.rand(n, 36, h, w, device="cuda", generator=g) * 0.1)
    ref = _oracle(state, x, build, wts)
    got = _run(m, x, build, wts)
    assert relmax(got[0].cpu(), ref[0]) < 1e-4
    if ref[1] is not None:
        assert relmax(got[1].cpu(), ref[1]) < 1e-4
    # input gradient: only the depth / confidence channels carry one through the NConv pyramid and conv1
    assert relmax(got[2][:, 3:].cpu(), ref[2][:, 3:]) < 2e-2, relmax(got[2][:, 3:].cpu(), ref[2][:, 3:])
    assert set(got[3]) == set(ref[3])
    bad = []
    for k in ref[3]:
        e = relmax(got[3][k].cpu(), ref[3][k])
        # golden_util.GRAD_ABS["bf16x3"] for the dense layers (measured 5e-6 .. 3e-3); the NConv parameters and the input gradient are
        # cancelling sums that amplify the split mode's unit round-off (golden_util.GRAD_AMP; measured up to 2.4e-2 at 48x64: nconv7.weight)
        if e >= (5e-2 if k.startswith("d_net.") else 5e-3):
            bad.append((k, e))
    assert not bad, bad


def test_pack_input_is_cat_plus_layout():
    from aerial_depth_completion_b200 import dense
    n, h, w = 2, 21, 30
    g = torch.Generator(device="cuda").manual_seed(1)
    full = torch.randn(n, 5, h, w, device="cuda", generator=g)
    xo, co = torch.randn(n, 1, h, w, device="cuda", generator=g), torch.rand(n, 1, h, w, device="cuda", generator=g)
    try:
        dense.set_precision("bf16x3")
        a = dense.new_act(n, h, w, 16, "cuda", pad=True, zero=True)
        b = dense.new_act(n, h, w, 16, "cuda", pad=True, zero=True)
        dense.ged_pack_input(xo, co, full[:, :3], a)
        dense.to_nhwc_bf16(torch.cat((xo, full[:, :3], co), 1), cpad=16, dst=b, dst_off=0)
    finally:
        dense.set_precision("bf16x3")
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)


@pytest.mark.parametrize("a,b", [(80, 80), (80, 2), (2, 32), (1, 32), (64, 129)])
def test_quad_tap_mix_vs_tap_maps(a, b):
    """adc_quad_tap_mix / adc_quad_tap_mix_adjoint against the broadcast product with the constant 0/1 maps they replace
    (dense.quad_maps): 's2' copies single taps (bit-exact), 'up' sums up to four (fp32 association may differ)."""
    from aerial_depth_completion_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(7 + a + b)
    w = torch.randn(a, b, 3, 3, device="cuda", generator=g)
    up, s2 = dense.quad_maps(w.device)
    for which, maps in (("up", up), ("s2", s2)):
        ref = (w.reshape(1, a, b, 1, 9) * maps.reshape(4, 1, 1, 9, 9)).sum(-1)            # [4, a, b, 9]
        got = dense.quad_tap_mix(w, which)
        got_t = dense.quad_tap_mix(w, which, transposed=True)
        if which == "s2":
            assert torch.equal(got, ref) and torch.equal(got_t, ref.permute(0, 2, 1, 3).contiguous())
        else:
            assert relmax(got, ref) < 1e-6 and relmax(got_t, ref.permute(0, 2, 1, 3).contiguous()) < 1e-6
    # a strided slice (conv10's first 32 input channels) goes through the same entry point
    w33 = torch.randn(a, b + 1, 3, 3, device="cuda", generator=g)
    ref = (w33[:, :b].reshape(1, a, b, 1, 9) * up.reshape(4, 1, 1, 9, 9)).sum(-1)
    assert relmax(dense.quad_tap_mix(w33[:, :b], "up"), ref) < 1e-6
    gq = torch.randn(4, a * b, 9, device="cuda", generator=g)
    ref = (gq.reshape(4, a * b, 9, 1) * up.reshape(4, 1, 9, 9)).sum((0, 2))
    assert relmax(dense.quad_tap_mix_adjoint(gq), ref) < 1e-6

"""GPU: the cbr3-c1 confidence head on the B200 path (tcgen05 conv1 + fused BN / LeakyReLU / 1x1 / sigmoid kernels, forward and
backward) against the same module evaluated with plain torch ops in fp64 on the CPU (the reference's definition,
confidence_depth_framework.py:299-329).  bf16x3 mode must meet the fp32 bars: outputs 1e-4, gradients 2e-3 (max-norm relative)."""
import copy

import pytest
import torch

from golden_util import relmax

gpu = pytest.mark.gpu


def _ref_forward(head, x):
    """The reference's forward (conv1 -> bn1 -> LeakyReLU(0.2) -> conv2 -> sigmoid) with plain torch ops on the CPU."""
    import torch.nn.functional as F
    y = F.conv2d(x, head.conv1.weight, None, 1, 1)
    y = F.leaky_relu(head.bn1(y), 0.2)
    return torch.sigmoid(F.conv2d(y, head.conv2.weight, head.conv2.bias))


@gpu
@pytest.mark.parametrize("c,shape", [(36, (2, 30, 44)), (50, (1, 57, 76)), (36, (3, 64, 80))])
@pytest.mark.parametrize("training", [True, False])
def test_conf_head_tc_vs_fp64_reference(c, shape, training):
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense, _lib
    n, h, w = shape
    torch.manual_seed(5 + c)
    head = mc.CBR3C1Confidence(c)
    with torch.no_grad():      # weights large enough that the sigmoid is not flat, non-trivial BN affine and running statistics
        head.conv1.weight.mul_(8.0)
        head.conv2.weight.mul_(20.0)
        head.bn1.weight.uniform_(0.5, 1.5)
        head.bn1.bias.uniform_(-0.3, 0.3)
        head.bn1.running_mean.uniform_(-0.2, 0.2)
        head.bn1.running_var.uniform_(0.5, 2.0)
    ref = copy.deepcopy(head).double().train(training)
    dev = copy.deepcopy(head).cuda().train(training)
    x = torch.randn(n, c, h, w)
    gout = torch.randn(n, 1, h, w)
    xr = x.double().requires_grad_(True)
    yr = _ref_forward(ref, xr)
    yr.backward(gout.double())
    xg = x.cuda().requires_grad_(True)
    n0 = _lib.lib.adc_launch_count()
    try:
        dense.set_precision("bf16x3")
        yg = dev(xg)
        yg.backward(gout.cuda())
    finally:
        dense.set_precision("bf16x3")
    assert _lib.lib.adc_launch_count() - n0 >= 8, "the confidence head did not run on the library's kernels"
    assert relmax(yg.detach().cpu(), yr.detach()) < 1e-4
    assert relmax(xg.grad.cpu(), xr.grad) < 2e-3
    for (name, pr), pg in zip(ref.named_parameters(), dev.parameters()):
        assert relmax(pg.grad.cpu(), pr.grad) < 2e-3, name
    for name in ("running_mean", "running_var"):
        assert relmax(getattr(dev.bn1, name).cpu(), getattr(ref.bn1, name)) < 1e-5, name
    assert int(dev.bn1.num_batches_tracked) == int(ref.bn1.num_batches_tracked)


@gpu
def test_conf_head_no_input_grad_and_no_grad_mode():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense
    torch.manual_seed(3)
    head = mc.CBR3C1Confidence(36)
    ref = copy.deepcopy(head).double().eval()
    dev = copy.deepcopy(head).cuda().eval()
    x = torch.randn(2, 36, 20, 28)
    try:
        dense.set_precision("bf16x3")
        with torch.no_grad():
            y = dev(x.cuda())
        dev.train()
        y2 = dev(x.cuda())          # input without requires_grad: only parameter gradients
        y2.sum().backward()
    finally:
        dense.set_precision("bf16x3")
    assert relmax(y.cpu(), _ref_forward(ref, x.double()).detach()) < 1e-4
    assert dev.conv1.weight.grad is not None and torch.isfinite(dev.conv1.weight.grad).all()

"""GPU: the masked-loss and depth-metric kernels (csrc/loss_metrics.cu) against the CPU oracle (oracle/loss_metrics_oracle.py, the
reference's definitions confidence_depth_framework.py:829-914 and metrics.py:38-66) evaluated in fp64 on the same tensors.
Tolerance 1e-5 relative for the scalars (fp32 per-element arithmetic, double accumulation), 1e-5 max-norm for the gradient."""
import pytest
import torch

from golden_util import relmax
from oracle import loss_metrics_oracle as lmo

gpu = pytest.mark.gpu


def _frames(seed, shape=(3, 1, 57, 76), holes=0.3):
    g = torch.Generator().manual_seed(seed)
    target = torch.rand(shape, generator=g) * 9.0 + 0.5
    target[torch.rand(shape, generator=g) < holes] = 0.0        # invalid pixels
    pred = target + torch.randn(shape, generator=g) * 0.4
    pred = pred.abs() + 0.05                                     # positive predictions: the log / inverse metrics stay finite
    return pred, target


@gpu
@pytest.mark.parametrize("kind", ["l2", "l1", "il1", "absrel"])
def test_masked_loss_value_and_gradient(kind):
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    cls = {"l2": mc.MaskedMSELoss, "l1": mc.MaskedL1Loss, "il1": mc.InvertedMaskedL1Loss, "absrel": mc.MaskedAbsRelLoss}[kind]
    pred, target = _frames(7)
    pr = pred.double().requires_grad_(True)
    ref = lmo.masked_loss(kind, pr, target.double())
    (ref * 3.0).backward()
    pg = pred.cuda().requires_grad_(True)
    crit = cls()
    loss = crit(None, pg, target.cuda())
    (loss * 3.0).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert abs(crit.loss[0] - float(ref)) <= 1e-5 * abs(float(ref))
    assert relmax(pg.grad.cpu(), pr.grad) < 1e-5
    assert bool((pg.grad.cpu()[target == 0] == 0).all())


@gpu
def test_masked_loss_on_channel_slice_and_too_few_valid_pixels():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    pred, target = _frames(9, shape=(2, 1, 40, 50))
    two = torch.cat((pred, pred * 2), 1).cuda().requires_grad_(True)          # non-contiguous channel slice like pred[:, 0:1]
    loss = mc.MaskedMSELoss()(None, two[:, 0:1], target.cuda())
    loss.backward()
    ref = lmo.masked_loss("l2", pred.double(), target.double())
    assert abs(float(loss) - float(ref)) <= 1e-5 * float(ref)
    assert bool((two.grad[:, 1] == 0).all())
    sparse = torch.zeros_like(target)
    sparse[0, 0, 0, :50] = 1.0                                                # 50 valid pixels: the reference asserts > 100
    with pytest.raises(AssertionError):
        mc.MaskedMSELoss()(None, pred.cuda(), sparse.cuda())


@gpu
@pytest.mark.parametrize("with_conf", [False, True])
def test_depth_metrics_vs_oracle(with_conf):
    from aerial_depth_completion_b200.metrics import Result
    pred, target = _frames(11, shape=(2, 1, 114, 152))
    conf = torch.rand(pred.shape, generator=torch.Generator().manual_seed(3)) if with_conf else None
    ref = lmo.evaluate(pred.double(), target.double(), None if conf is None else conf.double(), 0.4)
    r = Result(threshold=0.4)
    r.evaluate(pred.cuda(), target.cuda(), None if conf is None else conf.cuda())
    for k in ("mse", "rmse", "mae", "lg10", "absrel", "delta1", "delta2", "delta3", "irmse", "imae"):
        assert abs(getattr(r, k) - ref[k]) <= 2e-5 * abs(ref[k]) + 1e-9, (k, getattr(r, k), ref[k])


@gpu
def test_lazy_loss_matches_eager():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    g = torch.Generator().manual_seed(3)
    t = torch.rand(2, 1, 24, 32, generator=g) * 5
    t[t < 1] = 0
    t = t.cuda()
    p = (t * 1.1 + 0.2).requires_grad_(True)
    for cls in (mc.MaskedMSELoss, mc.MaskedL1Loss, mc.InvertedMaskedL1Loss, mc.MaskedAbsRelLoss):
        a, b = cls(), cls()
        b.lazy = True                 # device-side bookkeeping: no .item(), asynchronous assert
        la, lb = a(None, p, t), b(None, p, t)
        assert torch.equal(la, lb) and isinstance(a.loss[0], float) and torch.is_tensor(b.loss[0])


@gpu
def test_losses_vs_reference_golden():
    """Loss values and gradients produced by the unmodified reference (tests/golden/make_golden.py)."""
    from golden_util import gold
    from oracle import synth
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    G = gold("sampling_metrics")
    for kind, cls in (("l2", mc.MaskedMSELoss), ("l1", mc.MaskedL1Loss), ("il1", mc.InvertedMaskedL1Loss), ("absrel", mc.MaskedAbsRelLoss)):
        inp, tgt = synth.synth_frames(1, 24, 32, seed=9)
        tgt = (tgt * (tgt > 2).float()).float().cuda()
        pred = (tgt * 1.1 + 0.3).requires_grad_(True)
        l = cls()(None, pred, tgt)
        l.backward()
        assert abs(float(l) - G["loss_" + kind][0]) < 1e-5 * abs(G["loss_" + kind][0])
        assert relmax(pred.grad.cpu(), G["loss_" + kind + "_grad"]) < 1e-5


@gpu
def test_metrics_vs_reference_golden():
    """KAT-6 of SURVEY.md 8c: Result.evaluate on five hand-checked pixels."""
    import numpy as np
    from golden_util import gold
    from aerial_depth_completion_b200.metrics import Result
    G = gold("sampling_metrics")
    r = Result()
    r.evaluate(torch.tensor([1., 2., 4., 8., 3.]).cuda(), torch.tensor([1.1, 0, 5.5, 8., 2.]).cuda())
    got = [r.rmse, r.mae, r.delta1, r.delta2, r.delta3, r.absrel, r.lg10, r.irmse, r.imae, r.mse]
    assert np.allclose(got, G["kat6_metrics"], rtol=1e-5)

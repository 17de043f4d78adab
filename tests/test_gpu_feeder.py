"""GPU: parallel.HostFeeder (host->device copies overlapped with the previous step, losses read one step late) must produce the
same training trajectory as stepping batch by batch the way the reference trainer does (trainer.py:194-233)."""
import contextlib
import copy
import io

import pytest
import torch

gpu = pytest.mark.gpu


@gpu
def test_host_feeder_matches_serial_stepping():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import parallel
    from oracle import synth
    f = mc.ConfidenceDepthFrameworkFactory()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m1 = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
        m2 = copy.deepcopy(m1)
        (m1, p1), (m2, p2) = f.to_device(m1), f.to_device(m2)
        m1.train()
        m2.train()
    batches = []
    for k in range(5):
        inp, tgt = synth.synth_frames(2, 64, 80, seed=40 + k, num_samples=150)
        batches.append((inp.float().pin_memory(), tgt.float().pin_memory()))
    t1 = parallel.DataParallelTrainer(m1, f.create_loss("l2", False, 0.5)[0], torch.optim.SGD(p1, lr=1e-6), p1)
    t2 = parallel.DataParallelTrainer(m2, f.create_loss("l2", False, 0.5)[0], torch.optim.SGD(p2, lr=1e-6), p2)
    serial = [float(t1.step(a.cuda(), b.cuda())) for a, b in batches]
    fed = parallel.HostFeeder(t2, *batches[0]).run(batches)
    assert len(fed) == len(serial)
    for a, b in zip(fed, serial):
        assert abs(a - b) <= 1e-5 * abs(b), (fed, serial)
    for a, b in zip(p1, p2):
        # SGD, not Adam: the first Adam steps are lr * sign(g), which turns run-to-run rounding noise of near-zero gradients
        # (library fp32 convolutions are not bit-reproducible) into O(lr) parameter differences
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)

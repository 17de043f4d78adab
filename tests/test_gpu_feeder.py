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


@gpu
@pytest.mark.parametrize("hw", [(64, 80), (61, 47)])
def test_device_assembly_is_bit_exact_with_the_host_loader(hw):
    """adc_sample_assemble: (r, g, b, gt) + per-frame seeds -> (r, g, b, d, w) on the device == the reference loader's host-side
    UniformSampling + modality assembly (dense_to_sparse.py:29-48, dataloader_ext.py:544-566), bit for bit, incl. an odd size
    (scalar copy path) and a frame with invalid pixels."""
    import numpy as np
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import UniformSampling, sample_assemble_cuda
    H, W = hw
    g = torch.Generator().manual_seed(5)
    rgb = torch.rand(3, 3, H, W, generator=g)
    gt = torch.rand(3, 1, H, W, generator=g) * 9 + 1
    gt[1, 0, : H // 3] = 0                                   # invalid region: n_keep < HW changes the probability
    seeds = [7, 0, 2 ** 32 - 1]
    sampler = UniformSampling(150, np.inf)
    want = []
    for n in range(3):
        np.random.seed(seeds[n])
        mask = torch.from_numpy(sampler.dense_to_sparse(None, gt[n, 0].numpy())).float().unsqueeze(0)
        want.append(torch.cat((rgb[n], gt[n] * mask, mask), 0))
    want = torch.stack(want)
    got = sample_assemble_cuda(torch.cat((rgb, gt), 1).cuda(), seeds, 150)
    assert torch.equal(got.cpu(), want)
    assert 100 < int(got[:, 4].sum()) < 3 * 220


@gpu
def test_device_assembly_feeder_matches_host_feeder():
    """parallel.DeviceAssemblyFeeder (4 planes over PCIe, masks drawn on the device) == parallel.HostFeeder fed with the host-assembled
    batches: identical losses (the inputs are bit-identical, the steps are the same kernels)."""
    import numpy as np
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import parallel
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import UniformSampling
    f = mc.ConfidenceDepthFrameworkFactory()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m1 = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
        m2 = copy.deepcopy(m1)
        (m1, p1), (m2, p2) = f.to_device(m1), f.to_device(m2)
        m1.train()
        m2.train()
    g = torch.Generator().manual_seed(9)
    sampler = UniformSampling(150, np.inf)
    host, raw = [], []
    for k in range(4):
        rgb = torch.rand(2, 3, 64, 80, generator=g)
        gt = torch.rand(2, 1, 64, 80, generator=g) * 9 + 1
        seeds = [100 + 2 * k, 101 + 2 * k]
        masks = []
        for n in range(2):
            np.random.seed(seeds[n])
            masks.append(torch.from_numpy(sampler.dense_to_sparse(None, gt[n, 0].numpy())).float().unsqueeze(0))
        mask = torch.stack(masks)
        host.append((torch.cat((rgb, gt * mask, mask), 1).pin_memory(), gt.clone().pin_memory()))
        raw.append((torch.cat((rgb, gt), 1).pin_memory(), torch.tensor(seeds, dtype=torch.int32).pin_memory()))
    t1 = parallel.DataParallelTrainer(m1, f.create_loss("l2", False, 0.5)[0], torch.optim.SGD(p1, lr=1e-6), p1)
    t2 = parallel.DataParallelTrainer(m2, f.create_loss("l2", False, 0.5)[0], torch.optim.SGD(p2, lr=1e-6), p2)
    a = parallel.HostFeeder(t1, *host[0]).run(host)
    b = parallel.DeviceAssemblyFeeder(t2, raw[0][0], 150).run(raw)
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        assert abs(x - y) <= 1e-6 * abs(x), (a, b)

"""CPU suite: the N>1 data-parallel path (flat gradient bucket + all-reduce) on gloo, world_size 2, and the masked
losses' device-side ('lazy') mode.  The model here is a tiny torch module: the logic under test is host-side."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _tiny_model():
    """Shaped like the framework model: dc_model with a d_net part and a body, plus a parameter no gradient ever reaches (its
    bucket must still be reduced)."""
    class DC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.d_net = torch.nn.Conv2d(5, 2, 3, 1, 1)
            self.conv1 = torch.nn.Conv2d(2, 4, 3, 1, 1)
            self.conv2 = torch.nn.Conv2d(4, 1, 3, 1, 1)

        def forward(self, x):
            return self.conv2(torch.relu(self.conv1(self.d_net(x))))

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dc_model = DC()
            self.conf_model = torch.nn.Conv2d(1, 1, 1)         # never used in forward: no gradient

        def forward(self, x):
            return self.dc_model(x), None, None
    return Tiny()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aerial_depth_completion_b200 import parallel
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc

    torch.manual_seed(rank)                       # deliberately different init: the trainer must broadcast rank 0's
    model = _tiny_model()
    params = list(model.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    tr = parallel.DataParallelTrainer(model, mc.MaskedMSELoss(), opt, params, world_size=world)
    out["buckets%d" % rank] = [(lo, hi, n) for lo, hi, n in tr._buckets]
    g = torch.Generator().manual_seed(100)
    x = torch.rand(4, 5, 16, 16, generator=g)
    t = torch.rand(4, 1, 16, 16, generator=g) + 0.5
    shard = slice(rank * 2, rank * 2 + 2)
    for _ in range(3):
        tr.step(x[shard], t[shard])
    out[rank] = [p.detach().clone() for p in model.parameters()]
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_on_concatenated_batch():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    # single process, full batch, same initial weights as rank 0
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    torch.manual_seed(0)
    c = _tiny_model()
    opt = torch.optim.SGD(c.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100)
    x = torch.rand(4, 5, 16, 16, generator=g)
    t = torch.rand(4, 1, 16, 16, generator=g) + 0.5
    crit = mc.MaskedMSELoss()
    for _ in range(3):
        opt.zero_grad()
        crit(x, c(x)[0], t).backward()
        opt.step()
    ref = [p.detach() for p in c.parameters()]
    # three gradient buckets: dc_model.d_net | dc_model body (conv1, conv2) | conf_model
    assert [n for _, _, n in out["buckets0"]] == [2, 4, 2] and out["buckets0"] == out["buckets1"]
    for r in (0, 1):
        for a, b in zip(out[r], ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), "rank %d diverged from the single-process run" % r


def test_lazy_loss_matches_eager():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    g = torch.Generator().manual_seed(3)
    t = torch.rand(2, 1, 24, 32, generator=g) * 5
    t[t < 1] = 0
    p = (t * 1.1 + 0.2).requires_grad_(True)
    for cls in (mc.MaskedMSELoss, mc.MaskedL1Loss, mc.InvertedMaskedL1Loss, mc.MaskedAbsRelLoss):
        a, b = cls(), cls()
        b.lazy = True
        la, lb = a(None, p, t), b(None, p, t)
        assert torch.equal(la, lb) and isinstance(a.loss[0], float) and torch.is_tensor(b.loss[0])
    with pytest.raises(AssertionError):
        mc.MaskedMSELoss()(None, p[:, :, :2, :2], t[:, :, :2, :2])


def test_losses_vs_golden():
    from golden_util import gold, relmax
    from oracle import synth
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    G = gold("sampling_metrics")
    for kind, cls in (("l2", mc.MaskedMSELoss), ("l1", mc.MaskedL1Loss), ("il1", mc.InvertedMaskedL1Loss), ("absrel", mc.MaskedAbsRelLoss)):
        inp, tgt = synth.synth_frames(1, 24, 32, seed=9)
        tgt = tgt * (tgt > 2).float()
        pred = (tgt * 1.1 + 0.3).requires_grad_(True)
        l = cls()(None, pred, tgt)
        l.backward()
        assert abs(float(l) - G["loss_" + kind][0]) < 1e-6 * abs(G["loss_" + kind][0])
        assert relmax(pred.grad, G["loss_" + kind + "_grad"]) < 1e-6


def test_metrics_vs_golden():
    from golden_util import gold
    import numpy as np
    from aerial_depth_completion_b200.metrics import Result
    G = gold("sampling_metrics")
    r = Result()
    r.evaluate(torch.tensor([1., 2., 4., 8., 3.]), torch.tensor([1.1, 0, 5.5, 8., 2.]))
    got = [r.rmse, r.mae, r.delta1, r.delta2, r.delta3, r.absrel, r.lg10, r.irmse, r.imae, r.mse]
    assert np.allclose(got, G["kat6_metrics"], rtol=1e-6)

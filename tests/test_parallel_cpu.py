"""CPU suite: the N>1 data-parallel path (gradient buckets + all-reduce) on gloo, world_size 2.  The model and the criterion here are
tiny torch modules: the logic under test is host-side (the product's losses are CUDA kernels and refuse host tensors)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _MaskedMSE(torch.nn.Module):
    """Host stand-in with the criterion call signature of the framework (input, prediction, target)."""

    def forward(self, inp, pred, target, epoch=None):
        mask = target > 0
        return ((target - pred) ** 2)[mask].mean()


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


def _data():
    """Sparse targets with DIFFERENT numbers of valid pixels per shard (frames 0-1: ~80 %, frames 2-3: ~25 %): the mean of the
    per-rank masked means is then not the masked mean of the whole batch, which is what the reference's DataParallel computes
    (trainer.py:204-208)."""
    g = torch.Generator().manual_seed(100)
    x = torch.rand(4, 5, 16, 16, generator=g)
    t = torch.rand(4, 1, 16, 16, generator=g) + 0.5
    keep = torch.rand(4, 1, 16, 16, generator=g) < torch.tensor([0.8, 0.8, 0.25, 0.25]).view(4, 1, 1, 1)
    return x, t * keep


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aerial_depth_completion_b200 import parallel
    torch.manual_seed(rank)                       # deliberately different init: the trainer must broadcast rank 0's
    model = _tiny_model()
    params = list(model.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    tr = parallel.DataParallelTrainer(model, _MaskedMSE(), opt, params, world_size=world)
    out["buckets%d" % rank] = [(lo, hi, n) for lo, hi, n in tr._buckets]
    x, t = _data()
    shard = slice(rank * 2, rank * 2 + 2)
    losses = []
    for _ in range(3):
        losses.append(float(tr.step(x[shard], t[shard])))
    out[rank] = [p.detach().clone() for p in model.parameters()]
    out["loss%d" % rank] = losses
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_on_concatenated_batch():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    # single process, full batch, same initial weights as rank 0
    torch.manual_seed(0)
    c = _tiny_model()
    opt = torch.optim.SGD(c.parameters(), lr=0.1)
    x, t = _data()
    assert int((t[:2] > 0).sum()) != int((t[2:] > 0).sum())
    crit = _MaskedMSE()
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = crit(x, c(x)[0], t)
        loss.backward()
        opt.step()
        ref_losses.append(float(loss))
    ref = [p.detach() for p in c.parameters()]
    # three gradient buckets: dc_model.d_net | dc_model body (conv1, conv2) | conf_model
    assert [n for _, _, n in out["buckets0"]] == [2, 4, 2] and out["buckets0"] == out["buckets1"]
    for r in (0, 1):
        for a, b in zip(out[r], ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), "rank %d diverged from the single-process run" % r
        # the returned loss is the masked mean over the GLOBAL batch, identical on both ranks
        assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(out["loss%d" % r], ref_losses))


def test_product_losses_and_metrics_refuse_host_tensors():
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200.metrics import Result
    t = torch.rand(2, 1, 24, 32) + 0.5
    with pytest.raises(RuntimeError):
        mc.MaskedMSELoss()(None, t * 1.1, t)
    with pytest.raises(RuntimeError):
        Result().evaluate(t * 1.1, t)


def test_single_process_gradient_adoption_equals_accumulation(monkeypatch):
    """World 1: the trainer lets autograd ADOPT the incoming gradient tensors (.grad = None before backward) and moves them into the
    flat bucket with one multi-tensor copy, instead of one accumulate launch per parameter into zeroed views (ADC_ADOPT_GRADS=0).
    Same parameters after three steps, the flat views are back in place for the optimizer, and a parameter no gradient reaches
    (conf_model) ends every step with a zero gradient like after zero_grad."""
    from aerial_depth_completion_b200 import parallel
    x, t = _data()
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ADC_ADOPT_GRADS", flag)
        torch.manual_seed(0)
        model = _tiny_model()
        params = list(model.parameters())
        opt = torch.optim.SGD(params, lr=0.1)
        tr = parallel.DataParallelTrainer(model, _MaskedMSE(), opt, params, world_size=1)
        assert tr._adopt == (flag == "1")
        losses = [float(tr.step(x, t)) for _ in range(3)]
        for p, v in zip(tr.params, tr._views):
            assert p.grad is v and v.data_ptr() >= tr.flat.data_ptr()
        assert float(model.conf_model.weight.grad.abs().max()) == 0.0
        res[flag] = ([p.detach().clone() for p in params], losses, tr.flat.clone())
    for a, b in zip(res["0"][0], res["1"][0]):
        assert torch.equal(a, b)
    assert res["0"][1] == res["1"][1] and torch.equal(res["0"][2], res["1"][2])

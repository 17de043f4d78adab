"""GPU parity tests of the path bench.py times: ``DataParallelTrainer.capture()`` -> ONE CUDA graph per step holding the forward, the
lazy loss, the backward with its two side streams (the pyramid's weight-gradient stream and tc_train's), and ``optim.FlatAdam``.

* the replayed graph == the same steps issued eagerly on ONE stream (losses and every parameter), at the headline shape
  (ged_depthcompnet dc1_only, 228x304, bf16x3) for B = 2 and B = 32;
* step 1 == the fp64 CPU oracle (outputs 1e-4, per-tensor gradients by the k-times-floor rule of SURVEY.md 8c);
* a learning-rate schedule takes effect under replay, and a non-finite loss skips the update instead of poisoning the moments;
* two ranks on half batches with different valid-pixel counts == one process on the whole batch, over NCCL (needs 2 GPUs).
"""
import contextlib
import io
import os
import subprocess
import sys

import pytest
import torch

from golden_util import gold, grad_rule_violations, net_case, ref_grad_errors, relmax, run_oracle_case

gpu = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 228, 304


def _make(state=None, arch="ged_depthcompnet", mode="dc1_only", lr=1e-3, seed=0):
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import parallel
    from aerial_depth_completion_b200.optim import FlatAdam
    torch.manual_seed(seed)
    f = mc.ConfidenceDepthFrameworkFactory()
    with contextlib.redirect_stdout(io.StringIO()):
        m = f.create_model("rgbd", mode, arch, "")
        if state is not None:
            m.load_state_dict(state)
        m, params = f.to_device(m)
        m.train()
    crit, _ = f.create_loss("l2", False, 0.5)
    opt = FlatAdam(params, lr=lr)
    return m, opt, parallel.DataParallelTrainer(m, crit, opt, params, world_size=1)


def _frames(B, n, seed=77):
    from oracle import synth
    out = []
    for i in range(n):
        inp, tgt = synth.synth_frames(B, H, W, seed=seed + 100 * i)
        out.append((inp.cuda(), tgt.cuda()))
    return out


@gpu
@pytest.mark.parametrize("B", [2, 32])
def test_graph_replay_with_side_streams_equals_eager_single_stream(B):
    from aerial_depth_completion_b200 import _lib, dense, tc_train
    assert dense.PRECISION == "bf16x3", "the B200 path must be the default"
    data = _frames(B, 3)
    # A: eager, every kernel on one stream
    prev_side, prev_lib = tc_train._SIDE_STREAM, _lib.lib.adc_set_side_stream(0)
    tc_train._SIDE_STREAM = False
    try:
        ma, _, ta = _make()
        state0 = {k: v.detach().clone() for k, v in ma.state_dict().items()}
        la = [float(ta.step(x, t)) for x, t in data]
    finally:
        tc_train._SIDE_STREAM = prev_side
        _lib.lib.adc_set_side_stream(prev_lib)
    # B: the benchmarked configuration -- side streams on, one captured graph, replayed
    tc_train._SIDE_STREAM = True
    _lib.lib.adc_set_side_stream(1)
    mb, ob, tb = _make(state0)
    # capture() runs warm-up steps that train the model: restore parameters, moments and the step counter afterwards
    tb.capture(data[0][0], data[0][1])
    with torch.no_grad():
        mb.load_state_dict(state0)
        ob.flat_exp_avg.zero_()
        ob.flat_exp_avg_sq.zero_()
        ob.step_t.zero_()
    lb = [float(tb.step(x, t)) for x, t in data]
    torch.cuda.synchronize()
    assert all(v == v for v in la + lb)
    for a, b in zip(la, lb):
        assert abs(a - b) <= 1e-6 * abs(a), (la, lb)
    sa, sb = ma.state_dict(), mb.state_dict()
    worst = max(relmax(sb[k], sa[k]) for k in sa)
    assert worst <= 1e-6, worst      # same kernels, same deterministic reductions: the schedules may not change a result
    assert float(ob.step_t) == 3.0


@gpu
def test_benchmarked_step_vs_fp64_oracle():
    """ged_nyu fixture (B = 2, 228x304, dc1_only, l2): captured-graph step against the fp64 oracle and the reference's goldens."""
    G = gold("nets")
    cfg, state, inp, tgt = net_case("ged_nyu")
    m, opt, tr = _make({k: v.clone() for k, v in state.items()}, lr=0.0)      # lr 0: the step leaves the parameters where the oracle has them
    x, t = inp.cuda(), tgt.cuda()
    tr.capture(x, t)
    with torch.no_grad():
        m.load_state_dict(state)       # the warm-up forwards applied the softplus hooks to the NConv weights: start again
    loss = float(tr.step(x, t))
    torch.cuda.synchronize()
    _, S, _, _, d1, _, _, loss64, _ = run_oracle_case("ged_nyu", torch.float64)
    assert abs(loss - float(loss64)) <= 1e-4 * abs(float(loss64)), (loss, float(loss64))
    assert abs(loss - float(G["ged_nyu_loss"][0])) <= 1e-4 * abs(float(G["ged_nyu_loss"][0]))
    floor = ref_grad_errors("ged_nyu")
    names = {id(p): n for n, p in m.named_parameters()}
    bad = grad_rule_violations([(names[id(p)], p.grad) for p in tr.params], S, floor, "bf16x3")
    assert not bad, bad
    # outputs of the first training forward of a fresh copy against the oracle and the reference's golden
    m2, _, _ = _make({k: v.clone() for k, v in state.items()})
    out = m2(x)[0]
    assert relmax(out.detach().cpu(), d1.detach()) <= 1e-4
    assert relmax(out.detach().cpu(), G["ged_nyu_d1"]) <= 1e-4


@gpu
def test_step_lr_takes_effect_under_graph_replay():
    """trainer.py:236-241 decays lr by 10 every 5 epochs; a captured step must follow ``param_groups``."""
    data = _frames(2, 1)
    x, t = data[0]
    m, opt, tr = _make(lr=1e-2)
    tr.capture(x, t)
    p0 = opt.flat_param.clone()
    tr.step(x, t)
    d1 = (opt.flat_param - p0).abs().max().item()
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.0)
    sched.step()                                   # lr -> 0
    assert opt.param_groups[0]["lr"] == 0.0
    p1 = opt.flat_param.clone()
    tr.step(x, t)
    torch.cuda.synchronize()
    # softplus hooks still rewrite the 7 NConv weights every forward; every other parameter must stand still at lr = 0
    names = {id(p): n for n, p in m.named_parameters()}
    off = 0
    for p in tr.params:
        k = p.numel()
        if "nconv" not in names[id(p)] or names[id(p)].endswith("bias"):
            assert torch.equal(opt.flat_param[off:off + k], p1[off:off + k]), names[id(p)]
        off += k
    assert d1 > 0


@gpu
def test_non_finite_loss_skips_the_update():
    """trainer.py:206-208 ignores a batch whose loss is NaN/inf; on the device the optimizer's guard does."""
    data = _frames(2, 2)
    m, opt, tr = _make(lr=1e-3)
    tr.step(*data[0])
    torch.cuda.synchronize()
    t_before = float(opt.step_t)
    mom = opt.flat_exp_avg.clone(), opt.flat_exp_avg_sq.clone()
    names = {id(p): n for n, p in m.named_parameters()}
    keep = [(off, p.numel()) for off, p in _offsets(tr.params) if "nconv" not in names[id(p)] or names[id(p)].endswith("bias")]
    before = opt.flat_param.clone()
    x, t = data[1]
    tb = t.clone()
    tb[0, 0, 5, 5] = float("inf")                    # one poisoned ground-truth value: the masked loss becomes inf
    loss = tr.step(x, tb)
    torch.cuda.synchronize()
    assert not torch.isfinite(loss).item()
    assert float(opt.step_t) == t_before
    assert torch.equal(opt.flat_exp_avg, mom[0]) and torch.equal(opt.flat_exp_avg_sq, mom[1])
    for off, k in keep:
        assert torch.equal(opt.flat_param[off:off + k], before[off:off + k])
    # too few valid pixels (<= 100): the reference asserts on the host; lazy mode skips the batch
    tz = torch.zeros_like(t)
    tz[:, :, :5, :5] = 1.0
    loss = tr.step(x, tz)
    torch.cuda.synchronize()
    assert not torch.isfinite(loss).item() and float(opt.step_t) == t_before
    # and training continues afterwards
    loss = tr.step(x, t)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item() and float(opt.step_t) == t_before + 1


def _offsets(params):
    off = 0
    for p in params:
        yield off, p
        off += p.numel()


_NCCL_WORKER = r"""
import contextlib, io, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from test_gpu_bench_path import _make, _frames_sparse
from aerial_depth_completion_b200 import parallel
from aerial_depth_completion_b200.optim import FlatAdam
m, opt, _ = _make(lr=1e-3, seed=rank)              # different seeds: the trainer must broadcast rank 0's replica
params = [p for p in m.parameters() if p.requires_grad]
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
crit, _ = mc.ConfidenceDepthFrameworkFactory().create_loss("l2", False, 0.5)
tr = parallel.DataParallelTrainer(m, crit, opt, params, world_size=world)
data = _frames_sparse(4, 3)
half = slice(rank * 2, rank * 2 + 2)
use_graph = os.environ.get("ADC_TEST_GRAPH") == "1"
if use_graph:
    st = {{k: v.detach().clone() for k, v in m.state_dict().items()}}
    tr.capture(data[0][0][half], data[0][1][half])
    with torch.no_grad():
        m.load_state_dict(st); opt.flat_exp_avg.zero_(); opt.flat_exp_avg_sq.zero_(); opt.step_t.zero_()
losses, grad1 = [], None
for x, t in data:
    losses.append(float(tr.step(x[half], t[half])))
    if grad1 is None:
        grad1 = opt.flat_grad.clone()              # the averaged gradient of step 1 (the next step clears the bucket)
torch.cuda.synchronize()
torch.save({{"losses": losses, "grad1": grad1.cpu(), "state": {{k: v.cpu() for k, v in m.state_dict().items()}}}},
           os.path.join({out!r}, "rank%d.pt" % rank))
dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush(); os._exit(0)
"""


def _frames_sparse(B, n, seed=91):
    """Targets with very different valid-pixel counts in the two halves of the batch (100 % / 30 %)."""
    out = []
    for x, t in _frames(B, n, seed):
        keep = torch.rand(t.shape, device=t.device, generator=torch.Generator(device=t.device).manual_seed(seed)) < 0.3
        t = t.clone()
        t[B // 2:] = t[B // 2:] * keep[B // 2:]
        out.append((x, t))
    return out


@gpu
@pytest.mark.parametrize("graph", [0, 1])
def test_two_rank_nccl_step_equals_single_process(tmp_path, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, ADC_TEST_GRAPH=str(graph))
    port = 29500 + (os.getpid() % 1500) + graph
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    # single process, whole batch, rank 0's initial weights
    m, opt, tr = _make(lr=1e-3, seed=0)
    data = _frames_sparse(4, 3)
    ref_losses, ref_grad1 = [], None
    for x, t in data:
        ref_losses.append(float(tr.step(x, t)))
        if ref_grad1 is None:
            ref_grad1 = opt.flat_grad.clone().cpu()
    torch.cuda.synchronize()
    ref = {k: v.cpu() for k, v in m.state_dict().items()}
    got = [torch.load(tmp_path / ("rank%d.pt" % r)) for r in (0, 1)]
    # (i) the replicas stay bit-identical
    assert got[0]["losses"] == got[1]["losses"]
    for k in ref:
        assert torch.equal(got[0]["state"][k], got[1]["state"][k]), k
    # (ii) step 1: the loss and every gradient tensor are those of the GLOBAL batch (one masked mean over all valid pixels of both
    # shards, trainer.py:204-208) -- the shards hold 100 % and 30 % valid pixels, so a mean of per-rank means would be off by > 10 %.
    # Summation order differs (two partial sums instead of one): NConv weight gradients are heavily cancelling sums whose fp32 noise
    # floor is 1e-4..2e-3 (ref_grad_errors), everything else agrees to 1e-4.
    assert abs(got[0]["losses"][0] - ref_losses[0]) <= 2e-5 * abs(ref_losses[0]), (got[0]["losses"], ref_losses)
    names = {id(p): n for n, p in m.named_parameters()}
    for off, p in _offsets(tr.params):
        k = names[id(p)]
        err = relmax(got[0]["grad1"][off:off + p.numel()], ref_grad1[off:off + p.numel()])
        assert err <= (5e-3 if "nconv" in k else 1e-4), (k, err)
    # (iii) after three Adam steps (+-lr per element and step: noise-sized gradients may flip sign) nothing drifted by more than 3 lr
    for a, b in zip(got[0]["losses"], ref_losses):
        assert abs(a - b) <= 1e-3 * abs(b), (got[0]["losses"], ref_losses)
    for k in ref:
        if "nconv" in k and k.endswith("weight"):
            continue      # rewritten by softplus on every forward on top of the update
        assert (got[0]["state"][k] - ref[k]).abs().max().item() <= 3.2e-3, k

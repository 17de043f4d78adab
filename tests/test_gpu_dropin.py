"""GPU: the drop-in claim, exercised with the reference's OWN callers (SURVEY.md 4 integration tier, 8b, 8f-3, 8f-4).

The unmodified reference (oracle/_ref, staged by oracle/make_ref.py; /root/reference in the build container) supplies trainer.py:
 * its train() / validate() loops drive this library's modules, criterion and factory exactly as main.py:96-99 would, next to the same
   loops driving the reference's own modules on the same GPU (PyTorch/cuDNN, TF32 off) from the same weights and batches;
 * its save_checkpoint() / resume() move a trained model between the two implementations in both directions;
 * its metrics.ConfidencePixelwiseThrAverageMeter is the oracle of the one-pass histogram meter.
Skipped when the reference is not staged (it is git-ignored; build() stages it whenever /root/reference exists)."""
import contextlib
import copy
import csv
import io
import os

import numpy as np
import pytest
import torch

import reference_shim

gpu = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not reference_shim.have_reference(), reason="reference not staged (oracle/make_ref.py)")

FIELDS = ['epoch', 'mse', 'rmse', 'absrel', 'lg10', 'mae', 'delta1', 'delta2', 'delta3', 'data_time', 'gpu_time', 'loss0', 'loss1', 'loss2']


def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def _loader(n_batches, B, H, W, seed, num_samples=200):
    from oracle import synth
    out = []
    for k in range(n_batches):
        inp, tgt = synth.synth_frames(B, H, W, seed=seed + k, num_samples=num_samples)
        out.append((inp.float(), tgt.float(), torch.ones(B)))
    return out


def _fresh(loader):
    return [(a.clone(), b.clone(), c.clone()) for a, b, c in loader]      # the trainer scales targets in place


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _pair(ns, args):
    """(reference factory, reference model), (new factory, new model): same seeded weights, both on the GPU."""
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    torch.manual_seed(0)
    fr = ns.cdf.ConfidenceDepthFrameworkFactory()
    mr = _quiet(fr.create_model, *args)
    fn = mc.ConfidenceDepthFrameworkFactory()
    mn = _quiet(fn.create_model, *args)
    mn.load_state_dict(copy.deepcopy(mr.state_dict()))                    # key-for-key: part of the drop-in contract
    return (fr, mr), (fn, mn)


def _rows(path):
    with open(path) as fh:
        return [dict(zip(FIELDS, r)) for r in csv.reader(fh)]


@gpu
@needs_ref
@pytest.mark.parametrize("args,dual", [
    (("rgbd", "dc1_only", "ged_depthcompnet", ""), False),
    (("rgbdw", "dc1-cf1-ln1", "ged_depthcompnet", "", "cbr3-c1", "", "ged_depthcompnet", ""), True),
], ids=["ged_dc1_only", "ged_dc1_cf1_ln1"])
def test_reference_trainer_drives_new_modules(tmp_path, args, dual):
    """trainer.train() for 6 iterations and trainer.validate() (precision/recall table on) with main.py's construction order:
    create_model -> to_device -> create_loss -> create_optimizer.  Per-epoch csv rows (rmse, mae, delta1, losses) of the two
    implementations agree to 1e-4 relative (outputs are fp32-grade; six Adam steps of lr 1e-5 in between)."""
    trainer, ns = reference_shim.load_trainer()
    _strict_fp32()
    import types
    # validation image grids (utils.merge_into_row_with_gt2: matplotlib colour maps, SURVEY 2 row 14, out of scope) are not drawn
    trainer.ResultSampleImage = lambda *a, **k: types.SimpleNamespace(update=lambda *aa, **kk: None)
    (fr, mr), (fn, mn) = _pair(ns, args)
    train_batches = _loader(6, 2, 64, 80, seed=300)
    val_batches = _loader(3, 2, 64, 80, seed=400)
    results = {}
    for tag, f, m in (("ref", fr, mr), ("new", fn, mn)):
        out = tmp_path / tag
        out.mkdir()
        m, params = _quiet(f.to_device, m)
        crit, loss_def = f.create_loss("l2", dual, 0.5)
        opt, sched = trainer.create_optimizer("adam", params, 0, 0, 1e-5, 5, 0.1)
        _quiet(trainer.train, _fresh(train_batches), m, crit, opt, str(out), 0)
        res = _quiet(trainer.validate, _fresh(val_batches), m, crit, 0, 4, 10, str(out), dual, 0)
        results[tag] = (m, f, crit, loss_def, opt, sched, res, _rows(out / "train.csv"), _rows(out / "val.csv"),
                        open(out / "pr.csv").read() if dual else "")
    rt, rv = results["ref"][7], results["ref"][8]
    nt, nv = results["new"][7], results["new"][8]
    assert len(rt) == len(nt) == (2 if dual else 1) and len(rv) == len(nv)
    for a, b in zip(rt + rv, nt + nv):
        for k in ("mse", "rmse", "absrel", "lg10", "mae", "delta1", "delta2", "delta3", "loss0", "loss1"):
            x, y = float(a[k]), float(b[k])
            if x != x or y != y:                  # lg10 of a random-init prediction that crosses zero: NaN in both or in neither
                assert x != x and y != y, (k, x, y)
                continue
            assert abs(x - y) <= 1e-4 * max(abs(x), 1e-3), (k, x, y)
    assert abs(results["ref"][6].rmse - results["new"][6].rmse) <= 1e-4 * results["ref"][6].rmse
    if dual:
        # precision/recall table written by the reference's meter from this library's confidences: same thresholds survive, figures agree
        pr_r = [tuple(map(float, l.split(","))) for l in results["ref"][9].splitlines()]
        pr_n = [tuple(map(float, l.split(","))) for l in results["new"][9].splitlines()]
        assert len(pr_r) == len(pr_n) and len(pr_r) > 10
        for a, b in zip(pr_r, pr_n):
            assert a[0] == b[0] and abs(a[1] - b[1]) <= 2e-3 and abs(a[2] - b[2]) <= 2e-3 * max(1.0, a[2]), (a, b)

    # ---- checkpoint wire format, both directions, through the reference's own save_checkpoint / resume (trainer.py:161-184) ----
    m_new, f_new, crit, loss_def, opt, sched = results["new"][:6]
    m_ref, f_ref = results["ref"][:2]
    x = val_batches[0][0].cuda()
    ck = tmp_path / "ck_new"
    ck.mkdir()
    trainer.save_checkpoint(f_new, m_new, loss_def, opt, sched, float(results["new"][6].rmse), True, 0, str(ck))
    back = _quiet(trainer.resume, str(ck / "model_best.pth.tar"), f_ref, False)        # new modules -> file -> REFERENCE modules
    ref_from_new, opt2 = back[0], back[4]
    assert isinstance(opt2, torch.optim.Adam) and len(opt2.state_dict()["state"]) == len(opt.state_dict()["state"])
    m_new.eval()
    ref_from_new.eval()
    with torch.no_grad():
        a, b = m_new(x), ref_from_new(x)
    for u, v in zip(a, b):
        assert (u is None) == (v is None)
        if u is not None:
            assert float((u - v).abs().max()) <= 1e-4 * float(v.abs().max())
    ck2 = tmp_path / "ck_ref"
    ck2.mkdir()
    trainer.save_checkpoint(f_ref, m_ref, loss_def, results["ref"][4], results["ref"][5], float(results["ref"][6].rmse), True, 0, str(ck2))
    new_from_ref = _quiet(trainer.resume, str(ck2 / "model_best.pth.tar"), f_new, True)[0]   # reference -> file -> new modules
    new_from_ref.eval()
    m_ref.eval()
    with torch.no_grad():
        a, b = new_from_ref(x), m_ref(x)
    for u, v in zip(a, b):
        if u is not None:
            assert float((u - v).abs().max()) <= 1e-4 * float(v.abs().max())
    # --dcnet-pretrained path:key partial loading (confidence_depth_framework.py:84-110) from the reference's checkpoint
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    part = _quiet(mc.ConfidenceDepthFrameworkFactory().create_model, args[0], "dc1_only", args[2],
                  str(ck2 / "model_best.pth.tar") + ":dc_weights")
    for k, v in m_ref.dc_model.state_dict().items():
        if k in part.dc_model.state_dict() and part.dc_model.state_dict()[k].shape == v.shape:
            assert torch.equal(part.dc_model.state_dict()[k].cpu(), v.cpu()), k


@gpu
@needs_ref
@pytest.mark.parametrize("seed", [0, 1])
def test_pr_meter_histogram_vs_reference(seed):
    """metrics.ConfidencePixelwiseThrAverageMeter: one histogram kernel + suffix sums against the reference's 500 masked means.
    Recall counts are exact; the truncated integer mean may differ by one unit per call where the fp32 mean sits on an integer."""
    ns = reference_shim.load()
    from aerial_depth_completion_b200 import metrics as new_metrics
    g = torch.Generator().manual_seed(seed)
    mr, mn = ns.metrics.ConfidencePixelwiseThrAverageMeter(), new_metrics.ConfidencePixelwiseThrAverageMeter()
    calls = 3
    for _ in range(calls):
        tgt = torch.rand(2, 1, 61, 47, generator=g) * 9 + 1
        tgt[torch.rand(tgt.shape, generator=g) < 0.3] = 0                       # invalid pixels
        dep = tgt * (1 + 0.3 * torch.randn(tgt.shape, generator=g)) + 0.01
        conf = torch.rand(tgt.shape, generator=g)
        if seed == 1:
            conf = (conf * 20).round() / 20                                      # confidences exactly on thresholds' neighbours, incl. 0 and 1
        d, c, t = dep.cuda(), conf.cuda(), tgt.cuda()
        mr.evaluate(d, c, t)
        mn.evaluate(d, c, t)
    assert np.array_equal(mr.count, mn.count)
    assert np.array_equal(mr.recall, mn.recall)
    diff = np.abs(mr.absrel.astype(np.int64) - mn.absrel.astype(np.int64))
    assert diff.max() <= calls and (diff > 0).mean() < 0.05, (diff.max(), (diff > 0).mean())


@gpu
def test_pr_meter_edge_cases():
    """No valid pixel, NaN confidence, non-finite depth: the meter skips what the reference's isfinite test skips."""
    from aerial_depth_completion_b200 import metrics as new_metrics
    m = new_metrics.ConfidencePixelwiseThrAverageMeter(num_bins=11)
    z = torch.zeros(1, 1, 4, 4).cuda()
    m.evaluate(z + 1, z + 0.5, z)                                                # no valid pixel: nothing counted
    assert int(m.count.sum()) == 0
    t = z + 2
    c = z + 0.55
    c[0, 0, 0, 0] = float("nan")                                                 # NaN > thr is false for every threshold
    d = z + 3
    m.evaluate(d, c, t)
    assert list(m.count) == [1] * 6 + [0] * 5                                    # thresholds 0 .. 0.5 are below 0.55
    assert int(m.recall[0]) == int(np.float32(15) / (np.float32(16) * (np.float32(1) / np.float32(1000)))) and int(m.absrel[0]) == 500
    d[0, 0, 1, 1] = float("inf")
    m.reset()
    m.evaluate(d, c, t)
    assert int(m.count.sum()) == 0                                               # every selection contains the non-finite term

#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
Inputs and weights are formula / seed defined (oracle/synth.py), so the fixtures hold outputs only.
Each case stores full small outputs plus, per parameter gradient, [sum, abs-sum, first 8 entries].
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reference_shim  # noqa: E402
from oracle import synth  # noqa: E402

R = reference_shim.load()
torch.set_num_threads(4)


def digest(t):
    f = t.detach().double().flatten()
    return np.concatenate([[float(f.sum()), float(f.abs().sum())], f[:8].numpy(), np.zeros(max(0, 8 - f.numel()))])


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def g_like(t, seed):
    return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64).to(t.dtype)


def case_nconv_layers(out):
    """NConv2d single layers, eval and train, fp64.  Grad seeds: gy=11, gcout=12."""
    torch.set_default_dtype(torch.float64)
    for (ci, co, k) in ((1, 2, 5), (2, 2, 5), (4, 2, 3), (2, 1, 1)):
        for mode in ("eval", "train"):
            m = R.nconv.NConv2d(ci, co, (k, k), "softplus", "p", padding=k // 2).double()
            m.weight.data = synth.kat_weight(co, ci, k, torch.float64)
            m.bias.data = torch.tensor([0.01, 0.02][:co], dtype=torch.float64)
            x, c = synth.kat_sparse(2, 24, 32, torch.float64)
            x = torch.cat([x * (1 + 0.25 * j) for j in range(ci)], 1)
            c = torch.cat([c * (1 - 0.125 * j) for j in range(ci)], 1)
            x.requires_grad_(True), c.requires_grad_(True)
            m.train(mode == "train")
            y, cout = m(x, c)
            ((y * g_like(y, 11)).sum() + (cout * g_like(cout, 12)).sum()).backward()
            tag = "layer_%d_%d_%d_%s" % (ci, co, k, mode)
            out[tag + "_y"], out[tag + "_cout"] = y.detach().numpy(), cout.detach().numpy()
            out[tag + "_gx"], out[tag + "_gc"] = x.grad.numpy(), c.grad.numpy()
            out[tag + "_gw"], out[tag + "_gb"] = m.weight.grad.numpy(), m.bias.grad.numpy()
            out[tag + "_w_after"] = m.weight.detach().numpy()
    torch.set_default_dtype(torch.float32)


def case_pyramid(out):
    """CNN (unguided pyramid) 2x1x57x76, KAT weights; fp32 and clean fp64; train + eval.  Grad seeds 21/22."""
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        torch.set_default_dtype(dt)
        for mode in ("train", "eval"):
            m = R.nconv_sd.CNN().to(dt)
            m.load_state_dict(synth.kat_pyramid_state(dt))
            x0, c0 = synth.kat_sparse(2, 57, 76, dt)
            c0 = c0 * 0.75
            x0.requires_grad_(True), c0.requires_grad_(True)
            m.train(mode == "train")
            hooks = {}
            idxs = []
            orig = torch.nn.functional.max_pool2d

            def spy(*a, **k):
                r = orig(*a, **k)
                idxs.append(r[1])
                return r
            R.nconv_sd.F.max_pool2d = spy
            try:
                xo, co = m(x0, c0)
            finally:
                R.nconv_sd.F.max_pool2d = orig
            ((xo * g_like(xo, 21)).sum() + (co * g_like(co, 22)).sum()).backward()
            t = "pyr_%s_%s_" % (tag, mode)
            out[t + "xout"], out[t + "cout"] = xo.detach().numpy(), co.detach().numpy()
            out[t + "gx0"], out[t + "gc0"] = x0.grad.numpy(), c0.grad.numpy()
            for i, ix in enumerate(idxs):
                out[t + "idx%d" % i] = ix.numpy().astype(np.int32)
            for k_, p in m.named_parameters():
                out[t + "g_" + k_] = p.grad.numpy()
                out[t + "w_" + k_] = p.detach().numpy()
    torch.set_default_dtype(torch.float32)


def run_framework(out, tag, mode, dc_arch, hw, batch=2, conf_arch=None, ln_arch=None, input_type="rgbd", seed=3,
                  criterion="l2", train=True):
    f = R.cdf.ConfidenceDepthFrameworkFactory()
    model = quiet(f.create_model, input_type, mode, dc_arch, "", conf_arch, "", ln_arch, "")
    out[tag + "_template"] = np.array(json.dumps([[k_, list(v.shape)] for k_, v in model.state_dict().items()]))
    out[tag + "_cfg"] = np.array(json.dumps(dict(mode=mode, dc_arch=dc_arch, hw=list(hw), batch=batch, conf_arch=conf_arch,
                                                 ln_arch=ln_arch, input_type=input_type, seed=seed, criterion=criterion,
                                                 train=train)))
    model.load_state_dict(synth.seeded_state(model.state_dict(), seed))
    quiet(model.opt_params)
    quiet(model.train, train)
    inp, tgt = synth.synth_frames(batch, hw[0], hw[1], seed=100 + seed, num_samples=max(60, hw[0] * hw[1] // 140))
    d1, c1, d2 = model(inp)
    res = {"d1": d1, "c1": c1, "d2": d2}
    if train:
        loss, _ = f.create_loss(criterion, "ln" in mode, 0.5 if "dc1" in mode else 1.0)
        l = loss(inp, d1, d2, tgt) if d2 is not None else loss(inp, d1, tgt)
        l.backward()
        out[tag + "_loss"] = np.array([float(l)] + [float(v) for v in loss.loss])
        for k_, p in model.named_parameters():
            if p.grad is not None:
                out[tag + "_g_" + k_] = digest(p.grad)
        GERR_TODO.append((tag, {k_: p.grad.detach().clone() for k_, p in model.named_parameters() if p.grad is not None}))
        for k_, v in model.state_dict().items():
            if "nconv" in k_ and k_.endswith("weight"):
                out[tag + "_w_" + k_] = digest(v)
            if k_.endswith("running_mean") or k_.endswith("running_var"):
                out[tag + "_s_" + k_] = digest(v)
    for k_, v in res.items():
        if v is not None:
            out[tag + "_" + k_] = v.detach().numpy()
    m = R.metrics.Result()
    m.evaluate(d1.detach(), tgt)
    out[tag + "_metrics"] = np.array([m.rmse, m.mae, m.delta1, m.delta2, m.delta3, m.absrel, m.lg10, m.irmse, m.imae, m.mse])


GERR_TODO = []


def grad_noise_floor(out):
    """Per parameter tensor: max|g32 - g64| / max|g64| of the reference's OWN fp32 gradients against the fp64 oracle (which equals
    the reference to 1e-11 in fp64, tests/test_oracle_golden.py) -- the floor of the k-times-floor gradient rule."""
    import golden_util
    golden_util._cache["nets"] = {k: v for k, v in out.items()}           # the cases were just produced: read them from memory
    golden_util._cache["nets"] = type("G", (dict,), {"files": property(lambda self: list(self.keys()))})(out)
    for tag, grads in GERR_TODO:
        _, S, *_ = golden_util.run_oracle_case(tag, torch.float64)
        err = {}
        for k_, g in grads.items():
            g64 = S[k_].grad
            den = float(g64.abs().max())
            err[k_] = float((g.double() - g64).abs().max()) / (den if den > 0 else 1.0)
        out[tag + "_gerr"] = np.array(json.dumps(err))


def case_nets(out):
    run_framework(out, "ged_nyu", "dc1_only", "ged_depthcompnet", (228, 304), seed=5)
    run_framework(out, "ged_only", "dc1_only", "ged_depthcompnet", (64, 80))
    run_framework(out, "ged_only_eval", "dc1_only", "ged_depthcompnet", (64, 80), train=False)
    run_framework(out, "ged_dc1", "dc1", "ged_depthcompnet", (64, 80))
    run_framework(out, "ged_full", "dc1-cf1-ln1", "ged_depthcompnet", (64, 80), conf_arch="cbr3-c1",
                  ln_arch="ged_depthcompnet", input_type="rgbdw")
    run_framework(out, "ged_ln0", "dc1-cf1-ln0", "ged_depthcompnet", (64, 80), conf_arch="cbr3-c1",
                  ln_arch="ged_depthcompnet")
    run_framework(out, "ged_dc0", "dc0-cf1-ln1", "ged_depthcompnet", (64, 80), conf_arch="cbr3-c1",
                  ln_arch="ged_depthcompnet")
    run_framework(out, "ged_dcln", "dc1-ln1", "ged_depthcompnet", (64, 80), ln_arch="ged_depthcompnet")
    run_framework(out, "ged_odd", "dc1_only", "ged_depthcompnet", (57, 76), batch=1, criterion="l1")
    run_framework(out, "gms_only", "dc1_only", "gms_depthcompnet", (32, 40))
    run_framework(out, "gms_full", "dc1-cf1-ln1", "gms_depthcompnet", (32, 40), conf_arch="cbr3-c1",
                  ln_arch="gms_depthcompnet")
    run_framework(out, "gud_only", "dc1_only", "gudepthcompnet18", (32, 48))
    run_framework(out, "gud_only_eval", "dc1_only", "gudepthcompnet18", (32, 48), train=False)
    run_framework(out, "gud_ln0", "dc1-cf1-ln0", "gudepthcompnet18", (32, 48), conf_arch="cbr3-c1",
                  ln_arch="ged_depthcompnet")
    run_framework(out, "ud_only", "dc1_only", "udepthcompnet18", (32, 48), input_type="rgbdw")
    # a BatchNorm-net case large enough for its statistics to be well conditioned (bottleneck 6 x 8 x 2 = 96 values per channel; the
    # 32 x 48 cases above normalise over 12): BASELINE config 3's composition
    run_framework(out, "gud_mid", "dc1-cf1-ln0", "gudepthcompnet18", (96, 128), conf_arch="cbr3-c1", ln_arch="ged_depthcompnet", seed=7)
    run_framework(out, "s2d_only", "dc1_only", "resnet18", (64, 80))
    run_framework(out, "s2d_cf", "dc1-cf1-ln0", "resnet18", (64, 80), conf_arch="cbr3-c1", ln_arch="ged_depthcompnet")
    # BASELINE config 5's shape (odd sizes all the way down: 114x152 -> 57x76 -> 29x38 -> 15x19 -> 8x10, bilinear 128x160 -> 228x304)
    run_framework(out, "s2d_nyu", "dc1_only", "resnet18", (228, 304), seed=11)
    # two-channel depth-completion nets (depth + own confidence, confidence_depth_framework.py:512 / s2d_u_resnet.py:181 / s2d_resnet.py:239)
    run_framework(out, "s2d_dc1ln0", "dc1-ln0", "resnet18", (64, 80), ln_arch="ged_depthcompnet", seed=5)
    run_framework(out, "gud_dc1ln0", "dc1-ln0", "gudepthcompnet18", (32, 48), ln_arch="ged_depthcompnet", seed=6)
    # S2DResNet as the loss net (--lossnet-arch resnet18): its input rgb | depth1 | conf1 needs a gradient (5-channel 7x7 stem)
    run_framework(out, "ged_s2dln", "dc1-ln1", "ged_depthcompnet", (64, 80), ln_arch="resnet18", seed=8)
    run_framework(out, "s2d_eval", "dc1_only", "resnet18", (64, 80), train=False, seed=3)       # running statistics, no batch pass
    run_framework(out, "s2d_odd", "dc1_only", "resnet18", (93, 121), batch=3, criterion="l1", seed=4)   # odd at every level
    grad_noise_floor(out)


def case_sampling_metrics(out):
    depth = (np.arange(228 * 304).reshape(228, 304) % 17).astype(np.float64)
    for s in (0, 7, 1234):
        np.random.seed(s)
        mask = R.d2s.UniformSampling(500, np.inf).dense_to_sparse(None, depth)
        out["mask_seed%d" % s] = np.flatnonzero(mask).astype(np.int32)
    np.random.seed(5)
    d2 = np.random.uniform(0, 3, (60, 94)).astype(np.float32)
    d2[d2 < 0.5] = 0
    mask = R.d2s.UniformSampling(200, 2.5).dense_to_sparse(None, d2)
    out["mask_maxdepth"] = np.flatnonzero(mask).astype(np.int32)
    m = R.metrics.Result()
    m.evaluate(torch.tensor([1., 2., 4., 8., 3.]), torch.tensor([1.1, 0, 5.5, 8., 2.]))
    out["kat6_metrics"] = np.array([m.rmse, m.mae, m.delta1, m.delta2, m.delta3, m.absrel, m.lg10, m.irmse, m.imae, m.mse])
    for kind, cls in (("l2", R.cdf.MaskedMSELoss), ("l1", R.cdf.MaskedL1Loss), ("il1", R.cdf.InvertedMaskedL1Loss),
                      ("absrel", R.cdf.MaskedAbsRelLoss)):
        inp, tgt = synth.synth_frames(1, 24, 32, seed=9)
        tgt = tgt * (tgt > 2).float()
        pred = (tgt * 1.1 + 0.3).requires_grad_(True)
        l = cls()(None, pred, tgt)
        l.backward()
        out["loss_" + kind] = np.array([float(l)])
        out["loss_" + kind + "_grad"] = pred.grad.numpy()


if __name__ == "__main__":
    for name, fn in (("nconv_layers", case_nconv_layers), ("pyramid", case_pyramid), ("nets", case_nets),
                     ("sampling_metrics", case_sampling_metrics)):
        d = {}
        fn(d)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, len(d), "arrays", os.path.getsize(path) // 1024, "KiB")

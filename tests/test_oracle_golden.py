"""CPU suite: pins the oracle (oracle/) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the live reference."""
import json

import numpy as np
import pytest
import torch

import reference_shim
from golden_util import NET_TAGS, digest, g_like, gold, net_case, relmax, run_oracle_case
from oracle import loss_metrics_oracle as lm
from oracle import nconv_oracle as no
from oracle import nets_oracle as nets
from oracle import sampling_oracle as so
from oracle import synth

needs_ref = pytest.mark.skipif(not reference_shim.have_reference(), reason="/root/reference not present")


@pytest.mark.parametrize("cfg", [(1, 2, 5), (2, 2, 5), (4, 2, 3), (2, 1, 1)])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_layer_vs_golden(cfg, mode):
    ci, co, k = cfg
    G = gold("nconv_layers")
    tag = "layer_%d_%d_%d_%s" % (ci, co, k, mode)
    dt = torch.float64
    w = synth.kat_weight(co, ci, k, dt)
    b = torch.tensor([0.01, 0.02][:co], dtype=dt)
    x, c = synth.kat_sparse(2, 24, 32, dt)
    x = torch.cat([x * (1 + 0.25 * j) for j in range(ci)], 1)
    c = torch.cat([c * (1 - 0.125 * j) for j in range(ci)], 1)
    wf = no.softplus_pos(w) if mode == "train" else w
    y, cout, D = no.nconv2d_forward(x, c, wf, b)
    assert relmax(y, G[tag + "_y"]) < 1e-12 and relmax(cout, G[tag + "_cout"]) < 1e-12
    gx, gc, gw, gb = no.nconv2d_backward(g_like(y, 11), g_like(cout, 12), x, c, y, D, wf, wf, b)
    for name, v in (("gx", gx), ("gc", gc), ("gw", gw), ("gb", gb)):
        assert relmax(v, G[tag + "_" + name]) < 1e-11, name
    assert relmax(wf, G[tag + "_w_after"]) < 1e-14


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pyramid_vs_golden(prec, mode):
    G = gold("pyramid")
    dt = torch.float32 if prec == "f32" else torch.float64
    tol = 2e-6 if prec == "f32" else 1e-12
    P = synth.kat_pyramid_state(dt)
    x0, c0 = synth.kat_sparse(2, 57, 76, dt)
    c0 = c0 * 0.75
    xo, co, cache = no.pyramid_forward(P, x0, c0, mode == "train")
    t = "pyr_%s_%s_" % (prec, mode)
    assert relmax(xo, G[t + "xout"]) < tol and relmax(co, G[t + "cout"]) < tol
    for i in range(3):
        assert np.array_equal(cache["idx"][i].numpy(), G[t + "idx%d" % i].astype(np.int64)), "pool indices must be bit-exact"
    gx, gc, grads = no.pyramid_backward(cache, g_like(xo, 21), g_like(co, 22))
    gtol = 2e-3 if prec == "f32" else 1e-10      # fp32: the reference's own fp32-vs-fp64 floor is 4e-4..1.5e-3
    assert relmax(gx, G[t + "gx0"]) < gtol and relmax(gc, G[t + "gc0"]) < gtol
    for m in no.PYRAMID_MODULES:
        assert relmax(grads[m + ".weight"], G[t + "g_" + m + ".weight"]) < gtol, m
        assert relmax(grads[m + ".bias"], G[t + "g_" + m + ".bias"]) < gtol, m
        assert relmax(cache["final"][m], G[t + "w_" + m + ".weight"]) < 1e-6, m


def test_softplus_iteration_counts():
    """After one training forward: nconv2 = softplus^4, nconv3 = softplus^2, the rest softplus^1."""
    P = synth.kat_pyramid_state(torch.float64)
    x0, c0 = synth.kat_sparse(1, 32, 40, torch.float64)
    _, _, cache = no.pyramid_forward(P, x0, c0, True)
    for m, n in (("nconv1", 1), ("nconv2", 4), ("nconv3", 2), ("nconv4", 1), ("nconv7", 1)):
        w = P[m + ".weight"]
        for _ in range(n):
            w = no.softplus_pos(w)
        assert torch.equal(w, cache["final"][m])


@pytest.mark.parametrize("tag", NET_TAGS)
def test_nets_vs_golden(tag):
    G = gold("nets")
    cfg, S, inp, tgt, d1, c1, d2, loss, parts = run_oracle_case(tag)
    for name, v in (("d1", d1), ("c1", c1), ("d2", d2)):
        if v is None:
            assert (tag + "_" + name) not in G.files
        else:
            assert relmax(v, G[tag + "_" + name]) < 2e-5, name
    met = lm.evaluate(d1.detach(), tgt)
    ref = G[tag + "_metrics"]
    crosses_zero = bool((d1.detach() <= 0.05).any())     # 1/output and log(output) metrics are then dominated by rounding noise
    for i, k in enumerate(("rmse", "mae", "delta1", "delta2", "delta3", "absrel", "lg10", "irmse", "imae", "mse")):
        if np.isfinite(ref[i]):
            if crosses_zero and k in ("lg10", "irmse", "imae"):
                continue
            assert abs(met[k] - ref[i]) <= 1e-4 * max(1e-3, abs(ref[i])), k
    if cfg["train"]:
        gl = G[tag + "_loss"]
        assert abs(float(loss) - gl[0]) <= 1e-5 * abs(gl[0])
        assert abs(parts[0] - gl[1]) <= 1e-5 * abs(gl[1]) and abs(parts[1] - gl[2]) <= 1e-5 * max(abs(gl[2]), 1e-9)
        n = 0
        for k in S:
            gk = tag + "_g_" + k
            if gk in G.files:
                assert S[k].grad is not None, k
                d = digest(S[k].grad)
                scale = max(G[gk][1], 1e-30)
                # sum / abs-sum digests; fp32 grads are noise-limited, abs-sum is the stable one
                assert abs(d[1] - G[gk][1]) <= 2e-3 * scale, k
                n += 1
            elif S[k].requires_grad:
                assert S[k].grad is None or float(S[k].grad.abs().sum()) == 0 or ("g_" + k) not in "".join(G.files), k
        assert n > 0
        for k in S:
            wk = tag + "_w_" + k
            if wk in G.files:
                assert np.allclose(digest(S[k]), G[wk], rtol=1e-5, atol=1e-7), k
            sk = tag + "_s_" + k
            if sk in G.files:
                assert np.allclose(digest(S[k]), G[sk], rtol=1e-4, atol=1e-6), k


def test_sampling_vs_golden():
    G = gold("sampling_metrics")
    depth = (np.arange(228 * 304).reshape(228, 304) % 17).astype(np.float64)
    for s in (0, 7, 1234):
        np.random.seed(s)
        m = so.uniform_mask(depth, 500)
        assert np.array_equal(np.flatnonzero(m), G["mask_seed%d" % s])
        assert np.array_equal(np.flatnonzero(so.uniform_mask_mt(depth, 500, s)), G["mask_seed%d" % s])
    # KAT-1 of SURVEY.md 8c
    idx = G["mask_seed0"]
    assert len(idx) == 525 and list(idx[:5]) == [99, 540, 742, 757, 821] and int(idx.astype(np.int64).sum()) == 17701305
    idx = G["mask_seed7"]
    assert len(idx) == 497 and list(idx[:5]) == [55, 108, 302, 373, 558] and int(idx.astype(np.int64).sum()) == 16306862
    np.random.seed(5)
    d2 = np.random.uniform(0, 3, (60, 94)).astype(np.float32)
    d2[d2 < 0.5] = 0
    assert np.array_equal(np.flatnonzero(so.uniform_mask(d2, 200, 2.5)), G["mask_maxdepth"])


def test_mt19937_restatement():
    for seed in (0, 7, 123456):
        np.random.seed(seed)
        assert np.array_equal(so.mt19937_uniform(seed, 1500), np.random.uniform(0, 1, 1500))


def test_metrics_losses_vs_golden():
    G = gold("sampling_metrics")
    r = lm.evaluate(torch.tensor([1., 2., 4., 8., 3.]), torch.tensor([1.1, 0, 5.5, 8., 2.]))
    got = [r[k] for k in ("rmse", "mae", "delta1", "delta2", "delta3", "absrel", "lg10", "irmse", "imae", "mse")]
    assert np.allclose(got, G["kat6_metrics"], rtol=1e-6)
    assert abs(r["rmse"] - 0.90277350) < 1e-6 and abs(r["mae"] - 0.65) < 1e-6 and r["delta1"] == 0.5   # KAT-6
    for kind in ("l2", "l1", "il1", "absrel"):
        inp, tgt = synth.synth_frames(1, 24, 32, seed=9)
        tgt = tgt * (tgt > 2).float()
        pred = (tgt * 1.1 + 0.3).requires_grad_(True)
        l = lm.masked_loss(kind, pred, tgt)
        l.backward()
        assert abs(float(l) - G["loss_" + kind][0]) < 1e-6 * abs(G["loss_" + kind][0])
        assert relmax(pred.grad, G["loss_" + kind + "_grad"]) < 1e-6


def test_nearest_index_rule():
    import torch.nn.functional as F
    for (i, o) in ((28, 57), (38, 76), (57, 114), (76, 152), (114, 228), (152, 304), (30, 60), (60, 94), (94, 188), (7, 15), (15, 29), (8, 15), (10, 19)):
        t = torch.arange(i, dtype=torch.float32).view(1, 1, 1, i)
        ref = F.interpolate(t, size=(1, o), mode="nearest").flatten().long()
        assert torch.equal(no.nearest_src_index(o, i), ref), (i, o)


@needs_ref
def test_pyramid_vs_live_reference():
    R = reference_shim.load()
    torch.manual_seed(3)
    for dt in (torch.float32, torch.float64):
        torch.set_default_dtype(dt)
        try:
            m = R.nconv_sd.CNN().to(dt)
            S = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
            x0, c0 = synth.kat_sparse(2, 60, 47, dt)
            x0 = x0.clone().requires_grad_(True)
            c0 = (c0 * 0.9).requires_grad_(True)
            m.train()
            xo, co = m(x0, c0)
            gxo, gco = torch.randn_like(xo), torch.randn_like(co)
            ((xo * gxo).sum() + (co * gco).sum()).backward()
            xb, cb = x0.detach().clone().requires_grad_(True), c0.detach().clone().requires_grad_(True)
            xo2, co2 = no.pyramid_apply(S, "", xb, cb, True)
            ((xo2 * gxo).sum() + (co2 * gco).sum()).backward()
            tol = 1e-4 if dt == torch.float32 else 1e-11
            assert relmax(xo2, xo) < 1e-6 and relmax(co2, co) < 1e-6
            assert relmax(xb.grad, x0.grad) < tol and relmax(cb.grad, c0.grad) < tol
            for k, p in m.named_parameters():
                assert relmax(S[k].grad, p.grad) < tol, k
                assert torch.equal(S[k].data, p.data), k
        finally:
            torch.set_default_dtype(torch.float32)

"""GPU parity tests (run on the B200 box: pytest -m gpu): the CUDA NConv path, called through the C ABI via the
drop-in modules, against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances: outputs max|a-b|/max|b| <= 1e-4 in fp32 (we hold 2e-5); gradients are judged against the
fp64 oracle with the fp32 noise floor of the reference itself (4e-4..1.5e-3 for NConv weights) -> 2e-3;
pooling indices bit-exact."""
import numpy as np
import pytest
import torch

from golden_util import g_like, gold, relmax
from oracle import nconv_oracle as no
from oracle import synth

gpu = pytest.mark.gpu
OUT_TOL = 2e-5
GRAD_TOL = 2e-3


def _mods():
    import aerial_depth_completion_b200.model_zoo.nconv as nconv
    import aerial_depth_completion_b200.model_zoo.nconv_sd as nconv_sd
    from aerial_depth_completion_b200 import ops
    return nconv, nconv_sd, ops


def _layer_inputs(ci, n=2, h=24, w=32, dt=torch.float64):
    x, c = synth.kat_sparse(n, h, w, dt)
    x = torch.cat([x * (1 + 0.25 * j) for j in range(ci)], 1)
    c = torch.cat([c * (1 - 0.125 * j) for j in range(ci)], 1)
    return x, c


@gpu
@pytest.mark.parametrize("cfg", [(1, 2, 5), (2, 2, 5), (4, 2, 3), (2, 1, 1)])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_layer_vs_golden_and_oracle(cfg, mode):
    nconv, _, _ = _mods()
    ci, co, k = cfg
    G = gold("nconv_layers")
    tag = "layer_%d_%d_%d_%s" % (ci, co, k, mode)
    m = nconv.NConv2d(ci, co, (k, k), "softplus", "p", padding=k // 2)
    m.weight.data = synth.kat_weight(co, ci, k)
    m.bias.data = torch.tensor([0.01, 0.02][:co])
    m = m.cuda().train(mode == "train")
    x, c = _layer_inputs(ci)
    xg = x.float().cuda().requires_grad_(True)
    cg = c.float().cuda().requires_grad_(True)
    y, cout = m(xg, cg)
    assert relmax(y.cpu(), G[tag + "_y"]) < OUT_TOL and relmax(cout.cpu(), G[tag + "_cout"]) < OUT_TOL
    ((y * g_like(y, 11).cuda()).sum() + (cout * g_like(cout, 12).cuda()).sum()).backward()
    for name, v in (("gx", xg.grad), ("gc", cg.grad), ("gw", m.weight.grad), ("gb", m.bias.grad)):
        assert relmax(v.cpu(), G[tag + "_" + name]) < GRAD_TOL, name
    assert relmax(m.weight.detach().cpu(), G[tag + "_w_after"]) < 1e-6


@gpu
@pytest.mark.parametrize("shape", [(3, 2, 57, 76), (2, 2, 61, 47), (1, 2, 228, 304), (2, 1, 9, 8), (1, 2, 114, 152)])
def test_layer_random_sizes_vs_oracle(shape):
    """Odd sizes, partial tiles, non-multiple-of-4 widths (scalar load/store paths), batch-strided channel slices."""
    nconv, _, _ = _mods()
    n, _, h, w = shape
    g = torch.Generator().manual_seed(h * 1000 + w)
    for (ci, co, k) in ((1, 2, 5), (2, 2, 5), (4, 2, 3), (2, 1, 1), (2, 2, 3), (1, 2, 3)):
        x = torch.rand(n, ci, h, w, generator=g, dtype=torch.float64) * 5
        c = (torch.rand(n, ci, h, w, generator=g, dtype=torch.float64) < 0.3).double() * torch.rand(n, ci, h, w, generator=g, dtype=torch.float64)
        wt = 0.05 + torch.rand(co, ci, k, k, generator=g, dtype=torch.float64)
        b = torch.rand(co, generator=g, dtype=torch.float64) * 0.1
        m = nconv.NConv2d(ci, co, (k, k), "softplus", "p", padding=k // 2)
        m.weight.data, m.bias.data = wt.float(), b.float()
        m = m.cuda().eval()
        big = torch.zeros(n, ci + 3, h, w).cuda()           # feed a channel slice of a wider tensor (batch stride != C*H*W)
        big[:, 1:1 + ci] = x.float().cuda()
        xg = big[:, 1:1 + ci].detach().requires_grad_(True)
        cg = c.float().cuda().requires_grad_(True)
        y, cout = m(xg, cg)
        yo, co_, D = no.nconv2d_forward(x, c, wt, b)
        assert relmax(y.cpu(), yo) < OUT_TOL and relmax(cout.cpu(), co_) < OUT_TOL, (ci, co, k)
        gy, gc_ = g_like(yo, 5), g_like(co_, 6)
        (y * gy.float().cuda()).sum().add((cout * gc_.float().cuda()).sum()).backward()
        gx, gcin, gw, gb = no.nconv2d_backward(gy, gc_, x, c, yo, D, wt, wt, b)
        for name, a, r in (("gx", xg.grad, gx), ("gc", cg.grad, gcin), ("gw", m.weight.grad, gw), ("gb", m.bias.grad, gb)):
            assert relmax(a.cpu(), r) < GRAD_TOL, (ci, co, k, name, relmax(a.cpu(), r))


@gpu
def test_layer_zero_confidence_gives_bias():
    """Property: where no confidence falls in the receptive field, nconv == b and cout == 0."""
    nconv, _, _ = _mods()
    m = nconv.NConv2d(2, 2, (5, 5), "softplus", "p", padding=2).cuda().eval()
    x = torch.rand(1, 2, 40, 72).cuda()
    c = torch.zeros(1, 2, 40, 72).cuda()
    c[0, :, 5, 7] = 1.0
    y, cout = m(x, c)
    far = torch.ones(40, 72, dtype=torch.bool)
    far[3:8, 5:10] = False
    for o in range(2):
        assert torch.equal(y[0, o].cpu()[far], m.bias[o].cpu().expand(int(far.sum())))
        assert float(cout[0, o].cpu()[far].abs().max()) == 0.0
    assert float(cout.max()) <= 1.0 + 1e-6 and float(cout.min()) >= 0.0


@gpu
@pytest.mark.parametrize("shape", [(2, 2, 57, 76), (1, 3, 28, 38), (2, 2, 228, 304), (1, 1, 9, 11), (3, 2, 16, 16)])
def test_pool_bit_exact(shape):
    """Indices bit-exact with ATen max_pool2d incl. all-zero windows (tie -> first), NaN, odd sizes (floor)."""
    _, _, ops = _mods()
    g = torch.Generator().manual_seed(sum(shape))
    c = torch.rand(shape, generator=g)
    c[torch.rand(shape, generator=g) < 0.5] = 0.0            # many exact ties
    c = torch.round(c * 8) / 8                               # and coarse values -> non-zero ties
    if shape[2] > 20:
        c[0, 0, 4, 5] = float("nan")
        c[0, 0, 11, 10] = float("nan")
        c[0, 0, 10, 10] = float("inf")
        c[-1, -1, 6, 6] = -float("inf")
    x = torch.rand(shape, generator=g)
    x_ds, c_ds, idx = ops.conf_pool(x.cuda(), c.cuda())
    xr, cr, ir = no.conf_pool_forward(x, c)
    assert torch.equal(idx.cpu(), ir)
    assert torch.equal(x_ds.cpu(), xr)
    assert torch.equal(torch.nan_to_num(c_ds.cpu(), nan=-7.0), torch.nan_to_num(cr, nan=-7.0))
    # backward: scatter to the argmax (+ skip addend)
    gxd, gcd = torch.rand(xr.shape, generator=g), torch.rand(xr.shape, generator=g)
    ax, ac = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    gx, gc = ops.conf_pool_backward(gxd.cuda(), gcd.cuda(), idx, shape[2:], ax.cuda(), ac.cuda())
    rx, rc = no.conf_pool_backward(gxd, gcd, ir, shape[2:])
    assert torch.equal(gx.cpu(), rx + ax) and torch.equal(gc.cpu(), rc + ac)
    gx, gc = ops.conf_pool_backward(gxd.cuda(), gcd.cuda(), idx, shape[2:])
    assert torch.equal(gx.cpu(), rx) and torch.equal(gc.cpu(), rc)


def _run_pyramid(P, x0, c0, training, gseed=(21, 22)):
    _, nconv_sd, _ = _mods()
    m = nconv_sd.CNN()
    m.load_state_dict({k: v.float() for k, v in P.items()})
    m = m.cuda().train(training)
    xg = x0.float().cuda().requires_grad_(True)
    cg = c0.float().cuda().requires_grad_(True)
    xo, co = m(xg, cg)
    ((xo * g_like(xo, gseed[0]).cuda()).sum() + (co * g_like(co, gseed[1]).cuda()).sum()).backward()
    return m, xo, co, xg.grad, cg.grad


@gpu
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pyramid_vs_golden(mode):
    """CNN 2x1x57x76 with the KAT weights against the reference's own fp64 outputs (tests/golden/pyramid.npz)."""
    G = gold("pyramid")
    t = "pyr_f64_%s_" % mode
    P = synth.kat_pyramid_state(torch.float64)
    x0, c0 = synth.kat_sparse(2, 57, 76, torch.float64)
    m, xo, co, gx0, gc0 = _run_pyramid(P, x0, c0 * 0.75, mode == "train")
    assert relmax(xo.cpu(), G[t + "xout"]) < OUT_TOL and relmax(co.cpu(), G[t + "cout"]) < OUT_TOL
    for i, ix in enumerate(m.last_pool_indices()):
        assert np.array_equal(ix.cpu().numpy(), G[t + "idx%d" % i].astype(np.int64)), "pool indices must be bit-exact"
    assert relmax(gx0.cpu(), G[t + "gx0"]) < GRAD_TOL and relmax(gc0.cpu(), G[t + "gc0"]) < GRAD_TOL
    for name, p in m.named_parameters():
        assert relmax(p.grad.cpu(), G[t + "g_" + name]) < GRAD_TOL, name
        if name.endswith("weight"):
            assert relmax(p.detach().cpu(), G[t + "w_" + name]) < 1e-6, name      # softplus^calls landed in the Parameter


@gpu
@pytest.mark.parametrize("shape", [(2, 228, 304), (1, 240, 320), (2, 61, 47), (3, 64, 80), (1, 120, 188)])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pyramid_vs_oracle(shape, mode):
    """Seeded frames at the configs' resolutions (incl. odd pyramid levels) against the fp64 oracle."""
    n, h, w = shape
    inp, _ = synth.synth_frames(n, h, w, seed=7, num_samples=max(100, h * w // 140), dtype=torch.float64)
    x0, c0 = inp[:, 3:4], inp[:, 4:5]
    P = synth.seeded_state({k: v.shape for k, v in synth.kat_pyramid_state().items()}, seed=5, dtype=torch.float64)
    xo_r, co_r, cache = no.pyramid_forward(P, x0, c0, mode == "train")
    m, xo, co, gx0, gc0 = _run_pyramid(P, x0, c0, mode == "train")
    assert relmax(xo.cpu(), xo_r) < OUT_TOL and relmax(co.cpu(), co_r) < OUT_TOL
    mism = sum(int((a.cpu() != b).sum()) for a, b in zip(m.last_pool_indices(), cache["idx"]))
    assert mism == 0, "%d pooling indices differ from the oracle" % mism
    gx_r, gc_r, grads = no.pyramid_backward(cache, g_like(xo_r, 21), g_like(co_r, 22))
    assert relmax(gx0.cpu(), gx_r) < GRAD_TOL and relmax(gc0.cpu(), gc_r) < GRAD_TOL
    for name, p in m.named_parameters():
        assert relmax(p.grad.cpu(), grads[name]) < GRAD_TOL, (name, relmax(p.grad.cpu(), grads[name]))
    # properties that hold at any size
    assert float(co.min()) >= 0.0 and float(co.max()) <= 1.0 + 1e-5
    assert bool(torch.isfinite(xo).all()) and all(bool(torch.isfinite(p.grad).all()) for p in m.parameters())


@gpu
def test_pyramid_backward_reads_weights_at_backward_time():
    """data gradients use whatever the Parameters hold when backward runs."""
    P = synth.kat_pyramid_state(torch.float64)
    x0, c0 = synth.kat_sparse(1, 32, 40, torch.float64)
    _, nconv_sd, _ = _mods()
    m = nconv_sd.CNN()
    m.load_state_dict({k: v.float() for k, v in P.items()})
    m = m.cuda().train()
    xg = x0.float().cuda().requires_grad_(True)
    cg = (c0 * 0.5).float().cuda().requires_grad_(True)
    xo, co = m(xg, cg)
    with torch.no_grad():
        m.nconv2.weight.data = m.nconv2.weight.data * 1.5       # e.g. an optimizer step between forward and backward
    (xo.sum() + co.sum()).backward()
    xo_r, co_r, cache = no.pyramid_forward(P, x0, c0 * 0.5, True)
    w_bwd = dict(cache["final"])
    w_bwd["nconv2"] = w_bwd["nconv2"] * 1.5
    gx_r, gc_r, grads = no.pyramid_backward(cache, torch.ones_like(xo_r), torch.ones_like(co_r), w_bwd)
    assert relmax(xg.grad.cpu(), gx_r) < GRAD_TOL and relmax(cg.grad.cpu(), gc_r) < GRAD_TOL
    assert relmax(m.nconv1.weight.grad.cpu(), grads["nconv1.weight"]) < GRAD_TOL


@gpu
def test_no_cpu_fallback():
    nconv, nconv_sd, _ = _mods()
    m = nconv_sd.CNN()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))
    with pytest.raises(NotImplementedError):
        nconv.NConv2d(2, 2, (5, 5), padding=0).cuda()(torch.rand(1, 2, 16, 16).cuda(), torch.rand(1, 2, 16, 16).cuda())

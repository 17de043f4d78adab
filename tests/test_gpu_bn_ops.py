"""GPU: the building blocks of the BatchNorm nets' tensor-core path (tc_bn.py, csrc/bn_act.cu, adc_conv_tc) one by one against fp64
torch CPU references: BatchNorm (+ residual) + ReLU / LeakyReLU forward and backward incl. the running statistics, every
convolution form (row-segment kernel with a K tail and accumulation, stride-2 box kernel, 1x1 stride 2, ConvTranspose2d stride 2 and
stride 1) with its data and weight gradients, the 1x1 head, the im2col packer and the general cbr3-c1 head at 129 channels."""
import pytest
import torch
import torch.nn.functional as F

from golden_util import relmax

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def _to_pair(x, cs=None):
    """fp64 NCHW -> zero-bordered (hi, lo) activation."""
    from aerial_depth_completion_b200 import dense
    n, c, h, w = x.shape
    return dense.to_nhwc_bf16(x.float().cuda(), cpad=cs or dense._r64(c), pad=True)


def _to_f32(x, cs=None):
    from aerial_depth_completion_b200 import dense, tc_bn
    n, c, h, w = x.shape
    a = tc_bn.f32_act(n, h, w, cs or dense._r64(c), "cuda")
    a.hi[:, 1:-1, 1:-1, :c] = x.float().cuda().permute(0, 2, 3, 1)
    return a


def _from(act, c, off=0):
    from aerial_depth_completion_b200 import dense
    return dense.to_nchw_f32(act, c, off).double().cpu()


@pytest.mark.parametrize("C,act,res,training", [(16, 2, False, True), (48, 2, False, True), (64, 1, True, True), (129, 2, False, True),
                                                (512, 1, True, True), (64, 0, False, True), (64, 1, True, False)])
def test_bn_act_forward_backward(C, act, res, training):
    from aerial_depth_completion_b200 import dense, tc_bn
    n, h, w = 2, 9, 14
    y = _rand(n, C, h, w, seed=1, scale=2.0) + 0.5
    r = _rand(n, C, h, w, seed=2) if res else None
    g = _rand(n, C, h, w, seed=3)
    bn = torch.nn.BatchNorm2d(C).double()
    bn.weight.data = 1.0 + 0.3 * _rand(C, seed=4)
    bn.bias.data = 0.2 * _rand(C, seed=5)
    bn.running_mean.data = 0.1 * _rand(C, seed=6)
    bn.running_var.data = 1.0 + 0.2 * torch.rand(C, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    bn_dev = torch.nn.BatchNorm2d(C)
    bn_dev.load_state_dict({k: v.float() if v.dtype.is_floating_point else v for k, v in bn.state_dict().items()})
    bn_dev = bn_dev.cuda()
    bn.train(training)
    # reference
    yr = y.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    z = bn(yr) + (rr if res else 0.0)
    o = F.relu(z) if act == 1 else (F.leaky_relu(z, 0.2) if act == 2 else z)
    (o * g).sum().backward()
    # device
    cs = dense._r64(C)
    Y = _to_f32(y, cs)
    R = _to_pair(r, cs) if res else None
    out = tc_bn.pair_act(n, h, w, cs + 64, "cuda")           # written into a channel slice at offset 64
    st = tc_bn.bn_forward(Y, 0, C, bn_dev, training, out, 64, act, res=R, res_off=0)
    tol = 3e-5
    assert relmax(_from(out, C, 64), o.detach()) < tol
    assert float(out.hi[:, 0].abs().max()) == 0.0 and float(out.hi[:, 1:-1, 1:-1, :64].abs().max()) == 0.0     # border / other slices untouched
    if training:
        assert relmax(bn_dev.running_mean.double().cpu(), bn.running_mean) < 1e-5
        assert relmax(bn_dev.running_var.double().cpu(), bn.running_var) < 1e-5
        assert int(bn_dev.num_batches_tracked) == 1
    G = _to_f32(g, cs + 64)
    G.hi = torch.roll(G.hi, 64, 3)                             # the gradient lives at the same slice offset
    dY = tc_bn.pair_act(n, h, w, cs, "cuda")
    GR = tc_bn.f32_act(n, h, w, cs, "cuda") if res else None
    if res:
        GR.hi[:, 1:-1, 1:-1, :C] = 1.0                         # accumulate on top of an existing gradient
    gg, gb = tc_bn.bn_backward(G, 64, Y, 0, out, 64, st, act, dY, 0, C, gres=GR, gres_off=0, gres_accum=True)
    assert relmax(_from(dY, C), yr.grad) < 5e-5
    assert relmax(gg.double().cpu(), bn.weight.grad) < 2e-5 and relmax(gb.double().cpu(), bn.bias.grad) < 2e-5
    if res:
        assert relmax(_from(GR, C), rr.grad + 1.0) < 1e-5


def _prep(w, mode, cin_pad, off, cnt):
    from aerial_depth_completion_b200 import dense
    return dense.prep_conv3x3_weights_batch([(w.float().cuda(), mode, cin_pad, off, cnt)])[0]


@pytest.mark.parametrize("cin,cout", [(64, 64), (144, 80), (128, 49)])
def test_rowseg_ktail_and_accumulate(cin, cout):
    """3x3 stride 1 with cin = 64*k + 16*t and accumulation into an fp32 slice."""
    from aerial_depth_completion_b200 import dense, tc_bn
    n, h, w = 2, 10, 13
    creal = cin - 15 if cin % 64 else cin
    x = _rand(n, creal, h, w, seed=1)
    wt = _rand(cout, creal, 3, 3, seed=2, scale=0.1)
    ref = F.conv2d(x, wt, None, 1, 1)
    X = _to_pair(x, dense._r64(cin))
    wp = _prep(wt, 0, cin, 0, creal)
    out = tc_bn.f32_act(n, h, w, 192, "cuda")
    tc_bn.conv(tc_bn.K_S1, X, 0, cin, wp, out, 64, cout, h, w)
    assert relmax(_from(out, cout, 64), ref) < 3e-5
    tc_bn.conv(tc_bn.K_S1, X, 0, cin, wp, out, 64, cout, h, w, accumulate=True)
    assert relmax(_from(out, cout, 64), 2 * ref) < 3e-5
    assert float(out.hi[:, 0].abs().max()) == 0.0 and float(out.hi[..., :64].abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,h,w", [(64, 128, 16, 32), (128, 64, 12, 20)])
def test_stride2_conv_and_shortcut_with_gradients(cin, cout, h, w):
    from aerial_depth_completion_b200 import dense, tc_bn
    n = 2
    x = _rand(n, cin, h, w, seed=1).requires_grad_(True)
    w3 = _rand(cout, cin, 3, 3, seed=2, scale=0.1).requires_grad_(True)
    w1 = _rand(cout, cin, 1, 1, seed=3, scale=0.2).requires_grad_(True)
    g3, g1 = _rand(n, cout, h // 2, w // 2, seed=4), _rand(n, cout, h // 2, w // 2, seed=5)
    y3, y1 = F.conv2d(x, w3, None, 2, 1), F.conv2d(x, w1, None, 2, 0)
    ((y3 * g3).sum() + (y1 * g1).sum()).backward()
    X = _to_pair(x.detach())
    W = tc_bn.WPrep()
    W.fwd("c3", w3.detach().float().cuda(), cin)
    W.fwd("c1", tc_bn.embed_1x1(w1.detach().float().cuda()), cin)
    Wf = W.run()
    Y3 = tc_bn.f32_act(n, h // 2, w // 2, cout, "cuda")
    Y1 = tc_bn.f32_act(n, h // 2, w // 2, cout, "cuda")
    tc_bn.fwd_box(X, 0, cin, Wf, "c3", Y3, 0, cout, 2)
    tc_bn.fwd_box(X, 0, cin, Wf, "c1", Y1, 0, cout, 2, kind=tc_bn.K_1X1S2)
    assert relmax(_from(Y3, cout), y3.detach()) < 3e-5 and relmax(_from(Y1, cout), y1.detach()) < 3e-5
    # gradients
    D3, D1 = _to_pair(g3), _to_pair(g1)
    W = tc_bn.WPrep()
    W.quad_s2("c3", w3.detach().float().cuda(), dense._r64(cout))          # parity-block form on the row-segment kernel
    W.transposed("c1", tc_bn.embed_1x1(w1.detach().float().cuda()), dense._r64(cout), flip=False)
    Wb = W.run()
    Gx = tc_bn.f32_act(n, h, w, cin, "cuda")
    tc_bn.dgrad_s2(D3, dense._r64(cout), Wb, "c3", Gx, 0, cin, False)
    tc_bn.dgrad_s2(D1, dense._r64(cout), Wb, "c1", Gx, 0, cin, True, kind=tc_bn.K_1X1S2_DGRAD)
    assert relmax(_from(Gx, cin), x.grad) < 5e-5
    assert relmax(tc_bn.wgrad_box(D3, 0, cout, X, 0, cin, 2).double().cpu(), w3.grad) < 5e-5
    assert relmax(tc_bn.wgrad_1x1_box(D1, 0, cout, X, 0, cin, 2).double().cpu(), w1.grad) < 5e-5


@pytest.mark.parametrize("cin,cout,h,w", [(192, 64, 8, 12), (512, 256, 4, 6)])
def test_conv_transpose_stride2_with_gradients(cin, cout, h, w):
    from aerial_depth_completion_b200 import dense, tc_bn
    n = 2
    x = _rand(n, cin, h, w, seed=1).requires_grad_(True)
    wt = _rand(cin, cout, 3, 3, seed=2, scale=0.1).requires_grad_(True)
    g = _rand(n, cout, 2 * h, 2 * w, seed=3)
    y = F.conv_transpose2d(x, wt, None, 2, 1, 1)
    (y * g).sum().backward()
    X = _to_pair(x.detach())
    wd = wt.detach().float().cuda()
    W = tc_bn.WPrep()
    W.quad_s2("t", wd, cin)
    Wf = W.run()
    Y = tc_bn.f32_act(n, 2 * h, 2 * w, dense._r64(cout), "cuda")
    tc_bn.fwd_convt2(X, 0, cin, Wf, "t", Y, 0, cout)
    assert relmax(_from(Y, cout), y.detach()) < 3e-5
    D = _to_pair(g)
    W = tc_bn.WPrep()
    W.fwd("t", wd, dense._r64(cout))
    Wb = W.run()
    Gx = tc_bn.f32_act(n, h, w, cin, "cuda")
    tc_bn.dgrad_convt2(D, dense._r64(cout), Wb, "t", Gx, 0, cin, False)
    assert relmax(_from(Gx, cin), x.grad) < 5e-5
    assert relmax(tc_bn.wgrad_box(X, 0, cin, D, 0, cout, 2).double().cpu(), wt.grad) < 5e-5


def test_conv_transpose_stride1_with_gradients():
    from aerial_depth_completion_b200 import dense, tc_bn
    n, cin, cout, h, w = 2, 128, 64, 10, 12
    x = _rand(n, cin, h, w, seed=1).requires_grad_(True)
    wt = _rand(cin, cout, 3, 3, seed=2, scale=0.1).requires_grad_(True)
    g = _rand(n, cout, h, w, seed=3)
    y = F.conv_transpose2d(x, wt, None, 1, 1)
    (y * g).sum().backward()
    X = _to_pair(x.detach())
    wd = wt.detach().float().cuda()
    W = tc_bn.WPrep()
    W.transposed("t", wd, cin, flip=True)
    Wf = W.run()
    Y = tc_bn.f32_act(n, h, w, 64, "cuda")
    tc_bn.fwd_s1(X, 0, cin, Wf, "t", Y, 0, cout)
    assert relmax(_from(Y, cout), y.detach()) < 3e-5
    D = _to_pair(g)
    W = tc_bn.WPrep()
    W.fwd("t", wd, 64)
    Wb = W.run()
    Gx = tc_bn.f32_act(n, h, w, cin, "cuda")
    tc_bn.dgrad_s1(D, 64, Wb, "t", Gx, 0, cin, False)
    assert relmax(_from(Gx, cin), x.grad) < 5e-5
    dwc = tc_bn.wgrad_s1(D, 0, cout, X, 0, cin)
    assert relmax(dwc.flip(2, 3).transpose(0, 1).double().cpu(), wt.grad) < 5e-5


def test_wide_wgrad_blocks():
    """256 -> 512 channels: (128 x 128) channel blocks through slice offsets."""
    from aerial_depth_completion_b200 import tc_bn
    n, cin, cout, h, w = 1, 256, 512, 6, 8
    x = _rand(n, cin, h, w, seed=1)
    g = _rand(n, cout, h, w, seed=2)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x, wt, None, 1, 1) * g).sum().backward()
    assert relmax(tc_bn.wgrad_s1(_to_pair(g), 0, cout, _to_pair(x), 0, cin).double().cpu(), wt.grad) < 5e-5


@pytest.mark.parametrize("oc", [1, 2])
def test_dot_head(oc):
    from aerial_depth_completion_b200 import tc_bn
    n, C, h, w = 2, 128, 7, 9
    x = _rand(n, C, h, w, seed=1).requires_grad_(True)
    wt = _rand(oc, C, 1, 1, seed=2, scale=0.2).requires_grad_(True)
    b = _rand(oc, seed=3).requires_grad_(True)
    g = _rand(n, oc, h, w, seed=4)
    y = F.conv2d(x, wt, b)
    (y * g).sum().backward()
    X = _to_pair(x.detach(), 192)
    X.hi, X.lo = torch.roll(X.hi, 64, 3), torch.roll(X.lo, 64, 3)
    got = tc_bn.dot_head(X, 64, C, wt.detach().float().cuda(), b.detach().float().cuda())
    assert relmax(got.double().cpu(), y.detach()) < 2e-5
    G = tc_bn.f32_act(n, h, w, 192, "cuda")
    G.hi[:, 1:-1, 1:-1, 64:64 + C] = 0.5
    gw, gb = tc_bn.dot_head_backward(X, 64, C, wt.detach().float().cuda(), g.float().cuda(), G, 64, True)
    assert relmax(_from(G, C, 64), x.grad + 0.5) < 2e-5
    assert relmax(gw.double().cpu(), wt.grad) < 2e-5 and relmax(gb.double().cpu(), b.grad) < 2e-5


def test_im2col_pack_and_stem_gemm():
    from aerial_depth_completion_b200 import dense, tc_bn
    from aerial_depth_completion_b200._lib import lib, check
    n, h, w = 2, 9, 11
    full = _rand(n, 6, h, w, seed=1).float().cuda()
    d0, d1, rgb = full[:, 4:5].contiguous(), full[:, 5:6].contiguous(), full[:, :3]          # rgb: a channel-slice view
    dst = tc_bn.pair_act(n, h, w, 64, "cuda")
    check(lib.adc_im2col3x3_pack(d0.data_ptr(), 1, h * w, d1.data_ptr(), 1, h * w, rgb.data_ptr(), 3, rgb.stride(0), dst.hi.data_ptr(),
                                 dst.lo.data_ptr(), 64, n, h, w, 1, dense._stream()), "pack")
    src = torch.cat((d0, d1, rgb), 1).double().cpu()
    cols = F.unfold(src, 3, padding=1).reshape(n, 5, 9, h, w).permute(0, 2, 1, 3, 4).reshape(n, 45, h, w)      # channel (3u+v)*5 + c
    assert relmax(_from(dst, 45), cols) < 1e-5
    assert float(dst.hi[..., 45:].abs().max()) == 0.0


@pytest.mark.parametrize("C", [129, 36])
def test_general_conf_head_vs_fp64(C):
    """cbr3-c1 at the S2DU feature width (129 = 2 blocks + one K step; output channels 80 + 49) against the fp64 module."""
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import tc_bn
    torch.manual_seed(3)
    n, h, w = 2, 12, 16
    head = mc.CBR3C1Confidence(C)
    head.conv1.weight.data.mul_(5.0)
    ref = mc.CBR3C1Confidence(C).double()
    ref.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in head.state_dict().items()})
    ref.train()
    head = head.cuda().train()
    x = _rand(n, C, h, w, seed=5)
    g = _rand(n, 1, h, w, seed=6)
    xr = x.clone().requires_grad_(True)
    yr = torch.sigmoid(F.conv2d(F.leaky_relu(ref.bn1(F.conv2d(xr, ref.conv1.weight, None, 1, 1)), 0.2), ref.conv2.weight, ref.conv2.bias))
    (yr * g).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    yg = tc_bn.conf_head(head, xg)
    (yg * g.float().cuda()).sum().backward()
    assert relmax(yg.detach().double().cpu(), yr.detach()) < 1e-4
    assert relmax(xg.grad.double().cpu(), xr.grad) < 2e-3
    for k, p in head.named_parameters():
        assert relmax(p.grad.double().cpu(), dict(ref.named_parameters())[k].grad) < 2e-3, k
    assert relmax(head.bn1.running_var.double().cpu(), ref.bn1.running_var) < 1e-5

"""CPU suite: the C-ABI library loads and exports what include/adc_b200.h declares, host-side size/layout logic,
drop-in surface (state_dict keys, seeded init) and loud failure without CUDA.  No kernel is launched here."""
import ctypes
import json
import os
import re

import pytest
import torch

import reference_shim
from golden_util import gold

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not reference_shim.have_reference(), reason="/root/reference not present")


def test_library_exports_every_declared_symbol():
    from aerial_depth_completion_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "adc_b200.h")).read()
    declared = set(re.findall(r"\b(adc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), "libadc_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert _lib.lib.adc_abi_version() == 2


def test_host_side_layout_functions():
    from aerial_depth_completion_b200._lib import lib
    n, h, w = 2, 228, 304
    off = (ctypes.c_longlong * 3)()
    hs = (ctypes.c_int * 3)()
    ws = (ctypes.c_int * 3)()
    assert lib.adc_pyramid_idx_layout(n, h, w, off, hs, ws) == 0
    assert list(hs) == [114, 57, 28] and list(ws) == [152, 76, 38]            # floor pooling
    assert list(off) == [0, 2 * n * 114 * 152, 2 * n * (114 * 152 + 57 * 76)]
    assert lib.adc_pyramid_idx_count(n, h, w) == 2 * n * (114 * 152 + 57 * 76 + 28 * 38)
    assert lib.adc_pyramid_weight_floats() == 468                              # 481 params = 468 weights + 13 biases
    px = [n * 228 * 304, n * 114 * 152, n * 57 * 76, n * 28 * 38]
    # 13 saved pair tensors [N,2,h,w] + the packed input [N,1,H,W], after the 11 weight images
    assert lib.adc_pyramid_saved_floats(n, h, w) == 11 * 128 + 4 * (4 * px[0] + 4 * px[1] + 3 * px[2] + 2 * px[3]) + 2 * px[0]
    assert lib.adc_pyramid_bwd_scratch_floats(n, h, w) > 0
    # argument errors come back as negative status + message, not a crash
    assert lib.adc_pyramid_idx_layout(0, h, w, off, hs, ws) == -1
    assert b"bad argument" in lib.adc_last_error()


def test_weight_count_matches_reference_pyramid():
    from aerial_depth_completion_b200.model_zoo.nconv_sd import CNN
    m = CNN()
    assert sum(p.numel() for p in m.parameters()) == 481
    assert sum(p.numel() for n_, p in m.named_parameters() if n_.endswith("weight")) == 468
    keys = list(m.state_dict().keys())
    assert keys == ["nconv%d.%s" % (i, s) for i in range(1, 8) for s in ("weight", "bias")]


def test_cpu_tensors_raise_no_fallback():
    from aerial_depth_completion_b200.model_zoo.nconv import NConv2d
    from aerial_depth_completion_b200.model_zoo.nconv_sd import CNN
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CNN()(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NConv2d(2, 2, (5, 5), padding=2).eval()(torch.rand(1, 2, 16, 16), torch.rand(1, 2, 16, 16))
    with pytest.raises(NotImplementedError):
        NConv2d(2, 2, (5, 5), padding=0).eval()(torch.rand(1, 2, 16, 16), torch.rand(1, 2, 16, 16))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aerial_depth_completion_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


@needs_ref
def test_seeded_construction_equals_reference():
    R = reference_shim.load()
    from aerial_depth_completion_b200.model_zoo.nconv_sd import CNN
    torch.manual_seed(0)
    mine = CNN()
    torch.manual_seed(0)
    ref = R.nconv_sd.CNN()
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_sampling_mask_bit_exact():
    import numpy as np
    G = gold("sampling_metrics")
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import UniformSampling
    depth = (np.arange(228 * 304).reshape(228, 304) % 17).astype(np.float64)
    for s in (0, 7, 1234):
        np.random.seed(s)
        assert np.array_equal(np.flatnonzero(UniformSampling(500, np.inf).dense_to_sparse(None, depth)), G["mask_seed%d" % s])


@pytest.mark.skipif(not reference_shim.have_reference(), reason="reference not available")
@pytest.mark.parametrize("args", [
    ("rgbd", "dc1_only", "ged_depthcompnet", ""),
    ("rgbdw", "dc1-cf1-ln1", "ged_depthcompnet", "", "cbr3-c1", "", "ged_depthcompnet", ""),
    ("rgbd", "dc1-cf1-ln0", "gudepthcompnet18", "", "cbr3-c1", "", "ged_depthcompnet", ""),
    ("rgbd", "dc1_only", "gms_depthcompnet", ""),
    ("rgbd", "dc1_only", "resnet18", ""),
], ids=["ged", "ged_cf_ln", "gud_cf_ln0", "gms", "resnet18"])
def test_checkpoint_wire_format_round_trip(tmp_path, args):
    """SURVEY 8f-3 (trainer.py:175-184, confidence_depth_framework.py:84-110,161-187, utils.py:17-26): the reference's get_state ->
    torch.save -> this package's create_model_from_state restores every tensor bit for bit (post-softplus NConv weights, BN buffers
    and num_batches_tracked included), the reverse direction too, and `file:key` partial loading picks one sub-net."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    ns = reference_shim.load()
    q = contextlib.redirect_stdout(io.StringIO())
    with q:
        torch.manual_seed(3)
        fr, fn = ns.cdf.ConfidenceDepthFrameworkFactory(), mc.ConfidenceDepthFrameworkFactory()
        mr = fr.create_model(*args)
        for p in mr.parameters():                       # a "trained" state: every tensor distinct from its initialisation
            p.data.add_(0.01 * torch.randn_like(p))
        ck = str(tmp_path / "model_best.pth.tar")
        torch.save({"model_state": fr.get_state(mr), "epoch": 3}, ck)
        mn = fn.create_model_from_state(torch.load(ck)["model_state"])
    sr, sn = mr.state_dict(), mn.state_dict()
    assert list(sr.keys()) == list(sn.keys())
    for k in sr:
        assert sr[k].dtype == sn[k].dtype and torch.equal(sr[k], sn[k]), k
    for attr in ("input_type", "overall_arch", "dc_arch", "conf_arch", "loss_dc_arch", "input_size"):
        assert getattr(mr, attr) == getattr(mn, attr), attr
    with q:
        ck2 = str(tmp_path / "from_new.pth.tar")
        torch.save({"model_state": fn.get_state(mn)}, ck2)
        back = fr.create_model_from_state(torch.load(ck2)["model_state"])
        part = fn.create_model(args[0], "dc1_only", args[2], ck + ":dc_weights")
    for k, v in back.state_dict().items():
        assert torch.equal(v, sr[k]), k
    for k, v in part.dc_model.state_dict().items():
        if v.shape == mr.dc_model.state_dict()[k].shape:
            assert torch.equal(v, mr.dc_model.state_dict()[k]), k


def test_quad_tap_maps_are_the_identities_the_kernels_rely_on():
    """dense.quad_maps (host constants behind adc_quad_tap_mix, the tap masks and five kernel uses) against the
    convolutions they stand for, in fp64 on the CPU:
      'up': conv3x3(nearest_up2(x), w)[2i+py, 2j+px] == conv3x3(x, sum of the taps of w that land on the same low-resolution pixel)[i, j]
            (conv10 behind the 2x up-sampling, confidence_depth_framework.py:504-506);
      's2': the input gradient of a 3x3 / stride-2 / pad-1 convolution at (2i+py, 2j+px) == a stride-1 3x3 convolution of the
            output gradient with the taps that reach that parity (conv1..conv5, :442-452; ConvTranspose2d(3, 2, 1, 1) of
            model_zoo/s2d_u_resnet.py:45-61 is the same map with the forward weights)."""
    import torch.nn.functional as F
    from aerial_depth_completion_b200 import dense
    up, s2 = dense.quad_maps("cpu")
    up, s2 = up.double(), s2.double()
    g = torch.Generator().manual_seed(5)
    ci, co, h, w = 3, 4, 6, 7
    x = torch.randn(2, ci, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    # 'up'
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wt, padding=1)                     # [2, co, 2h, 2w]
    wq = (wt.reshape(1, co, ci, 1, 9) * up.reshape(4, 1, 1, 9, 9)).sum(-1).reshape(4, co, ci, 3, 3)
    for q in range(4):
        py, px = q >> 1, q & 1
        got = F.conv2d(x, wq[q], padding=1)
        assert torch.allclose(got, ref[:, :, py::2, px::2], atol=1e-12)
    # 's2' (even input size: every parity class has h x w pixels)
    xin = torch.randn(2, ci, 2 * h, 2 * w, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xin, wt, stride=2, padding=1)                                                          # [2, co, h, w]
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad(y, xin, gy)
    wd = (wt.reshape(1, co, ci, 1, 9) * s2.reshape(4, 1, 1, 9, 9)).sum(-1).permute(0, 2, 1, 3).reshape(4, ci, co, 3, 3)
    masks, mask_all = dense.quad_s2_tap_masks()
    for q in range(4):
        py, px = q >> 1, q & 1
        got = F.conv2d(gy, wd[q], padding=1)
        assert torch.allclose(got, gx[:, :, py::2, px::2], atol=1e-12)
        live = sum(1 << t for t in range(9) if float(wd[q].reshape(ci, co, 9)[..., t].abs().max()) > 0)
        assert live == masks[q]                       # the tap mask of a parity block names exactly its non-zero taps
    assert [bin(m).count("1") for m in masks] == [1, 2, 2, 4] and mask_all == 0b111111111 & (masks[0] | masks[1] | masks[2] | masks[3])
    # ConvTranspose2d(3, stride 2, padding 1, output_padding 1) with weight [in][out] = the same map
    wio = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)                                   # [in = co][out = ci]
    yt = F.conv_transpose2d(gy, wio, stride=2, padding=1, output_padding=1)
    wtq = (wio.reshape(1, co, ci, 1, 9) * s2.reshape(4, 1, 1, 9, 9)).sum(-1).permute(0, 2, 1, 3).reshape(4, ci, co, 3, 3)
    for q in range(4):
        assert torch.allclose(F.conv2d(gy, wtq[q], padding=1), yt[:, :, (q >> 1)::2, (q & 1)::2], atol=1e-12)


def test_quad_tap_mix_index_formulas_equal_the_tap_maps():
    """The closed forms adc_quad_tap_mix computes on the device (csrc/tc_glue.cu: quad_up_slot and the 's2' tap (py+3-2a, px+3-2b)),
    restated here, against the 0/1 tensors of dense.quad_maps (the GPU test compares the kernel's OUTPUT with the same tensors)."""
    from aerial_depth_completion_b200 import dense
    up, s2 = dense.quad_maps("cpu")
    for q in range(4):
        py, px = q >> 1, q & 1
        for t in range(9):
            u, v = divmod(t, 3)
            slot = (((py + u - 1) >> 1) + 1) * 3 + ((px + v - 1) >> 1) + 1
            assert [int(up[q, s, t]) for s in range(9)] == [int(s == slot) for s in range(9)]
        for s in range(9):
            a, b = divmod(s, 3)
            u, v = py + 3 - 2 * a, px + 3 - 2 * b
            want = [0] * 9
            if 0 <= u < 3 and 0 <= v < 3:
                want[u * 3 + v] = 1
            assert [int(s2[q, s, t]) for t in range(9)] == want

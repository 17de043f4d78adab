"""Oracle (TEST INFRASTRUCTURE): the dense encoder-decoders around the NConv pyramid.

Functional restatement over a flat ``state`` dict that uses the reference's own
``state_dict`` keys (so a reference checkpoint / golden fixture drops in).  Plain
torch CPU ops + autograd; the pyramid goes through ``nconv_oracle.pyramid_apply``
(explicit reference-faithful backward).

Reference lines followed (relative to /root/reference/model_zoo):
  confidence_depth_framework.py:217-235  framework forward (dc -> conf -> loss net)
  confidence_depth_framework.py:237-260  per-part train/eval mode
  confidence_depth_framework.py:317-329  CBR3C1Confidence.forward
  confidence_depth_framework.py:464-518  GEDNet.forward
  confidence_depth_framework.py:630-669  GMSNet.forward
  s2d_u_resnet.py:132-190                S2DUResNet.forward
  s2d_resnet.py:11-24,105-143,212-242    Unpool / UpProj / S2DResNet.forward
"""
import torch
import torch.nn.functional as F

from .nconv_oracle import pyramid_apply


def _conv(S, key, x, stride=1, pad=1):
    return F.conv2d(x, S[key + ".weight"], S.get(key + ".bias"), stride, pad)


def _convt(S, key, x, stride, pad, out_pad):
    return F.conv_transpose2d(x, S[key + ".weight"], S.get(key + ".bias"), stride, pad, out_pad)


def _bn(S, key, x, training):
    if training and (key + ".num_batches_tracked") in S:
        S[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, S[key + ".running_mean"], S[key + ".running_var"], S[key + ".weight"],
                        S[key + ".bias"], training, 0.1, 1e-5)


def _conf0(x0):
    """c0 = (d > 0) for 4-channel input, else channel 4 (confidence_depth_framework.py:469-472)."""
    x0_d = x0[:, 3:4]
    c0 = (x0_d > 0).to(x0.dtype) if x0.shape[1] == 4 else x0[:, 4:5]
    return x0_d, c0


def ged_forward(S, pre, x0, build_conf_input, training):
    """confidence_depth_framework.py:464-518."""
    rgb = x0[:, :3]
    x0_d, c0 = _conf0(x0)
    xout_d, cout_d = pyramid_apply(S, pre + "d_net.", x0_d, c0, training)
    up = lambda t, ref: F.interpolate(t, ref.shape[2:], mode="nearest")
    x1 = F.relu(_conv(S, pre + "conv1", torch.cat((xout_d, rgb, cout_d), 1), 2))
    x2 = F.relu(_conv(S, pre + "conv2", x1, 2))
    x3 = F.relu(_conv(S, pre + "conv3", x2, 2))
    x4 = F.relu(_conv(S, pre + "conv4", x3, 2))
    x5 = F.relu(_conv(S, pre + "conv5", x4, 2))
    x6 = F.leaky_relu(_conv(S, pre + "conv6", torch.cat((up(x5, x4), x4), 1)), 0.2)
    x7 = F.leaky_relu(_conv(S, pre + "conv7", torch.cat((up(x6, x3), x3), 1)), 0.2)
    x8 = F.leaky_relu(_conv(S, pre + "conv8", torch.cat((up(x7, x2), x2), 1)), 0.2)
    x9 = F.leaky_relu(_conv(S, pre + "conv9", torch.cat((up(x8, x1), x1), 1)), 0.2)
    last_in = torch.cat((up(x9, x0_d), x0_d), 1)
    x10 = F.leaky_relu(_conv(S, pre + "conv10", last_in), 0.2)
    feats = None
    if build_conf_input:
        feats = torch.cat((last_in, xout_d, cout_d, x10), 1) if x10.shape[1] == 1 else torch.sigmoid(x10[:, 1:2])
    return x10[:, 0:1], feats


def _stack(S, pre, x, n):
    for i in range(n):
        x = F.relu(_conv(S, "%s%d" % (pre, 2 * i), x))
    return x


def gms_forward(S, pre, x0, build_conf_input, training):
    """confidence_depth_framework.py:630-669."""
    rgb = x0[:, :3]
    x0_d, c0 = _conf0(x0)
    xout_d, cout_d = pyramid_apply(S, pre + "d_net.", x0_d, c0, training)
    xout_d = _stack(S, pre + "d.", xout_d, 6)
    xout_rgb = _stack(S, pre + "rgb.", torch.cat((rgb, cout_d), 1), 6)
    last_in = _stack(S, pre + "fuse.", torch.cat((xout_rgb, xout_d), 1), 6)
    xout = _conv(S, pre + "last_layer", last_in, 1, 0)
    feats = None
    if build_conf_input:
        feats = torch.cat((last_in, xout_d, cout_d, xout), 1) if xout.shape[1] == 1 else torch.sigmoid(xout[:, 1:2])
    return xout[:, 0:1], feats


def cbr3c1_forward(S, pre, x, training):
    """confidence_depth_framework.py:317-329."""
    x = _conv(S, pre + "conv1", x)
    x = F.leaky_relu(_bn(S, pre + "bn1", x, training), 0.2)
    return torch.sigmoid(_conv(S, pre + "conv2", x, 1, 0))


def _basic_block(S, pre, x, stride, training):
    """torchvision BasicBlock topology (resnet.py), as used by s2d_u_resnet.py:99-107."""
    out = F.relu(_bn(S, pre + "bn1", _conv(S, pre + "conv1", x, stride), training))
    out = _bn(S, pre + "bn2", _conv(S, pre + "conv2", out), training)
    if (pre + "downsample.0.weight") in S:
        x = _bn(S, pre + "downsample.1", _conv(S, pre + "downsample.0", x, stride, 0), training)
    return F.relu(out + x)


def _res_layer(S, pre, x, stride, training):
    x = _basic_block(S, pre + "0.", x, stride, training)
    return _basic_block(S, pre + "1.", x, 1, training)


def _cbr(S, pre, x, stride, pad, training):
    return F.leaky_relu(_bn(S, pre + "1", _conv(S, pre + "0", x, stride, pad), training), 0.2)


def _ctbr(S, pre, x, stride, out_pad, training):
    return F.leaky_relu(_bn(S, pre + "1", _convt(S, pre + "0", x, stride, 1, out_pad), training), 0.2)


def s2du_forward(S, pre, x, build_conf_input, training):
    """s2d_u_resnet.py:132-190 (rgbd / rgbdw inputs; unguided iff 'unguidednet.*' keys exist)."""
    cin = x.shape[1]
    img = _cbr(S, pre + "conv1_img.", x[:, :3], 1, 1, training)
    if cin == 3:
        conv1 = img
    else:
        if (pre + "unguidednet.nconv1.weight") in S:
            c0 = (x[:, 3:4] > 0).to(x.dtype) if cin == 4 else x[:, 4:5]
            densex, coutx = pyramid_apply(S, pre + "unguidednet.", x[:, 3:4], c0, training)
            d_in = torch.cat((densex, coutx), 1)
        else:
            d_in = x[:, 3:4] if cin == 4 else x[:, 3:5]
        conv1 = torch.cat((_cbr(S, pre + "conv1_d.", d_in, 1, 1, training), img), 1)
    conv2 = _res_layer(S, pre + "conv2.", conv1, 1, training)
    conv3 = _res_layer(S, pre + "conv3.", conv2, 2, training)
    conv4 = _res_layer(S, pre + "conv4.", conv3, 2, training)
    conv5 = _res_layer(S, pre + "conv5.", conv4, 2, training)
    conv6 = _cbr(S, pre + "conv6.", conv5, 2, 1, training)
    y = torch.cat((_ctbr(S, pre + "convt5.", conv6, 2, 1, training), conv5), 1)
    y = torch.cat((_ctbr(S, pre + "convt4.", y, 2, 1, training), conv4), 1)
    y = torch.cat((_ctbr(S, pre + "convt3.", y, 2, 1, training), conv3), 1)
    y = torch.cat((_ctbr(S, pre + "convt2.", y, 2, 1, training), conv2), 1)
    last_in = torch.cat((_ctbr(S, pre + "convt1.", y, 1, 0, training), conv1), 1)
    y = _conv(S, pre + "convtf.0", last_in, 1, 0)
    feats = None
    if build_conf_input:
        feats = torch.cat((last_in, y), 1) if y.shape[1] == 1 else torch.sigmoid(y[:, 1:2])
    depth = y[:, 0:1] if training else F.relu(y[:, 0:1])          # :185-188
    return depth, feats


def _unpool(x):
    """s2d_resnet.py:11-24: depthwise stride-2 transposed conv with kernel [[1,0],[0,0]] = zero stuffing."""
    N, C, H, W = x.shape
    out = x.new_zeros(N, C, 2 * H, 2 * W)
    out[:, :, ::2, ::2] = x
    return out


def _upproj(S, pre, x, training):
    """s2d_resnet.py:105-143."""
    x = _unpool(x)
    u = F.relu(_bn(S, pre + "upper_branch.batchnorm1", _conv(S, pre + "upper_branch.conv1", x, 1, 2), training))
    u = _bn(S, pre + "upper_branch.batchnorm2", _conv(S, pre + "upper_branch.conv2", u, 1, 1), training)
    b = _bn(S, pre + "bottom_branch.batchnorm", _conv(S, pre + "bottom_branch.conv", x, 1, 2), training)
    return F.relu(u + b)


def s2d_forward(S, pre, x, build_conf_input, training):
    """s2d_resnet.py:212-242 (resnet18 + upproj decoder)."""
    res = x.shape[2:]
    x = F.relu(_bn(S, pre + "bn1", _conv(S, pre + "conv1", x, 2, 3), training))
    x = F.max_pool2d(x, 3, 2, 1)
    x = _res_layer(S, pre + "layer1.", x, 1, training)
    x = _res_layer(S, pre + "layer2.", x, 2, training)
    x = _res_layer(S, pre + "layer3.", x, 2, training)
    x = _res_layer(S, pre + "layer4.", x, 2, training)
    x = _bn(S, pre + "bn2", _conv(S, pre + "conv2", x, 1, 0), training)
    for i in range(1, 5):
        x = _upproj(S, pre + "decoder.layer%d." % i, x, training)
    feats_lo = x
    x = F.interpolate(_conv(S, pre + "conv3", x), size=res, mode="bilinear", align_corners=True)
    feats = None
    if build_conf_input:
        feats = (F.interpolate(feats_lo, size=res, mode="bilinear", align_corners=True)
                 if x.shape[1] == 1 else torch.sigmoid(x[:, 1:2]))
    return x[:, 0:1], feats


DC_FORWARD = {
    "ged_depthcompnet": ged_forward,
    "gms_depthcompnet": gms_forward,
    "gudepthcompnet18": s2du_forward,
    "udepthcompnet18": s2du_forward,
    "resnet18": s2d_forward,
}


def framework_forward(S, inp, overall_arch, dc_arch, input_size, conf_arch=None, lossdc_arch=None, training=True):
    """confidence_depth_framework.py:217-235 with the per-part modes of :237-260.

    ``S`` uses the framework prefixes 'dc_model.', 'conf_model.', 'loss_dc_model.'."""
    has_ln = "ln" in overall_arch
    dc_train = training and ("dc1" in overall_arch)
    cf_train = training and ("cf1" in overall_arch)
    ln_train = training and ("ln1" in overall_arch)
    depth1, conf_x = DC_FORWARD[dc_arch](S, "dc_model.", inp[:, :input_size], has_ln, dc_train)
    if "cf" in overall_arch:
        assert conf_x is not None
        assert conf_arch == "cbr3-c1"
        conf1 = cbr3c1_forward(S, "conf_model.", conf_x, cf_train)
    else:
        conf1 = conf_x
    depth2 = None
    if has_ln:
        rgbd1c1 = torch.cat([inp[:, :3], depth1, conf1], 1)
        depth2, _ = DC_FORWARD[lossdc_arch](S, "loss_dc_model.", rgbd1c1, False, ln_train)
    return depth1, conf1, depth2


def trainable_keys(S, overall_arch):
    """confidence_depth_framework.py:262-292 (opt_params): which state entries are optimised."""
    keys = []
    for part, flag in (("dc_model.", "dc1"), ("conf_model.", "cf1"), ("loss_dc_model.", "ln1")):
        if flag in overall_arch:
            keys += [k for k in S if k.startswith(part) and S[k].dtype.is_floating_point
                     and not k.endswith(("running_mean", "running_var"))]
    return keys

"""``S2DResNet`` (the ``resnet18`` baseline of --dcnet-arch): drop-in for the reference's ``model_zoo/s2d_resnet.py:159-242``.

What the drop-in contract fixes is the constructor signature, ``out_channels`` / ``out_feature_channels`` (16, or 1 with
two output channels), the ``state_dict`` names -- ``conv1``, ``bn1``, torchvision BasicBlock names under ``layer1..layer4``, ``conv2``,
``bn2``, ``decoder.layer{1..4}.upper_branch.{conv1,batchnorm1,conv2,batchnorm2}``, ``decoder.layer{k}.bottom_branch.{conv,batchnorm}``,
``conv3`` -- and the N(0, sqrt(2/n)) initialisation of everything outside the ResNet stages (:26-40).  The module below is a parameter
container described by two tables; its forward runs on the tensor-core path of ``tc_s2d``: the 7x7 / stride-2 stem as one GEMM over an
im2col tensor, max-pooling, the BasicBlock stages, and the UpProj decoder (:105-143) whose 5x5 convolutions over the zero-stuffed
``Unpool`` output (:11-24) are evaluated as 3x3 convolutions of the low-resolution tensor with four parity blocks of output channels --
the zero-stuffed tensor is never formed.  ``dense.set_precision('fp32' | 'tf32')`` routes the same parameters through the ATen library
(the baseline the kernels are measured against).  Only what the reference's factory builds is provided: the ResNet-18 body with the
``upproj`` decoder (confidence_depth_framework.py:24-27).
"""
import collections
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock

from .. import dense

# (attribute, planes, stride of the first block) of the four stages; blocks per stage by depth
_STAGES = (("layer1", 64, 1), ("layer2", 128, 2), ("layer3", 256, 2), ("layer4", 512, 2))
_BLOCKS = {18: (2, 2, 2, 2)}


def _he_init(module):
    """N(0, sqrt(2 / (k*k*out_channels))) for convolutions, (1, 0) for BatchNorm (s2d_resnet.py:26-40)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data.normal_(0, math.sqrt(2. / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
    return module


def _stage(inplanes, planes, stride, blocks):
    """torchvision's ``_make_layer`` with its default initialisation (kaiming_normal_ fan_out; the reference keeps it for the stages)."""
    down = None
    if stride != 1 or inplanes != planes:
        down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
    st = nn.Sequential(BasicBlock(inplanes, planes, stride, down), *[BasicBlock(planes, planes) for _ in range(blocks - 1)])
    for m in st.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    return st


class _UpProjModule(nn.Module):
    """Parameters of one UpProj module (s2d_resnet.py:109-136): Unpool -> [5x5 -> BN -> ReLU -> 3x3 -> BN] + [5x5 -> BN] -> ReLU."""

    def __init__(self, cin):
        super().__init__()
        cout = cin // 2
        self.upper_branch = nn.Sequential(collections.OrderedDict([
            ("conv1", nn.Conv2d(cin, cout, 5, 1, 2, bias=False)), ("batchnorm1", nn.BatchNorm2d(cout)), ("relu", nn.ReLU()),
            ("conv2", nn.Conv2d(cout, cout, 3, 1, 1, bias=False)), ("batchnorm2", nn.BatchNorm2d(cout))]))
        self.bottom_branch = nn.Sequential(collections.OrderedDict([
            ("conv", nn.Conv2d(cin, cout, 5, 1, 2, bias=False)), ("batchnorm", nn.BatchNorm2d(cout))]))

    def forward(self, x):
        """Library path only: zero stuffing written out."""
        n, c, h, w = x.shape
        z = x.new_zeros(n, c, 2 * h, 2 * w)
        z[:, :, ::2, ::2] = x
        return F.relu(self.upper_branch(z) + self.bottom_branch(z))


class _UpProj(nn.Module):
    def __init__(self, cin):
        super().__init__()
        for k in range(4):
            setattr(self, "layer%d" % (k + 1), _UpProjModule(cin >> k))

    def forward(self, x):
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class S2DResNet(nn.Module):

    def __init__(self, layers, decoder, in_channels=3, out_channels=1, pretrained=True):
        if layers not in [18, 34, 50, 101, 152]:
            raise RuntimeError('Only 18, 34, 50, 101, and 152 layer model are defined for ResNet. Got {}'.format(layers))
        if layers not in _BLOCKS or decoder != "upproj":
            raise NotImplementedError("the B200 path implements the ResNet-18 body with the upproj decoder -- what the "
                                      "reference's factory builds (confidence_depth_framework.py:24-27); got layers=%r decoder=%r"
                                      % (layers, decoder))
        if pretrained:
            raise RuntimeError("ImageNet-pretrained ResNet weights need a download; load them with --dcnet-pretrained path:key")
        super(S2DResNet, self).__init__()
        self.layers, self.in_channels, self.out_channels = layers, in_channels, out_channels
        self.out_feature_channels = 1 if out_channels == 2 else 16
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        if in_channels == 3:
            nn.init.kaiming_normal_(self.conv1.weight, mode="fan_out", nonlinearity="relu")
        else:
            _he_init(self.conv1)
        inplanes = 64
        for (name, planes, stride), blocks in zip(_STAGES, _BLOCKS[layers]):
            setattr(self, name, _stage(inplanes, planes, stride, blocks))
            inplanes = planes
        self.conv2 = _he_init(nn.Conv2d(512, 256, kernel_size=1, bias=False))
        self.bn2 = nn.BatchNorm2d(256)
        self.decoder = _he_init(_UpProj(256))
        self.conv3 = _he_init(nn.Conv2d(16, out_channels, kernel_size=3, stride=1, padding=1, bias=False))

    def forward(self, x, build_conf_input):
        if dense.tc_enabled():
            dense._need_cuda(x)
            from .. import tc_s2d
            return tc_s2d.s2d_forward(self, x, build_conf_input)
        return self._forward_library(x, build_conf_input)

    def _forward_library(self, x, build_conf_input):
        """The same graph through ATen / cuDNN (precision modes 'fp32' / 'tf32'): measurement baseline, not the product path."""
        dense._need_cuda(x)
        size = (x.size(2), x.size(3))
        t = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        for name, _, _ in _STAGES:
            t = getattr(self, name)(t)
        feat = self.decoder(self.bn2(self.conv2(t)))
        y = F.interpolate(self.conv3(feat), size=size, mode='bilinear', align_corners=True)
        features = None
        if build_conf_input:
            features = (F.interpolate(feat, size=size, mode='bilinear', align_corners=True) if self.out_channels == 1
                        else torch.sigmoid(y[:, 1:2]))
        return y[:, 0:1], features

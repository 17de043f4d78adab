"""``S2DUResNet`` (udepthcompnet18 / gudepthcompnet18): drop-in for the reference's ``model_zoo/s2d_u_resnet.py:64-190``.

What the drop-in contract fixes is the constructor signature, the attribute / ``state_dict`` names (``conv1_img.{0,1}``,
``conv1_d``, torchvision BasicBlock names under ``conv2..conv5``, ``conv6``, ``convt5..convt1``, ``convtf.0``), the N(0, 0.02)
initialisation, train / eval behaviour (raw y in training, relu(y) in eval, :185-188) and the H, W = 0 mod 16 requirement.  The
module below is a parameter container described by two layer tables; its forward runs on the tensor-core path of ``tc_bn``
(every convolution, transposed convolution, BatchNorm, activation and residual add forward and backward on this library's kernels).
``dense.set_precision('fp32' | 'tf32')`` routes the same parameters through the ATen library instead (the baseline the kernels are
measured against).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock

from .. import dense
from .nconv_sd import CNN as unguided_net

# (attribute, transposed?, channels in, channels out, kernel, stride, padding, output_padding) of the layers around the ResNet stages
_DECODER = (("conv6", False, 512, 512, 3, 2, 1, 0), ("convt5", True, 512, 256, 3, 2, 1, 1), ("convt4", True, 768, 128, 3, 2, 1, 1),
            ("convt3", True, 384, 64, 3, 2, 1, 1), ("convt2", True, 192, 64, 3, 2, 1, 1), ("convt1", True, 128, 64, 3, 1, 1, 0))
# (attribute, planes, stride of the first block) of the four encoder stages: ResNet-18 without its stem stride and max-pool
_ENCODER = (("conv2", 64, 1), ("conv3", 128, 2), ("conv4", 256, 2), ("conv5", 512, 2))


def _normal_init(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            m.weight.data.normal_(0, 0.02)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
    return module


def _unit(cin, cout, k, stride=1, padding=0, output_padding=0, transposed=False, bn=True, act=True):
    """Sequential(conv | convT, [BatchNorm2d], [LeakyReLU(0.2)]) with the reference's child indices 0 / 1 / 2."""
    if transposed:
        mods = [nn.ConvTranspose2d(cin, cout, k, stride, padding, output_padding, bias=not bn)]
    else:
        mods = [nn.Conv2d(cin, cout, k, stride, padding, bias=not bn)]
    if bn:
        mods.append(nn.BatchNorm2d(cout))
    if act:
        mods.append(nn.LeakyReLU(0.2, inplace=True))
    return _normal_init(nn.Sequential(*mods))


def _stage(inplanes, planes, stride):
    """torchvision's ``_make_layer`` for two BasicBlocks (keeps the names conv1 / bn1 / conv2 / bn2 / downsample.{0,1})."""
    down = None
    if stride != 1 or inplanes != planes:
        down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
    return _normal_init(nn.Sequential(BasicBlock(inplanes, planes, stride, down), BasicBlock(planes, planes)))


class S2DUResNet(nn.Module):

    def __init__(self, layers=18, in_channels=3, out_channels=1, pretrained=True, unguided=False):
        if layers != 18:
            raise NotImplementedError("the B200 path implements the ResNet-18 body (the only depth the reference's factory builds, "
                                      "confidence_depth_framework.py:35-40); got layers=%r" % (layers,))
        if pretrained:
            raise RuntimeError("ImageNet-pretrained ResNet weights need a download; load them with --dcnet-pretrained path:key")
        super(S2DUResNet, self).__init__()
        self.in_channels, self.out_channels, self.unguided = in_channels, out_channels, unguided
        self.out_feature_channels = 1 if out_channels == 2 else 129
        self._adc_lazy_features = False      # set by the framework model when a cbr3-c1 head consumes the features in place
        if unguided:
            self.unguidednet = unguided_net()
        if in_channels == 3:
            self.conv1_img = _unit(3, 64, 3, 1, 1)
        elif in_channels in (4, 5):
            self.conv1_img = _unit(3, 48, 3, 1, 1)
            self.conv1_d = _unit(2 if (unguided or in_channels == 5) else 1, 16, 3, 1, 1)
        inplanes = 64
        for name, planes, stride in _ENCODER:
            setattr(self, name, _stage(inplanes, planes, stride))
            inplanes = planes
        for name, transposed, cin, cout, k, s, p, op in _DECODER:
            setattr(self, name, _unit(cin, cout, k, s, p, op, transposed))
        self.convtf = _unit(128, out_channels, 1, 1, bn=False, act=False)

    def forward(self, x, build_conf_input):
        if dense.tc_enabled():
            dense._need_cuda(x)
            from .. import tc_bn
            return tc_bn.s2du_forward(self, x, build_conf_input)
        return self._forward_library(x, build_conf_input)

    def _forward_library(self, x, build_conf_input):
        """The same graph through ATen / cuDNN (precision modes 'fp32' / 'tf32'): measurement baseline, not the product path."""
        dense._need_cuda(x)
        first = self.conv1_img(x[:, :3])
        if self.in_channels != 3:
            if self.unguided:
                c0 = (x[:, 3:4] > 0).float() if self.in_channels == 4 else x[:, 4:5]
                d_in = torch.cat(self.unguidednet(x[:, 3:4], c0), 1)
            else:
                d_in = x[:, 3:4] if self.in_channels == 4 else x[:, 3:5]
            first = torch.cat((self.conv1_d(d_in), first), 1)
        skips, t = [first], first
        for name, _, _ in _ENCODER:
            t = getattr(self, name)(t)
            skips.append(t)
        t = self.conv6(t)
        for name in ("convt5", "convt4", "convt3", "convt2", "convt1"):
            t = torch.cat((getattr(self, name)(t), skips.pop()), 1)
        y = self.convtf(t)
        features = None
        if build_conf_input:
            features = torch.cat((t, y), 1) if self.out_channels == 1 else torch.sigmoid(y[:, 1:2])
        return (y[:, 0:1] if self.training else F.relu(y[:, 0:1])), features

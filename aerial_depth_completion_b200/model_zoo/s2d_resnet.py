"""Sparse-to-dense ResNet baseline -- drop-in for the reference's ``model_zoo/s2d_resnet.py`` (``S2DResNet`` with the
deconv / upconv / upproj decoders; the framework uses resnet18 + upproj).  Same constructors and ``state_dict`` keys.
``Unpool`` (a depthwise stride-2 transposed conv with the fixed kernel [[1,0],[0,0]], s2d_resnet.py:11-24) is plain zero
stuffing and is written as such (no weight tensor, no .cuda() at construction)."""
import collections
import math

import torch
import torch.nn as nn
import torchvision.models

from .. import dense


class Unpool(nn.Module):
    def __init__(self, num_channels, stride=2):
        super(Unpool, self).__init__()
        self.num_channels = num_channels
        self.stride = stride

    def forward(self, x):
        n, c, h, w = x.shape
        out = x.new_zeros(n, c, h * self.stride, w * self.stride)
        out[:, :, ::self.stride, ::self.stride] = x
        return out


def weights_init(m):
    """He-style N(0, sqrt(2/n)) (s2d_resnet.py:26-40)."""
    if isinstance(m, nn.Conv2d):
        n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        m.weight.data.normal_(0, math.sqrt(2. / n))
        if m.bias is not None:
            m.bias.data.zero_()
    elif isinstance(m, nn.ConvTranspose2d):
        n = m.kernel_size[0] * m.kernel_size[1] * m.in_channels
        m.weight.data.normal_(0, math.sqrt(2. / n))
        if m.bias is not None:
            m.bias.data.zero_()
    elif isinstance(m, nn.BatchNorm2d):
        m.weight.data.fill_(1)
        m.bias.data.zero_()


class Decoder(nn.Module):
    names = ['deconv2', 'deconv3', 'upconv', 'upproj']

    def __init__(self):
        super(Decoder, self).__init__()
        self.layer1 = self.layer2 = self.layer3 = self.layer4 = None

    def forward(self, x):
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class DeConv(Decoder):
    def __init__(self, in_channels, kernel_size):
        assert kernel_size >= 2, "kernel_size out of range: {}".format(kernel_size)
        super(DeConv, self).__init__()

        def convt(c):
            padding = (kernel_size - 1) // 2
            output_padding = kernel_size % 2
            assert -2 - 2 * padding + kernel_size + output_padding == 0, "deconv parameters incorrect"
            return nn.Sequential(collections.OrderedDict([
                ("deconv{}".format(kernel_size), dense.ConvTranspose2d(c, c // 2, kernel_size, 2, padding, output_padding, bias=False)),
                ('batchnorm', nn.BatchNorm2d(c // 2)),
                ('relu', nn.ReLU(inplace=True)),
            ]))

        self.layer1, self.layer2 = convt(in_channels), convt(in_channels // 2)
        self.layer3, self.layer4 = convt(in_channels // 4), convt(in_channels // 8)


class UpConv(Decoder):
    def upconv_module(self, in_channels):
        return nn.Sequential(collections.OrderedDict([
            ('unpool', Unpool(in_channels)),
            ('conv', dense.Conv2d(in_channels, in_channels // 2, kernel_size=5, stride=1, padding=2, bias=False)),
            ('batchnorm', nn.BatchNorm2d(in_channels // 2)),
            ('relu', nn.ReLU()),
        ]))

    def __init__(self, in_channels):
        super(UpConv, self).__init__()
        self.layer1, self.layer2 = self.upconv_module(in_channels), self.upconv_module(in_channels // 2)
        self.layer3, self.layer4 = self.upconv_module(in_channels // 4), self.upconv_module(in_channels // 8)


class UpProj(Decoder):
    class UpProjModule(nn.Module):
        def __init__(self, in_channels):
            super(UpProj.UpProjModule, self).__init__()
            out_channels = in_channels // 2
            self.unpool = Unpool(in_channels)
            self.upper_branch = nn.Sequential(collections.OrderedDict([
                ('conv1', dense.Conv2d(in_channels, out_channels, kernel_size=5, stride=1, padding=2, bias=False)),
                ('batchnorm1', nn.BatchNorm2d(out_channels)),
                ('relu', nn.ReLU()),
                ('conv2', dense.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)),
                ('batchnorm2', nn.BatchNorm2d(out_channels)),
            ]))
            self.bottom_branch = nn.Sequential(collections.OrderedDict([
                ('conv', dense.Conv2d(in_channels, out_channels, kernel_size=5, stride=1, padding=2, bias=False)),
                ('batchnorm', nn.BatchNorm2d(out_channels)),
            ]))
            self.relu = nn.ReLU()

        def forward(self, x):
            x = self.unpool(x)
            return self.relu(self.upper_branch(x) + self.bottom_branch(x))

    def __init__(self, in_channels):
        super(UpProj, self).__init__()
        self.layer1, self.layer2 = self.UpProjModule(in_channels), self.UpProjModule(in_channels // 2)
        self.layer3, self.layer4 = self.UpProjModule(in_channels // 4), self.UpProjModule(in_channels // 8)


def choose_decoder(decoder, in_channels):
    if decoder[:6] == 'deconv':
        assert len(decoder) == 7
        return DeConv(in_channels, int(decoder[6]))
    elif decoder == "upproj":
        return UpProj(in_channels)
    elif decoder == "upconv":
        return UpConv(in_channels)
    assert False, "invalid option for decoder: {}".format(decoder)


class S2DResNet(nn.Module):
    def __init__(self, layers, decoder, in_channels=3, out_channels=1, pretrained=True):
        if layers not in [18, 34, 50, 101, 152]:
            raise RuntimeError('Only 18, 34, 50, 101, and 152 layer model are defined for ResNet. Got {}'.format(layers))
        super(S2DResNet, self).__init__()
        self.out_channels = out_channels
        self.out_feature_channels = 1 if out_channels == 2 else 16
        if pretrained:
            raise RuntimeError("ImageNet-pretrained ResNet weights need a download; load them with --dcnet-pretrained path:key")
        body = torchvision.models.__dict__['resnet{}'.format(layers)](weights=None)
        if in_channels == 3:
            self.conv1, self.bn1 = body._modules['conv1'], body._modules['bn1']
        else:
            self.conv1 = dense.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
            self.bn1 = nn.BatchNorm2d(64)
            weights_init(self.conv1)
            weights_init(self.bn1)
        self.relu = body._modules['relu']
        self.maxpool = body._modules['maxpool']
        self.layer1, self.layer2 = body._modules['layer1'], body._modules['layer2']
        self.layer3, self.layer4 = body._modules['layer3'], body._modules['layer4']
        del body
        num_channels = 512 if layers <= 34 else 2048
        self.conv2 = dense.Conv2d(num_channels, num_channels // 2, kernel_size=1, bias=False)
        self.bn2 = nn.BatchNorm2d(num_channels // 2)
        self.decoder = choose_decoder(decoder, num_channels // 2)
        self.conv3 = dense.Conv2d(num_channels // 32, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv2.apply(weights_init)
        self.bn2.apply(weights_init)
        self.decoder.apply(weights_init)
        self.conv3.apply(weights_init)

    def forward(self, x, build_conf_input):
        resolution = (x.size()[2], x.size()[3])
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.bn2(self.conv2(x))
        features = self.decoder(x)
        x = self.conv3(features)
        x = nn.functional.interpolate(x, size=resolution, mode='bilinear', align_corners=True)
        if build_conf_input:
            if self.out_channels == 1:
                features = nn.functional.interpolate(features, size=resolution, mode='bilinear', align_corners=True)
            else:
                features = torch.sigmoid(x[:, 1:2, :, :])
        else:
            features = None
        return x[:, 0:1, :, :], features

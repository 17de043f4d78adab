"""Full-resolution ResNet-18 U-Net -- drop-in for the reference's ``model_zoo/s2d_u_resnet.py`` (``S2DUResNet``:
``udepthcompnet18`` / ``gudepthcompnet18``).  Same constructor, ``state_dict`` keys (torchvision BasicBlock names for
conv2..conv5), train/eval behaviour (raw y in training, relu(y) in eval, s2d_u_resnet.py:185-188) and the H,W % 16 == 0
requirement.  ``unguided=True`` prepends the fused CUDA pyramid (``unguidednet``)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models import resnet

from .. import dense
from .nconv_sd import CNN as unguided_net


def init_weights(m):
    if isinstance(m, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d)):
        m.weight.data.normal_(0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()
    elif isinstance(m, nn.BatchNorm2d):
        m.weight.data.fill_(1)
        m.bias.data.zero_()


def conv_bn_relu(in_channels, out_channels, kernel_size, stride=1, padding=0, bn=True, relu=True):
    layers = [dense.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=not bn)]
    if bn:
        layers.append(nn.BatchNorm2d(out_channels))
    if relu:
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    layers = nn.Sequential(*layers)
    for m in layers.modules():
        init_weights(m)
    return layers


def convt_bn_relu(in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bn=True, relu=True):
    layers = [dense.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, output_padding, bias=not bn)]
    if bn:
        layers.append(nn.BatchNorm2d(out_channels))
    if relu:
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    layers = nn.Sequential(*layers)
    for m in layers.modules():
        init_weights(m)
    return layers


def _resnet_body(layers, pretrained):
    if pretrained:
        raise RuntimeError("ImageNet-pretrained ResNet weights need a download; load them with --dcnet-pretrained path:key")
    model = resnet.__dict__['resnet{}'.format(layers)](weights=None)
    model.apply(init_weights)
    return model


class S2DUResNet(nn.Module):

    def __init__(self, layers=18, in_channels=3, out_channels=1, pretrained=True, unguided=False):
        assert (layers in [18, 34, 50, 101, 152]), \
            'Only layers 18, 34, 50, 101, and 152 are defined, but got {}'.format(layers)
        super(S2DUResNet, self).__init__()
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.out_feature_channels = 1 if out_channels == 2 else 129
        self.unguided = unguided
        extra_channel = 0
        if unguided:
            self.unguidednet = unguided_net()
            extra_channel = 1
        if in_channels == 3:
            self.conv1_img = conv_bn_relu(3, 64, kernel_size=3, stride=1, padding=1)
        if in_channels == 4:
            self.conv1_img = conv_bn_relu(3, 48, kernel_size=3, stride=1, padding=1)
            self.conv1_d = conv_bn_relu(1 + extra_channel, 16, kernel_size=3, stride=1, padding=1)
        if in_channels == 5:
            self.conv1_img = conv_bn_relu(3, 48, kernel_size=3, stride=1, padding=1)
            self.conv1_d = conv_bn_relu(2, 16, kernel_size=3, stride=1, padding=1)
        body = _resnet_body(layers, pretrained)
        self.conv2 = body._modules['layer1']
        self.conv3 = body._modules['layer2']
        self.conv4 = body._modules['layer3']
        self.conv5 = body._modules['layer4']
        del body
        num_channels = 512 if layers <= 34 else 2048
        self.conv6 = conv_bn_relu(num_channels, 512, kernel_size=3, stride=2, padding=1)
        self.convt5 = convt_bn_relu(512, 256, kernel_size=3, stride=2, padding=1, output_padding=1)
        self.convt4 = convt_bn_relu(768, 128, kernel_size=3, stride=2, padding=1, output_padding=1)
        self.convt3 = convt_bn_relu(256 + 128, 64, kernel_size=3, stride=2, padding=1, output_padding=1)
        self.convt2 = convt_bn_relu(128 + 64, 64, kernel_size=3, stride=2, padding=1, output_padding=1)
        self.convt1 = convt_bn_relu(128, 64, kernel_size=3, stride=1, padding=1)
        self.convtf = conv_bn_relu(128, self.out_channels, kernel_size=1, stride=1, bn=False, relu=False)

    def forward(self, x, build_conf_input):
        if self.in_channels == 3:
            conv1 = self.conv1_img(x[:, :3, :, :])
        else:
            conv1_img = self.conv1_img(x[:, :3, :, :])
            if self.unguided:
                c0 = (x[:, 3:4, :, :] > 0).float() if self.in_channels == 4 else x[:, 4:5, :, :]
                densex, coutx = self.unguidednet(x[:, 3:4, :, :], c0)
                conv1_d = self.conv1_d(torch.cat((densex, coutx), 1))
            else:
                conv1_d = self.conv1_d(x[:, 3:4, :, :] if self.in_channels == 4 else x[:, 3:5, :, :])
            conv1 = torch.cat((conv1_d, conv1_img), 1)
        conv2 = self.conv2(conv1)
        conv3 = self.conv3(conv2)
        conv4 = self.conv4(conv3)
        conv5 = self.conv5(conv4)
        conv6 = self.conv6(conv5)
        y = torch.cat((self.convt5(conv6), conv5), 1)
        y = torch.cat((self.convt4(y), conv4), 1)
        y = torch.cat((self.convt3(y), conv3), 1)
        y = torch.cat((self.convt2(y), conv2), 1)
        last_layer_input = torch.cat((self.convt1(y), conv1), 1)
        y = self.convtf(last_layer_input)
        features = None
        if build_conf_input:
            if self.out_channels == 1:
                features = torch.cat((last_layer_input, y), 1)   # 128 + 1
            else:
                features = torch.sigmoid(y[:, 1:2, :, :])
        depth = y[:, 0:1, :, :] if self.training else F.relu(y[:, 0:1, :, :])
        return depth, features

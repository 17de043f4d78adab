"""Framework composition -- drop-in for the reference's ``model_zoo/confidence_depth_framework.py``.

Same public surface: ``ConfidenceDepthFrameworkFactory`` (create_model / create_dc_model /
create_conf_model / create_loss / create_loss_fromstate / load_weights / to_device / get_state /
create_model_from_state), ``ConfidenceDepthFrameworkModel`` (forward / train / opt_params), the nets ``GEDNet``,
``GMSNet``, the confidence heads and the masked losses, with identical constructor arguments, attribute names,
``state_dict`` keys and error messages.  The unguided pyramid inside every net is the fused CUDA op
(``nconv_sd.CNN``); one process drives one GPU (multi-GPU = ``parallel.DataParallelTrainer``, not nn.DataParallel).
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import dense
from .nconv_sd import CNN as unguided_net
from .s2d_resnet import S2DResNet
from .s2d_u_resnet import S2DUResNet


class ConfidenceDepthFrameworkFactory():
    """confidence_depth_framework.py:13-187."""

    def __init__(self):
        pass

    def create_dc_model(self, model_arch, pretrained_args, input_type, output_type):
        use_resnet_pretrained = pretrained_args == 'resnet'
        n_in, n_out = len(input_type), len(output_type)
        if model_arch == 'resnet18':
            model = S2DResNet(layers=18, decoder='upproj', in_channels=n_in, out_channels=n_out,
                              pretrained=use_resnet_pretrained)
        elif model_arch == 'udepthcompnet18':
            model = S2DUResNet(layers=18, in_channels=n_in, out_channels=n_out, pretrained=use_resnet_pretrained,
                               unguided=False)
        elif model_arch == 'gudepthcompnet18':
            model = S2DUResNet(layers=18, in_channels=n_in, out_channels=n_out, pretrained=use_resnet_pretrained,
                               unguided=True)
        elif model_arch == 'gms_depthcompnet':
            model = GMSNet(out_channels=n_out)
        elif model_arch == 'ged_depthcompnet':
            model = GEDNet(out_channels=n_out)
        else:
            raise RuntimeError('Model: {} not found.'.format(model_arch))
        if not use_resnet_pretrained and pretrained_args:
            self.load_weights(pretrained_args, model)
        return model

    def create_conf_model(self, model_arch, pretrained_args, dc_model):
        in_channels = dc_model.out_feature_channels
        if model_arch == 'cbr3-c1':
            model = CBR3C1Confidence(in_channels=in_channels)
        elif model_arch in ('cbr3-cbr1-c1', 'cbr3-cbr1-c1res'):       # both names map to the Res head (:51-54)
            model = CBR3CBR1C1ResConfidence(in_channels=in_channels)
        elif model_arch == 'forward':
            model = ForwardConfidence(in_channels=in_channels)
        else:
            raise RuntimeError('Confidence net does not exist: {}'.format(model_arch))
        if pretrained_args:
            self.load_weights(pretrained_args, model)
        return model

    def create_loss(self, criterion, dual=False, weight1=0):
        if criterion == 'l2':
            loss = MaskedMSELoss()
        elif criterion == 'l1':
            loss = MaskedL1Loss()
        elif criterion == 'il1':
            loss = InvertedMaskedL1Loss()
        elif criterion == 'absrel':
            loss = MaskedAbsRelLoss()
        if dual:
            loss = DualLoss(loss, loss, weight1)
        loss_definition = {'criterion': criterion, 'dual': dual, 'weight1': weight1}
        return loss, loss_definition

    def create_loss_fromstate(self, definition):
        return self.create_loss(definition['criterion'], definition['dual'], definition['weight1'])

    @staticmethod
    def load_weights(filename, model):
        """'path:key' (key in dc_weights / conf_weights / lossdc_weights) or an OrderedDict of tensors (:84-110)."""
        if isinstance(filename, OrderedDict):
            weights, key = filename, 'dict'
        else:
            parts = filename.split(':')
            if len(parts) != 2 or not os.path.exists(parts[0]):
                return
            file, key = parts
            weights = torch.load(file)['model_state'][key]
        model_dict = model.state_dict()
        used = ignored = 0
        for k, v in weights.items():
            if k in model_dict:
                model_dict[k] = v
                used += 1
            else:
                ignored += 1
        print('=>loading weights: key {} - used {} - ignored {}'.format(key, used, ignored))
        model.load_state_dict(model_dict)

    def create_model(self, input_type, overall_arch, dc_arch, dc_weights, conf_arch=None, conf_weights=None,
                     lossdc_arch=None, lossdc_weights=None):
        cdfmodel = ConfidenceDepthFrameworkModel()
        cdfmodel.dc_arch, cdfmodel.conf_arch, cdfmodel.loss_dc_arch = dc_arch, conf_arch, lossdc_arch
        cdfmodel.overall_arch, cdfmodel.input_type = overall_arch, input_type
        if 'dc' in overall_arch:
            output_type = 'd' if ('only' in overall_arch or 'cf' in overall_arch) else 'dc'
            cdfmodel.dc_model = self.create_dc_model(dc_arch, dc_weights, input_type, output_type)
        if 'cf' in overall_arch:
            cdfmodel.conf_model = self.create_conf_model(model_arch=conf_arch, pretrained_args=conf_weights,
                                                         dc_model=cdfmodel.dc_model)
        if 'ln' in overall_arch:
            cdfmodel.loss_dc_model = self.create_dc_model(model_arch=lossdc_arch, pretrained_args=lossdc_weights,
                                                          input_type='rgbdc', output_type='d')
        cdfmodel.input_size = len(input_type)
        return cdfmodel

    def to_device(self, cdfmodel):
        """One process = one GPU: the model moves to the current CUDA device.  (The reference wraps it in
        nn.DataParallel when several GPUs are visible, :145-159; here multi-GPU runs are launched one rank per GPU and
        gradients are all-reduced by ``aerial_depth_completion_b200.parallel``.)"""
        if not torch.cuda.is_available():
            raise RuntimeError('the B200 path needs a CUDA device (no CPU fallback)')
        cdfmodel.cuda()
        return cdfmodel, cdfmodel.opt_params()

    def get_state(self, x):
        cdfmodel = x.module if hasattr(x, 'module') else x
        has_cf, has_ln = cdfmodel.conf_model is not None, cdfmodel.loss_dc_model is not None
        return {'input_type': cdfmodel.input_type,
                'overall_arch': cdfmodel.overall_arch,
                'dc_arch': cdfmodel.dc_arch,
                'dc_weights': cdfmodel.dc_model.state_dict(),
                'conf_arch': cdfmodel.conf_arch if has_cf else None,
                'conf_weights': cdfmodel.conf_model.state_dict() if has_cf else None,
                'loss_dc_arch': cdfmodel.loss_dc_arch if has_ln else None,
                'lossdc_weights': cdfmodel.loss_dc_model.state_dict() if has_ln else None}

    def create_model_from_state(self, state):
        return self.create_model(state['input_type'], state['overall_arch'], state['dc_arch'], state['dc_weights'],
                                 state['conf_arch'], state['conf_weights'], state['loss_dc_arch'], state['lossdc_weights'])


def init_weights(m):
    """N(0, 0.02) conv / convT weights, zero biases, unit BN (:190-202; from Ma & Karaman 2018)."""
    if isinstance(m, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d)):
        m.weight.data.normal_(0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()
    elif isinstance(m, nn.BatchNorm2d):
        m.weight.data.fill_(1)
        m.bias.data.zero_()


class ConfidenceDepthFrameworkModel(torch.nn.Module):
    """dc net -> confidence head -> loss net (confidence_depth_framework.py:205-292)."""

    def __init__(self):
        super(ConfidenceDepthFrameworkModel, self).__init__()
        self.dc_model = None
        self.conf_model = None
        self.loss_dc_model = None
        self.overall_arch = ''
        self.input_size = 0   # 3: rgb, 4: rgbd, 5: rgbdc

    def forward(self, input):
        dc_x = input[:, :self.input_size, :, :]
        if hasattr(self.dc_model, '_adc_lazy_features'):
            # a cbr3-c1 head behind a BN net reads the 129 feature channels in place from the net's NHWC buffer (tc_bn.LazyFeatures)
            self.dc_model._adc_lazy_features = isinstance(self.conf_model, CBR3C1Confidence) and dense.tc_enabled()
        depth1, conf_x = self.dc_model(dc_x, (self.loss_dc_model is not None))
        if self.conf_model is not None:
            assert (conf_x is not None), 'dc_model does not support extern confidence net'
            conf1 = self.conf_model(conf_x)
        else:
            conf1 = conf_x
        depth2 = None
        if self.loss_dc_model is not None:
            rgbd1c1 = torch.cat([input[:, :3, :, :], depth1, conf1], dim=1)
            depth2, _ = self.loss_dc_model(rgbd1c1, False)
        return depth1, conf1, depth2

    def train(self, mode=True):
        print(mode)    # the reference prints the flag on every call (:238); kept so logs line up
        self.training = mode
        self.dc_model.train(mode if 'dc1' in self.overall_arch else False)
        if self.conf_model is not None:
            self.conf_model.train(mode if 'cf1' in self.overall_arch else False)
        if self.loss_dc_model is not None:
            self.loss_dc_model.train(mode if 'ln1' in self.overall_arch else False)
        return self

    def opt_params(self):
        def all_no_grad(model):
            for param in model.parameters():
                param.requires_grad = False

        opt_parameters = []
        for flag_on, flag_off, model in (('dc1', 'dc0', self.dc_model), ('cf1', 'cf0', self.conf_model),
                                         ('ln1', 'ln0', self.loss_dc_model)):
            if flag_on in self.overall_arch:
                assert model is not None
                opt_parameters += model.parameters()
            if flag_off in self.overall_arch:
                assert model is not None
                all_no_grad(model)
        return opt_parameters


###########################################
# Confidence heads (confidence_depth_framework.py:299-419)
###########################################

class CBR3C1Confidence(nn.Module):
    """conv3x3 (no bias) -> BN -> LeakyReLU(0.2) -> conv1x1 -> sigmoid (:299-329)."""

    def __init__(self, in_channels):
        super(CBR3C1Confidence, self).__init__()
        self.conv1 = nn.Conv2d(in_channels, in_channels, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.relu1 = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv2d(in_channels, 1, 1, 1, 0, bias=True)
        for m in (self.conv1, self.bn1, self.conv2):
            init_weights(m)

    def forward(self, x):
        if dense.tc_enabled():
            from .. import tc_bn
            if isinstance(x, tc_bn.LazyFeatures):
                return tc_bn.conf_head(self, x)
            dense._need_cuda(x)
            if self.conv1.in_channels <= 80:
                # conv1 + its gradients on the tcgen05 kernels, BN + LeakyReLU + conv2 + sigmoid fused (csrc/conf_head.cu)
                from .. import tc_train
                return tc_train.conf_head(self, x)
            # wider features (129 channels behind gudepthcompnet18): output-channel chunks + K tail, general BN passes (tc_bn.py)
            return tc_bn.conf_head(self, x)
        return torch.sigmoid(self.conv2(self.relu1(self.bn1(self.conv1(x)))))


class CBR3CBR1C1ResConfidence(nn.Module):
    """(:372-409) conv3-BN-LReLU -> conv1-BN-LReLU -> conv1 over cat(second, first) -> sigmoid.  Outside the hot-path scope (SURVEY 8a
    lists cbr3-c1 only): kept so that --confnet-arch cbr3-cbr1-c1 / cbr3-cbr1-c1res resolve as in the reference (:51-54); it runs on ATen
    modules, not on this library's kernels, and no measurement or parity claim covers it."""

    def __init__(self, in_channels):
        super(CBR3CBR1C1ResConfidence, self).__init__()
        self.conv1 = nn.Conv2d(in_channels, in_channels, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.relu1 = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv2d(in_channels, in_channels // 2, 1, 1, 0, bias=False)
        self.bn2 = nn.BatchNorm2d(in_channels // 2)
        self.relu2 = nn.LeakyReLU(0.2, inplace=True)
        self.conv3 = nn.Conv2d(in_channels + (in_channels // 2), 1, 1, 1, 0, bias=True)
        for m in (self.conv1, self.bn1, self.conv2, self.bn2, self.conv3):
            init_weights(m)

    def forward(self, x):
        y = self.relu1(self.bn1(self.conv1(x)))
        x = self.relu2(self.bn2(self.conv2(y)))
        return torch.sigmoid(self.conv3(torch.cat([x, y], dim=1)))


class ForwardConfidence(nn.Module):
    def __init__(self, in_channels):
        super(ForwardConfidence, self).__init__()

    def forward(self, x):
        assert (x.size()[1] == 1), 'forward confidence need to have only one channel'
        return x


################################################
# depth completion nets
################################################

def _c0_of(x0):
    """c0 = (d > 0) for rgbd input, channel 4 for rgbdw (confidence_depth_framework.py:469-472)."""
    x0_d = x0[:, 3:4, :, :]
    if x0.shape[1] != 4:
        return x0_d, x0[:, 4:5, :, :]
    if x0.is_cuda and not (torch.is_grad_enabled() and x0.requires_grad):
        # one copy + one compare written as fp32 (the pyramid and the dense net both want the depth plane contiguous: the slice
        # used to be copied twice, and the mask went through a bool tensor)
        x0_d = x0_d.contiguous()
        return x0_d, torch.gt(x0_d, 0, out=torch.empty_like(x0_d))
    return x0_d, (x0_d > 0).float()


class GEDNet(nn.Module):
    """Unguided pyramid + 80-channel encoder-decoder (confidence_depth_framework.py:426-518)."""

    def __init__(self, pos_fn='SoftPlus', pretrained=None, out_channels=1):
        super(GEDNet, self).__init__()
        self.out_channels = out_channels
        self.out_feature_channels = 1 if out_channels == 2 else 36
        self.d_net = unguided_net(pos_fn)
        self.conv1 = nn.Conv2d(5, 80, (3, 3), 2, 1, bias=True)
        self.conv2 = nn.Conv2d(80, 80, (3, 3), 2, 1, bias=True)
        self.conv3 = nn.Conv2d(80, 80, (3, 3), 2, 1, bias=True)
        self.conv4 = nn.Conv2d(80, 80, (3, 3), 2, 1, bias=True)
        self.conv5 = nn.Conv2d(80, 80, (3, 3), 2, 1, bias=True)
        self.conv6 = nn.Conv2d(80 + 80, 64, (3, 3), 1, 1, bias=True)
        self.conv7 = nn.Conv2d(64 + 80, 64, (3, 3), 1, 1, bias=True)
        self.conv8 = nn.Conv2d(64 + 80, 32, (3, 3), 1, 1, bias=True)
        self.conv9 = nn.Conv2d(32 + 80, 32, (3, 3), 1, 1, bias=True)
        self.conv10 = nn.Conv2d(32 + 1, self.out_channels, (3, 3), 1, 1, bias=True)
        for i in range(1, 11):
            m = getattr(self, 'conv%d' % i)
            nn.init.kaiming_normal_(m.weight)
            nn.init.constant_(m.bias, 0.01)

    def forward(self, x0, build_conf_input):
        x0_rgb = x0[:, :3, :, :]
        x0_d, c0 = _c0_of(x0)
        xout_d, cout_d = self.d_net(x0_d, c0)
        if dense.tc_enabled():
            # conv1..conv10 forward + backward on the hand-written tcgen05 kernels (one autograd node, NHWC bf16 / bf16x3)
            from .. import tc_train
            want = bool(build_conf_input) and self.out_channels == 1
            x10, last_layer_input = tc_train.ged_dense(self, xout_d, cout_d, x0_rgb, x0_d.contiguous(), want)
            features = None
            if build_conf_input:
                features = torch.cat((last_layer_input, xout_d, cout_d, x10), 1) if want else torch.sigmoid(x10[:, 1:2, :, :])
            return x10[:, 0:1, :, :], features
        dense._need_cuda(x0)
        up = lambda t, size: F.interpolate(t, size, mode='nearest')      # library baseline path (precision 'fp32' / 'tf32')
        x1 = F.relu(self.conv1(torch.cat((xout_d, x0_rgb, cout_d), 1)))
        x2 = F.relu(self.conv2(x1))
        x3 = F.relu(self.conv3(x2))
        x4 = F.relu(self.conv4(x3))
        x5 = F.relu(self.conv5(x4))
        x6 = F.leaky_relu(self.conv6(torch.cat((up(x5, x4.shape[2:]), x4), 1)), 0.2)
        x7 = F.leaky_relu(self.conv7(torch.cat((up(x6, x3.shape[2:]), x3), 1)), 0.2)
        x8 = F.leaky_relu(self.conv8(torch.cat((up(x7, x2.shape[2:]), x2), 1)), 0.2)
        x9 = F.leaky_relu(self.conv9(torch.cat((up(x8, x1.shape[2:]), x1), 1)), 0.2)
        last_layer_input = torch.cat((up(x9, x0_d.shape[2:]), x0_d), 1)
        x10 = F.leaky_relu(self.conv10(last_layer_input), 0.2)
        features = None
        if build_conf_input:
            if self.out_channels == 1:
                features = torch.cat((last_layer_input, xout_d, cout_d, x10), 1)   # 33 + 1 + 1 + 1
            else:
                features = torch.sigmoid(x10[:, 1:2, :, :])                        # already the confidence
        return x10[:, 0:1, :, :], features


def _conv_stack(first_in, widths):
    layers, cin = [], first_in
    for cout in widths:
        layers += [nn.Conv2d(cin, cout, 3, 1, 1), nn.ReLU()]
        cin = cout
    return nn.Sequential(*layers)


class GMSNet(nn.Module):
    """Unguided pyramid + three full-resolution 3x3 stacks (confidence_depth_framework.py:521-669)."""

    def __init__(self, pos_fn='SoftPlus', out_channels=1):
        super(GMSNet, self).__init__()
        self.out_channels = out_channels
        self.out_feature_channels = 1 if out_channels == 2 else 50
        self.d_net = unguided_net(pos_fn)
        self.d = _conv_stack(1, [16] * 6)
        self.rgb = _conv_stack(4, [64] * 6)
        self.fuse = _conv_stack(80, [64, 64, 64, 32, 32, 32])
        self.last_layer = nn.Conv2d(32, out_channels, 1, 1)
        for seq in (self.d, self.rgb, self.fuse):
            for p in seq:
                if isinstance(p, nn.Conv2d):
                    nn.init.xavier_normal_(p.weight)
                    nn.init.constant_(p.bias, 0.01)
        nn.init.xavier_normal_(self.last_layer.weight)
        nn.init.constant_(self.last_layer.bias, 0.01)

    def _forward_tc(self, x0, build_conf_input):
        """Inference forward on the tensor cores: every 3x3 layer is the tcgen05 implicit-GEMM kernel (tap-reuse variant) on
        zero-bordered NHWC bf16 activations whose channel count is padded to a multiple of 64 (pad channels are zeros and
        meet zero weights); ReLU fused; both torch.cat calls folded into channel-slice writes
        (confidence_depth_framework.py:630-669)."""
        x0_rgb = x0[:, :3, :, :]
        x0_d, c0 = _c0_of(x0)
        xout_d, cout_d = self.d_net(x0_d, c0)
        n, _, h, w = x0.shape
        dev = x0.device
        conv = lambda a, m, cout, relu=1, **kw: dense.conv3x3_tc(a, a.cs, dense.cached_weights(m, cin_pad=a.cs), m.bias, cout, 1,
                                                                  relu, **kw)
        # [rgb stream 64 | depth stream 16 | zeros 48]: both 64-channel slices (borders and pad channels included) are written
        # whole by the convolutions that produce them, so the buffer needs no zero fill
        convs = [m for m in self.d if isinstance(m, nn.Conv2d)]
        if w % 4 == 0:
            fuse_in = self._d_branch_packed(xout_d, convs, n, h, w, dev)
        else:
            fuse_in = dense.new_act(n, h, w, 128, dev, pad=True, zero=False)
            act = dense.to_nhwc_bf16(xout_d, cpad=64, pad=True)                 # channel 0 real
            for i, m in enumerate(convs):
                last = i == len(convs) - 1
                act = conv(act, m, 16, out=fuse_in if last else None, out_off=64 if last else 0, out_zero_to=64 if last else 0)
        # cat(rgb, cout_d) (confidence_depth_framework.py:648) as ONE conversion that writes the whole 64-channel block (pad channels
        # as zeros) with 16-byte stores: no zero fill of the buffer, no 2-byte scattered writes
        act = dense.to_nhwc_bf16(torch.cat((x0_rgb, cout_d), 1), cpad=64, pad=True)
        convs = [m for m in self.rgb if isinstance(m, nn.Conv2d)]
        for i, m in enumerate(convs):
            last = i == len(convs) - 1
            act = conv(act, m, 64, out=fuse_in if last else None, out_off=0, out_zero_to=64 if last else 0)
        act = fuse_in
        for m in [m for m in self.fuse if isinstance(m, nn.Conv2d)]:
            act = conv(act, m, m.out_channels)
        last_in = act
        # last_layer is 1x1: the GEMM mode of the row-segment kernel reads the centre tap only (1/9 of the MMAs of the 3x3 form)
        xout_n = dense.new_act(n, h, w, 16, dev, pad=True, border_only=True, dtype=torch.float32)
        dense.conv1x1_tc(act, act.cs, dense.cached_weights(self.last_layer, embed_1x1=True, cin_pad=act.cs), self.last_layer.bias,
                         self.out_channels, 0, out=xout_n)
        xout = dense.to_nchw_f32(xout_n, self.out_channels)
        features = None
        if build_conf_input:
            if self.out_channels == 1:
                features = torch.cat((dense.to_nchw_f32(last_in, 32), dense.to_nchw_f32(fuse_in, 16, 64), cout_d, xout), 1)
            else:
                features = torch.sigmoid(xout[:, 1:2, :, :])
        return xout[:, 0:1, :, :], features

    def _d_branch_packed(self, xout_d, convs, n, h, w, dev):
        """The depth stream (1 -> 16 -> ... -> 16 channels at full resolution) in PACKED form.  One tcgen05.mma of the row-segment
        kernel costs ~91 cycles whatever its N, and a 16-channel tensor padded to the kernel's 64-channel blocks
        spends 4 K-steps on one real one: the 16 -> 16 layers cost as much as 64 -> 64 ones (476 us each at B=32).  Packing 4
        horizontally adjacent pixels into the channel dimension (channel = c*4 + q for pixel 4g+q) turns the layer into a 3x3
        convolution on the [H, W/4] grid with 64 -> 64 block-structured weights: Wp[(co,q)][(ci,r)][u][dj] = W[co][ci][u][4dj+r-q]
        where that tap exists -- the same kernel on a quarter of the pixels.  Returns the fuse input buffer with channels [64,80)
        filled (its pad channels and border are zeroed once: the buffer is cached per shape)."""
        key = (n, h, w, dense.PRECISION, str(dev))
        cache = self.__dict__.setdefault("_adc_fuse_in", {})
        if key not in cache:
            cache.clear()
            cache[key] = dense.new_act(n, h, w, 128, dev, pad=True, zero=True)
        fuse_in = cache[key]
        xp = xout_d.reshape(n, h, w // 4, 4).permute(0, 3, 1, 2).contiguous()       # [n, q, h, w/4]
        act = dense.to_nhwc_bf16(xp, cpad=64, pad=True)
        for m in convs:
            wp, bp = self._packed_weights(m)
            act = dense.conv3x3_tc(act, 64, wp, bp, 64, 1, 1)
        for src, dst in ((act.hi, fuse_in.hi), (act.lo, fuse_in.lo)):
            if src is not None:
                dst[:, 1:-1, 1:-1, 64:80] = src[:, 1:-1, 1:-1, :].reshape(n, h, w // 4, 16, 4).permute(0, 1, 2, 4, 3).reshape(n, h, w, 16)
        return fuse_in

    @staticmethod
    def _packed_weights(conv):
        """Prepared weights / bias of one depth-stream layer in the packed form above, cached until the Parameter changes."""
        wt = conv.weight
        tag = (wt.data_ptr(), wt._version, conv.bias._version, dense.PRECISION)
        hit = getattr(conv, "_adc_packed", None)
        if hit is not None and hit[0] == tag:
            return hit[1], hit[2]
        W = wt.detach()
        co, ci = W.shape[:2]
        Wp = torch.zeros(co, 4, ci, 4, 3, 3, device=W.device, dtype=W.dtype)
        for q in range(4):
            for r in range(4):
                for dj in (-1, 0, 1):
                    v = 4 * dj + r - q
                    if -1 <= v <= 1:
                        Wp[:, q, :, r, :, dj + 1] = W[:, :, :, v + 1]
        wp = dense.prep_conv3x3_weights(Wp.reshape(co * 4, ci * 4, 3, 3), cin_pad=64)
        bp = conv.bias.detach().repeat_interleave(4).contiguous()
        conv._adc_packed = (tag, wp, bp)
        return wp, bp

    def forward(self, x0, build_conf_input):
        assert x0.shape[1] > 3, "The input is not RGB-D or rgb-dc"
        if dense.tc_inference_enabled():
            return self._forward_tc(x0, build_conf_input)
        if dense.tc_enabled():
            # training: the three stacks + last_layer forward and backward on the tcgen05 kernels, one autograd node (tc_gms.py)
            from .. import tc_gms
            x0_d, c0 = _c0_of(x0)
            xout_d, cout_d = self.d_net(x0_d, c0)
            want = bool(build_conf_input) and self.out_channels == 1
            xout, lli, d16 = tc_gms.gms_dense(self, xout_d, cout_d, x0[:, :3, :, :], want)
            features = None
            if build_conf_input:
                features = torch.cat((lli, d16, cout_d, xout), 1) if want else torch.sigmoid(xout[:, 1:2, :, :])
            return xout[:, 0:1, :, :], features
        x0_rgb = x0[:, :3, :, :]
        x0_d, c0 = _c0_of(x0)
        xout_d, cout_d = self.d_net(x0_d, c0)
        xout_d = self.d(xout_d)
        xout_rgb = self.rgb(torch.cat((x0_rgb, cout_d), 1))
        last_layer_input = self.fuse(torch.cat((xout_rgb, xout_d), 1))
        xout = self.last_layer(last_layer_input)
        features = None
        if build_conf_input:
            if self.out_channels == 1:
                features = torch.cat((last_layer_input, xout_d, cout_d, xout), 1)   # 32 + 16 + 1 + 1
            else:
                features = torch.sigmoid(xout[:, 1:2, :, :])
        return xout[:, 0:1, :, :], features


################################################
# criteria (confidence_depth_framework.py:808-914)
################################################

class _MaskedLoss(nn.Module):
    """Mean of a per-pixel error over ``target > 0`` (confidence_depth_framework.py:829-914): one reduction kernel forward, one
    elementwise kernel backward (csrc/loss_metrics.cu), no boolean-index compaction.  CUDA tensors only (no CPU fallback).
    ``lazy`` = keep ``self.loss`` and the >100-valid-pixels check on the device, so a training step has no host synchronisation:
    a batch with <= 100 valid pixels then yields a NaN loss, which ``optim.FlatAdam``'s guard turns into a skipped update (the
    reference raises an AssertionError on the host; a device-side assert would kill the CUDA context).  The default mirrors the
    reference (python floats, assert).  ``num_valid`` keeps the valid-pixel count of the last call on the device."""
    lazy = False
    kind = None

    def __init__(self):
        super(_MaskedLoss, self).__init__()
        self.loss = -1

    def forward(self, depth_input, depth_prediction, depth_target, epoch=None):
        assert depth_prediction.dim() == depth_target.dim(), "inconsistent dimensions"
        from .. import ops
        final_loss, num_valids = ops.masked_loss(depth_prediction, depth_target.detach(), self.kind)
        self.num_valid = num_valids
        if self.lazy:
            nan = self.__dict__.get('_nan')
            if nan is None or nan.device != final_loss.device:
                nan = self.__dict__['_nan'] = final_loss.new_full((), float('nan'))
            final_loss = torch.where(num_valids > 100, final_loss, nan)
        else:
            assert (num_valids > 100), 'training image has less than 100 valid pixels'
        self.loss = [final_loss.detach() if self.lazy else final_loss.item(), 0, 0]
        return final_loss


class DualLoss(nn.Module):
    def __init__(self, net_a, net_b, weight_b):
        super(DualLoss, self).__init__()
        self.net_a, self.net_b, self.weight_b = net_a, net_b, weight_b

    def forward(self, input, depth_a, depth_b, target_depth, epoch=None):
        error_a = self.net_a(input, depth_a, target_depth, epoch)
        self.loss = self.net_a.loss
        if self.weight_b > 0:
            error_b = self.net_b(input, depth_b, target_depth, epoch)
            self.loss[1] = self.net_b.loss[0]
            return (self.weight_b * error_b) + ((1.0 - self.weight_b) * error_a)
        return error_a


class MaskedMSELoss(_MaskedLoss):
    kind = 'l2'


class InvertedMaskedL1Loss(_MaskedLoss):
    kind = 'il1'


class MaskedAbsRelLoss(_MaskedLoss):
    kind = 'absrel'


class MaskedL1Loss(_MaskedLoss):
    kind = 'l1'

"""Normalized convolution layer -- drop-in for the reference's ``model_zoo/nconv.py`` on the B200 kernels.

Same constructor signature, attribute names and ``state_dict`` keys as ``NConv2d`` (model_zoo/nconv.py:19-96) and
the same ``EnforcePos`` pre-hook semantics (model_zoo/nconv.py:100-136): in training mode every *call* replaces
``weight.data`` by ``pos_fn(weight)`` and gradients do not chain through ``pos_fn``.  The forward/backward math
runs in the fused sm_100a kernels behind ``adc_nconv2d_fwd`` / ``adc_nconv2d_bwd``.
"""
import math

import torch
import torch.nn.functional as F
from torch.nn.modules.conv import _ConvNd
from torch.nn.modules.utils import _pair

from .. import ops


def _poisson_pmf(k, mu):
    return [math.exp(-mu) * mu ** i / math.factorial(i) for i in range(k)]


class NConv2d(_ConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, pos_fn='softplus', init_method='k', stride=1, padding=0,
                 dilation=1, groups=1, bias=True):
        # _ConvNd.__init__ draws the same default-init random numbers as in the reference, so a seeded
        # construction yields the same parameters (model_zoo/nconv.py:23).
        super().__init__(in_channels, out_channels, _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation),
                         False, _pair(0), groups, bias, 'zeros')
        self.eps = 1e-20
        self.pos_fn = pos_fn
        self.init_method = init_method
        self.init_parameters()
        if self.pos_fn is not None:
            EnforcePos.apply(self, 'weight', pos_fn)

    def _check_supported(self):
        k = self.kernel_size
        ok = (k[0] == k[1] and k[0] % 2 == 1 and self.stride == (1, 1) and self.dilation == (1, 1) and self.groups == 1
              and self.padding == (k[0] // 2, k[0] // 2))
        if not ok:
            raise NotImplementedError(
                "NConv2d on the B200 path supports square odd kernels with stride 1, dilation 1, groups 1 and "
                "padding k//2 (the only configuration the reference nets use, model_zoo/nconv_sd.py:21-32); got "
                "kernel %s stride %s padding %s dilation %s groups %d" % (k, self.stride, self.padding, self.dilation, self.groups))

    def forward(self, data, conf):
        self._check_supported()
        return ops.NConv2dFunction.apply(data, conf, self.weight, self.bias, self)

    def init_parameters(self):
        """model_zoo/nconv.py:70-96: 'x' Xavier, 'k' Kaiming, 'n' normal(2, sqrt(2/n)), 'p' Poisson bump + U[0,1)."""
        if self.init_method == 'x':
            torch.nn.init.xavier_uniform_(self.weight)
        elif self.init_method == 'k':
            torch.nn.init.kaiming_uniform_(self.weight)
        elif self.init_method == 'n':
            n = self.kernel_size[0] * self.kernel_size[1] * self.out_channels
            self.weight.data.normal_(2, math.sqrt(2. / n))
        elif self.init_method == 'p':
            k = self.kernel_size[0]
            pmf = torch.tensor(_poisson_pmf(k, k / 2), dtype=torch.float64)
            bump = torch.outer(pmf, pmf).to(self.weight.dtype)      # = convolve2d(col, row, 'full') of the pmf
            w = bump.expand(self.out_channels, self.in_channels, k, k)
            self.weight.data = w + torch.rand(w.shape)
        else:
            raise ValueError('Undefined Initialization method!')
        self.bias = torch.nn.Parameter(torch.zeros(self.out_channels) + 0.01)


class EnforcePos(object):
    """Forward-pre-hook keeping a parameter positive by overwriting its data (no re-parametrisation)."""

    def __init__(self, pos_fn, name):
        self.name = name
        self.pos_fn = pos_fn

    @staticmethod
    def apply(module, name, pos_fn):
        fn = EnforcePos(pos_fn, name)
        module.register_forward_pre_hook(fn)
        return fn

    def __call__(self, module, inputs):
        if module.training:
            # the reference rebinds ``weight.data``; the values are copied INTO the parameter's storage here so that flat
            # parameter buckets (optim.FlatAdam, parallel.DataParallelTrainer) and captured CUDA graphs keep their addresses
            weight = getattr(module, self.name)
            weight.data.copy_(self._pos(weight).data)

    def _pos(self, p):
        pos_fn = self.pos_fn.lower()
        if pos_fn == 'softplus':
            return ops.softplus_pow(p.detach(), 1)          # CUDA kernel; a CPU tensor raises (no CPU path)
        if pos_fn == 'softmax':
            sz = p.size()
            return F.softmax(p.view(sz[0], sz[1], -1), -1).view(sz)
        if pos_fn == 'exp':
            return torch.exp(p)
        if pos_fn == 'sigmoid':
            return torch.sigmoid(p)
        print('Undefined positive function!')
        return

"""Unguided NConv pyramid -- drop-in for the reference's ``model_zoo/nconv_sd.py`` (class ``CNN``).

Seven ``NConv2d`` sub-modules with the reference's names (``nconv1`` .. ``nconv7``, model_zoo/nconv_sd.py:21-32), so
``state_dict`` keys match; the forward (11 NConv2d calls, 3 confidence-argmax poolings, 3 nearest-upsample+concat,
model_zoo/nconv_sd.py:35-101) is ONE fused call into the CUDA library (``ops.PyramidFunction``).  The per-call softplus
of ``EnforcePos`` (nconv2 is called 4x, nconv3 2x per forward) is reproduced on the device.
"""
import torch.nn as nn

from .. import ops
from .nconv import NConv2d


class CNN(nn.Module):

    def __init__(self, pos_fn='SoftPlus', num_channels=2):
        super().__init__()
        self.pos_fn = pos_fn
        if num_channels != 2:
            raise NotImplementedError("the B200 pyramid kernels are instantiated for num_channels=2 (the reference default)")
        if pos_fn is None or pos_fn.lower() != 'softplus':
            raise NotImplementedError("the fused pyramid implements the SoftPlus positivity constraint every reference net uses")
        self.nconv1 = NConv2d(1, num_channels, (5, 5), pos_fn, 'p', padding=2)
        self.nconv2 = NConv2d(num_channels, num_channels, (5, 5), pos_fn, 'p', padding=2)
        self.nconv3 = NConv2d(num_channels, num_channels, (5, 5), pos_fn, 'p', padding=2)
        self.nconv4 = NConv2d(2 * num_channels, num_channels, (3, 3), pos_fn, 'p', padding=1)
        self.nconv5 = NConv2d(2 * num_channels, num_channels, (3, 3), pos_fn, 'p', padding=1)
        self.nconv6 = NConv2d(2 * num_channels, num_channels, (3, 3), pos_fn, 'p', padding=1)
        self.nconv7 = NConv2d(num_channels, 1, (1, 1), pos_fn, 'k')
        self._adc_slot = ops.new_weight_slot()
        self._adc_last_idx = None

    def forward(self, x0, c0):
        params = []
        for name in ops.PYR_MODULES:
            m = getattr(self, name)
            params += [m.weight, m.bias]
        return ops.PyramidFunction.apply(x0, c0, self, self.training, *params)

    def last_pool_indices(self):
        """The three int64 argmax index maps of the most recent forward (what max_pool2d(return_indices) returned
        in the reference, model_zoo/nconv_sd.py:44,57,70)."""
        idx, n, h, w = self._adc_last_idx
        return ops.pyramid_idx_views(idx, n, h, w)

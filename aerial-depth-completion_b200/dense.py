"""Dense (wide-channel) layers around the NConv pyramid: one dispatch point per op.

BACKEND 'cudnn' routes to the ATen/cuDNN library call (what the reference itself executes on a GPU); it is the
baseline these layers are measured against and is replaced op by op by the hand-written sm_100a kernels
('native').  Nothing here runs on the CPU: inputs must be CUDA tensors (no CPU fallback).
"""
import torch
import torch.nn.functional as F

BACKEND = "cudnn"
PRECISION = "fp32"


def set_precision(mode):
    """'fp32' (strict: parity 1e-4 with the reference, TF32 off in library calls) or 'tf32' (parity 1e-2)."""
    global PRECISION
    if mode not in ("fp32", "tf32"):
        raise ValueError("precision must be 'fp32' or 'tf32'")
    PRECISION = mode
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"


set_precision("fp32")


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("dense layer input must be a CUDA tensor: the B200 path has no CPU fallback")


def conv2d(x, weight, bias=None, stride=1, padding=0):
    _need_cuda(x)
    return F.conv2d(x, weight, bias, stride, padding)


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, output_padding=0):
    _need_cuda(x)
    return F.conv_transpose2d(x, weight, bias, stride, padding, output_padding)


def upsample_nearest(x, size):
    _need_cuda(x)
    return F.interpolate(x, size, mode="nearest")


class Conv2d(torch.nn.Conv2d):
    """nn.Conv2d with identical parameters / state_dict keys whose forward goes through the dispatch above."""

    def forward(self, x):
        return conv2d(x, self.weight, self.bias, self.stride, self.padding)


class ConvTranspose2d(torch.nn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        return conv_transpose2d(x, self.weight, self.bias, self.stride, self.padding, self.output_padding)

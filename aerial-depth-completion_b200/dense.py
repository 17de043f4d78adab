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


# ---- tensor-core path (tcgen05 implicit GEMM, NHWC bf16) ---------------------------------------------------------------
def _r16(c):
    return (c + 15) // 16 * 16


def _stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_nhwc_bf16(x, cpad=None, dst=None, dst_off=0):
    """NCHW fp32 (any batch stride, contiguous planes) -> NHWC bf16 with the channel count padded to a multiple of 16;
    optionally into channels [dst_off, dst_off+cpad) of an existing NHWC tensor (folds torch.cat)."""
    from ._lib import lib, check
    _need_cuda(x)
    n, c, h, w = x.shape
    if x.stride(3) != 1 or x.stride(2) != w or (c > 1 and x.stride(1) != h * w):
        x = x.contiguous()
    cpad = _r16(c) if cpad is None else cpad
    if dst is None:
        dst = torch.empty((n, h, w, cpad), device=x.device, dtype=torch.bfloat16)
    check(lib.adc_nchw_to_nhwc_bf16(x.data_ptr(), dst.data_ptr(), n, c, cpad, h, w, x.stride(0) if n > 1 else c * h * w,
                                    dst.shape[3], dst_off, _stream()), "adc_nchw_to_nhwc_bf16")
    return dst


def to_nchw_f32(x_nhwc, c, off=0):
    from ._lib import lib, check
    n, h, w, cs = x_nhwc.shape
    out = torch.empty((n, c, h, w), device=x_nhwc.device, dtype=torch.float32)
    check(lib.adc_nhwc_to_nchw_f32(x_nhwc.data_ptr(), int(x_nhwc.dtype == torch.float32), out.data_ptr(), n, c, h, w, cs, off,
                                   _stream()), "adc_nhwc_to_nchw_f32")
    return out


def prep_conv3x3_weights(weight, flip_transpose=False):
    """fp32 [Cout,Cin,3,3] -> bf16 [9][round16(Cout)][round16(Cin)] as the tensor-core kernel reads it."""
    from ._lib import lib, check
    co, ci = weight.shape[:2]
    oc, ic = (ci, co) if flip_transpose else (co, ci)
    out = torch.empty((9, _r16(oc), _r16(ic)), device=weight.device, dtype=torch.bfloat16)
    check(lib.adc_conv3x3_prep_weights(weight.detach().contiguous().data_ptr(), out.data_ptr(), co, ci, int(flip_transpose),
                                       _stream()), "adc_conv3x3_prep_weights")
    return out


def conv3x3_tc(x_nhwc, cin, w_prep, bias, cout, stride=1, act=0, x_off=0, out=None, out_off=0, out_f32=False):
    """3x3 / pad 1 convolution on the tensor cores.  x_nhwc: [N,H,W,Cs] bf16, channels [x_off, x_off+cin) are read
    (cin a multiple of 16, zero padded); result written to channels [out_off, out_off+cout) of ``out`` (NHWC)."""
    from ._lib import lib, check
    n, h, w, cs = x_nhwc.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, _r16(cout)), device=x_nhwc.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if _r16(cout) != cout:
            out.zero_()
    b = bias.detach().contiguous().data_ptr() if bias is not None else None
    check(lib.adc_conv3x3_fprop_tc(x_nhwc.data_ptr(), cs, x_off, w_prep.data_ptr(), b, out.data_ptr(), int(out.dtype == torch.float32),
                                   out.shape[3], out_off, n, h, w, cin, cout, stride, act, _stream()), "adc_conv3x3_fprop_tc")
    return out

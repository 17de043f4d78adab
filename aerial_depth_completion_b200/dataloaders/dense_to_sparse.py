"""Sparse-depth sampling -- drop-in for ``UniformSampling`` of the reference's ``dataloaders/dense_to_sparse.py``.

Bit-exact with the reference (dense_to_sparse.py:29-48): the Bernoulli mask is drawn from numpy's legacy global
MT19937 stream (``np.random.uniform``), so ``np.random.seed(s)`` reproduces the reference's masks.  ``SimulatedStereo``
(never selected by the reference's main.py) is outside the hot path and not provided."""
import numpy as np


class DenseToSparse:
    def __init__(self):
        pass

    def dense_to_sparse(self, rgb, depth):
        pass

    def __repr__(self):
        pass


class UniformSampling(DenseToSparse):
    name = "uar"

    def __init__(self, num_samples, max_depth=np.inf):
        DenseToSparse.__init__(self)
        self.num_samples = num_samples
        self.max_depth = max_depth

    def __repr__(self):
        return "%s{ns=%d,md=%f}" % (self.name, self.num_samples, self.max_depth)

    def dense_to_sparse(self, rgb, depth):
        mask_keep = depth > 0
        if self.max_depth is not np.inf:
            mask_keep = np.bitwise_and(mask_keep, depth <= self.max_depth)
        n_keep = np.count_nonzero(mask_keep)
        if n_keep == 0 or self.num_samples == 0:
            return mask_keep
        if self.num_samples == -1:
            prob = np.random.randint(5, 100, 1) / 100.0
        elif self.num_samples == -2:
            prob = np.random.randint(5, 100, 1) / 1000.0
        else:
            prob = float(self.num_samples) / n_keep
        return np.bitwise_and(mask_keep, np.random.uniform(0, 1, depth.shape) < prob)


def uniform_sampling_cuda(depth, seeds, num_samples, max_depth=np.inf):
    """``UniformSampling(num_samples, max_depth).dense_to_sparse`` for a batch of frames on the device (adc_uniform_sampling):
    frame n gets exactly the mask the reference computes after ``np.random.seed(seeds[n])``.  depth: CUDA float32 [N, H, W] or
    [N, 1, H, W]; seeds: N integers in [0, 2**32); returns a CUDA bool tensor of depth's shape.  The random-probability modes
    (num_samples -1 / -2) stay on the host path."""
    import ctypes
    import torch
    from .._lib import lib, check
    if not torch.is_tensor(depth) or not depth.is_cuda or depth.dtype != torch.float32:
        raise RuntimeError("uniform_sampling_cuda needs a CUDA float32 depth tensor: the B200 path has no CPU fallback")
    if num_samples < 0:
        raise ValueError("uniform_sampling_cuda: num_samples must be >= 0")
    d = depth.contiguous()
    n = d.shape[0]
    hw = d.numel() // n
    seeds = [int(s) for s in seeds]
    if len(seeds) != n or any(s < 0 or s > 0xFFFFFFFF for s in seeds):
        raise ValueError("uniform_sampling_cuda: one seed in [0, 2**32) per frame")
    sd = torch.from_numpy(np.array(seeds, dtype=np.uint32).view(np.int32)).to(d.device)      # same 32 bits; the kernel reads them unsigned
    mask = torch.empty(d.shape, device=d.device, dtype=torch.uint8)
    use_max = max_depth is not np.inf and np.isfinite(max_depth)
    check(lib.adc_uniform_sampling(d.data_ptr(), sd.data_ptr(), int(num_samples), float(max_depth) if use_max else 0.0, int(use_max),
                                   mask.data_ptr(), n, hw, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "adc_uniform_sampling")
    return mask.bool()


def sample_assemble_cuda(rgbgt, seeds, num_samples, max_depth=np.inf, out=None):
    """The loader's sparse-depth synthesis and modality assembly on the device (dense_to_sparse.py:29-48 +
    dataloader_ext.py:544-566): ``rgbgt`` CUDA float32 [N, 4, H, W] = (r, g, b, ground-truth depth); returns the network input
    [N, 5, H, W] = (r, g, b, d, w) with d = gt * mask, w = mask, mask = the reference's ``UniformSampling(num_samples, max_depth)`` mask of
    frame n after ``np.random.seed(seeds[n])`` (bit-exact; adc_sample_assemble).  ``seeds``: N integers in [0, 2**32) or a CUDA int32
    tensor holding the same 32 bits.  The training target is the view ``rgbgt[:, 3:4]``."""
    import ctypes
    import torch
    from .._lib import lib, check
    if not torch.is_tensor(rgbgt) or not rgbgt.is_cuda or rgbgt.dtype != torch.float32 or rgbgt.dim() != 4 or rgbgt.shape[1] != 4:
        raise RuntimeError("sample_assemble_cuda needs a CUDA float32 [N, 4, H, W] tensor: the B200 path has no CPU fallback")
    if num_samples < 0:
        raise ValueError("sample_assemble_cuda: num_samples must be >= 0")
    x = rgbgt.contiguous()
    n, _, h, w = x.shape
    if torch.is_tensor(seeds):
        sd = seeds
        if not sd.is_cuda or sd.dtype != torch.int32 or sd.numel() != n:
            raise ValueError("sample_assemble_cuda: seeds tensor must be CUDA int32 with one entry per frame")
    else:
        seeds = [int(v) for v in seeds]
        if len(seeds) != n or any(v < 0 or v > 0xFFFFFFFF for v in seeds):
            raise ValueError("sample_assemble_cuda: one seed in [0, 2**32) per frame")
        sd = torch.from_numpy(np.array(seeds, dtype=np.uint32).view(np.int32)).to(x.device)
    if out is None:
        out = torch.empty(n, 5, h, w, device=x.device, dtype=torch.float32)
    elif out.shape != (n, 5, h, w) or not out.is_contiguous() or out.dtype != torch.float32 or out.device != x.device:
        raise ValueError("sample_assemble_cuda: out must be a contiguous CUDA float32 [N, 5, H, W] tensor")
    use_max = max_depth is not np.inf and np.isfinite(max_depth)
    check(lib.adc_sample_assemble(x.data_ptr(), sd.data_ptr(), int(num_samples), float(max_depth) if use_max else 0.0, int(use_max),
                                  out.data_ptr(), n, h * w, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "adc_sample_assemble")
    return out

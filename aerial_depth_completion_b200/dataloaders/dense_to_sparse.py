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

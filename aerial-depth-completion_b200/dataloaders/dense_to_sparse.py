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

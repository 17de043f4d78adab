"""GPU: the device-side UniformSampling (csrc/sampler.cu, MT19937 in shared memory) must give EXACTLY the mask the reference's
numpy path gives after np.random.seed(seed) (dataloaders/dense_to_sparse.py:29-48)."""
import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu


def _host_masks(depth, seeds, num_samples, max_depth=np.inf):
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import UniformSampling
    s = UniformSampling(num_samples, max_depth)
    out = []
    for d, seed in zip(depth, seeds):
        np.random.seed(seed)
        out.append(s.dense_to_sparse(None, d))
    return np.stack(out)


@gpu
@pytest.mark.parametrize("shape,num_samples", [((3, 228, 304), 500), ((2, 57, 76), 100), ((1, 480, 752), 500), ((2, 61, 47), 20)])
def test_device_sampler_is_bit_exact_with_numpy(shape, num_samples):
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import uniform_sampling_cuda
    g = np.random.RandomState(5)
    depth = (g.rand(*shape) * 10).astype(np.float32)
    depth[g.rand(*shape) < 0.2] = 0.0                      # holes
    seeds = [0, 7, 4294967295][: shape[0]]
    ref = _host_masks(depth, seeds, num_samples)
    got = uniform_sampling_cuda(torch.from_numpy(depth).cuda(), seeds, num_samples)
    assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), ref)
    # max_depth variant and the [N, 1, H, W] layout
    ref2 = _host_masks(depth, seeds, num_samples, 6.5)
    got2 = uniform_sampling_cuda(torch.from_numpy(depth).cuda().unsqueeze(1), seeds, num_samples, 6.5)
    assert np.array_equal(got2[:, 0].cpu().numpy(), ref2)


@gpu
def test_device_sampler_kat1_and_degenerate_cases():
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import uniform_sampling_cuda
    depth = (np.arange(228 * 304).reshape(228, 304) % 17).astype(np.float32)      # KAT-1 of SURVEY.md 8c
    got = uniform_sampling_cuda(torch.from_numpy(depth[None]).cuda().repeat(2, 1, 1), [0, 7], 500).cpu().numpy()
    idx0, idx7 = np.flatnonzero(got[0]), np.flatnonzero(got[1])
    assert got[0].sum() == 525 and list(idx0[:5]) == [99, 540, 742, 757, 821] and int(idx0.sum()) == 17701305
    assert got[1].sum() == 497 and list(idx7[:5]) == [55, 108, 302, 373, 558] and int(idx7.sum()) == 16306862
    zeros = torch.zeros(1, 20, 30, device="cuda")
    assert not bool(uniform_sampling_cuda(zeros, [3], 500).any())                  # no valid pixel: mask = keep = all False
    some = torch.from_numpy(depth[None]).cuda()
    assert torch.equal(uniform_sampling_cuda(some, [3], 0), some > 0)              # num_samples == 0: mask = keep
    with pytest.raises(RuntimeError):
        uniform_sampling_cuda(torch.zeros(1, 4, 4), [1], 10)

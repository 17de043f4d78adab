"""Oracle (TEST INFRASTRUCTURE): deterministic synthetic inputs shared by tests, goldens and bench.

SURVEY.md 8(c)/(d) fixtures: formula-defined weights / sparse depth for the
known-answer tests, and the seeded NYU-shape frames for the timed configs
(rgb ~ U[0,1), gt ~ U[1,10), sparse mask = dense_to_sparse.py:29-48 with
np.random.seed(seed) per frame, d = gt*mask, w = mask).
"""
import numpy as np
import torch

from .sampling_oracle import uniform_mask


def kat_weight(co, ci, k, dtype=torch.float32):
    """W[o,i,kh,kw] = 0.05 + 0.01*((7o+5i+3kh+kw) mod 11)."""
    o, i, u, v = torch.meshgrid(torch.arange(co), torch.arange(ci), torch.arange(k), torch.arange(k), indexing="ij")
    return (0.05 + 0.01 * ((7 * o + 5 * i + 3 * u + v) % 11).to(torch.float64)).to(dtype)


def kat_sparse(n, h, w, dtype=torch.float32):
    """c[n,0,h,w] = 1 if (31h+17w+5n) mod 97 == 0 else 0 ;  x = (1 + 0.5*((h*W+w+n) mod 13)) * c."""
    nn, hh, ww = torch.meshgrid(torch.arange(n), torch.arange(h), torch.arange(w), indexing="ij")
    c = (((31 * hh + 17 * ww + 5 * nn) % 97) == 0).to(torch.float64)
    x = (1.0 + 0.5 * ((hh * w + ww + nn) % 13).to(torch.float64)) * c
    return x.unsqueeze(1).to(dtype), c.unsqueeze(1).to(dtype)


def kat_pyramid_state(dtype=torch.float32, prefix=""):
    """Formula weights for the 7 NConv modules, all biases 0.01 (positive, so eval mode is well-posed)."""
    shapes = dict(nconv1=(2, 1, 5), nconv2=(2, 2, 5), nconv3=(2, 2, 5), nconv4=(2, 4, 3), nconv5=(2, 4, 3),
                  nconv6=(2, 4, 3), nconv7=(1, 2, 1))
    S = {}
    for m, (co, ci, k) in shapes.items():
        S[prefix + m + ".weight"] = kat_weight(co, ci, k, dtype)
        S[prefix + m + ".bias"] = torch.full((co,), 0.01, dtype=dtype)
    return S


def synth_frames(batch, h, w, seed=1234, num_samples=500, dtype=torch.float32):
    """Returns (input[B,5,H,W] = rgb|d|w, target[B,1,H,W]).  SURVEY.md 8(d) 'Synthetic inputs'."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(batch, 3, h, w, generator=g, dtype=torch.float32)
    gt = torch.rand(batch, 1, h, w, generator=g, dtype=torch.float32) * 9.0 + 1.0
    masks = []
    for b in range(batch):
        np.random.seed(seed + b)
        masks.append(uniform_mask(gt[b, 0].numpy(), num_samples))
    mask = torch.from_numpy(np.stack(masks)).unsqueeze(1).to(torch.float32)
    inp = torch.cat((rgb, gt * mask, mask), 1)
    return inp.to(dtype), gt.to(dtype)


def seeded_state(template, seed=0, dtype=torch.float32):
    """Deterministic, well-conditioned values for every entry of a state_dict-like ``template``
    (name -> tensor or shape).  Used instead of storing MBs of random weights in fixtures: the golden
    script loads these into the *reference* modules, the tests load them into the oracle / product.
    NConv weights are positive (eval mode is then well-posed), conv weights are fan-in scaled."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(template.keys()):
        t = template[name]
        shape = tuple(t.shape) if hasattr(t, "shape") else tuple(t)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        r = torch.rand(shape, generator=g, dtype=torch.float64)
        n = torch.randn(shape, generator=g, dtype=torch.float64)
        if "nconv" in name and leaf == "weight":
            v = 0.1 + 0.9 * r
        elif "nconv" in name and leaf == "bias":
            v = 0.01 + 0.01 * r
        elif leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            v = n * (2.0 / fan_in) ** 0.5
        elif leaf == "weight":            # batch-norm scale
            v = 1.0 + 0.1 * n
        elif leaf == "running_var":
            v = 1.0 + 0.2 * r
        elif leaf == "running_mean":
            v = 0.1 * n
        else:                             # conv / bn bias
            v = 0.05 * n
        out[name] = v.to(dtype)
    return out

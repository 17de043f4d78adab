"""Helpers shared by the oracle and product parity tests: load fixtures, rebuild the seeded cases."""
import json
import os

import numpy as np
import torch

from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def gold(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return _cache[name]


def relmax(a, b):
    """max|a-b| / max|b|  (the parity norm of SURVEY.md 8c; never element-wise rtol)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def digest(t):
    f = t.detach().double().flatten().cpu()
    return np.concatenate([[float(f.sum()), float(f.abs().sum())], f[:8].numpy(), np.zeros(max(0, 8 - f.numel()))])


def g_like(t, seed):
    return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64).to(t.dtype)


NET_TAGS = ("ged_only", "ged_only_eval", "ged_dc1", "ged_full", "ged_ln0", "ged_dc0", "ged_dcln", "ged_odd",
            "gms_only", "gms_full", "gud_only", "gud_only_eval", "gud_ln0", "ud_only", "s2d_only", "s2d_cf")


def net_case(tag, dtype=torch.float32):
    """Returns (cfg, state, inp, tgt) for a nets.npz case: same seeded weights / frames the reference was run on."""
    G = gold("nets")
    cfg = json.loads(str(G[tag + "_cfg"]))
    template = {k: tuple(s) for k, s in json.loads(str(G[tag + "_template"]))}
    state = synth.seeded_state(template, cfg["seed"], dtype)
    h, w = cfg["hw"]
    inp, tgt = synth.synth_frames(cfg["batch"], h, w, seed=100 + cfg["seed"], num_samples=max(60, h * w // 140), dtype=dtype)
    return cfg, state, inp, tgt

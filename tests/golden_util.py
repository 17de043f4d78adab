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


NET_TAGS = ("ged_nyu", "ged_only", "ged_only_eval", "ged_dc1", "ged_full", "ged_ln0", "ged_dc0", "ged_dcln", "ged_odd",
            "gms_only", "gms_full", "gud_only", "gud_only_eval", "gud_ln0", "gud_mid", "ud_only", "s2d_only", "s2d_cf", "s2d_nyu", "s2d_eval", "s2d_odd", "s2d_dc1ln0", "gud_dc1ln0", "ged_s2dln")


def net_case(tag, dtype=torch.float32):
    """Returns (cfg, state, inp, tgt) for a nets.npz case: same seeded weights / frames the reference was run on."""
    G = gold("nets")
    cfg = json.loads(str(G[tag + "_cfg"]))
    template = {k: tuple(s) for k, s in json.loads(str(G[tag + "_template"]))}
    state = synth.seeded_state(template, cfg["seed"], dtype)
    h, w = cfg["hw"]
    inp, tgt = synth.synth_frames(cfg["batch"], h, w, seed=100 + cfg["seed"], num_samples=max(60, h * w // 140), dtype=dtype)
    return cfg, state, inp, tgt


def run_oracle_case(tag, dtype=torch.float32):
    """The nets.npz case ``tag`` through the CPU oracle in ``dtype`` (forward, loss, backward): returns (cfg, S, inp, tgt, d1, c1, d2,
    loss, parts) with the gradients in ``S[k].grad``.  In float64 this is the gradient reference of the k-times-floor rule
    (SURVEY.md 8c): the oracle equals the reference to 1e-11 in fp64 (tests/test_oracle_golden.py)."""
    from oracle import loss_metrics_oracle as lm
    from oracle import nets_oracle as nets
    cfg, S, inp, tgt = net_case(tag, dtype)
    for k in S:
        if S[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            S[k].requires_grad_(True)
    frozen = [p for p, f in (("dc_model.", "dc0"), ("conf_model.", "cf0"), ("loss_dc_model.", "ln0")) if f in cfg["mode"]]
    for k in S:
        if any(k.startswith(p) for p in frozen) and S[k].requires_grad:
            S[k].requires_grad_(False)
    d1, c1, d2 = nets.framework_forward(S, inp, cfg["mode"], cfg["dc_arch"], len(cfg["input_type"]), cfg["conf_arch"],
                                        cfg["ln_arch"], cfg["train"])
    loss = parts = None
    if cfg["train"]:
        if d2 is not None:
            loss, parts = lm.dual_loss(cfg["criterion"], d1, d2, tgt, 0.5 if "dc1" in cfg["mode"] else 1.0)
        else:
            loss = lm.masked_loss(cfg["criterion"], d1, tgt)
            parts = [float(loss), 0, 0]
        loss.backward()
    return cfg, S, inp, tgt, d1, c1, d2, loss, parts


def ref_grad_errors(tag):
    """{parameter name: max|g32 - g64| / max|g64|} of the REFERENCE's own fp32 gradients against the fp64 oracle, measured when the
    fixtures were generated (tests/golden/make_golden.py): the noise floor the k-times-floor rule multiplies."""
    G = gold("nets")
    return json.loads(str(G[tag + "_gerr"])) if (tag + "_gerr") in G.files else {}


# Gradient parity rule, per parameter tensor over the WHOLE tensor:  max|g - g64| / max|g64|  <=  max(GRAD_K * GRAD_AMP[mode]
# * floor_t, GRAD_ABS[mode]) where floor_t is the reference's own fp32-vs-fp64 error for that tensor (ref_grad_errors) and GRAD_ABS is the
# absolute term of the arithmetic mode:
#   fp32    2e-5  kernels that compute in fp32 (NConv pyramid, losses; dense layers on the fp32 library path): the strict rule.
#   bf16x3  5e-3  hi/lo split bf16 operands carry 16-17 mantissa bits, so the FORWARD activations are exact to ~2e-5 (bar 1e-4) instead
#                 of fp32's ~1e-7.  ReLU / LeakyReLU masks of near-zero activations then flip ~100x more often than in an fp32 run, and
#                 every flip changes one element of a gradient by O(1): measured per-tensor errors 5e-6 (conv9/conv10) .. 2.6e-3 (conv6
#                 at 15x19 pixels, NConv weights), profiles/r2_grad_noise_*.txt -- between the fp32 bar and north_star's 1e-2 bf16 bar.
#   bf16    2e-1  plain bf16 operands (the 1e-2 output mode).
#
# Tensors whose gradient is a heavily cancelling sum (the NConv weights: the reference's own floor_t is 1e-4..2e-3, i.e. a condition
# number floor_t / 2^-24 of 1e3..3e4) amplify the arithmetic's unit round-off, not its absolute error: for them the floor is scaled by the
# ratio of the unit round-offs, GRAD_AMP = 2^-17 / 2^-24 / GRAD_K = 32 for bf16x3 (measured err / floor_t: 60..90 for nconv1..7 of
# gms_full, profiles/r2_grad_noise_gms_full.txt).
GRAD_K = 4.0
GRAD_AMP = {"fp32": 1.0, "bf16x3": 32.0, "bf16": 32.0 * 256.0, "bn_tiny": 32.0}
#   bn_tiny 3e-1  BatchNorm nets at the tiny fixture sizes (32x48 frames: the bottleneck normalises over 2 x 2 x 3 = 12 values), where any
#                 rounding difference in a convolution is amplified by 1/sigma of a near-degenerate batch.
GRAD_ABS = {"fp32": 2e-5, "bf16x3": 5e-3, "bf16": 2e-1, "bn_tiny": 3e-1}


def grad_rule_violations(named_grads, oracle_state, floor, mode):
    """[(name, err, floor)] of the tensors that break the rule above; ``named_grads``: iterable of (name, gradient tensor)."""
    bad = []
    for k, g in named_grads:
        err = relmax(g.detach().cpu(), oracle_state[k].grad)
        if not err <= max(GRAD_K * GRAD_AMP[mode] * floor.get(k, 0.0), GRAD_ABS[mode]):
            bad.append((k, "%.2e" % err, "%.2e" % floor.get(k, 0.0)))
    return bad

"""Oracle (TEST INFRASTRUCTURE): masked losses and depth metrics.

Reference lines followed (relative to /root/reference):
  model_zoo/confidence_depth_framework.py:808-826  DualLoss
  model_zoo/confidence_depth_framework.py:829-848  MaskedMSELoss
  model_zoo/confidence_depth_framework.py:851-914  inverted-L1 / AbsRel / L1
  metrics.py:38-66                                 Result.evaluate
"""
import math

import torch


def masked_loss(kind, pred, target):
    """kind in {'l2','l1','il1','absrel'}; mean over target>0; asserts >100 valid pixels (:839)."""
    assert pred.dim() == target.dim(), "inconsistent dimensions"
    valid = (target > 0).detach()
    assert int(valid.sum()) > 100, "training image has less than 100 valid pixels"
    t, p = target[valid], pred[valid]
    if kind == "l2":
        return ((t - p) ** 2).mean()
    if kind == "l1":
        return (t - p).abs().mean()
    if kind == "il1":
        return (1.0 / t - 1.0 / p).abs().mean()
    if kind == "absrel":
        return ((t - p) / t).abs().mean()
    raise ValueError(kind)


def dual_loss(kind, depth_a, depth_b, target, weight_b):
    """confidence_depth_framework.py:808-826.  Returns (total, [L(a), L(b), 0])."""
    ea = masked_loss(kind, depth_a, target)
    parts = [float(ea), 0, 0]
    if weight_b > 0:
        eb = masked_loss(kind, depth_b, target)
        parts[1] = float(eb)
        return weight_b * eb + (1.0 - weight_b) * ea, parts
    return ea, parts


def evaluate(output, target, confidence=None, threshold=0.0):
    """metrics.py:38-66.  Returns a dict of python floats."""
    valid = target > 0 if confidence is None else (target > 0) & (confidence > threshold)
    o, t = output[valid], target[valid]
    ad = (o - t).abs()
    r = {}
    r["mse"] = float((ad ** 2).mean())
    r["rmse"] = math.sqrt(r["mse"])
    r["mae"] = float(ad.mean())
    r["lg10"] = float((torch.log(o) / math.log(10) - torch.log(t) / math.log(10)).abs().mean())
    r["absrel"] = float((ad / t).mean())
    ratio = torch.max(o / t, t / o)
    r["delta1"] = float((ratio < 1.25).to(o.dtype).mean())
    r["delta2"] = float((ratio < 1.25 ** 2).to(o.dtype).mean())
    r["delta3"] = float((ratio < 1.25 ** 3).to(o.dtype).mean())
    aid = (1 / o - 1 / t).abs()
    r["irmse"] = math.sqrt(float((aid ** 2).mean()))
    r["imae"] = float(aid.mean())
    return r

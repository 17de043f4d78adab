"""Single-layer timing of the tcgen05 v2 convolution (64->64 and 128->64, batch 32, 228x304) with the experiment switches."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200 import dense
n, h, w = 32, 228, 304
res = {}
for prec in ("bf16", "bf16x3"):
    dense.set_precision(prec)
    for ci in (64, 128):
        x = dense.new_act(n, h, w, ci, "cuda", pad=True, zero=True)
        x.hi.normal_()
        if x.lo is not None: x.lo.normal_(std=0.01)
        wt = torch.randn(64, ci, 3, 3, device="cuda") * 0.05
        bias = torch.randn(64, device="cuda")
        wp = dense.prep_conv3x3_weights(wt, cin_pad=ci)
        for tag, b, act in (("bias_relu", bias, 1), ("nobias_noact", None, 0)):
            for _ in range(3): o = dense.conv3x3_tc(x, ci, wp, b, 64, 1, act)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): o = dense.conv3x3_tc(x, ci, wp, b, 64, 1, act, out=o, out_zero_to=64)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            fl = 2.0 * n * h * w * 9 * ci * 64
            res[f"{prec}_{ci}_{tag}"] = dict(us=round(us, 1), tflops=round(fl / us / 1e6, 1))
print(json.dumps(res))

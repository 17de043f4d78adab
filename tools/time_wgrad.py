"""Timing of the tcgen05 wgrad kernel at the layer shapes of GED / GMS (batch 32)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200 import dense
n = 32
shapes = [("gms64", 228, 304, 64, 64, 64), ("gms_fuse0", 228, 304, 128, 128, 64), ("ged_conv9", 114, 152, 112, 128, 32),
          ("ged_conv8", 57, 76, 144, 192, 32), ("ged_conv2_dil", 114, 152, 80, 128, 80), ("ged_conv10", 228, 304, 33, 64, 1)]
res = {}
for prec in ("bf16", "bf16x3"):
    dense.set_precision(prec)
    for tag, h, w, cin, cs, cout in shapes:
        x = dense.new_act(n, h, w, cs, "cuda", pad=True, zero=True); x.hi.normal_()
        gy = dense.new_act(n, h, w, dense._r64(cout), "cuda", pad=True, zero=True); gy.hi.normal_()
        cp = min(dense._r16(cin), 160)
        for _ in range(2): dense.conv3x3_wgrad_tc(gy, x, cin, cout, cin_pad=cp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dense.conv3x3_wgrad_tc(gy, x, cin, cout, cin_pad=cp)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        res[f"{prec}_{tag}"] = dict(us=round(us, 1), tflops=round(2.0 * n * h * w * 9 * cin * cout / us / 1e6, 1))
print(json.dumps(res))

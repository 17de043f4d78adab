import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from aerial_depth_completion_b200 import dense
dense.set_precision("bf16")
g = torch.Generator().manual_seed(1)
n, ci, co, h, w = 1, 64, 64, 12, 20
x = torch.randn(n, ci, h, w, generator=g); wt = torch.randn(co, ci, 3, 3, generator=g) * 0.05
ref = F.conv2d(x.bfloat16().double(), wt.bfloat16().double(), None, 1, 1).float()
xa = dense.to_nhwc_bf16(x.cuda(), pad=True)
# check padded conversion round trip
back = dense.to_nchw_f32(xa, ci).cpu()
print("roundtrip err", float((back - x.bfloat16().float()).abs().max()), "border", float(xa.hi[:, 0].abs().max()), float(xa.hi[:, :, 0].abs().max()))
wp = dense.prep_conv3x3_weights(wt.cuda())
for v2 in (False, True):
    dense.USE_V2 = v2
    o = dense.conv3x3_tc(xa, ci, wp, None, co, 1, 0, out_f32=True)
    got = dense.to_nchw_f32(o, co).cpu()
    print("v2" if v2 else "v1-padded", "relerr", float((got - ref).abs().max() / ref.abs().max()))
    if v2:
        # per-tap diagnosis: weights with a single non-zero tap
        for tap in range(9):
            w1 = torch.zeros_like(wt); w1[:, :, tap // 3, tap % 3] = wt[:, :, tap // 3, tap % 3]
            r1 = F.conv2d(x.bfloat16().double(), w1.bfloat16().double(), None, 1, 1).float()
            o1 = dense.to_nchw_f32(dense.conv3x3_tc(xa, ci, dense.prep_conv3x3_weights(w1.cuda()), None, co, 1, 0, out_f32=True), co).cpu()
            print("  tap", tap, float((o1 - r1).abs().max() / r1.abs().max()))

"""One tcgen05 v2 layer (bf16, 64->64, batch 32, 228x304) launched 6 times: target of the ncu capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200 import dense
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ci = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dense.set_precision(prec)
n, h, w = 32, 228, 304
x = dense.new_act(n, h, w, ci, "cuda", pad=True, zero=True)
x.hi.normal_()
wt = torch.randn(64, ci, 3, 3, device="cuda") * 0.05
bias = torch.randn(64, device="cuda")
wp = dense.prep_conv3x3_weights(wt, cin_pad=ci)
o = None
for _ in range(6):
    o = dense.conv3x3_tc(x, ci, wp, bias, 64, 1, 1, out=o, out_zero_to=64)
torch.cuda.synchronize()
print("ok")

"""One GED / GMS shaped layer on the row-segment kernel, three launches (for ncu: -k regex:conv3x3_s1 -s 1 -c 1).
LAYER = dgrad9 | dgrad9w | fwd9 | gms64 | quad4 | conv1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200 import dense
dense.set_precision("bf16x3")
n = int(os.environ.get("B", 128))
layer = os.environ.get("LAYER", "dgrad9")


def act(hw, c):
    a = dense.new_act(n, hw[0], hw[1], c, "cuda", pad=True, zero=True)
    a.hi[:, 1:-1, 1:-1].normal_()
    a.lo[:, 1:-1, 1:-1].normal_(std=0.004)
    return a


if layer == "dgrad9":
    x, cin, cout, out, off, zt, mask, hw = act((114, 152), 64), 32, 64, act((114, 152), 128), 0, 64, 0, None
elif layer == "dgrad9w":      # the whole 112-channel data gradient of conv9 in one launch (128-channel chunks)
    x, cin, cout, out, off, zt, mask = act((114, 152), 64), 32, 112, act((114, 152), 128), 0, 0, 0
elif layer == "gms64":        # a 64 -> 64 layer of GMSNet at B = 32
    n = int(os.environ.get("B", 32))
    x, cin, cout, out, off, zt, mask = act((228, 304), 64), 64, 64, act((228, 304), 64), 0, 64, 0
elif layer == "fwd9":
    x, cin, cout, out, off, zt, mask = act((114, 152), 128), 112, 32, act((114, 152), 64), 0, 64, 0
else:
    x, cin, cout, out, off, zt, mask = act((57, 76), 128), 80, 80, act((57, 76), 320), 240, 0, 0b110110000
if layer == "conv1":          # GED conv1 as a 1x1 GEMM over the im2col tensor: K = 48, 80 output channels + bias + ReLU into cat9[32:112]
    x45, cat9 = act((114, 152), 64), act((114, 152), 128)
    w1 = torch.zeros(80, 48, 3, 3, device="cuda")
    w1[:, :45, 1, 1] = torch.randn(80, 45, device="cuda") * 0.05
    wp1 = dense.prep_conv3x3_weights(w1, cin_pad=48)
    bias = torch.randn(80, device="cuda")
    for _ in range(3):
        dense.conv1x1_tc(x45, 48, wp1, bias, 80, 1, out=cat9, out_off=32)
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)
wp = dense.prep_conv3x3_weights(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05, cin_pad=cin)
for _ in range(3):
    dense.conv3x3_tc(x, cin, wp, None, cout, 1, 0, out=out, out_off=off, out_zero_to=zt, tap_mask=mask)
torch.cuda.synchronize()
print("ok")

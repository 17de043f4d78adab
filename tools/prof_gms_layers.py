"""Per-launch times of one GMS inference forward (B=32, 228x304) from the library's own event records."""
import sys, os, contextlib, io, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense
from aerial_depth_completion_b200._lib import lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
f = mc.ConfidenceDepthFrameworkFactory()
with contextlib.redirect_stdout(io.StringIO()):
    m = f.create_model("rgbd", "dc1_only", "gms_depthcompnet", "")
    m, _ = f.to_device(m)
    m.train(False)
x = torch.rand(B, 5, 228, 304, device="cuda")
mask = (torch.rand(B, 1, 228, 304, device="cuda") < 0.007).float()
x[:, 3:4] *= mask; x[:, 4:5] = mask
dense.set_precision(prec)
with torch.no_grad():
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    lib.adc_profile_enable(1)
    m(x)
    torch.cuda.synchronize()
cap = 512
n = ctypes.c_int(0)
kinds = (ctypes.c_int * cap)(); cl = (ctypes.c_int * cap)(); ms = (ctypes.c_float * cap)(); by = (ctypes.c_double * cap)()
lib.adc_profile_read(cap, ctypes.byref(n), kinds, cl, ms, by)
lib.adc_profile_enable(0)
tot = 0.0
for j in range(n.value):
    tot += ms[j] * 1e3
    if kinds[j] >= 20:
        print("kind %2d  work %10.3g  %8.1f us  %s" % (kinds[j], by[j], ms[j] * 1e3, ("%.0f TF/s" % (by[j] / ms[j] / 1e9)) if kinds[j] in (20, 21) else ("%.0f GB/s" % (by[j] / ms[j] / 1e6))))
print("library kernels: %d launches, %.1f us" % (n.value, tot))

"""GMS inference (BASELINE config 1: batch 32, 228x304) timing: library fp32 vs hand-written tcgen05 modes."""
import sys, os, contextlib, io, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
f = mc.ConfidenceDepthFrameworkFactory()
with contextlib.redirect_stdout(io.StringIO()):
    m = f.create_model("rgbd", "dc1_only", "gms_depthcompnet", "")
    m, _ = f.to_device(m)
    m.train(False)
x = torch.rand(B, 5, 228, 304, device="cuda")
mask = (torch.rand(B, 1, 228, 304, device="cuda") < 0.007).float()
x[:, 3:4] *= mask; x[:, 4:5] = mask
out = {}
for prec in ("fp32", "bf16x3", "bf16"):
    dense.set_precision(prec)
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): m(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[prec] = dict(ms=round(ms, 3), frames_s=round(B / ms * 1e3, 1), tflops=round(49.3e9 * B / ms / 1e9, 1))
dense.set_precision("fp32")
print(json.dumps(out))

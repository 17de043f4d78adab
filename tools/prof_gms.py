import sys, os, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
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
print("done")

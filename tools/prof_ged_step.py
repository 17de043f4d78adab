"""One eager GED training step at bench size (for ncu: --kernel-name regex filters pick the kernels)."""
import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense
B, H, W = int(os.environ.get("B", 32)), 228, 304
torch.manual_seed(0)
f = mc.ConfidenceDepthFrameworkFactory()
with contextlib.redirect_stdout(io.StringIO()):
    m = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
    m, params = f.to_device(m)
    m.train()
crit, _ = f.create_loss("l2", False, 0.5)
dense.set_precision(os.environ.get("PREC", "bf16x3"))
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(B, 5, H, W, device="cuda", generator=g)
mask = (torch.rand(B, 1, H, W, device="cuda", generator=g) < 0.0072).float()
x[:, 3:4] = (x[:, 3:4] * 9 + 1) * mask
x[:, 4:5] = mask
t = torch.rand(B, 1, H, W, device="cuda", generator=g) * 9 + 1
for _ in range(int(os.environ.get("STEPS", 2))):
    d1, _, _ = m(x)
    loss = crit(x, d1, t)
    for p in params:
        p.grad = None
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))

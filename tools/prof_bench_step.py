"""Eager training steps of the bench workload (ged_depthcompnet dc1_only, B=32, 228x304, bf16x3, FlatAdam) for an ncu launch list:
STEPS steps after the plans exist, so `-s <launches of the warm-up steps>` can skip to a steady-state step."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense, parallel
from aerial_depth_completion_b200.optim import FlatAdam
B, H, W = int(os.environ.get("B", 32)), 228, 304
torch.manual_seed(0)
f = mc.ConfidenceDepthFrameworkFactory()
with contextlib.redirect_stdout(io.StringIO()):
    m = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
    m, params = f.to_device(m)
    m.train()
crit, _ = f.create_loss("l2", False, 0.5)
dense.set_precision(os.environ.get("PREC", "bf16x3"))
tr = parallel.DataParallelTrainer(m, crit, FlatAdam(params, lr=1e-5), params, world_size=1)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(B, 5, H, W, device="cuda", generator=g)
mask = (torch.rand(B, 1, H, W, device="cuda", generator=g) < 0.0072).float()
x[:, 3:4] = (x[:, 3:4] * 9 + 1) * mask
x[:, 4:5] = mask
t = torch.rand(B, 1, H, W, device="cuda", generator=g) * 9 + 1
for i in range(int(os.environ.get("STEPS", 3))):
    if i == int(os.environ.get("MARK", -1)):
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    loss = tr._eager_step(x, t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(loss))

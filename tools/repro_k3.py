import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200.model_zoo.nconv import NConv2d
ci, k = int(sys.argv[1]), int(sys.argv[2])
m = NConv2d(ci, 2, (k, k), "softplus", "p", padding=k // 2).cuda().eval()
x = torch.rand(2, ci, 24, 32, device="cuda"); c = torch.rand(2, ci, 24, 32, device="cuda")
y, co = m(x, c); torch.cuda.synchronize(); print(ci, k, "ok", float(y.sum()))

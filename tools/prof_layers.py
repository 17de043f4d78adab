"""Launch every NConv kernel variant a few times at the bench shape (for an ncu launch list / full capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200.model_zoo.nconv import NConv2d
from aerial_depth_completion_b200.model_zoo.nconv_sd import CNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W = 228, 304
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
which = sys.argv[3] if len(sys.argv) > 3 else "all"
if which in ("all", "layers"):
    for (ci, co, k, h, w) in ((1, 2, 5, H, W), (2, 2, 5, H, W), (4, 2, 3, H, W), (2, 1, 1, H, W)):
        m = NConv2d(ci, co, (k, k), "softplus", "p", padding=k // 2).cuda().eval()
        x = torch.rand(B, ci, h, w, device="cuda").requires_grad_(True)
        c = torch.rand(B, ci, h, w, device="cuda").requires_grad_(True)
        for _ in range(reps):
            y, cout = m(x, c)
            torch.autograd.grad((y, cout), (x, c, m.weight, m.bias), (torch.ones_like(y), torch.ones_like(cout)))
if which in ("all", "pyramid"):
    net = CNN().cuda().train()
    x0 = (torch.rand(B, 1, H, W, device="cuda") * (torch.rand(B, 1, H, W, device="cuda") < 0.007)).requires_grad_(True)
    c0 = (x0.detach() > 0).float().requires_grad_(True)
    for _ in range(reps):
        xo, co = net(x0, c0)
        torch.autograd.grad((xo, co), (x0, c0) + tuple(net.parameters()), (torch.ones_like(xo), torch.ones_like(co)))
torch.cuda.synchronize()
print("done")

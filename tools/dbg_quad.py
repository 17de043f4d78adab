import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import relmax
from oracle import synth
from aerial_depth_completion_b200 import dense
import test_gpu_quad as T
n, h, w = 1, 228, 304
oc, build = 1, False
m = T._ged(oc, 5)
inp, _ = synth.synth_frames(n, h, w, seed=31, num_samples=max(60, h * w // 140))
x = inp.cuda()
g = torch.Generator(device="cuda").manual_seed(3)
wts = (torch.rand(n, 1, h, w, device="cuda", generator=g) + 0.5, torch.rand(n, 36, h, w, device="cuda", generator=g) * 0.1)
dense.set_precision("fp32")
ref = T._run(m, x, build, wts)
ref2 = T._run(m, x, build, wts)
print("fp32 rerun", max(relmax(ref2[3][k], ref[3][k]) for k in ref[3]))
dense.set_precision("bf16x3")
runs = [T._run(m, x, build, wts) for _ in range(3)]
for r in runs:
    errs = {k: relmax(r[3][k], ref[3][k]) for k in ref[3]}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print("out", relmax(r[0], ref[0]), "gin", relmax(r[2][:, 3:], ref[2][:, 3:]), worst)
print("tc rerun diff", max(relmax(runs[1][3][k], runs[0][3][k]) for k in ref[3]), relmax(runs[1][2][:, 3:], runs[0][2][:, 3:]))
# fp64 oracle on the CPU: which of the two GPU paths is closer to the truth?
from oracle import nets_oracle as nets
st = {("dc_model." + k): v.detach().double().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
d1, _, _ = nets.framework_forward(st, inp.double(), "dc1_only", "ged_depthcompnet", 4, training=True)
(d1 * wts[0].double().cpu()).sum().backward()
for k in ("d_net.nconv7.weight", "d_net.nconv1.weight", "d_net.nconv6.weight", "conv1.weight", "conv10.weight"):
    t = st["dc_model." + k].grad
    print(k, "fp32-lib vs fp64", relmax(ref[3][k].cpu().double(), t), "tc vs fp64", relmax(runs[0][3][k].cpu().double(), t), "val", t.flatten()[:2].tolist())

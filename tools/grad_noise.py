"""Per-tensor gradient error of the GED training step against the fp64 oracle, next to the reference's own fp32 floor (nets.npz),
for each arithmetic mode.  Usage: python tools/grad_noise.py [tag] > gpurun_out/grad_noise.txt"""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import net_case, ref_grad_errors, relmax, run_oracle_case
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense

tag = sys.argv[1] if len(sys.argv) > 1 else "ged_nyu"
cfg, state, inp, tgt = net_case(tag)
_, S, _, _, d1, c1, d2, loss64, _ = run_oracle_case(tag, torch.float64)
floor = ref_grad_errors(tag)
res = {}
for mode in ("fp32", "bf16x3", "bf16"):
    dense.set_precision(mode)
    f = mc.ConfidenceDepthFrameworkFactory()
    with contextlib.redirect_stdout(io.StringIO()):
        m = f.create_model(cfg["input_type"], cfg["mode"], cfg["dc_arch"], "", cfg["conf_arch"], "", cfg["ln_arch"], "")
        m.load_state_dict(state)
        m, params = f.to_device(m)
        m.train(cfg["train"])
    o1, oc, o2 = m(inp.cuda())
    crit, _ = f.create_loss(cfg["criterion"], "ln" in cfg["mode"], 0.5 if "dc1" in cfg["mode"] else 1.0)
    l = crit(inp, o1, o2, tgt.cuda()) if o2 is not None else crit(inp, o1, tgt.cuda())
    l.backward()
    res[mode] = {k: relmax(p.grad.cpu(), S[k].grad) for k, p in m.named_parameters() if p.grad is not None and k in floor}
    print(mode, "d1 err %.2e" % relmax(o1.detach().cpu(), d1.detach()), "loss rel %.2e" % abs(float(l) / float(loss64) - 1))
print("%-44s %9s %9s %9s %9s" % ("tensor", "ref_fp32", "fp32lib", "bf16x3", "bf16"))
for k in floor:
    print("%-44s %9.2e %9.2e %9.2e %9.2e" % (k, floor[k], res["fp32"].get(k, -1), res["bf16x3"].get(k, -1), res["bf16"].get(k, -1)))

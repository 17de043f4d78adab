"""Per-kernel CUDA-event profile (the library's own instrumentation) of ONE eager training step of a framework configuration.
Usage: python tools/prof_step.py gud|ged|gms|full|s2d|ud [batch] > gpurun_out/prof_<cfg>.txt"""
import contextlib, ctypes, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
from aerial_depth_completion_b200 import dense, parallel, tc_train
from aerial_depth_completion_b200._lib import lib
from aerial_depth_completion_b200.optim import FlatAdam

cfg = sys.argv[1] if len(sys.argv) > 1 else "gud"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"gud": 16, "ged": 32, "gms": 32, "full": 8, "s2d": 64, "ud": 32}[cfg]
H, W, args, dual = {"gud": (240, 320, ("rgbd", "dc1-cf1-ln0", "gudepthcompnet18", "", "cbr3-c1", "", "ged_depthcompnet", ""), True),
                    "ged": (228, 304, ("rgbd", "dc1_only", "ged_depthcompnet", ""), False),
                    "gms": (228, 304, ("rgbd", "dc1_only", "gms_depthcompnet", ""), False),
                    "s2d": (228, 304, ("rgbd", "dc1_only", "resnet18", ""), False),
                    "ud": (240, 320, ("rgbdw", "dc1_only", "udepthcompnet18", ""), False),
                    "full": (480, 752, ("rgbdw", "dc1-cf1-ln1", "ged_depthcompnet", "", "cbr3-c1", "", "ged_depthcompnet", ""), True)}[cfg]
torch.manual_seed(0)
f = mc.ConfidenceDepthFrameworkFactory()
with contextlib.redirect_stdout(io.StringIO()):
    m = f.create_model(*args)
    m, params = f.to_device(m)
    m.train()
crit, _ = f.create_loss("l2", dual, 0.5)
tr = parallel.DataParallelTrainer(m, crit, FlatAdam(params, lr=1e-5), params, world_size=1)
g = torch.Generator(device="cuda").manual_seed(13)
x = torch.rand(B, 5, H, W, device="cuda", generator=g)
gt = torch.rand(B, 1, H, W, device="cuda", generator=g) * 9.0 + 1.0
mask = (torch.rand(B, 1, H, W, device="cuda", generator=g) < 500.0 / (H * W)).float()
x[:, 3:4] = gt * mask
x[:, 4:5] = mask
tc_train._SIDE_STREAM = False
for _ in range(2):
    tr._eager_step(x, gt)
torch.cuda.synchronize()
lib.adc_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr._eager_step(x, gt)
e1.record()
torch.cuda.synchronize()
cap = 8192
n = ctypes.c_int(0)
kinds, cl, ms, by = (ctypes.c_int * cap)(), (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_double * cap)()
lib.adc_profile_read(cap, ctypes.byref(n), kinds, cl, ms, by)
lib.adc_profile_enable(0)
names = {0: "nconv_fwd", 1: "nconv_bwd_data", 2: "nconv_bwd_weight", 3: "pool_fwd", 4: "pool_bwd", 5: "upsample_bwd", 6: "head_fwd", 7: "head_bwd",
         8: "pack", 9: "reduce", 10: "prep", 20: "tc_conv_box", 21: "tc_conv_rowseg", 22: "tc_dgrad_s2/convT", 23: "tc_wgrad_flat", 24: "tc_wgrad_box",
         25: "tc_glue", 26: "tc_layout", 27: "tc_prep_weights", 28: "quad_head", 29: "loss+adam", 30: "conf_head_fused", 31: "bn/act/dot"}
tot = {}
rows = []
for j in range(n.value):
    k = names.get(kinds[j], str(kinds[j]))
    a = tot.setdefault(k, [0.0, 0, 0.0])
    a[0] += ms[j] * 1e3
    a[1] += 1
    a[2] += by[j]
    rows.append((ms[j] * 1e3, k, by[j]))
wall = e0.elapsed_time(e1) * 1e3
print("config %s B=%d %dx%d: eager step wall %.0f us (with per-kernel events), sum of library kernels %.0f us, %d records" %
      (cfg, B, H, W, wall, sum(a[0] for a in tot.values()), n.value))
print("%-22s %6s %10s %12s" % ("kind", "count", "us", "work/us"))
for k, (us, cnt, work) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%-22s %6d %10.1f %12.4g   (%s)" % (k, cnt, us, work / max(us, 1e-9) / 1e6, "TFLOP/s" if k.startswith("tc_conv") or k.startswith("tc_wgrad") or k.startswith("tc_dgrad") else "TB/s"))
print("top launches:")
for us, k, work in sorted(rows, key=lambda r: -r[0])[:40]:
    print("  %8.1f us  %-20s work %.4g  -> %.1f per us (T)" % (us, k, work, work / max(us, 1e-9) / 1e6))
if os.environ.get("ALL_KIND"):
    print("all launches of kind %s, in issue order:" % os.environ["ALL_KIND"])
    for us, k, work in rows:
        if k == os.environ["ALL_KIND"]:
            print("  %8.1f us  work %.4g  -> %.2f per us (T)" % (us, work, work / max(us, 1e-9) / 1e6))

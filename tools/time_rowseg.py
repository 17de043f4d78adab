"""Row-segment kernel timings of GED's layer shapes at B = 128 (228x304 -> levels 114x152, 57x76) under the experiment switches:
ADC_V2_NO_PAIR (single-CTA MMAs), ADC_V2_NO_HSTACK, ADC_V2_DBG=1 (epilogue without stores) / 2 (no epilogue), ADC_V2_NO_RESIDENT."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200 import dense
dense.set_precision(os.environ.get("PREC", "bf16x3"))
n = int(os.environ.get("B", 128))
L1, L2 = (114, 152), (57, 76)


def act(hw, c, f32=False):
    a = dense.new_act(n, hw[0], hw[1], c, "cuda", pad=True, zero=True, dtype=torch.float32 if f32 else torch.bfloat16)
    if not f32:
        a.hi[:, 1:-1, 1:-1].normal_()
        if a.lo is not None:
            a.lo[:, 1:-1, 1:-1].normal_(std=0.004)
    return a


def run(name, x, cin, cout, out, out_off, zero_to, act_fn=0, one=False, mask=0, x_off=0, bias=True):
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(cout, device="cuda") if bias else None
    wp = dense.prep_conv3x3_weights(wt, cin_pad=cin)
    f = (lambda: dense.conv1x1_tc(x, cin, wp, b, cout, act_fn, x_off=x_off, out=out, out_off=out_off, out_zero_to=zero_to)) if one else \
        (lambda: dense.conv3x3_tc(x, cin, wp, b, cout, 1, act_fn, x_off=x_off, out=out, out_off=out_off, out_zero_to=zero_to, tap_mask=mask))
    res = {}
    S = {"ADC_V2_NO_PAIR": "1"}
    for tag, env in (("pair", {}), ("single", S), ("pair_notri", {"ADC_V2_NO_TRI": "1"}), ("pair_tri64", {"ADC_V2_TRI_MIN": "64"}), ("pair_noepi", {"ADC_V2_DBG": "2"}),
                     ("pair_notri_noepi", {"ADC_V2_DBG": "2", "ADC_V2_NO_TRI": "1"})):
        for k in ("ADC_V2_NO_PAIR", "ADC_V2_DBG", "ADC_V2_PF", "ADC_V2_MAXST", "ADC_V2_NO_WIDE_TMA", "ADC_V2_EPI4", "ADC_V2_HSTACK_MAX", "ADC_V2_NO_TRI", "ADC_V2_TRI_MIN"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            for _ in range(2):
                f()
        except RuntimeError:              # (a form that does not fit: e.g. 128 output channels on single CTAs in the split mode)
            res[tag] = None
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        res[tag] = round(e0.elapsed_time(e1) * 100, 1)
    for k in ("ADC_V2_NO_PAIR", "ADC_V2_DBG", "ADC_V2_PF", "ADC_V2_MAXST", "ADC_V2_NO_WIDE_TMA", "ADC_V2_EPI4", "ADC_V2_HSTACK_MAX", "ADC_V2_NO_TRI", "ADC_V2_TRI_MIN"):
        os.environ.pop(k, None)
    print(name, json.dumps(res), flush=True)


if os.environ.get("SET", "ged") == "wide":
    def act2(nn, hw, c):
        global n
        n0, n = n, nn
        a = act(hw, c)
        n = n0
        return a
    for nn, hw, ci, co in ((16, (120, 160), 128, 128), (16, (60, 80), 256, 128), (16, (30, 40), 512, 128), (16, (240, 320), 144, 128),
                           (32, (228, 304), 64, 64), (16, (120, 160), 64, 64), (16, (60, 80), 128, 64), (16, (30, 40), 256, 64),
                           (16, (15, 20), 512, 64), (16, (240, 320), 144, 64)):
        n = nn
        x, o = act(hw, ci), act(hw, max(64, co))
        run("wide %d->%d %dx%d B=%d" % (ci, co, hw[0], hw[1], nn), x, ci, co, o, 0, max(64, co), act_fn=1)
        del x, o
    sys.exit(0)
G9, gcat9 = act(L1, 64), act(L1, 128)
run("conv9_dgrad_chunk0 32->64 L1", G9, 32, 64, gcat9, 0, 64, bias=False)
cat9, x9 = act(L1, 128), act(L1, 64)
run("conv9_fwd 112->32 L1", cat9, 112, 32, x9, 0, 64, act_fn=2)
x45 = act(L1, 64)
run("conv1_gemm 48->80 L1", x45, 48, 80, cat9, 32, 0, act_fn=1, one=True)
q10 = act(L1, 16, f32=True)
run("conv10_quad 32->8 L1 f32", x9, 32, 8, q10, 0, 0, bias=False)
G2, gx1 = act(L2, 128), act(L2, 320)
run("conv2_quad_dgrad 80->80 L2 mask 4 taps", G2, 80, 80, gx1, 240, 0, mask=0b110110000, bias=False)
run("conv2_quad_dgrad 80->80 L2 mask 1 tap", G2, 80, 80, gx1, 0, 0, mask=0b000010000, bias=False)
cat8, x8 = act(L2, 192), act(L2, 64)
run("conv8_fwd 144->32 L2", cat8, 144, 32, x8, 0, 64, act_fn=2)
G1, qin0 = act(L1, 128), act(L1, 16, f32=True)
run("conv1_dgrad 80->8 L1 f32", G1, 80, 8, qin0, 0, 0, bias=False)

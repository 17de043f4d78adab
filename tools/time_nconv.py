"""Quick CUDA-event timings of the NConv kernels (development aid, not the bench): per-layer and whole pyramid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200.model_zoo.nconv import NConv2d
from aerial_depth_completion_b200.model_zoo.nconv_sd import CNN

PEAK = 6535.4


def timeit(fn, warm=5, it=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    H, W = 228, 304
    out = {}
    for (ci, co, k, h, w) in ((1, 2, 5, H, W), (2, 2, 5, H, W), (4, 2, 3, H, W), (2, 1, 1, H, W), (2, 2, 5, H // 2, W // 2)):
        m = NConv2d(ci, co, (k, k), "softplus", "p", padding=k // 2).cuda().eval()
        x = torch.rand(B, ci, h, w, device="cuda").requires_grad_(True)
        c = torch.rand(B, ci, h, w, device="cuda").requires_grad_(True)
        with torch.no_grad():
            t_f = timeit(lambda: m(x, c))
        y, cout = m(x, c)
        gy, gc = torch.rand_like(y), torch.rand_like(cout)

        def bwd():
            torch.autograd.grad((y, cout), (x, c, m.weight, m.bias), (gy, gc), retain_graph=True)
        t_b = timeit(bwd)
        px = B * h * w
        bf, bb = 8 * (ci + co) * px, 16 * (ci + co) * px
        out["%d->%d k%d %dx%d" % (ci, co, k, h, w)] = dict(fwd_ms=round(t_f, 4), fwd_GBs=round(bf / t_f / 1e6, 1), fwd_frac=round(bf / t_f / 1e6 / PEAK, 3),
                                                     bwd_ms=round(t_b, 4), bwd_GBs=round(bb / t_b / 1e6, 1), bwd_frac=round(bb / t_b / 1e6 / PEAK, 3))
    net = CNN().cuda().train()
    x0 = (torch.rand(B, 1, H, W, device="cuda") * (torch.rand(B, 1, H, W, device="cuda") < 0.007)).requires_grad_(True)
    c0 = (x0.detach() > 0).float().requires_grad_(True)
    with torch.no_grad():
        t_f = timeit(lambda: net(x0, c0))

    def step():
        xo, co = net(x0, c0)
        torch.autograd.grad((xo, co), (x0, c0) + tuple(net.parameters()), (torch.ones_like(xo), torch.ones_like(co)))
    t_s = timeit(step)
    out["pyramid B=%d" % B] = dict(fwd_ms=round(t_f, 3), fwd_frames_s=round(B / t_f * 1e3, 1), fwd_GBs_model=round(13.4e6 * B / t_f / 1e6, 1),
                                  step_ms=round(t_s, 3), step_frames_s=round(B / t_s * 1e3, 1))
    for k, v in out.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()

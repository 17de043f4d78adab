#!/usr/bin/env python
"""Benchmark of the NConv depth-completion hot path (BASELINE.json metric: frames/sec forward+backward at NYU
228x304, batch per GPU, 1/2/4/8 B200; NConv achieved HBM GB/s as a fraction of the measured peak).

  python bench.py --gpus N --steps K --warmup W            # this repo's arm (one JSON line from rank 0)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU implementation (oracle port)
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): ged_depthcompnet, --training-mode dc1_only, masked-L2, fp32, synthetic NYU-shape frames
(SURVEY.md 8d), one step = forward + loss + zero_grad + backward (+ gradient all-reduce for N > 1) + Adam step over
`--batch` frames per GPU.  `value` times steps whose inputs are resident in HBM; `e2e` times the same step through the
public API with the frames coming from pinned host memory (H2D inside the timed region) and the loss read back (D2H).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 228, 304
METRIC = "frames/sec fwd+bwd (NYU 228x304, bs per GPU)"
WORKLOAD = "ged_depthcompnet dc1_only masked-l2 train step (fwd+loss+bwd+adam), rgbd 228x304 synthetic NYUv2-shape"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--batch", type=int, default=128, help="frames per GPU per step (extra.batch_sweep reports 8 = the reference's default, 32, 64)")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--cpu-frames", type=int, default=0, help="frames per step of the CPU arms (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    p.add_argument("--no-extras", action="store_true", help="skip the extra (non-headline) measurements")
    p.add_argument("--kernel-detail", default="", help="write one step's per-launch records (kind, work, us) to this file")
    p.add_argument("--precision", default="bf16x3", choices=["fp32", "tf32", "bf16x3", "bf16"],
                   help="dense-layer arithmetic: fp32/tf32 = library backend; bf16x3 (fp32-grade, hi/lo split bf16 operands) and "
                        "bf16 = hand-written tcgen05 kernels for conv1..conv10 forward and backward")
    return p.parse_args()


def _load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


PEAKS = _load_peaks()


def _load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the NConv kernels from the committed `ncu --set full` capture
    (profiles/r2_nconv_traffic.json at the default batch, written by tools/ncu_traffic.py from the .ncu-rep): kernel name as in the roofline
    table.  Only reported when the capture's batch size is this run's (the source string names it)."""
    for name in ("r2c_nconv_traffic.json", "r2_nconv_traffic.json", "r1_nconv_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return {}


TRAFFIC = _load_traffic()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated in oracle/ (the Python reference cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------------------------------
def _reference_root():
    """/root/reference (build container) or the byte-identical copy staged by oracle/make_ref.py (oracle/_ref: travels to the GPU box)."""
    for p in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if os.path.isdir(os.path.join(p, "model_zoo")):
            return p
    return None


def _load_reference():
    """The unmodified reference modules through the two-patch shim of SURVEY.md 8c (tests/reference_shim.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reference_shim
    return reference_shim.load()


def cpu_reference_run(steps, warmup, frames):
    """GED dc1_only train step (forward + masked-L2 + zero_grad + backward + Adam) on the host cores with all threads: returns
    (frames/s, ms/step, cores, frames, kind).  kind "reference" = the UNMODIFIED reference modules (model_zoo/*, MaskedMSELoss) on CPU
    torch when the reference is present (/root/reference or oracle/_ref); "port" = the oracle's restatement otherwise."""
    import contextlib
    import io
    import torch
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp, tgt = synth.synth_frames(frames, H, W, seed=1234)
    if _reference_root() is not None:
        assert not torch.cuda.is_available(), "the CPU arm runs with CUDA_VISIBLE_DEVICES='' (the reference hard-codes .cuda())"
        ns = _load_reference()
        torch.manual_seed(0)
        f = ns.cdf.ConfidenceDepthFrameworkFactory()
        with contextlib.redirect_stdout(io.StringIO()):
            model = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
            model, params = f.to_device(model)
            model.train()
        crit, _ = f.create_loss("l2", False, 0.5)
        opt = torch.optim.Adam(params, lr=1e-5)
        inp, tgt = inp.float(), tgt.float()

        def step():
            with contextlib.redirect_stdout(io.StringIO()):
                d1, _, _ = model(inp)
            loss = crit(inp, d1[:, 0:1], tgt, 0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.detach())
        kind = "reference"
    else:
        from oracle import loss_metrics_oracle as lm
        from oracle import nets_oracle as nets
        template = ged_template()
        S = synth.seeded_state({"dc_model." + k: v for k, v in template.items()}, seed=0)
        for k in S:
            S[k].requires_grad_(True)
        opt = torch.optim.Adam([S[k] for k in S], lr=1e-5)

        def step():
            d1, _, _ = nets.framework_forward(S, inp, "dc1_only", "ged_depthcompnet", 4, training=True)
            loss = lm.masked_loss("l2", d1, tgt)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.detach())
        kind = "port"

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return frames * steps / dt, dt / steps * 1e3, cores, frames, kind


_KIND_NOTE = {"reference": "the UNMODIFIED reference (model_zoo.ConfidenceDepthFrameworkFactory, MaskedMSELoss, torch.optim.Adam) on CPU torch, "
                           "all host threads",
              "port": "torch CPU ops (oracle/, the reference's algorithm restated; no staged reference on this box)"}


def ged_template():
    """state_dict shapes of ged_depthcompnet with one output channel (confidence_depth_framework.py:426-452)."""
    t = {}
    for i, (co, ci, k) in enumerate(((2, 1, 5), (2, 2, 5), (2, 2, 5), (2, 4, 3), (2, 4, 3), (2, 4, 3), (1, 2, 1)), 1):
        t["d_net.nconv%d.weight" % i] = (co, ci, k, k)
        t["d_net.nconv%d.bias" % i] = (co,)
    for i, (ci, co) in enumerate(((5, 80), (80, 80), (80, 80), (80, 80), (80, 80), (160, 64), (144, 64), (144, 32), (112, 32), (33, 1)), 1):
        t["conv%d.weight" % i] = (co, ci, 3, 3)
        t["conv%d.bias" % i] = (co,)
    return t


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same workload and config, exactly --steps / --warmup steps; each step is a BOUNDED SAMPLE of the arm's batch (32 of the 128 frames
    # by default: ~1.5 s per step on 16 cores, so the driver's 20 + 5 steps end in well under a minute; frames/s does not depend on it
    # beyond a few per cent -- measured 19 frames/s at 128 frames per step, 22 at 32, 27-34 at 8)
    frames = args.cpu_frames or min(args.batch, 32)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    fps, ms, cores, frames, kind = cpu_reference_run(steps, warmup, frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision": "fp32", "batch_per_gpu": args.batch, "global_batch": args.batch, "device": "host CPU",
                   "sample_frames_per_step": frames},
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d timed steps of %d frame(s) (a bounded sample of the %d-frame batch) of the same workload: %s"
                                   % (steps, frames, args.batch, _KIND_NOTE[kind])},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(frames, steps=3, warmup=1):
    """The CPU arm in its own process with CUDA hidden (the reference's pyramid hard-codes .cuda(); the shim makes it a no-op only when
    no device is visible): returns the parsed cpu_baseline object of that run, or None."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup),
                            "--cpu-frames", str(frames)], capture_output=True, text=True, timeout=600, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled WHILE the timed regions run: through NVML (nvidia_ml_py, ~10 ms per sample) when it
    loads, else one nvidia-smi call per sample (~100 ms)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    BITS = (0x8, 0x40, 0x20, 0x4)          # NVML clocks-event-reason bits, in the order of NAMES

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.source = index, [], False, "nvidia-smi"

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain list of indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            ids = [int(t) for t in vis.split(",")]
            if self.index < len(ids):
                idx = ids[self.index]
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        return pynvml, h, reasons, mx

    def run(self):
        try:
            nv, h, reasons, mx = self._nvml()
            self.source = "nvml"
            while not self.stop_flag:
                mask = int(reasons(h))
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                                 ["Active" if mask & b else "Not Active" for b in self.BITS])
                time.sleep(0.01)
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "source": self.source}


def _finish(world, trainer=None):
    """End of a multi-rank run.  Round 1 measured that dist.destroy_process_group() never returned after the step graph (which holds
    the captured NCCL all-reduces) had been replayed.  The ranks now release the graph first (parallel.DataParallelTrainer.release),
    meet at a barrier and tear the communicator down under a watchdog: if the teardown has not returned after 20 s the process
    leaves with os._exit(0) as before (the JSON line is already flushed either way); the outcome is printed on stderr."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    if trainer is not None:
        trainer.release()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(20.0, lambda: (sys.stderr.write("bench: destroy_process_group() did not return within 20 s; leaving\n"),
                                           sys.stderr.flush(), os._exit(0)))
    timer.daemon = True
    timer.start()
    t0 = time.perf_counter()
    dist.destroy_process_group()
    timer.cancel()
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("bench: destroy_process_group() returned after %.2f s\n" % (time.perf_counter() - t0))
    sys.stderr.flush()
    os._exit(0)


def b200_arm(args):
    import torch
    import torch.distributed as dist
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import _lib, parallel
    from aerial_depth_completion_b200.dataloaders.dense_to_sparse import UniformSampling
    import numpy as np
    import contextlib
    import io

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    # model: reference constructors' distributions under manual_seed(0) (same seed on every rank => identical replicas)
    torch.manual_seed(0)
    factory = mc.ConfidenceDepthFrameworkFactory()
    with contextlib.redirect_stdout(io.StringIO()):
        model = factory.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
        model, params = factory.to_device(model)
        model.train()
    criterion, _ = factory.create_loss("l2", False, 0.5)
    from aerial_depth_completion_b200.optim import FlatAdam
    optimizer = FlatAdam(params, lr=1e-5)       # torch.optim.Adam's update as ONE kernel over flat buffers (device-side step counter)
    from aerial_depth_completion_b200 import dense as _dense
    _dense.set_precision(args.precision)
    trainer = parallel.DataParallelTrainer(model, criterion, optimizer, params, world_size=world)

    # synthetic frames: a pool larger than L2 so no step finds its inputs cached
    pool_n = max(2, min(8, (512 << 20) // (B * 6 * H * W * 4) + 1))      # distinct batches: > 400 MB in total, well beyond the 126 MB L2
    g = torch.Generator().manual_seed(1234 + rank)
    sampler = UniformSampling(500, np.inf)
    host_in, host_tg, host_raw = [], [], []
    for i in range(pool_n):
        rgb = torch.rand(B, 3, H, W, generator=g)
        gt = torch.rand(B, 1, H, W, generator=g) * 9.0 + 1.0
        masks, seeds = [], []
        for b in range(B):
            seeds.append(1234 + 1000 * rank + i * B + b)
            np.random.seed(seeds[-1])
            masks.append(sampler.dense_to_sparse(None, gt[b, 0].numpy()))
        mask = torch.from_numpy(np.stack(masks)).unsqueeze(1).float()
        host_in.append(torch.cat((rgb, gt * mask, mask), 1).pin_memory())
        host_tg.append(gt.pin_memory())
        # what a loader has to ship when the sparse depth is drawn on the device: rgb + gt + one seed per frame
        host_raw.append((torch.cat((rgb, gt), 1).pin_memory(), torch.tensor(seeds, dtype=torch.int32).pin_memory()))
    dev_in = [t.to(dev) for t in host_in]
    dev_tg = [t.to(dev) for t in host_tg]
    pool_mb = sum(t.numel() * 4 for t in dev_in + dev_tg) / 1e6

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    resident = lambda i: trainer.step(dev_in[i % pool_n], dev_tg[i % pool_n])

    def from_host(i):
        x = host_in[i % pool_n].to(dev, non_blocking=True)
        t = host_tg[i % pool_n].to(dev, non_blocking=True)
        loss = trainer.step(x, t)
        return float(loss)          # D2H read of the step's result

    launches_per_step = None
    if not args.no_graph:
        l0 = _lib.lib.adc_launch_count()
        trainer._eager_step(dev_in[0], dev_tg[0])
        launches_per_step = _lib.lib.adc_launch_count() - l0      # the graph replays exactly these launches
        trainer.capture(dev_in[0], dev_tg[0])
    for i in range(args.warmup):
        resident(i)
    launches0 = _lib.lib.adc_launch_count()
    sampler_thread = ClockSampler(local)
    if rank == 0:
        sampler_thread.start()
    ms_total = timed(resident, args.steps)
    launches = _lib.lib.adc_launch_count() - launches0
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    # e2e: every step's frames come from pinned host memory and every step's loss is read back on the host, through the public
    # feeder (parallel.HostFeeder: copy of batch i+1 overlaps step i, the loss of step i is read one step later)
    def run_feeder(feeder, batches):
        feeder.run([batches[i % pool_n] for i in range(min(3, args.warmup))])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses = feeder.run([batches[i % pool_n] for i in range(args.steps)])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert len(losses) == args.steps and all(v == v for v in losses), "e2e: a step's loss was not read back"
        return float(t.item()), losses

    # headline e2e: DeviceAssemblyFeeder -- rgb + gt + seeds cross PCIe, the UniformSampling masks and the 5-channel input are built
    # on the device (bit-identical to the host-assembled batches: tests/test_gpu_feeder.py), copy + assembly of batch i+1 overlap step i
    # (measured after the host-assembled variant: 6 planes per frame over PCIe, the loader doing the assembly on the host)
    ms_e2e_host, _ = run_feeder(parallel.HostFeeder(trainer, host_in[0], host_tg[0]), list(zip(host_in, host_tg)))
    ms_e2e, e2e_losses = run_feeder(parallel.DeviceAssemblyFeeder(trainer, host_raw[0][0], 500), host_raw)
    ms_serial = timed(from_host, min(args.steps, 5)) / min(args.steps, 5)      # the reference trainer's order: copy, step, read, next
    sampler_thread.stop_flag = True

    # replicas must still be identical after the timed steps: every rank applied the same all-reduced gradients to the same weights
    ranks_agree = None
    if world > 1:
        cs = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        ranks_agree = all(bool(torch.equal(c, allcs[0])) for c in allcs)
        assert ranks_agree, "parameter checksums differ across ranks after the timed steps: %s" % [float(c) for c in allcs]

    # per-kernel CUDA-event timing of the same step (profiling pass after the timed region)
    roof = kernel_roofline(trainer, dev_in, dev_tg, pool_n, B, args.kernel_detail if rank == 0 else "")

    if rank != 0:
        _finish(world, trainer)
        return
    peak, peak_src = peaks()
    frames = B * world * args.steps
    line = {
        "metric": METRIC, "value": round(frames / ms_total * 1e3, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16x3": "f32 (NConv pyramid) + bf16x3 (dense 3x3 layers: "
                                       "hi/lo split bf16 operands, fp32 accumulation, fp32-grade parity)",
                                       "bf16": "f32 (NConv pyramid) + bf16 (dense 3x3 layers)"}[args.precision], "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision": args.precision, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": "dp%d (one process per GPU, NCCL all-reduce of flat fp32 gradients)" % world,
                   "l2": "inputs rotate through a %.0f MB pool of distinct batches (> 126 MB L2)" % pool_mb,
                   "dense_backend": ("hand-written tcgen05 implicit GEMM (conv1..conv9 forward, data and weight gradients; conv10 = quad-form "
                                     "tensor-core forward + direct CUDA-core backward)"
                                     if args.precision in ("bf16", "bf16x3") else _dense.BACKEND),
                   "launch": "eager" if args.no_graph else "one CUDA graph per step (forward+loss+backward+allreduce+adam)",
                   "timing": "CUDA events, barrier + synchronize on both sides, max over ranks"},
        "clocks": sampler_thread.summary(),
        "e2e": {"value": round(frames / ms_e2e * 1e3, 2), "unit": "frames/s", "h2d_bytes_per_step": B * 4 * H * W * 4 + B * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 4),
                "how": "parallel.DeviceAssemblyFeeder: pinned-host (r, g, b, gt) planes + one int32 seed per frame copied on a copy stream "
                       "while step i runs; adc_sample_assemble draws the reference-exact UniformSampling(500) masks and builds the 5-channel "
                       "input on the device behind the copy; loss of step i read on the host one step later; every batch copied, every "
                       "loss read, all inside the timed region",
                "host_assembled": {"value": round(frames / ms_e2e_host * 1e3, 2), "ms_per_step": round(ms_e2e_host / args.steps, 4),
                                   "h2d_bytes_per_step": B * 6 * H * W * 4,
                                   "how": "parallel.HostFeeder: 5 input planes + target assembled by the loader on the host"},
                "serialised_ms_per_step": round(ms_serial, 4)},
        "gpu_launches": int(launches),
    }
    if ranks_agree is not None:
        line["param_checksum_equal_across_ranks"] = ranks_agree
    if roof:
        dom = roof["dominant"]
        tr_ = TRAFFIC.get(dom["name"]) or {}
        if ("(B=%d," % B) not in (tr_.get("source") or ""):      # a capture at another batch size says nothing about this launch
            tr_ = {}
        line["roofline"] = {"bound": "hbm", "achieved": round(dom["gbs"], 1), "peak": peak, "unit": "GB/s",
                            "frac": round(dom["gbs"] / peak, 4), "traffic": tr_.get("dram_bytes_per_launch"),
                            "traffic_source": tr_.get("source"), "kernel": dom["name"],
                            "launch_us": round(dom["us"], 2), "algorithmic_mb_per_launch": round(dom["mb"], 2),
                            "peak_source": peak_src, "nconv_kernels": roof["table"],
                            "nconv_layers_survey_byte_model": roof.get("layers")}
        td = roof.get("tensor_dominant")
        if td:
            pk = PEAKS.get("bf16_tflops_sustained") or 1389.3
            ach = td["flop_per_step"] / td["us_per_step"] / 1e6
            line["roofline"]["tensor"] = {
                "bound": "tensor", "kernel": td["name"], "achieved": round(ach, 1), "peak": pk, "unit": "TFLOP/s", "frac": round(ach / pk, 4),
                "us_per_step": round(td["us_per_step"], 1),
                "basis": "FLOPs of the launched problem (2*M*9*Cin_padded*Cout per launch, one per product: the second MMA of bf16x3 is not "
                         "counted) over the CUDA-event time of the step-dominant kernel; peak = sustained bf16 (MEASURED_PEAKS.json)"}
        if roof.get("tensor"):
            alg = ged_algorithmic_flops(B)
            mult = 2 if args.precision == "bf16x3" else 1
            line["roofline"]["tensor_kernels"] = {
                "rows": roof["tensor"], "mma_kernels_us_per_step": round(roof["mma_us"], 1),
                "dense_path_us_per_step": round(roof["dense_us"], 1),
                "algorithmic_gflop_per_step": round(alg / 1e9, 1),
                "algorithmic_tflops_over_mma_kernels": round(alg / roof["mma_us"] / 1e6, 1),
                "peak_tflops_bf16_sustained": PEAKS.get("bf16_tflops_sustained"), "mma_per_product": mult,
                "note": "bf16x3 executes 2 MMAs per product (hi and lo activations against concatenated [W_hi | W_lo]); padded channels and the "
                        "zero taps of the quad forms are executed but not counted as algorithmic FLOPs"}
    if world == 1 and not args.no_extras:
        line["extra"] = {"gms_inference_b32": gms_inference(dev)}
        try:
            line["extra"]["batch_sweep_headline_workload"] = batch_sweep(dev, args.precision)
        except Exception as e:
            line["extra"]["batch_sweep_headline_workload"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        try:
            line["extra"]["full_framework_dc1_cf1_ln1_480x752"] = full_framework_training(dev)
        except Exception as e:       # an extra measurement must never cost the headline line
            line["extra"]["full_framework_dc1_cf1_ln1_480x752"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        try:
            line["extra"]["gudepthcompnet18_dc1_cf1_ln0_240x320"] = gudepthcompnet_training(dev)
        except Exception as e:
            line["extra"]["gudepthcompnet18_dc1_cf1_ln0_240x320"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if world == 1 and not args.no_extras:
        try:
            line["extra"]["baseline_nets_dc1_only_train_sweep"] = baseline_nets_sweep(dev)
        except Exception as e:
            line["extra"]["baseline_nets_dc1_only_train_sweep"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        try:
            line["extra"]["reference_on_gpu_cudnn"] = reference_on_gpu(dev, B)
        except Exception as e:
            line["extra"]["reference_on_gpu_cudnn"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline_subprocess(args.cpu_frames or 8)
        if cb is not None:
            line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    _finish(world, trainer)


def reference_on_gpu(dev, B):
    """The "kernel to beat" of BASELINE.md 3: the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) on this B200 through
    PyTorch/cuDNN with the reference's own settings (cudnn.benchmark = True, trainer.py:9; TF32 convolutions allowed = PyTorch's
    default), same workloads, same batch sizes, eager launches as trainer.py runs them.  Never on the product path."""
    import contextlib
    import io
    import torch
    if _reference_root() is None:
        return {"unavailable": "reference not staged on this box (oracle/make_ref.py)"}
    ns = _load_reference()
    torch.backends.cudnn.benchmark = True
    q = lambda: contextlib.redirect_stdout(io.StringIO())
    out = {"settings": "unmodified reference modules, cudnn.benchmark=True, TF32 convolutions allowed (PyTorch default), eager"}

    def frames(b, hh, ww, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        x = torch.rand(b, 5, hh, ww, device=dev, generator=g)
        gt = torch.rand(b, 1, hh, ww, device=dev, generator=g) * 9.0 + 1.0
        mask = (torch.rand(b, 1, hh, ww, device=dev, generator=g) < 500.0 / (hh * ww)).float()
        x[:, 3:4] = gt * mask
        x[:, 4:5] = mask
        return x, gt

    def train_case(args, dual, b, hh, ww, steps):
        torch.manual_seed(0)
        f = ns.cdf.ConfidenceDepthFrameworkFactory()
        with q():
            m = f.create_model(*args)
            m, params = f.to_device(m)
            m.train()
        crit, _ = f.create_loss("l2", dual, 0.5)
        opt = torch.optim.Adam(params, lr=1e-5)
        x, gt = frames(b, hh, ww, 21)

        def step():
            with q():
                pred = m(x)
            loss = crit(x, pred[0][:, 0:1], pred[2][:, 0:1], gt, 0) if dual else crit(x, pred[0][:, 0:1], gt, 0)
            opt.zero_grad()
            loss.backward()
            opt.step()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del m, opt
        torch.cuda.empty_cache()
        return {"batch": b, "ms_per_step": round(ms, 2), "frames_s": round(b / ms * 1e3, 1)}

    out["ged_dc1_only_228x304_train"] = train_case(("rgbd", "dc1_only", "ged_depthcompnet", ""), False, B, H, W, 5)
    out["gudepthcompnet18_dc1_cf1_ln0_240x320_train"] = train_case(
        ("rgbd", "dc1-cf1-ln0", "gudepthcompnet18", "", "cbr3-c1", "", "ged_depthcompnet", ""), True, 16, 240, 320, 3)
    out["full_framework_dc1_cf1_ln1_480x752_train"] = train_case(
        ("rgbdw", "dc1-cf1-ln1", "ged_depthcompnet", "", "cbr3-c1", "", "ged_depthcompnet", ""), True, 8, 480, 752, 3)
    out["resnet18_dc1_only_228x304_train"] = train_case(("rgbd", "dc1_only", "resnet18", ""), False, 64, H, W, 3)
    out["udepthcompnet18_dc1_only_240x320_train"] = train_case(("rgbdw", "dc1_only", "udepthcompnet18", ""), False, 32, 240, 320, 3)
    torch.manual_seed(0)
    f = ns.cdf.ConfidenceDepthFrameworkFactory()
    with q():
        m = f.create_model("rgbd", "dc1_only", "gms_depthcompnet", "")
        m, _ = f.to_device(m)
        m.eval()
    x, _ = frames(32, H, W, 22)
    with torch.no_grad():
        for _ in range(3):
            with q():
                m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            with q():
                m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["gms_inference_228x304"] = {"batch": 32, "ms_per_batch": round(ms, 2), "frames_s": round(32 / ms * 1e3, 1)}
    return out


def gms_inference(dev, B=32):
    """BASELINE.json configs[1]: gms_depthcompnet inference, batch 32, 228x304, one GPU (49.3 GFLOP of 3x3 convolutions per
    frame).  Three arithmetic modes of the same drop-in module: library fp32, and the hand-written tcgen05 implicit-GEMM
    kernels in bf16x3 (hi/lo split operands, fp32-grade results: the 1e-4 parity mode) and plain bf16 (the 1e-2 mode)."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense
    torch.manual_seed(0)
    f = mc.ConfidenceDepthFrameworkFactory()
    with contextlib.redirect_stdout(io.StringIO()):
        m = f.create_model("rgbd", "dc1_only", "gms_depthcompnet", "")
        m, _ = f.to_device(m)
        m.train(False)
    g = torch.Generator(device=dev).manual_seed(7)
    xs = []
    for _ in range(3):                                     # 3 x 44 MB inputs; activations (GBs per pass) flush the L2 anyway
        x = torch.rand(B, 5, H, W, device=dev, generator=g)
        mask = (torch.rand(B, 1, H, W, device=dev, generator=g) < 0.0072).float()
        x[:, 3:4] = (x[:, 3:4] * 9.0 + 1.0) * mask
        x[:, 4:5] = mask
        xs.append(x)
    out = {"flop_per_frame": 49.3e9}
    prev = dense.PRECISION
    try:
        for prec in ("fp32", "bf16x3", "bf16"):
            dense.set_precision(prec)
            with torch.no_grad():
                for i in range(3):
                    m(xs[i % 3])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(10):
                    m(xs[i % 3])
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[prec] = {"ms_per_batch": round(ms, 3), "frames_s": round(B / ms * 1e3, 1), "tflops": round(49.3e9 * B / ms / 1e9, 1),
                         "dense_path": "library (cuDNN fp32)" if prec == "fp32" else "tcgen05 implicit GEMM (libadc_b200)"}
    finally:
        dense.set_precision(prev)
    return out


def full_framework_training(dev, B=8, HH=480, WW=752, steps=8):
    """BASELINE.json configs[3]: ged_depthcompnet dc + cbr3-c1 confidence head + ged_depthcompnet loss net, training mode
    dc1-cf1-ln1, DualLoss(l2, 0.5), rgbdw frames of the aerial shape 480x752, one GPU of the box.  Every convolution of the three
    nets (NConv pyramids, both encoder-decoders, the head's 3x3) and the head's BN/LeakyReLU/1x1/sigmoid run on the library's
    kernels (bf16x3 mode); one CUDA graph per step."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense, parallel, _lib
    torch.manual_seed(0)
    f = mc.ConfidenceDepthFrameworkFactory()
    prev = dense.PRECISION
    try:
        dense.set_precision("bf16x3")
        with contextlib.redirect_stdout(io.StringIO()):
            m = f.create_model("rgbdw", "dc1-cf1-ln1", "ged_depthcompnet", "", "cbr3-c1", "", "ged_depthcompnet", "")
            m, params = f.to_device(m)
            m.train()
        crit, _ = f.create_loss("l2", True, 0.5)
        from aerial_depth_completion_b200.optim import FlatAdam
        opt = FlatAdam(params, lr=1e-5)
        tr = parallel.DataParallelTrainer(m, crit, opt, params, world_size=1)
        g = torch.Generator(device=dev).manual_seed(11)
        xs, ts = [], []
        for _ in range(3):
            x = torch.rand(B, 5, HH, WW, device=dev, generator=g)
            gt = torch.rand(B, 1, HH, WW, device=dev, generator=g) * 9.0 + 1.0
            mask = (torch.rand(B, 1, HH, WW, device=dev, generator=g) < 500.0 / (HH * WW)).float()
            x[:, 3:4] = gt * mask
            x[:, 4:5] = mask
            xs.append(x)
            ts.append(gt)
        l0 = _lib.lib.adc_launch_count()
        tr._eager_step(xs[0], ts[0])
        launches = _lib.lib.adc_launch_count() - l0
        tr.capture(xs[0], ts[0])
        for i in range(3):
            tr.step(xs[i % 3], ts[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = tr.step(xs[i % 3], ts[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"workload": "ged_depthcompnet + cbr3-c1 + ged_depthcompnet loss net, dc1-cf1-ln1, DualLoss(l2, 0.5), rgbdw %dx%d" % (HH, WW),
                "batch": B, "ms_per_step": round(ms, 3), "frames_s": round(B / ms * 1e3, 1), "library_launches_per_step": int(launches),
                "trainable_parameters": int(sum(p.numel() for p in params)), "loss": round(float(loss), 5), "precision": "bf16x3",
                "launch": "one CUDA graph per step"}
    finally:
        dense.set_precision(prev)


def gudepthcompnet_training(dev, B=16, HH=240, WW=320, steps=5, graph=True):
    """BASELINE.json configs[2]: gudepthcompnet18 + cbr3-c1 confidence head + frozen ged_depthcompnet loss net, training mode
    dc1-cf1-ln0, DualLoss(l2, 0.5), batch 16 per GPU at 240x320 (the net needs H, W = 0 mod 16: SURVEY.md 0.4).  Every layer of the three
    nets runs on this library's kernels in the bf16x3 mode: the ResNet-18 U-Net body (tc_bn.S2DUBodyTC: tcgen05 convolutions and
    transposed convolutions, BatchNorm / residual / activation passes), the 129-channel head (tc_bn.ConfHeadBN), the NConv pyramids, the
    loss net (tc_train.GEDDenseTC), the losses and the optimizer; one CUDA graph per step."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense, parallel, _lib
    from aerial_depth_completion_b200.optim import FlatAdam
    torch.manual_seed(0)
    f = mc.ConfidenceDepthFrameworkFactory()
    prev = dense.PRECISION
    try:
        dense.set_precision("bf16x3")
        with contextlib.redirect_stdout(io.StringIO()):
            m = f.create_model("rgbd", "dc1-cf1-ln0", "gudepthcompnet18", "", "cbr3-c1", "", "ged_depthcompnet", "")
            m, params = f.to_device(m)
            m.train()
        crit, _ = f.create_loss("l2", True, 0.5)
        tr = parallel.DataParallelTrainer(m, crit, FlatAdam(params, lr=1e-5), params, world_size=1)
        g = torch.Generator(device=dev).manual_seed(13)
        xs, ts = [], []
        for _ in range(2):
            x = torch.rand(B, 5, HH, WW, device=dev, generator=g)
            gt = torch.rand(B, 1, HH, WW, device=dev, generator=g) * 9.0 + 1.0
            mask = (torch.rand(B, 1, HH, WW, device=dev, generator=g) < 500.0 / (HH * WW)).float()
            x[:, 3:4] = gt * mask
            x[:, 4:5] = mask
            xs.append(x)
            ts.append(gt)
        l0 = _lib.lib.adc_launch_count()
        tr._eager_step(xs[0], ts[0])
        launches = _lib.lib.adc_launch_count() - l0
        if graph:
            tr.capture(xs[0], ts[0], warmup=1)
        for i in range(2):
            tr.step(xs[i % 2], ts[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = tr.step(xs[i % 2], ts[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flop = 3.0 * (105.0e9 + 23.0e9) * B + 2.0 * 2.95e9 * B       # trained body + head (fwd, dgrad, wgrad) + frozen loss net (fwd, dgrad)
        return {"workload": "gudepthcompnet18 + cbr3-c1 + frozen ged_depthcompnet loss net, dc1-cf1-ln0, DualLoss(l2, 0.5), rgbd %dx%d" % (HH, WW),
                "batch": B, "ms_per_step": round(ms, 2), "frames_s": round(B / ms * 1e3, 1), "library_launches_per_step": int(launches),
                "algorithmic_tflops": round(flop / ms / 1e9, 1),
                "trainable_parameters": int(sum(p.numel() for p in params)), "loss": round(float(loss), 5), "precision": "bf16x3",
                "dense_path": "every convolution / transposed convolution / BatchNorm / activation of the ResNet-18 U-Net and the 129-channel "
                              "head: libadc_b200 (tcgen05 implicit GEMM + fused BN passes); no cuDNN kernel in the step",
                "launch": "one CUDA graph per step" if graph else "eager"}
    finally:
        dense.set_precision(prev)


def batch_sweep(dev, precision, batches=(8, 32, 64)):
    """The headline workload (GED dc1_only train step, resident inputs, one CUDA graph per step) at other batch sizes per GPU: 8 is the
    reference's --batch-size default (trainer.py:101).  Small batches leave the 148 SMs partly idle in the pyramid's kernels (1280 tiles
    at B = 32 are 1.7 waves of the resident CTAs) and pay the launch latency of the small pyramid levels once per step."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense, parallel
    from aerial_depth_completion_b200.optim import FlatAdam
    out = {}
    for B in batches:
        torch.manual_seed(0)
        f = mc.ConfidenceDepthFrameworkFactory()
        with contextlib.redirect_stdout(io.StringIO()):
            m = f.create_model("rgbd", "dc1_only", "ged_depthcompnet", "")
            m, params = f.to_device(m)
            m.train()
        crit, _ = f.create_loss("l2", False, 0.5)
        tr = parallel.DataParallelTrainer(m, crit, FlatAdam(params, lr=1e-5), params, world_size=1)
        g = torch.Generator(device=dev).manual_seed(19)
        xs, ts = [], []
        for _ in range(max(2, min(8, (256 << 20) // (B * 6 * H * W * 4) + 1))):
            x = torch.rand(B, 5, H, W, device=dev, generator=g)
            gt = torch.rand(B, 1, H, W, device=dev, generator=g) * 9.0 + 1.0
            mask = (torch.rand(B, 1, H, W, device=dev, generator=g) < 500.0 / (H * W)).float()
            x[:, 3:4] = gt * mask
            x[:, 4:5] = mask
            xs.append(x)
            ts.append(gt)
        tr._eager_step(xs[0], ts[0])
        tr.capture(xs[0], ts[0])
        for i in range(3):
            tr.step(xs[i % len(xs)], ts[i % len(xs)])
        torch.cuda.synchronize()
        steps = 12
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.step(xs[i % len(xs)], ts[i % len(xs)])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["b%d" % B] = {"ms_per_step": round(ms, 3), "frames_s": round(B / ms * 1e3, 1)}
        del tr, m, params, xs, ts
        torch.cuda.empty_cache()
    return out


def baseline_nets_sweep(dev):
    """BASELINE.json configs[4]: the sparse-to-dense baseline nets -- resnet18 (S2DResNet, 228x304) and udepthcompnet18 (S2DUResNet
    on the NConv pyramid's output, 240x320: the net needs H, W = 0 mod 16) -- dc1_only, masked-L2, forward + backward + Adam, batch sweep up
    to 64 frames.  Every layer on this library's kernels (tc_s2d.S2DBodyTC / tc_bn.S2DUBodyTC, bf16x3), one CUDA graph per step."""
    import contextlib
    import io
    import torch
    import aerial_depth_completion_b200.model_zoo.confidence_depth_framework as mc
    from aerial_depth_completion_b200 import dense, parallel, _lib
    from aerial_depth_completion_b200.optim import FlatAdam
    out = {}
    prev = dense.PRECISION
    try:
        dense.set_precision("bf16x3")
        for arch, itype, HH, WW, gflop, batches in (("resnet18", "rgbd", 228, 304, 10.0, (16, 32, 64)),
                                                    ("udepthcompnet18", "rgbdw", 240, 320, 105.0, (16, 32, 64))):
            rows = {}
            for B in batches:
                try:
                    torch.manual_seed(0)
                    f = mc.ConfidenceDepthFrameworkFactory()
                    with contextlib.redirect_stdout(io.StringIO()):
                        m = f.create_model(itype, "dc1_only", arch, "")
                        m, params = f.to_device(m)
                        m.train()
                    crit, _ = f.create_loss("l2", False, 0.5)
                    tr = parallel.DataParallelTrainer(m, crit, FlatAdam(params, lr=1e-5), params, world_size=1)
                    g = torch.Generator(device=dev).manual_seed(17)
                    x = torch.rand(B, 5, HH, WW, device=dev, generator=g)
                    gt = torch.rand(B, 1, HH, WW, device=dev, generator=g) * 9.0 + 1.0
                    mask = (torch.rand(B, 1, HH, WW, device=dev, generator=g) < 500.0 / (HH * WW)).float()
                    x[:, 3:4] = gt * mask
                    x[:, 4:5] = mask
                    l0 = _lib.lib.adc_launch_count()
                    tr._eager_step(x, gt)
                    launches = _lib.lib.adc_launch_count() - l0
                    tr.capture(x, gt, warmup=1)
                    for _ in range(2):
                        tr.step(x, gt)
                    torch.cuda.synchronize()
                    steps = 4
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        loss = tr.step(x, gt)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    rows["b%d" % B] = {"ms_per_step": round(ms, 2), "frames_s": round(B / ms * 1e3, 1), "library_launches_per_step": int(launches),
                                       "algorithmic_tflops": round(3.0 * gflop * B / ms, 1), "loss": round(float(loss), 4)}
                    del tr, m, params, x, gt, mask
                    torch.cuda.empty_cache()
                except Exception as e:
                    rows["b%d" % B] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
                    torch.cuda.empty_cache()
            out[arch + "_%dx%d" % (HH, WW)] = rows
    finally:
        dense.set_precision(prev)
    return out


def ged_algorithmic_flops(B):
    """FLOPs per training step of GEDNet's dense layers conv1..conv10 (2*N*Ho*Wo*9*Cin*Cout per convolution, real channel counts:
    no padding, no quad-form zero taps, no bf16x3 double counting; confidence_depth_framework.py:442-452): forward + data gradient
    + weight gradient = 3x the forward, minus conv1's data gradient for the three rgb channels nobody needs."""
    hs, ws = [H], [W]
    for _ in range(5):
        hs.append((hs[-1] - 1) // 2 + 1)
        ws.append((ws[-1] - 1) // 2 + 1)
    f = lambda lvl, ci, co: 2.0 * B * hs[lvl] * ws[lvl] * 9 * ci * co
    fwd = [f(1, 5, 80), f(2, 80, 80), f(3, 80, 80), f(4, 80, 80), f(5, 80, 80), f(4, 160, 64), f(3, 144, 64), f(2, 144, 32), f(1, 112, 32),
           f(0, 33, 1)]
    return 3.0 * sum(fwd) - f(1, 3, 80)


def kernel_roofline(trainer, dev_in, dev_tg, pool_n, B, detail_path=""):
    """Average launch duration (CUDA events on the launching stream, recorded by the library around every kernel) of every
    kernel of the library inside the training step; algorithmic HBM bytes for the NConv kernels, algorithmic FLOPs for the
    tensor-core kernels."""
    import ctypes
    import torch
    from aerial_depth_completion_b200._lib import lib
    names = {0: "nconv_fwd", 1: "nconv_bwd_data", 2: "nconv_bwd_weight", 3: "pool_fwd", 4: "pool_bwd", 5: "upsample_bwd",
             6: "head_fwd", 7: "head_bwd"}
    tc_names = {20: "tc_conv_box(stride-2 fprop)", 21: "tc_conv_rowseg(fprop, 1x1, data gradients incl. quad forms)", 22: "tc_dgrad_s2_parity",
                23: "tc_wgrad_flat(stride-1, 1x1)", 24: "tc_wgrad_box(stride-2)", 25: "tc_glue(upsample, grad assemble + bias grad)",
                26: "tc_layout(im2col pack, unshuffle)", 27: "tc_prep_weights", 28: "conv10 quad head (direct kernels)",
                29: "loss + optimizer (masked loss fwd/bwd, flat Adam)", 30: "confidence head (BN/LeakyReLU/1x1/sigmoid passes)"}
    mma_kinds = (20, 21, 22, 23, 24)
    calls = ("nconv1@L0", "nconv2@L0", "nconv3@L0", "nconv2@L1", "nconv3@L1", "nconv2@L2", "nconv2@L3", "nconv4@L2", "nconv5@L1",
             "nconv6@L0", "nconv7@L0")
    agg, tc = {}, {}
    reps = 3
    cap = 1024
    from aerial_depth_completion_b200 import tc_train
    side_prev = tc_train._SIDE_STREAM
    tc_train._SIDE_STREAM = False       # one stream while profiling: an event pair must time its kernel alone (the library does the
    for i in range(reps):               # same for the pyramid's weight-gradient stream when profiling is enabled)
        lib.adc_profile_enable(1)
        try:
            trainer._eager_step(dev_in[i % pool_n], dev_tg[i % pool_n])
        finally:
            tc_train._SIDE_STREAM = side_prev if i == reps - 1 else False
        torch.cuda.synchronize()
        n = ctypes.c_int(0)
        kinds = (ctypes.c_int * cap)()
        cl = (ctypes.c_int * cap)()
        ms = (ctypes.c_float * cap)()
        by = (ctypes.c_double * cap)()
        lib.adc_profile_read(cap, ctypes.byref(n), kinds, cl, ms, by)
        lib.adc_profile_enable(0)
        if detail_path and i == reps - 1:
            with open(detail_path, "w") as fh:
                fh.write("# one training step, launches in issue order: kind, name, work (bytes or FLOPs as the kernel sees them), us\n")
                for j in range(n.value):
                    nm = tc_names.get(kinds[j]) or names.get(kinds[j]) or {8: "pack", 9: "reduce", 10: "prep"}.get(kinds[j], "?")
                    fh.write("%d,\"%s\",%.4g,%.2f\n" % (kinds[j], nm, by[j], ms[j] * 1e3))
        for j in range(n.value):
            if kinds[j] in tc_names:
                a = tc.setdefault(tc_names[kinds[j]], [0.0, 0.0, 0])
                a[0] += ms[j] * 1e3
                a[1] += by[j]
                a[2] += 1
                continue
            if kinds[j] not in names or by[j] <= 0:
                continue
            key = names[kinds[j]] + (":" + calls[cl[j]] if kinds[j] < 3 and 0 <= cl[j] < 11 else "")
            a = agg.setdefault(key, [0.0, 0.0, 0])
            a[0] += ms[j] * 1e3
            a[1] += by[j]
            a[2] += 1
    if not agg:
        return None
    table = []
    for k, (us, byts, cnt) in agg.items():
        table.append({"name": k, "us": round(us / cnt, 2), "mb": round(byts / cnt / 1e6, 2), "gbs": round(byts / us / 1e3, 1),
                      "launches_per_step": cnt // reps, "us_per_step": round(us / reps, 2)})
    table.sort(key=lambda r: -r["us_per_step"])
    # the headline kernel is the NConv2d layer kernel (forward / data gradient / weight gradient of one call) with the largest share of
    # the step; pooling / up-sampling rows aggregate three pyramid levels of very different sizes and stay in the table only
    d = next((r for r in table if r["name"].startswith("nconv_")), table[0])
    out = {"dominant": {"name": d["name"], "us": d["us"], "mb": d["mb"], "gbs": d["gbs"]}, "table": table[:12]}
    # per LAYER on SURVEY.md 8(d)'s byte model: forward 8*(Ci+Co) B/px, backward (data + weight gradient kernels together)
    # 16*(Ci+Co) B/px -- no credit for the split of the backward into two kernels that both read x, c and G
    chan = {"nconv1": (1, 2), "nconv2": (2, 2), "nconv3": (2, 2), "nconv4": (4, 2), "nconv5": (4, 2), "nconv6": (4, 2)}
    px = [B * 228 * 304, B * 114 * 152, B * 57 * 76, B * 28 * 38]
    hbm = peaks()[0]
    layers = []
    for call in calls[:10]:
        mod, lvl = call.split("@L")
        ci, co = chan[mod]
        f, bd, bw = agg.get("nconv_fwd:" + call), agg.get("nconv_bwd_data:" + call), agg.get("nconv_bwd_weight:" + call)
        if not f:
            continue
        row = {"layer": call, "fwd_us": round(f[0] / f[2], 2),
               "fwd_frac": round(8.0 * (ci + co) * px[int(lvl)] / (f[0] / f[2]) / 1e3 / hbm, 3)}
        if bw:
            bus = bw[0] / bw[2] + (bd[0] / bd[2] if bd else 0.0)      # nconv1 has no data gradient (x0, c0 need none)
            row["bwd_us"] = round(bus, 2)
            row["bwd_frac"] = round(16.0 * (ci + co) * px[int(lvl)] / bus / 1e3 / hbm, 3)
        layers.append(row)
    out["layers"] = layers
    if tc:
        rows = []
        for k, (us, work, cnt) in tc.items():
            rows.append({"name": k, "launches_per_step": cnt // reps, "us_per_step": round(us / reps, 1)})
        rows.sort(key=lambda r: -r["us_per_step"])
        out["tensor"] = rows
        top = max(tc.items(), key=lambda kv: kv[1][0])
        if top[1][1] > 0 and top[0] in [tc_names[q] for q in mma_kinds]:
            out["tensor_dominant"] = {"name": top[0], "us_per_step": top[1][0] / reps, "flop_per_step": top[1][1] / reps}
        out["mma_us"] = sum(us for k, (us, work, cnt) in tc.items() if k in [tc_names[q] for q in mma_kinds]) / reps
        out["dense_us"] = sum(us for us, work, cnt in tc.values()) / reps
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""      # host cores only (set before torch initialises CUDA)
        reference_arm(a)
    else:
        b200_arm(a)

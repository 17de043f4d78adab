"""Data-parallel training step: one process per GPU, gradients of all parameters live in ONE flat fp32 buffer that is
all-reduced over NCCL (NVLink 5 / NVSwitch), bucket by bucket WHILE backward is still producing the upstream gradients.

Replaces the reference's single-process ``nn.DataParallel`` (confidence_depth_framework.py:145-159, which re-broadcasts
every parameter and gathers every output each step) with the torch.distributed layout of SURVEY.md 8e: frames are
independent, the only coupling is the gradient sum (1.9 MB for GED .. 65 MB for gudepthcompnet18); BatchNorm statistics
stay per replica like under DataParallel.  The step is sync-free on the host (the loss is returned as a device tensor).
"""
import os

import torch
import torch.distributed as dist


class DataParallelTrainer:
    def __init__(self, model, criterion, optimizer, params, world_size=1, process_group=None):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.world, self.group = int(world_size), process_group
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        own = getattr(optimizer, "flat_grad", None)
        if own is not None and [id(p) for p in getattr(optimizer, "_params", [])] == [id(p) for p in self.params]:
            self.flat = own          # optim.FlatAdam already keeps every .grad as a view into one bucket: all-reduce that one
        else:
            self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=self.params[0].dtype)
            off = 0
            for p in self.params:   # .grad of every parameter is a view into the flat buffer: backward fills the bucket
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        # One process, no bucket hooks: autograd's AccumulateGrad ADDS every incoming gradient into the (zeroed) flat views -- one
        # library launch per parameter (35 for GED) plus the memset.  With .grad = None it adopts the incoming tensors instead (no
        # launch), one multi-tensor copy moves them into the flat bucket, and the views are put back for the optimizer.
        self._views = [p.grad for p in self.params]
        self._adopt = self.world == 1 and os.environ.get("ADC_ADOPT_GRADS", "1") != "0"
        if hasattr(criterion, "lazy"):
            criterion.lazy = True   # keep loss bookkeeping on the device (no .item() per step)
        for sub in ("net_a", "net_b"):
            if hasattr(criterion, sub) and hasattr(getattr(criterion, sub), "lazy"):
                getattr(criterion, sub).lazy = True
        self._buckets, self._left, self._works = [], [], []
        self._scale = self._scale_work = None
        if self.world > 1:
            for p in list(model.parameters()) + list(model.buffers()):   # identical replicas, once (not per step)
                dist.broadcast(p.data, 0, group=self.group)
            self._avg = dist.get_backend(self.group) == "nccl"           # NCCL averages in the collective; gloo sums, then one div
            self._make_buckets()

    def _make_buckets(self):
        """Gradient buckets = contiguous slices of the flat buffer, one per sub-net part in parameter order (dc_model.d_net |
        dc_model body | conf_model | loss_dc_model.d_net | loss_dc_model body).  Backward produces them in the reverse order; a
        bucket's all-reduce is launched (async, NCCL's own stream) by the hook of its last gradient, so the exchange of the loss
        net's and the encoder-decoder's gradients overlaps the backward of everything upstream of them."""
        names = {id(p): n for n, p in self.model.named_parameters()}
        off, cur = 0, None
        for p in self.params:
            parts = names.get(id(p), "").split(".")
            key = (parts[0], parts[1] if len(parts) > 1 and parts[1] in ("d_net", "unguidednet") else "body")
            if key != cur:
                self._buckets.append([off, off, 0])
                cur = key
            b = len(self._buckets) - 1
            self._buckets[b][1] = off + p.numel()
            self._buckets[b][2] += 1
            p.register_post_accumulate_grad_hook(lambda _p, b=b: self._on_grad(b))
            off += p.numel()
        self._left = [0] * len(self._buckets)

    def _reduce_bucket(self, b, async_op):
        lo, hi, _ = self._buckets[b]
        if self._scale_work is not None:
            # Global masked mean (trainer.py:204-208 computes ONE loss over the gathered outputs of all replicas): this rank's
            # gradients were produced for its local mean S_r / n_r; the global one is sum_r S_r / sum_r n_r, so they are weighted by
            # n_r * world / sum n (exactly 1.0 when every rank holds the same number of valid pixels) before the average.
            self._scale_work.wait()
            self._scale_work = None
            self._scale = self._cnt * float(self.world) / self._stats[0]
        if self._scale is not None:
            self.flat[lo:hi].mul_(self._scale)
        return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group,
                               async_op=async_op)

    def _on_grad(self, b):
        if not self._left or self._left[b] <= 0:
            return
        self._left[b] -= 1
        if self._left[b] == 0:
            self._works.append(self._reduce_bucket(b, True))

    def forward_loss(self, inp, target):
        pred = self.model(inp)
        target_depth = target[:, 0:1, :, :]
        if pred[2] is not None:
            return self.criterion(inp, pred[0][:, 0:1], pred[2][:, 0:1], target_depth), pred
        return self.criterion(inp, pred[0][:, 0:1], target_depth), pred

    def _valid_count(self, target):
        """Number of valid target pixels of this rank's batch as a device scalar (the product's losses leave it in ``num_valid``)."""
        for c in (self.criterion, getattr(self.criterion, "net_a", None)):
            n = getattr(c, "num_valid", None)
            if n is not None:
                return n.detach().float().reshape(())
        return (target[:, 0:1] > 0).sum().float()

    def capture(self, inp, target, warmup=3):
        """Record the whole training step (forward, loss, backward, all-reduce, optimizer) into ONE CUDA graph.
        The step has static shapes and no host synchronisation, so every later ``step`` is a single graph launch: the
        ~450 kernel launches and the Python between them (4.4 ms of a 15 ms step at batch 32) disappear.
        Requires a capturable optimizer (optim.FlatAdam, or torch.optim.Adam(capturable=True)).  Inputs of later steps are copied
        into the static buffers; optim.FlatAdam reads lr / betas / weight_decay from a device buffer that ``step`` refreshes from
        ``param_groups`` before every replay, so learning-rate schedulers keep working (trainer.py:236-241)."""
        self._static_in, self._static_tg = inp.clone(), target.clone()
        if hasattr(self.optimizer, "refresh_hyper"):
            self.optimizer.refresh_hyper()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(self._static_in, self._static_tg)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._eager_step(self._static_in, self._static_tg)
        return self

    def release(self):
        """Drop the captured step graph and its static buffers (later steps run eagerly again).  A graph that recorded NCCL
        collectives keeps the communicator's work alive; release it -- and synchronise -- before ``dist.destroy_process_group()``."""
        self._graph = None
        self._static_in = self._static_tg = self._static_loss = None
        torch.cuda.synchronize()

    def step(self, inp, target):
        """forward + loss + zero_grad + backward + gradient all-reduce + optimizer step (trainer.py:202-216)."""
        if getattr(self, "_graph", None) is not None:
            self._static_in.copy_(inp, non_blocking=True)
            self._static_tg.copy_(target, non_blocking=True)
            if hasattr(self.optimizer, "refresh_hyper"):
                self.optimizer.refresh_hyper()     # a scheduler may have changed lr since the last replay
            self._graph.replay()
            for p in self.params:      # the replay updated the parameters behind autograd's back: caches keyed on the version
                torch.autograd.graph.increment_version(p)     # counter (dense.cached_weights) must see it
            return self._static_loss
        return self._eager_step(inp, target)

    def _eager_step(self, inp, target):
        loss, _ = self.forward_loss(inp, target)
        out = loss.detach()
        if self._adopt:
            for p in self.params:
                p.grad = None
            loss.backward()
            dst, src = [], []
            for p, v in zip(self.params, self._views):
                if p.grad is None:
                    v.zero_()                 # a parameter outside this step's graph: zero gradient, as after zero_grad
                else:
                    dst.append(v)
                    src.append(p.grad)
            if dst:
                torch._foreach_copy_(dst, src)
            for p, v in zip(self.params, self._views):
                p.grad = v
            if hasattr(self.optimizer, "guard"):
                self.optimizer.guard = out
            self.optimizer.step()
            return out
        self.flat.zero_()
        self._scale = self._scale_work = None
        if self.world > 1:
            self._left = [n for _, _, n in self._buckets]
            self._works = []
            # [n_r, S_r = n_r * loss_r] summed over the ranks while backward runs: the global loss, the per-rank gradient weight
            # and the collective "is the loss finite" verdict (a NaN on any rank reaches every rank)
            self._cnt = self._valid_count(target)
            self._stats = torch.stack((self._cnt, out * self._cnt))
            self._scale_work = dist.all_reduce(self._stats, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        loss.backward()
        if self.world > 1:
            for b, left in enumerate(self._left):      # a bucket with a parameter that got no gradient this step: reduce it now
                if left > 0:
                    self._reduce_bucket(b, False)
            self._left = [0] * len(self._buckets)
            for w in self._works:
                w.wait()                                # the compute stream waits for the collectives (no host synchronisation)
            self._works = []
            if not self._avg:
                self.flat.div_(self.world)
            out = self._stats[1] / self._stats[0]       # the loss of the global batch, identical on every rank
        if hasattr(self.optimizer, "guard"):
            self.optimizer.guard = out                  # non-finite loss (any rank): the update is skipped on the device
        self.optimizer.step()
        return out


class HostFeeder:
    """Feeds a ``DataParallelTrainer`` from pinned HOST batches the way a loader thread would (trainer.py:194-233 copies each batch
    with ``.cuda()`` and reads ``loss.item()`` before it touches the next one: copy, step and read-back are serialised).  Here the
    host->device copy of batch i+1 runs on a copy stream while the GPU computes step i (two device buffer pairs, events both ways),
    and the loss of step i is copied to pinned host memory asynchronously and read one step later, so the only thing the host waits
    for is a step that has already finished.  Every batch is still copied and every loss is still read."""

    def __init__(self, trainer, example_in, example_tg):
        dev = trainer.params[0].device
        self.trainer = trainer
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.buf = [(torch.empty_like(example_in, device=dev), torch.empty_like(example_tg, device=dev)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # H2D copy into buffer b finished
        self.free = [torch.cuda.Event() for _ in range(2)]       # the step that read buffer b finished
        self.host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_done = [torch.cuda.Event() for _ in range(2)]

    def _prefetch(self, b, host_in, host_tg, first_use):
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.free[b])
            self.buf[b][0].copy_(host_in, non_blocking=True)
            self.buf[b][1].copy_(host_tg, non_blocking=True)
            self.ready[b].record(self.copy_stream)

    def run(self, batches):
        """batches: sequence of (pinned host input, pinned host target).  Returns the list of per-step losses (python floats)."""
        main = torch.cuda.current_stream()
        losses, n = [], len(batches)
        if n == 0:
            return losses
        self._prefetch(0, batches[0][0], batches[0][1], True)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                self._prefetch(b ^ 1, batches[i + 1][0], batches[i + 1][1], i == 0)
            main.wait_event(self.ready[b])
            loss = self.trainer.step(self.buf[b][0], self.buf[b][1])
            self.free[b].record(main)
            self.host_loss[b].copy_(loss.detach().reshape(1), non_blocking=True)
            self.loss_done[b].record(main)
            if i > 0:                                            # read the PREVIOUS step's loss: it finished while this one was enqueued
                self.loss_done[b ^ 1].synchronize()
                losses.append(float(self.host_loss[b ^ 1][0]))
        last = (n - 1) & 1
        self.loss_done[last].synchronize()
        losses.append(float(self.host_loss[last][0]))
        return losses


class DeviceAssemblyFeeder(HostFeeder):
    """``HostFeeder`` with the loader's sparse-depth synthesis moved to the device.  A host batch is
    ``(rgbgt, seeds)``: pinned float32 [B, 4, H, W] = (r, g, b, ground-truth depth) and pinned int32 [B] (the per-frame
    ``np.random.seed`` values as 32-bit patterns).  Only these 4 planes cross PCIe (the reference's loader ships 5 input planes + the
    target = 6); on the copy stream, behind the copy, ``adc_sample_assemble`` draws the bit-exact ``UniformSampling`` mask of every frame
    and writes the 5-channel network input, so the assembly of batch i+1 also overlaps step i.  The target is the gt plane."""

    def __init__(self, trainer, example_rgbgt, num_samples, max_depth=float("inf")):
        dev = trainer.params[0].device
        n, _, h, w = example_rgbgt.shape
        super().__init__(trainer, torch.empty(n, 5, h, w), torch.empty(n, 1, h, w))
        self.num_samples, self.max_depth = num_samples, max_depth
        self.raw = [(torch.empty(n, 4, h, w, device=dev), torch.empty(n, dtype=torch.int32, device=dev)) for _ in range(2)]
        self.buf = [(self.buf[i][0], self.raw[i][0][:, 3:4]) for i in range(2)]      # the target is the uploaded gt plane itself

    def _prefetch(self, b, host_rgbgt, host_seeds, first_use):
        from .dataloaders.dense_to_sparse import sample_assemble_cuda
        import numpy as _np
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.free[b])
            raw, seeds = self.raw[b]
            raw.copy_(host_rgbgt, non_blocking=True)
            seeds.copy_(host_seeds, non_blocking=True)
            sample_assemble_cuda(raw, seeds, self.num_samples, _np.inf if self.max_depth == float("inf") else self.max_depth,
                                 out=self.buf[b][0])
            self.ready[b].record(self.copy_stream)

"""Optimizer of the training step on flat buffers.

The reference trainer builds ``torch.optim.Adam(params, lr, weight_decay)`` over the trainable parameters (trainer.py:134-146); its
foreach implementation launches seven multi-tensor kernels of ~27 us each for GEDNet's thirty small tensors.  ``FlatAdam`` has the
same constructor and ``step`` / ``zero_grad`` / ``state_dict`` surface, but re-homes every parameter, its gradient and both moments as
views into four flat fp32 buffers, so the update is ONE elementwise kernel (adc_adam_step) and the gradient buffer is exactly the
bucket the data-parallel trainer all-reduces.  The step counter AND the hyper-parameters (lr, betas, eps, weight_decay) live on the
device: a captured CUDA graph advances the bias correction on replay and follows a learning-rate schedule (``refresh_hyper`` copies
changed ``param_groups`` values before a replay; trainer.py:236-241 decays lr every 5 epochs).  ``guard`` (a device scalar, normally
the step's loss) makes the update a no-op when it is not finite -- the reference trainer skips such batches (trainer.py:206-208).
It subclasses ``torch.optim.Adam`` so the reference's checkpoint tagging ``isinstance(optimizer, torch.optim.Adam)``
(trainer.py:147) keeps saying 'adam'.
"""
import ctypes

import torch

from ._lib import lib, check


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise ValueError("FlatAdam: amsgrad is not supported (the reference trainer does not use it)")
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("FlatAdam takes one flat list of parameters (one group), like the reference trainer passes")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps:
            raise ValueError("FlatAdam: no trainable parameters")
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FlatAdam needs CUDA float32 parameters: the B200 path has no CPU fallback")
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self._params = ps
        self.flat_param = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self._state4 = torch.zeros(4, device=dev, dtype=torch.float32)      # {t, step_size, 1/sqrt(1-b2^t), apply}
        self.step_t = self._state4[0:1]
        self._hyper = torch.zeros(5, device=dev, dtype=torch.float64)       # {lr, beta1, beta2, eps, weight_decay}
        self._hyper_host = torch.zeros(5, dtype=torch.float64).pin_memory() if torch.cuda.is_available() else torch.zeros(5, dtype=torch.float64)
        self._hyper_seen = None
        self.guard = None            # device float scalar: the update is skipped while it is not finite
        self._views, off = [], 0
        with torch.no_grad():
            for p in ps:
                k = p.numel()
                view = self.flat_param[off:off + k].view(p.shape)
                view.copy_(p.data)
                p.data = view                                  # the module's Parameter now lives inside the flat vector
                g = self.flat_grad[off:off + k].view(p.shape)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
                self._views.append((g, off, k))
                self.state[p] = {"step": self.step_t, "exp_avg": self.flat_exp_avg[off:off + k].view(p.shape),
                                 "exp_avg_sq": self.flat_exp_avg_sq[off:off + k].view(p.shape)}
                off += k

    def zero_grad(self, set_to_none=False):
        """Clears the flat gradient buffer; the ``.grad`` views stay in place (set_to_none would break the flat layout)."""
        self.flat_grad.zero_()
        for p, (g, _, _) in zip(self._params, self._views):
            p.grad = g

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for p, (g, _, _) in zip(self._params, self._views):    # gradients assigned from outside are copied back into the bucket
            if p.grad is None:
                g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g
        for p, (_, off, k) in zip(self._params, self._views):   # a Parameter re-homed behind our back would be silently left behind
            if p.data_ptr() != self.flat_param.data_ptr() + 4 * off:
                raise RuntimeError("FlatAdam: a parameter's storage was rebound (p.data = ...) after construction; its updates would be "
                                   "lost.  Write into parameters with .data.copy_() instead")
        if not torch.cuda.is_current_stream_capturing():
            self.refresh_hyper()
        guard = self.guard
        check(lib.adc_adam_step_dev(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.flat_exp_avg.data_ptr(),
                                    self.flat_exp_avg_sq.data_ptr(), self._state4.data_ptr(), self._hyper.data_ptr(),
                                    guard.data_ptr() if guard is not None else None, self.flat_param.numel(),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "adc_adam_step_dev")
        for p in self._params:                                  # the kernel wrote the storage behind torch's back: caches keyed on
            torch.autograd.graph.increment_version(p)           # the version counter (dense.cached_weights) must see the update
        return loss

    def refresh_hyper(self):
        """Copies lr / betas / eps / weight_decay of ``param_groups[0]`` to the device when they changed (a scheduler wrote them).
        Called by ``step`` outside graph capture and by the trainer before every graph replay."""
        grp = self.param_groups[0]
        cur = (float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]))
        if cur != self._hyper_seen:
            self._hyper_host.copy_(torch.tensor(cur, dtype=torch.float64))
            self._hyper.copy_(self._hyper_host, non_blocking=True)
            self._hyper_seen = cur

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam / FlatAdam state dict (per-parameter 'step', 'exp_avg', 'exp_avg_sq'): the moments are copied
        INTO the flat buffers so the views stay valid."""
        groups = state_dict["param_groups"]
        ids = [i for grp in groups for i in grp["params"]]
        mine = self.param_groups[0]["params"]
        if len(ids) != len(mine):
            raise ValueError("loaded state dict has a different number of parameters")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        with torch.no_grad():
            for i, p in zip(ids, mine):
                st = state_dict["state"].get(i)
                if st is None or p not in self.state:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self.step_t.fill_(float(st["step"]))

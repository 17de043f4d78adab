"""Optimizer of the training step on flat buffers.

The reference trainer builds ``torch.optim.Adam(params, lr, weight_decay)`` over the trainable parameters (trainer.py:134-146); its
foreach implementation launches seven multi-tensor kernels of ~27 us each for GEDNet's thirty small tensors.  ``FlatAdam`` has the
same constructor and ``step`` / ``zero_grad`` / ``state_dict`` surface, but re-homes every parameter, its gradient and both moments as
views into four flat fp32 buffers, so the update is ONE elementwise kernel (adc_adam_step) and the gradient buffer is exactly the
bucket the data-parallel trainer all-reduces.  The step counter lives on the device: a captured CUDA graph advances it on replay.
"""
import ctypes

import torch

from ._lib import lib, check


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise ValueError("FlatAdam: amsgrad is not supported (the reference trainer does not use it)")
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("FlatAdam takes one flat list of parameters (one group), like the reference trainer passes")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
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
        self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
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
        grp = self.param_groups[0]
        b1, b2 = grp["betas"]
        check(lib.adc_adam_step(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.flat_exp_avg.data_ptr(),
                                self.flat_exp_avg_sq.data_ptr(), self.step_t.data_ptr(), self.flat_param.numel(), float(grp["lr"]),
                                float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]),
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "adc_adam_step")
        for p in self._params:                                  # the kernel wrote the storage behind torch's back: caches keyed on
            torch.autograd.graph.increment_version(p)           # the version counter (dense.cached_weights) must see the update
        return loss

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

"""Flat-bucket Adam on top of libsscg_b200.so (reference: torch.optim.Adam(itertools.chain(...), lr, betas=(0.5, 0.999))
at model.py:286-287, stepped at model.py:474,542, decayed by LambdaLR at model.py:289-290,659-660).

All parameters of one optimizer live in ONE flat fp32 buffer (the Parameters become views into it), their
gradients in the matching flat bucket of step.FlatGrads, and exp_avg / exp_avg_sq in two more flat buffers,
so an update is a single kernel launch (sscg_adam_flat) instead of a multi-tensor pass over ~100 tensors.
Step counter and learning rate are device scalars: the update is CUDA-graph capturable and a LambdaLR
scheduler only rewrites one float (`sync_lr`).  The optimizer state keeps torch.optim.Adam's layout
(state[p] = {step, exp_avg, exp_avg_sq}), so state_dict() / load_state_dict() interoperate with checkpoints
written by the reference (model.py:641-655)."""
import torch

from . import _lib as L
from .kernels import _ptr, _stream


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, flat_grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, post_step=()):
        params = list(params)
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), "FlatAdam: CUDA fp32 parameters only"
        assert [id(p) for p in params] == [id(p) for p in flat_grads.params], "FlatAdam: parameter order must match FlatGrads"
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=True, differentiable=False, fused=None)
        super().__init__(params, defaults)
        assert len(self.param_groups) == 1
        L.lib()                       # fail loudly if the extension is missing
        self.grads = flat_grads
        self.post_step = list(post_step)
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.n = n
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros((), dtype=torch.float32, device=dev)
        self.lr_t = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                view = self.flat_p[off:off + k].view_as(p)
                view.copy_(p.data)
                p.data = view
                self.state[p] = {"step": self.step_t, "exp_avg": self.flat_m[off:off + k].view_as(p),
                                 "exp_avg_sq": self.flat_v[off:off + k].view_as(p)}
                off += k
        self._checked = False

    def _check_layout(self):
        off = 0
        base_p, base_g = self.flat_p.data_ptr(), self.grads.flat.data_ptr()
        for p in self.param_groups[0]["params"]:
            if p.data_ptr() != base_p + 4 * off:
                raise RuntimeError("FlatAdam: a parameter was re-allocated outside the flat bucket (use load_state_dict / copy_)")
            if p.grad is None or p.grad.data_ptr() != base_g + 4 * off:
                raise RuntimeError("FlatAdam: a gradient is not the FlatGrads view (do not call zero_grad(set_to_none=True))")
            off += p.numel()
        self._checked = True

    def sync_lr(self):
        """Push param_groups[0]['lr'] (as rewritten by a LambdaLR scheduler) to the device scalar."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_t.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if not self._checked:
            self._check_layout()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        g = self.param_groups[0]
        self.step_t.add_(1.0)
        L.check(L.lib().sscg_adam_flat(_ptr(self.flat_p), _ptr(self.grads.flat), _ptr(self.flat_m), _ptr(self.flat_v),
                                       self.n, _ptr(self.lr_t), float(g["betas"][0]), float(g["betas"][1]),
                                       float(g["eps"]), _ptr(self.step_t), _stream()), "sscg_adam_flat")
        for cb in self.post_step:     # the raw update does not move the parameters' version counters
            cb()
        return loss

    def zero_grad(self, set_to_none=False):
        self.grads.zero()             # gradients are persistent views: never set them to None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the base class installs fresh tensors: fold them back into the flat buffers
        off = 0
        step = None
        with torch.no_grad():
            for p in self.param_groups[0]["params"]:
                k = p.numel()
                st = self.state.get(p, {})
                m, v = self.flat_m[off:off + k].view_as(p), self.flat_v[off:off + k].view_as(p)
                if "exp_avg" in st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    step = st["step"] if step is None else step
                self.state[p] = {"step": self.step_t, "exp_avg": m, "exp_avg_sq": v}
                off += k
            if step is not None:
                self.step_t.fill_(float(step))
        self._lr_host = None
        self.sync_lr()

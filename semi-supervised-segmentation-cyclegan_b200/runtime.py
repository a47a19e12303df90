"""Autograd bridge between the nn.Module boundary (arch.*) and the stage engine.

`NetRunner` is owned by one network module.  It builds the stage list once, keeps the bf16 weight
slabs in sync with the fp32 parameters (re-derived whenever a parameter's version counter moves,
i.e. after every optimizer step), caches one NetPlan per input shape and exposes the whole network
as a single torch.autograd.Function so callers can keep composing losses with stock torch
(reference model.py:398-468) and call .backward().
"""
import os
from typing import Callable, List

import torch

from . import _lib as L
from .engine import NetPlan, StageSpec, StageWeights

_DEFAULT_PRECISION = os.environ.get("SSCG_PRECISION", "bf16")


def set_default_precision(p: str):
    """'bf16' (fast path) or 'bf16x3' (parity mode: hi/lo operand split, fp32 storage)."""
    global _DEFAULT_PRECISION
    if p not in ("bf16", "bf16x3"):
        raise ValueError("precision must be 'bf16' or 'bf16x3'")
    _DEFAULT_PRECISION = p


def default_precision() -> str:
    return _DEFAULT_PRECISION


def _rank_salt() -> int:
    """Per-rank term of the dropout seeds: data-parallel ranks seed torch identically (same initial weights), but
    each must draw its own masks.  0 on rank 0 / single process."""
    import torch.distributed as dist
    rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else int(os.environ.get("RANK", "0"))
    return (rank * 0x9E3779B1) & 0x7FFFFFFF


class NetRunner:
    def __init__(self, build_specs: Callable[[], List[StageSpec]], residual_plan=None):
        self._build_specs = build_specs
        self._residual_plan = residual_plan
        self.specs = None
        self.weights = None
        self.plans = {}
        self.precision = None
        self._wkey = None
        self.device = None
        self.drop_ctr = None      # device int64 tensor; when set, dropout masks change with its value
        self.dw_flat = None
        # direct_grad: backward adds weight gradients straight into the (pre-allocated) p.grad tensors and
        # returns None to autograd for them — saves one zero-fill and one add per parameter per backward
        self.direct_grad = False
        # defer_unpack (with direct_grad): backward passes only ACCUMULATE the weight-gradient slabs; the
        # owner calls flush_wgrad() once per optimizer step (one re-layout per stage instead of one per pass)
        self.defer_unpack = False
        self._dw_dirty = False
        self._prep_table = None       # (key, device table, count, total) of sscg_wprep_batch
        self._unpack_table = None     # same for sscg_wgrad_unpack_batch

    def _setup(self, device, precision):
        if self.specs is not None and self.precision == precision and self.device == device:
            return
        L.lib()   # fail loudly here if the extension is missing
        self.specs = self._build_specs()
        self.precision = precision
        self.device = device
        split = precision == "bf16x3"
        self.weights = [StageWeights(s, split, True, device, first=(i == 0)) for i, s in enumerate(self.specs)]
        # one flat fp32 buffer behind all weight-gradient slabs: a single fill zeroes them
        total = sum(w.dw.numel() for w in self.weights)
        self.dw_flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for w in self.weights:
            n = w.dw.numel()
            w.dw = self.dw_flat[off:off + n]
            off += n
        self.plans = {}
        self._wkey = None
        self._prep_table = None
        self._unpack_table = None

    def params(self):
        ps = []
        for s in self.specs:
            ps.append(s.weight)
            if s.bias is not None:
                ps.append(s.bias)
        return ps

    @staticmethod
    def _slab_elems(a):
        """Work items of one slab in the batched kernels: one per (row, k) position (the kernel walks the taps)."""
        return a.rows_pad * a.Kc

    def _upload_table(self, entries):
        """entries: [(WprepArgs, slab ptr, grad ptr)] -> (device byte tensor holding SscgWbatchEntry[], count, total)"""
        import ctypes as C
        tab = (L.WbatchEntry * len(entries))()
        start = 0
        for i, (a, slab, grad) in enumerate(entries):
            C.memmove(C.byref(tab[i].a), C.byref(a), C.sizeof(L.WprepArgs))
            tab[i].slab, tab[i].grad, tab[i].start = slab, grad, start
            start += self._slab_elems(a)
        dev = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(self.device)
        return dev, len(entries), start

    def _prepare_weights(self):
        """All bf16 operand slabs of the network in ONE launch (sscg_wprep_batch over a device table)."""
        from . import kernels as K
        key = tuple(s.weight.data_ptr() for s in self.specs)
        if self._prep_table is None or self._prep_table[0] != key:
            entries = []
            for w in self.weights:
                entries += [(d, None, None) for d in w.prep_descs()]
            self._prep_table = (key,) + self._upload_table(entries)
        _, dev, count, total = self._prep_table
        L.check(L.lib().sscg_wprep_batch(dev.data_ptr(), count, total, K._stream()), "sscg_wprep_batch")
        for w in self.weights:
            if w.bias_pad is not None:
                w.bias_pad[: w.spec.Cout].copy_(w.spec.bias.detach())

    def ensure_weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.params())
        if key != self._wkey:
            with torch.no_grad():
                self._prepare_weights()
            self._wkey = key

    def invalidate_weights(self):
        """Force the bf16 slabs to be re-derived at the next forward (for updates that bypass autograd's
        version counters, e.g. optim.FlatAdam)."""
        self._wkey = None

    def flush_wgrad(self, scale=1.0):
        """Fold the accumulated weight-gradient slabs into the parameters' .grad and clear the slabs."""
        if not self._dw_dirty:
            return
        from . import kernels as K
        live = [(s, wt) for s, wt in zip(self.specs, self.weights) if s.weight.grad is not None]
        key = tuple((wt.dw.data_ptr(), s.weight.grad.data_ptr()) for s, wt in live)
        if live:
            if self._unpack_table is None or self._unpack_table[0] != key:
                entries = [(wt.unpack_wg, wt.dw.data_ptr(), s.weight.grad.data_ptr()) for s, wt in live]
                self._unpack_table = (key,) + self._upload_table(entries)
            _, dev, count, total = self._unpack_table
            L.check(L.lib().sscg_wgrad_unpack_batch(dev.data_ptr(), count, total, float(scale), K._stream()),
                    "sscg_wgrad_unpack_batch")
        self.dw_flat.zero_()
        self._dw_dirty = False

    def plan(self, N, H, W) -> NetPlan:
        k = (N, H, W)
        p = self.plans.get(k)
        if p is None:
            p = NetPlan(self.specs, self.weights, N, H, W, self.precision, self.device)
            if self._residual_plan is not None:
                p.res_bwd = self._residual_plan(self.specs)
            self.plans[k] = p
        p.drop_ctr = self.drop_ctr
        p.dw_flat = self.dw_flat
        return p

    def __call__(self, x, training, use_dropout, precision=None):
        """x: NCHW fp32, or an int64 label map N x 1 x H x W that stands for its one-hot encoding over the
        network's input channels (make_one_hot, utils.py:314-350): the first stage's operand buffer is then written
        straight from the labels (sscg_onehot_pack) and no N x C x H x W fp32 tensor is materialised.
        A list / tuple of such tensors runs ONE batched pass over the concatenation of their batches (InstanceNorm is
        per sample, so every sample's result is the one a separate call would give); the output is the concatenation."""
        parts = list(x) if isinstance(x, (list, tuple)) else [x]
        if not all(p.is_cuda for p in parts):
            raise RuntimeError("fused path needs CUDA tensors")
        self._setup(parts[0].device, precision or _DEFAULT_PRECISION)
        params = self.params()
        return _FusedNet.apply(self, training and use_dropout, len(parts), *parts, *params)


class _CtxLease:
    """Returns the activation context to its plan when the autograd graph that holds it is freed
    (also when backward is never run, e.g. model.py:409 `recon_lab_img`, which feeds no loss)."""

    def __init__(self, plan, c):
        self.plan, self.c = plan, c

    def release(self):
        if self.c is not None:
            self.plan.release_ctx(self.c)
            self.c = None

    def __del__(self):
        self.release()


class _FusedNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner: NetRunner, dropout_on, nparts, *rest):
        parts, params = rest[:nparts], rest[nparts:]
        prepared = []
        for x in parts:
            x = x.detach()
            if x.dtype != torch.int64 and x.dtype != torch.float32:
                x = x.float()
            if x.dtype == torch.int64:
                assert x.shape[1] == 1, "label maps are N x 1 x H x W"
            prepared.append(x.contiguous())
        N = sum(p.shape[0] for p in prepared)
        H, W = prepared[0].shape[2:]
        assert all(p.shape[2:] == prepared[0].shape[2:] for p in prepared), "batched parts must share H x W"
        runner.ensure_weights()
        plan = runner.plan(N, H, W)
        c = plan.acquire_ctx()
        seed = 0
        if dropout_on:
            seed = int(torch.randint(1, 2 ** 31 - 1, (1,)).item())   # CPU generator: follows torch.manual_seed
            seed = (seed ^ _rank_salt()) or 1                        # ... and differs between data-parallel ranks
        plan.forward(c, parts=prepared, training=dropout_on, drop_seed=seed)
        y = plan.output_nchw(c)
        # (grad mode is always off inside Function.forward; needs_input_grad already reflects no_grad())
        need_grad = any(ctx.needs_input_grad[3:])
        if need_grad:
            ctx.runner, ctx.plan, ctx.lease = runner, plan, _CtxLease(plan, c)
            ctx.nparts = nparts
            ctx.part_sizes = [p.shape[0] for p in prepared]
        else:
            plan.release_ctx(c)
            ctx.plan = None
        return y

    @staticmethod
    def backward(ctx, gy):
        plan, c, runner = ctx.plan, ctx.lease.c, ctx.runner
        if c is None:
            raise RuntimeError("fused network: backward called twice on the same graph (activations were released)")
        nparts = ctx.nparts
        part_need = list(ctx.needs_input_grad[3:3 + nparts])
        need_dx = any(part_need)
        need_dw = any(ctx.needs_input_grad[3 + nparts:])
        # input gradient only for the sample range spanned by the parts that want one
        starts = [sum(ctx.part_sizes[:j]) for j in range(nparts + 1)]
        dx_range = None
        if need_dx:
            want = [j for j in range(nparts) if part_need[j]]
            dx_range = (starts[want[0]], starts[want[-1] + 1])
        direct = need_dw and runner.direct_grad and all(p.grad is not None for p in runner.params() if p.requires_grad)
        defer = direct and runner.defer_unpack
        gx = plan.backward(c, gy.float(), need_dx=need_dx, need_dw=need_dw, accumulate_dw=defer, dx_range=dx_range)
        gparts = [None] * nparts
        if need_dx:
            for j in range(nparts):
                if part_need[j]:
                    a0 = starts[j] - dx_range[0]
                    gparts[j] = gx[a0:a0 + ctx.part_sizes[j]]
        grads = []
        nparams = len(ctx.needs_input_grad) - 3 - nparts
        if direct:
            into = [(s.weight.grad, s.bias.grad if s.bias is not None else None) for s in plan.specs]
            plan.param_grads(into=into, weights=not defer)
            runner._dw_dirty = runner._dw_dirty or defer
            grads = [None] * nparams
        elif need_dw:
            pg = plan.param_grads()
            flags = list(ctx.needs_input_grad[3 + nparts:])
            j = 0
            for (gw, gb), s in zip(pg, plan.specs):
                grads.append(gw if flags[j] else None)
                j += 1
                if s.bias is not None:
                    grads.append(gb if flags[j] else None)
                    j += 1
        else:
            grads = [None] * nparams
        ctx.lease.release()
        return (None, None, None, *gparts, *grads)

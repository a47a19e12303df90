"""Fused loss nodes on top of libsscg_b200.so, each one autograd.Function:
  * seg_head  — softmax + cross-entropy + argmax: nn.CrossEntropyLoss / nn.Softmax2d / .max(1)[1] of the
                reference step (model.py:272-273,398,401-402,435,455,509);
  * lsgan_loss — nn.MSELoss against an all-ones / all-zeros target (model.py:270,445-446,452,521-534);
  * l1_loss   — nn.L1Loss (model.py:271,453,461).
CUDA tensors only; the step uses the stock torch ops on CPU tensors (the reference's own gpu_ids=[] path)."""
import ctypes as C

import torch

from . import _lib as L
from .kernels import _ptr, _stream


_LOSS_WS = {}


def _loss_ws(device):
    """Scratch of the fixed-order loss reductions (SSCG_LOSS_WS_BYTES, zeroed once; every launch leaves its arrival
    counter zero).  One per (device, stream): launches that share it are ordered."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _LOSS_WS.get(key)
    if ws is None:
        ws = _LOSS_WS[key] = torch.zeros(L.SSCG_LOSS_WS_BYTES, dtype=torch.uint8, device=device)
    return ws


class _SegHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        ctx.set_materialize_grads(False)        # unused outputs arrive as None, not as zero tensors
        assert logits.is_cuda and logits.dtype == torch.float32, "seg_head: CUDA fp32 logits"
        logits = logits.contiguous()
        N, Cc, H, W = logits.shape
        dev = logits.device
        probs = torch.empty_like(logits)
        argmax = torch.empty((N, H, W), dtype=torch.int64, device=dev)
        out = torch.zeros(2, dtype=torch.float32, device=dev)      # (sum of -log p[label], counted pixels)
        lab = None
        if labels is not None:
            if labels.is_floating_point() or labels.numel() != N * H * W:
                raise TypeError("seg_head: labels must be an integer map with N*H*W elements (nn.CrossEntropyLoss target)")
            lab = labels.reshape(N, H, W).long().contiguous()        # int64 is what the kernel reads
        L.check(L.lib().sscg_seg_head_fwd(_ptr(logits), _ptr(lab), N, Cc, H * W, int(ignore_index), _ptr(probs),
                                          _ptr(argmax), _ptr(out) if lab is not None else None, _ptr(_loss_ws(dev)),
                                          _stream()), "sscg_seg_head_fwd")
        loss = out[0] / out[1] if lab is not None else out[0]      # mean over the counted pixels (reduction='mean')
        ctx.save_for_backward(probs, lab if lab is not None else torch.empty(0, device=dev), out)
        ctx.has_labels = lab is not None
        ctx.mark_non_differentiable(argmax)
        return loss, probs, argmax

    @staticmethod
    def backward(ctx, dloss, dprobs, _dargmax):
        probs, lab, out = ctx.saved_tensors
        N, Cc, H, W = probs.shape
        dlogits = torch.empty_like(probs)
        lab_p = lab if ctx.has_labels else None
        dl = dloss.contiguous().float() if (dloss is not None and ctx.has_labels) else None
        dp = dprobs.contiguous() if dprobs is not None else None
        L.check(L.lib().sscg_seg_head_bwd(_ptr(probs), _ptr(lab_p), _ptr(dl), _ptr(out, 4) if dl is not None else None,
                                          _ptr(dp), N, Cc, H * W, _ptr(dlogits), _stream()), "sscg_seg_head_bwd")
        return dlogits, None, None


def seg_head(logits, labels=None, ignore_index=-100):
    """-> (mean cross-entropy (0 if labels is None), softmax probabilities, argmax label map).
    nn.CrossEntropyLoss semantics for the labels: pixels equal to ignore_index are left out of the mean and of the
    gradient; any other label outside [0, C) raises the device error flag (kernels.device_error() -> code 31;
    torch device-asserts on those)."""
    return _SegHead.apply(logits, labels, ignore_index)


class _LsganLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        x = x.contiguous()
        assert x.dtype == torch.float32 and x.is_cuda
        n = x.numel()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        L.check(L.lib().sscg_lsgan_fwd(_ptr(x), n, float(target), 1.0 / n, _ptr(loss), _ptr(_loss_ws(x.device)), _stream()),
                "sscg_lsgan_fwd")
        ctx.save_for_backward(x)
        ctx.target = float(target)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        dl = dloss.contiguous().float()
        L.check(L.lib().sscg_lsgan_bwd(_ptr(x), x.numel(), ctx.target, _ptr(dl), _ptr(dx), _stream()), "sscg_lsgan_bwd")
        return dx, None


def lsgan_loss(x, target):
    """mean((x - target)^2) for a constant target (1.0 = "real", 0.0 = "fake")."""
    return _LsganLoss.apply(x, target)


class _L1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = x.contiguous(), y.contiguous()
        assert x.dtype == torch.float32 and y.dtype == torch.float32 and x.shape == y.shape and x.is_cuda
        n = x.numel()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        L.check(L.lib().sscg_l1_fwd(_ptr(x), _ptr(y), n, 1.0 / n, _ptr(loss), _ptr(_loss_ws(x.device)), _stream()),
                "sscg_l1_fwd")
        ctx.save_for_backward(x, y)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, y = ctx.saved_tensors
        dx = torch.empty_like(x)
        dl = dloss.contiguous().float()
        L.check(L.lib().sscg_l1_bwd(_ptr(x), _ptr(y), x.numel(), _ptr(dl), _ptr(dx), _stream()), "sscg_l1_bwd")
        return dx, None       # the second argument is data (model.py:453,461), never a graph tensor


def l1_loss(x, y):
    """mean(|x - y|); gradient flows to x only."""
    return _L1Loss.apply(x, y.detach())

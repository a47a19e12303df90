"""Fused segmentation-head loss (softmax + cross-entropy + argmax) as one autograd node on top of
libsscg_b200.so — replaces nn.CrossEntropyLoss / nn.Softmax2d / .max(1)[1] of the reference step
(model.py:272-273,398,401-402,435,455,509).  CUDA tensors only; the step uses the stock torch ops on
CPU tensors (the reference's own gpu_ids=[] path)."""
import ctypes as C

import torch

from . import _lib as L
from .kernels import _ptr, _stream


class _SegHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, want_probs):
        ctx.set_materialize_grads(False)        # unused outputs arrive as None, not as zero tensors
        logits = logits.contiguous()
        N, Cc, H, W = logits.shape
        dev = logits.device
        probs = torch.empty_like(logits)
        argmax = torch.empty((N, H, W), dtype=torch.int64, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        lab = None
        if labels is not None:
            lab = labels.reshape(N, H, W).contiguous()
        L.check(L.lib().sscg_seg_head_fwd(_ptr(logits), _ptr(lab), N, Cc, H * W, _ptr(probs), _ptr(argmax),
                                          _ptr(loss) if lab is not None else None, _stream()), "sscg_seg_head_fwd")
        if lab is not None:
            loss = loss / float(N * H * W)
        ctx.save_for_backward(probs, lab if lab is not None else torch.empty(0, device=dev))
        ctx.has_labels = lab is not None
        ctx.mark_non_differentiable(argmax)
        return loss, probs, argmax

    @staticmethod
    def backward(ctx, dloss, dprobs, _dargmax):
        probs, lab = ctx.saved_tensors
        N, Cc, H, W = probs.shape
        dlogits = torch.empty_like(probs)
        lab_p = lab if ctx.has_labels else None
        dl = dloss.contiguous().float() if (dloss is not None and ctx.has_labels) else None
        dp = dprobs.contiguous() if dprobs is not None else None
        L.check(L.lib().sscg_seg_head_bwd(_ptr(probs), _ptr(lab_p), _ptr(dl), _ptr(dp), N, Cc, H * W, _ptr(dlogits),
                                          _stream()), "sscg_seg_head_bwd")
        return dlogits, None, None


def seg_head(logits, labels=None):
    """-> (mean cross-entropy (0 if labels is None), softmax probabilities, argmax label map)."""
    return _SegHead.apply(logits, labels, True)

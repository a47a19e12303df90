"""`interp` of the reference's training loops: nn.Upsample(size=(crop_height, crop_width), mode='bilinear',
align_corners=True) (reference model.py:62-63,268), applied to every generator output (model.py:132,390-392,413-415,
562,581-594) so that a generator working at reduced resolution (deeplab: 1/8) meets the label / image size.

On CUDA fp32 tensors the resize runs in this repo's kernels (sscg_interp_bilinear_fwd / _bwd: one streaming pass each,
the backward a reproducible gather); when the input already has the target size — the ResNet generators of the hot
path — it is the identity and costs nothing.  Other tensors take torch's own nn.Upsample."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .kernels import _ptr, _stream


class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo):
        x = x.contiguous()
        N, Cc, Hi, Wi = x.shape
        y = torch.empty((N, Cc, Ho, Wo), dtype=torch.float32, device=x.device)
        L.check(L.lib().sscg_interp_bilinear_fwd(_ptr(x), N, Cc, Hi, Wi, _ptr(y), Ho, Wo, _stream()),
                "sscg_interp_bilinear_fwd")
        ctx.shape = (N, Cc, Hi, Wi, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, Cc, Hi, Wi, Ho, Wo = ctx.shape
        dy = dy.contiguous().float()
        dx = torch.empty((N, Cc, Hi, Wi), dtype=torch.float32, device=dy.device)
        L.check(L.lib().sscg_interp_bilinear_bwd(_ptr(dy), N, Cc, Hi, Wi, Ho, Wo, _ptr(dx), _stream()),
                "sscg_interp_bilinear_bwd")
        return dx, None, None


class Interp(nn.Module):
    """Drop-in for the reference's `self.interp` (same constructor meaning: target (height, width))."""

    def __init__(self, size, mode="bilinear", align_corners=True):
        super().__init__()
        assert mode == "bilinear" and align_corners, "the reference uses bilinear, align_corners=True"
        self.size = (int(size[0]), int(size[1]))

    def forward(self, x):
        if tuple(x.shape[2:]) == self.size:
            return x                                     # identity for the hot path's full-resolution generators
        if x.is_cuda and x.dtype == torch.float32:
            return _Bilinear.apply(x, self.size[0], self.size[1])
        return F.interpolate(x, size=self.size, mode="bilinear", align_corners=True)

"""Geometry of the implicit-GEMM convolutions: tap tables, tiles, channel padding.

Pure Python / no torch: everything here is host logic that the CPU test-suite checks against a
brute-force enumeration (tests/test_geometry.py).  A "tap" is (dh, dw, brow): the input pixel read
for output pixel (o_h, o_w) is (o_h*stride + dh + org_h, o_w*stride + dw + org_w) and `brow` names
the weight slab (filter position) that multiplies it — see include/sscg_b200.h.

Conventions follow torch.nn.Conv2d / ConvTranspose2d as used by the reference
(arch/ops.py:40-57): Conv2d out = floor((in + 2p - k)/s) + 1; ConvTranspose2d
out = (in - 1)*s - 2p + k + output_padding, o = s*i - p + kh.
"""
from dataclasses import dataclass, field
from typing import List, Tuple


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def next_pow2(x: int) -> int:
    p = 1
    while p < x:
        p *= 2
    return p


def pad_out_channels(c: int) -> int:
    """Output-channel padding to a legal tcgen05 N tile (16, 32, 64, 128 or a multiple of 256)."""
    for b in (16, 32, 64, 128, 256):
        if c <= b:
            return b
    return round_up(c, 256)


def pick_bn(co_pad: int) -> int:
    return co_pad if co_pad <= 256 else 256


def pad_in_channels(c: int) -> int:
    """Channel pitch of an activation buffer: multiple of 8 (16-byte pixels) for small C, of 64 above."""
    if c >= 64:
        return round_up(c, 64)
    return round_up(c, 8)


def pick_tile(w_out: int, pixels: int = 128, h_out: int = 0) -> Tuple[int, int]:
    """(TH, TW) with TH*TW == pixels and TW a power of two.  Without h_out: the widest TW that the row
    can use.  With h_out: the shape that covers h_out x w_out with the fewest tiles (ties -> wider TW,
    longer contiguous rows per TMA box); this matters for extents such as 66 x 66 (dgrad w.r.t. a
    reflect-padded buffer), where 1 x 128 tiles would leave half of every tile empty."""
    if h_out <= 0:
        tw = min(pixels, next_pow2(max(w_out, 1)))
        return pixels // tw, tw
    best = None
    tw = pixels
    while tw >= 4:
        th = pixels // tw
        tiles = ((h_out + th - 1) // th) * ((w_out + tw - 1) // tw)
        if best is None or tiles < best[0]:
            best = (tiles, th, tw)
        tw //= 2
    return best[1], best[2]


def conv_out(n: int, k: int, s: int, p: int) -> int:
    return (n + 2 * p - k) // s + 1


def convT_out(n: int, k: int, s: int, p: int, op: int) -> int:
    return (n - 1) * s - 2 * p + k + op


@dataclass
class TapTable:
    n_phases: int
    phase_start: List[int]
    taps: List[Tuple[int, int, int]] = field(default_factory=list)   # (dh, dw, brow)
    stride: int = 1
    org_h: int = 0
    org_w: int = 0


def taps_conv_fwd(kh: int, kw: int, stride: int, org: int) -> TapTable:
    """Forward Conv2d on a view whose coordinate 0 is input pixel `-org`... i.e. in = out*s + k + org
    (org = -padding when the halo is implicit zero fill, 0 when the halo is explicit)."""
    taps = [(a, b, a * kw + b) for a in range(kh) for b in range(kw)]
    return TapTable(1, [0, len(taps), len(taps), len(taps), len(taps)], taps, stride, org, org)


def taps_conv_fwd_window(kh: int, stride: int, org: int) -> TapTable:
    """Row-window mode: one tap per filter row; the (kw, c) run is contiguous in memory."""
    taps = [(a, 0, a) for a in range(kh)]
    return TapTable(1, [0, kh, kh, kh, kh], taps, stride, org, org)


def taps_conv_dgrad(kh: int, kw: int, stride: int, org: int) -> TapTable:
    """Gradient w.r.t. the input view of a forward conv (in = out*s + k + org): gather over dY.
    stride 1: dX[i] = sum_k dY[i - k - org]            -> one phase, dh = -k - org
    stride 2: i = 2a + ph, k == (ph - org) mod 2: dX[i] += dY[a + (ph - k - org)/2] -> four phases."""
    if stride == 1:
        taps = [(-a - org, -b - org, a * kw + b) for a in range(kh) for b in range(kw)]
        return TapTable(1, [0, len(taps), len(taps), len(taps), len(taps)], taps, 1, 0, 0)
    assert stride == 2
    taps, starts = [], [0]
    for ph in range(2):
        for pw in range(2):
            for a in range(kh):
                if (ph - a - org) % 2:
                    continue
                for b in range(kw):
                    if (pw - b - org) % 2:
                        continue
                    taps.append(((ph - a - org) // 2, (pw - b - org) // 2, a * kw + b))
            starts.append(len(taps))
    return TapTable(4, starts, taps, 1, 0, 0)


def taps_convT_fwd(kh: int, kw: int, stride: int, pad: int) -> TapTable:
    """Forward ConvTranspose2d (stride 2) as four interleaved gathers: o = 2a + ph receives
    in[a + (ph + pad - k)/2] * W[k] for k == (ph + pad) mod 2."""
    assert stride == 2
    taps, starts = [], [0]
    for ph in range(2):
        for pw in range(2):
            for a in range(kh):
                if (ph + pad - a) % 2:
                    continue
                for b in range(kw):
                    if (pw + pad - b) % 2:
                        continue
                    taps.append(((ph + pad - a) // 2, (pw + pad - b) // 2, a * kw + b))
            starts.append(len(taps))
    return TapTable(4, starts, taps, 1, 0, 0)


def taps_convT_dgrad(kh: int, kw: int, stride: int, pad: int) -> TapTable:
    """Gradient w.r.t. the input of ConvTranspose2d: dX[i] = sum_k dY[s*i - pad + k] W[k] — a strided
    forward-style gather over dY."""
    return taps_conv_fwd(kh, kw, stride, -pad)


def taps_rowshift_fwd(kh: int, kw: int, org: int) -> TapTable:
    """Row-shift mode, forward stride-1 conv: one entry per filter row = (dh, leftmost dw, slab of the
    leftmost tap); tap j of the row is slab brow + j (see SscgConvArgs.shift_kw)."""
    taps = [(a, 0, a * kw) for a in range(kh)]
    return TapTable(1, [0, kh, kh, kh, kh], taps, 1, org, org)


def taps_rowshift_dgrad(kh: int, kw: int, org: int) -> TapTable:
    """Row-shift mode, stride-1 dgrad: dX[i] = sum_k dY[i - k - org] W[k]; leftmost input column is
    dw = -(kw-1) - org and belongs to tap kw-1, so slabs run backwards (shift_brow_step = -1)."""
    taps = [(-a - org, -(kw - 1) - org, a * kw + kw - 1) for a in range(kh)]
    return TapTable(1, [0, kh, kh, kh, kh], taps, 1, 0, 0)

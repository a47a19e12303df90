"""Host-side builders for the C-ABI argument blocks, plus thin launch wrappers.

torch is used here only for device memory (tensor.data_ptr()) and the current stream; every
arithmetic kernel lives in libsscg_b200.so.
"""
import ctypes as C
import os

import torch

from . import _lib as L
from .geometry import TapTable, pick_tile

SLACK = 1024  # zeroed elements appended to every activation buffer (window reads run past the end)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, byte_off=0):
    if t is None:
        return None
    return t.data_ptr() + byte_off


class ActBuf:
    """NHWC activation buffer [N][H+2*pad][W+2*pad][C] (bf16 hi plane, optional lo plane, or fp32)."""

    def __init__(self, N, H, W, Cc, pad=0, device="cuda", split=False, fp32=False):
        self.N, self.H, self.W, self.C, self.pad = N, H, W, Cc, pad
        self.Hp, self.Wp = H + 2 * pad, W + 2 * pad
        self.fp32 = fp32
        n = N * self.Hp * self.Wp * Cc + SLACK
        dt = torch.float32 if fp32 else torch.bfloat16
        self.hi = torch.zeros(n, dtype=dt, device=device)
        self.lo = torch.zeros(n, dtype=dt, device=device) if (split and not fp32) else None
        self.esize = 4 if fp32 else 2

    # element strides
    @property
    def sW(self):
        return self.C

    @property
    def sH(self):
        return self.Wp * self.C

    @property
    def sN(self):
        return self.Hp * self.Wp * self.C

    def _off(self, interior):
        return (self.pad * self.sH + self.pad * self.sW) * self.esize if interior else 0

    def view(self, interior=False, lo=False):
        """Full (padded extents) or interior (H x W extents, zero fill outside) view."""
        t = self.lo if lo else self.hi
        H, W = (self.H, self.W) if interior else (self.Hp, self.Wp)
        return L.make_view(_ptr(t, self._off(interior)), self.N, H, W, self.C, self.sN, self.sH, self.sW)

    def lo_ptr(self, interior=False):
        return _ptr(self.lo, self._off(interior)) if self.lo is not None else None

    def window_view(self, kwpad, lo=False):
        """Row-window view over the padded buffer: inner extent = kwpad contiguous (kw, c) elements."""
        t = self.lo if lo else self.hi
        return L.make_view(_ptr(t), self.N, self.Hp, self.Wp, kwpad, self.sN, self.sH, self.sW)

    def as_nhwc(self, interior=True):
        """torch view [N, H(p), W(p), C] of the hi plane (tests / debugging)."""
        t = self.hi[: self.N * self.Hp * self.Wp * self.C].view(self.N, self.Hp, self.Wp, self.C)
        if interior and self.pad:
            t = t[:, self.pad:self.pad + self.H, self.pad:self.pad + self.W, :]
        return t

    def as_nhwc_f32(self, interior=True):
        t = self.as_nhwc(interior).float()
        if self.lo is not None:
            l = self.lo[: self.N * self.Hp * self.Wp * self.C].view(self.N, self.Hp, self.Wp, self.C)
            if interior and self.pad:
                l = l[:, self.pad:self.pad + self.H, self.pad:self.pad + self.W, :]
            t = t + l.float()
        return t


class WsPool:
    """Grow-only zeroed device scratch for the fixed-order reductions (statistics / plane sums / split-K partials):
    arrival counters at the head (every launch leaves them zero) + partial slots.  One pool serves launches that
    are ordered on one stream; an argument block keeps the tensor it points into alive (`_keep`)."""

    def __init__(self, device="cuda"):
        self.device, self.buf = device, None

    def get(self, nbytes):
        if nbytes is None or nbytes <= 0:
            return None
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.zeros(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
        return self.buf


_DEFAULT_POOLS = {}


def default_pool(kind, device="cuda"):
    """Pools used when a caller (kernel tests, tools) does not bring its own."""
    key = (kind, str(torch.device(device)))
    if key not in _DEFAULT_POOLS:
        _DEFAULT_POOLS[key] = WsPool(device)
    return _DEFAULT_POOLS[key]


def _attach_ws(a, field, nbytes, pool, kind):
    if nbytes < 0:
        raise L.SscgError("workspace query failed for " + kind)
    buf = (pool or default_pool(kind)).get(nbytes)
    setattr(a, field, _ptr(buf))
    a._keep = getattr(a, "_keep", []) + [buf]


def fill_taps(dst_taps, table: TapTable):
    assert len(table.taps) <= L.SSCG_MAX_TAPS, "too many taps"
    for i, (dh, dw, brow) in enumerate(table.taps):
        dst_taps[i].dh, dst_taps[i].dw, dst_taps[i].brow = dh, dw, brow


def conv_args(xview, x_lo, table: TapTable, Kc, w, w_lo, w_rows, Co_pad, y_ptr, y_fp32, y_strides, y_off, Ho, Wo,
              bias=None, act=L.ACT_NONE, slope=0.2, stats=None, BN=None, tile=None, split=1, tag=4,
              shift_kw=0, shift_brow_step=1, shift_base_mode=2, flat=None, ws_pool=None, rw_pitch=0):
    a = L.ConvArgs()
    a.x = xview
    a.x_lo = x_lo
    a.stride, a.Kc, a.org_h, a.org_w = table.stride, Kc, table.org_h, table.org_w
    a.n_phases = table.n_phases
    for i in range(5):
        a.phase_start[i] = table.phase_start[i]
    fill_taps(a.taps, table)
    a.w, a.w_lo, a.w_rows, a.Co_pad = _ptr(w), _ptr(w_lo), w_rows, Co_pad
    a.split = split
    a.y, a.y_fp32 = y_ptr, 1 if y_fp32 else 0
    a.y_sN, a.y_sH, a.y_sW = y_strides
    a.y_oh, a.y_ow = y_off
    a.Ho, a.Wo = Ho, Wo
    a.bias = _ptr(bias)
    a.act, a.slope = act, slope
    a.stats = _ptr(stats)
    if tile is None:
        w_phase = (Wo + 1) // 2 if table.n_phases == 4 else Wo
        h_phase = (Ho + 1) // 2 if table.n_phases == 4 else Ho
        tile = pick_tile(w_phase, 128, h_phase)
    a.TH, a.TW = tile
    a.BN = BN if BN is not None else (Co_pad if Co_pad <= 256 else 256)
    a.tag = tag
    a.shift_kw, a.shift_brow_step, a.shift_base_mode = shift_kw, shift_brow_step, shift_base_mode
    if shift_kw:
        a.TH, a.TW = 1, 128
    if flat is not None:          # (row pitch, positions per sample, samples) of the flattened zero-haloed input
        a.flat_pitch, a.flat_hw, a.flat_n = flat
        a.TH, a.TW = 1, 128
    a.rw_pitch = rw_pitch
    if rw_pitch:
        a.TH, a.TW = 1, 128
    return a


def wgrad_args(dyview, dy_lo, xview, x_lo, table: TapTable, Kc, Co_pad, dw, w_rows, BN=None, tile=None, ksplit=None,
               split=1, tag=6, ws_pool=None, rw_pitch=0):
    assert table.n_phases == 1
    a = L.WgradArgs()
    a.dy, a.dy_lo, a.x, a.x_lo = dyview, dy_lo, xview, x_lo
    a.stride, a.Kc, a.org_h, a.org_w = table.stride, Kc, table.org_h, table.org_w
    a.n_taps = len(table.taps)
    fill_taps(a.taps, table)
    a.Co_pad, a.split = Co_pad, split
    a.dw, a.w_rows = _ptr(dw), w_rows
    a.rw_pitch = rw_pitch
    if rw_pitch:
        tile, BN = (1, 64), Kc
    if tile is None:
        tile = pick_tile(dyview.W, 64, dyview.H)
    a.TH, a.TW = tile
    if BN is None:
        if split == 1 and Kc in (192, 448):
            BN = Kc                       # one wide tile: dY is read once per tap row
        else:
            BN = 256 if Kc % 256 == 0 else (128 if Kc % 128 == 0 else 64)
    a.BN = BN
    if ksplit is None:
        blocks = dyview.N * ((dyview.H + a.TH - 1) // a.TH) * ((dyview.W + a.TW - 1) // a.TW)
        ctas = len(table.taps) * ((Co_pad + 127) // 128) * (Kc // BN)
        ksplit = pick_ksplit(ctas, blocks, sms=148 * L.lib().sscg_conv_wgrad_ctas_per_sm(BN, split))
    a.ksplit = ksplit
    a.tag = tag
    _attach_ws(a, "ws", L.lib().sscg_conv_wgrad_ws_bytes(C.byref(a)), ws_pool, "wgrad")
    return a


def pick_ksplit(ctas, blocks, sms=148):
    """Split-K factor for the weight-gradient GEMM: one CTA per SM at a time (the tile needs most of
    the shared memory), so the grid should be a whole number of waves.  Prefer the fewest waves that
    still leaves every CTA a K loop of >= 8 pixel blocks; never exceed the number of blocks."""
    best = None
    for waves in (1, 2, 3, 4):
        ks = max(1, (sms * waves) // ctas)
        ks = min(ks, blocks)
        grid = ctas * ks
        eff = grid / (sms * ((grid + sms - 1) // sms))
        loop = blocks / ks
        score = eff - (0.15 if loop < 8 else 0.0) - 0.02 * (waves - 1)
        if best is None or score > best[0] + 1e-9:
            best = (score, ks)
    return best[1]


def run_conv(a):
    L.check(L.lib().sscg_conv_igemm(C.byref(a), _stream()), "sscg_conv_igemm")


def run_conv7(a):
    L.check(L.lib().sscg_conv7_nexp(C.byref(a), _stream()), "sscg_conv7_nexp")


def conv7_args(x_ptr, x_pitch, N, Hp, Wp, w, CoW, n_ntiles, ksteps, c_store, y_ptr, y_fp32, y_strides, bias=None,
               act=L.ACT_NONE, tag=4):
    a = L.Conv7Args()
    a.x, a.x_pitch, a.N, a.Hp, a.Wp = x_ptr, x_pitch, N, Hp, Wp
    a.w, a.CoW, a.n_ntiles, a.ksteps, a.c_store = _ptr(w), CoW, n_ntiles, ksteps, c_store
    a.y, a.y_fp32 = y_ptr, 1 if y_fp32 else 0
    a.y_sN, a.y_sH, a.y_sW = y_strides
    a.bias, a.act, a.tag = _ptr(bias), act, tag
    return a


def wgrad7_args(x: ActBuf, dy: ActBuf, dw, tag=6, ws_pool=None):
    """7x7 head weight gradient (csrc/conv_wgrad7.cu): x = the head's 64-channel input with its halo of 3, dy = dRaw in
    a zero-haloed (6) buffer of 16 / 32 channels, dw = the window-mode slab [7][64][448]."""
    assert x.C == 64 and x.pad == 3 and dy.pad == 6 and dy.C in (16, 32) and not x.fp32 and not dy.fp32
    assert (x.N, x.H, x.W) == (dy.N, dy.H, dy.W)
    a = L.Wgrad7Args()
    a.x, a.dy = _ptr(x.hi), _ptr(dy.hi)
    a.N, a.H, a.W, a.Cy = x.N, x.H, x.W, dy.C
    a.dw, a.tag = _ptr(dw), tag
    _attach_ws(a, "ws", L.lib().sscg_conv_wgrad7_ws_bytes(C.byref(a)), ws_pool, "wgrad7")
    return a


def run_wgrad(a):
    if isinstance(a, L.Wgrad7Args):
        L.check(L.lib().sscg_conv_wgrad7(C.byref(a), _stream()), "sscg_conv_wgrad7")
        return
    L.check(L.lib().sscg_conv_wgrad(C.byref(a), _stream()), "sscg_conv_wgrad")


def run_apply(a):
    L.check(L.lib().sscg_in_apply(C.byref(a), _stream()), "sscg_in_apply")


def run_bwd_prep(a):
    L.check(L.lib().sscg_in_bwd_prep(C.byref(a), _stream()), "sscg_in_bwd_prep")


def run_bwd_apply(a, draw, draw_lo=None):
    L.check(L.lib().sscg_in_bwd_apply(C.byref(a), _ptr(draw), _ptr(draw_lo), _stream()), "sscg_in_bwd_apply")


def stats_buffer(N, Cc, device="cuda"):
    """Zeroed plane-sum accumulators [N][C][2][SSCG_STAT_WORDS] (int64; binned fixed point, include/sscg_b200.h)."""
    return torch.zeros(N, Cc, 2, L.SSCG_STAT_WORDS, dtype=torch.int64, device=device)


def stats_encode(values):
    """float [..., 2] plane sums -> int64 [..., 2, SSCG_STAT_WORDS] in the kernels' accumulator format."""
    v = values.double()
    hi = torch.round(v * 256.0)
    lo = torch.round((v - hi / 256.0) * 2.0 ** 56)
    return torch.stack([hi, lo], dim=-1).to(torch.int64).contiguous()


def stats_decode(acc):
    """int64 [..., SSCG_STAT_WORDS] accumulators -> float32 values (the rounding the kernels apply)."""
    return (acc[..., 0].double() / 256.0 + acc[..., 1].double() * 2.0 ** -56).float()


def wprep_args(w, transposed, Co, Ci, KH, KW, mode, Cp, rows_pad, Kc, dst, dst_lo=None):
    a = L.WprepArgs()
    a.w, a.transposed = _ptr(w), 1 if transposed else 0
    a.Co, a.Ci, a.KH, a.KW = Co, Ci, KH, KW
    a.mode, a.Cp, a.rows_pad, a.Kc = mode, Cp, rows_pad, Kc
    a.dst, a.dst_lo = _ptr(dst), _ptr(dst_lo)
    return a


def run_wprep(a):
    L.check(L.lib().sscg_wprep(C.byref(a), _stream()), "sscg_wprep")


def run_wgrad_unpack(a, slab, grad, scale=1.0):
    L.check(L.lib().sscg_wgrad_unpack(C.byref(a), _ptr(slab), _ptr(grad), C.c_float(scale), _stream()),
            "sscg_wgrad_unpack")


def pack_nchw(src, dst: ActBuf, pad_mode, n_off=0):
    """NCHW fp32 -> samples [n_off, n_off + N) of the NHWC buffer (a batched pass packs its parts one after the other)."""
    N, Cc, H, W = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous() and n_off + N <= dst.N
    off = n_off * dst.sN * dst.esize
    L.check(L.lib().sscg_pack_nchw(_ptr(src), N, Cc, H, W, _ptr(dst.hi, off), _ptr(dst.lo, off), 1 if dst.fp32 else 0,
                                   dst.C, dst.pad, pad_mode, _stream()), "sscg_pack_nchw")


def onehot_pack(labels, Cc, dst: ActBuf, pad_mode, n_off=0):
    N, one, H, W = labels.shape
    assert labels.dtype == torch.int64 and labels.is_contiguous() and one == 1 and n_off + N <= dst.N
    off = n_off * dst.sN * dst.esize
    L.check(L.lib().sscg_onehot_pack(_ptr(labels), N, Cc, H, W, _ptr(dst.hi, off), _ptr(dst.lo, off), dst.C, dst.pad,
                                     pad_mode, _stream()), "sscg_onehot_pack")


def unpack_nhwc(src_f32, N, Cc, H, W, Cp, dst):
    L.check(L.lib().sscg_unpack_nhwc(_ptr(src_f32), N, Cc, H, W, Cp, _ptr(dst), _stream()), "sscg_unpack_nhwc")


def unpack_fold(src: ActBuf, Cc, dst, pad_mode, n0=0, n1=None):
    """Samples [n0, n1) of the padded NHWC gradient buffer -> NCHW fp32 (halo gradients folded back)."""
    n1 = src.N if n1 is None else n1
    off = n0 * src.sN * src.esize
    L.check(L.lib().sscg_unpack_fold(_ptr(src.hi, off), 1 if src.fp32 else 0, n1 - n0, Cc, src.H, src.W, src.C, src.pad,
                                     pad_mode, _ptr(dst), _stream()), "sscg_unpack_fold")


def sub_view(v, n0, n1, esize=2):
    """Samples [n0, n1) of a view."""
    return L.make_view(v.ptr + n0 * v.sN * esize, n1 - n0, v.H, v.W, v.C, v.sN, v.sH, v.sW)


def bias_grad(bstats, N, Cc, Cp, grad, scale=1.0):
    L.check(L.lib().sscg_bias_grad(_ptr(bstats), N, Cc, Cp, _ptr(grad), C.c_float(scale), _stream()),
            "sscg_bias_grad")


TAG_NAMES = {10: "loss_kernels", 11: "adam", 1: "res_conv_fwd", 2: "res_conv_dgrad", 3: "res_conv_wgrad", 4: "other_conv_fwd", 5: "other_conv_dgrad",
             6: "other_conv_wgrad", 7: "in_apply", 8: "in_bwd", 9: "pack_unpack_wprep"}


def launch_count():
    return int(L.lib().sscg_launch_count())


def prof_begin(tags=None):
    """tags: iterable of tag ids to bracket with CUDA events (None = all)."""
    mask = 0
    for t in (tags or []):
        mask |= 1 << t
    L.check(L.lib().sscg_prof_begin(mask), "sscg_prof_begin")


def prof_end():
    """-> ({tag name: (sum_ms, launches)}, complete flag)"""
    ms = (C.c_float * 16)()
    cnt = (C.c_int32 * 16)()
    rc = L.lib().sscg_prof_end(ms, cnt)
    if rc not in (0, 2):
        L.check(rc, "sscg_prof_end")
    return {TAG_NAMES.get(t, "tag%d" % t): (float(ms[t]), int(cnt[t])) for t in range(16) if cnt[t]}, rc == 0


def device_error():
    return L.lib().sscg_device_error()

"""Kernel-level parity cases: each function runs one C-ABI kernel configuration on cuda:0 and returns
(max_abs_err, ref_scale, tolerance) against a plain PyTorch fp32 computation of the same operator
(TF32 disabled) on the same bf16-rounded operands.  Used by tests/test_kernels_gpu.py and by
tools/kernel_probe.py (which isolates every case in a subprocess so a trap cannot hide the rest).
"""
import torch
import torch.nn.functional as F

import sscg_b200  # noqa: F401
from sscg_b200 import _lib as L
from sscg_b200 import geometry as G
from sscg_b200 import kernels as K

DEV = "cuda"


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)


def _bf(x):
    return x.to(torch.bfloat16).float()


def _fill_act(buf: K.ActBuf, x_nchw, pad_mode):
    """Write an NCHW fp32 tensor (already bf16-representable) into buf with halo (torch side)."""
    N, Cc, H, W = x_nchw.shape
    p = buf.pad
    if p:
        mode = "reflect" if pad_mode == L.PAD_REFLECT else "constant"
        xp = F.pad(x_nchw, (p, p, p, p), mode=mode)
    else:
        xp = x_nchw
    t = torch.zeros(N, buf.Hp, buf.Wp, buf.C, device=DEV)
    t[..., :Cc] = xp.permute(0, 2, 3, 1)
    buf.hi[: t.numel()].copy_(t.reshape(-1).to(torch.bfloat16))


def _wslab(w, transposed, mode, Cp, rows_pad, Kc, KH, KW, Co, Ci, split=False):
    ntaps = KH if mode in (1, 5) else KH * KW
    dst = torch.zeros(ntaps * rows_pad * Kc, dtype=torch.bfloat16, device=DEV)
    dst_lo = torch.zeros_like(dst) if split else None
    a = K.wprep_args(w, transposed, Co, Ci, KH, KW, mode, Cp, rows_pad, Kc, dst, dst_lo)
    K.run_wprep(a)
    return dst, dst_lo, a


def _result(got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale, tol


# ------------------------------------------------------------------------------------------------
def case_conv_fwd(N=2, H=16, W=16, Cin=64, Cout=64, k=3, stride=1, pad=1, reflect=True, explicit=True,
                  out_fp32=True, bias=False, act=L.ACT_NONE, stats=False, split=1, shift=0):
    """Conv2d forward, regular mode. explicit=True: halo materialised in the buffer (org 0);
    explicit=False: interior view + zero fill (org = -pad)."""
    _setup()
    x = torch.randn(N, Cin, H, W, device=DEV)
    w = torch.randn(Cout, Cin, k, k, device=DEV) * 0.05
    b = torch.randn(Cout, device=DEV) if bias else None
    if split == 1:
        x, w = _bf(x), _bf(w)
    Cp = G.pad_in_channels(Cin)
    Kc = G.round_up(Cp, 64)
    Co_pad = G.pad_out_channels(Cout)
    buf = K.ActBuf(N, H, W, Cp, pad if explicit else 0, DEV, split=(split == 3))
    if split == 3:
        K.pack_nchw(x.contiguous(), buf, L.PAD_REFLECT if reflect else L.PAD_ZERO)
    else:
        _fill_act(buf, x, L.PAD_REFLECT if reflect else L.PAD_ZERO)
    slab, slab_lo, _ = _wslab(w, False, 0, 0, Co_pad, Kc, k, k, Cout, Cin, split == 3)
    Ho, Wo = G.conv_out(H, k, stride, pad), G.conv_out(W, k, stride, pad)
    y = torch.zeros(N, Ho, Wo, Co_pad, device=DEV, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    st = K.stats_buffer(N, Co_pad, DEV) if stats else None      # binned plane-sum accumulators (zeroed)
    bias_pad = None
    if bias:
        bias_pad = torch.zeros(Co_pad, device=DEV)
        bias_pad[:Cout] = b
    table = G.taps_conv_fwd(k, k, stride, 0 if explicit else -pad)
    if shift:
        table = G.taps_rowshift_fwd(k, k, 0 if explicit else -pad)
    view = buf.view(interior=False) if explicit else buf.view(interior=True)
    a = K.conv_args(view, buf.lo_ptr(interior=not explicit), table, Kc, slab, slab_lo, k * k * Co_pad, Co_pad,
                    y.data_ptr(), out_fp32, (Ho * Wo * Co_pad, Wo * Co_pad, Co_pad), (0, 0), Ho, Wo, bias=bias_pad,
                    act=act, stats=st, split=split, shift_kw=(k if shift else 0), shift_base_mode=2)
    K.run_conv(a)
    torch.cuda.synchronize()
    if stats:                         # order-independent integer accumulation: a second launch reproduces every bit
        st1, y1 = st.clone(), y.clone()
        st.zero_()
        K.run_conv(a)
        torch.cuda.synchronize()
        assert torch.equal(st1, st) and torch.equal(y1, y), "conv_igemm statistics are not reproducible"
        st = K.stats_decode(st)
    xp = F.pad(x, (pad,) * 4, mode="reflect" if reflect else "constant") if pad else x
    ref = F.conv2d(xp, w, b, stride=stride)
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    elif act == L.ACT_LRELU:
        ref = F.leaky_relu(ref, 0.2)
    elif act == L.ACT_TANH:
        ref = torch.tanh(ref)
    got = y.float()[..., :Cout].permute(0, 3, 1, 2)
    err, scale, tol = _result(got, ref, 2e-4)
    if not out_fp32:
        tol = scale * 2.0 ** -8     # one bf16 rounding of the stored value
    if split == 3:
        tol = 1e-4
    if stats:
        src = got if out_fp32 else got
        s1 = src.sum(dim=(2, 3))
        s2 = (src * src).sum(dim=(2, 3))
        e1 = (st[:, :Cout, 0] - s1).abs().max().item() / max(1.0, s1.abs().max().item())
        e2 = (st[:, :Cout, 1] - s2).abs().max().item() / max(1.0, s2.abs().max().item())
        err = max(err, e1 * scale, e2 * scale)
    return err, scale, tol


def case_conv_window(N=2, H=16, W=16, Cin=3, Cout=64, k=7, stride=1, pad=3, reflect=True, pixel_row=False):
    """Row-window mode (small Cin): stem 7x7 / PatchGAN 4x4 s2 first layer.  pixel_row: the same layer through the
    pixel-row mode (plain 8-channel view, overlapping K-major rows, conv_igemm.cu RW)."""
    _setup()
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    w = _bf(torch.randn(Cout, Cin, k, k, device=DEV) * 0.05)
    Cp = G.pad_in_channels(Cin)
    if k * Cp <= 64 and 64 % Cp == 0:
        pass
    kwpad = G.round_up(k * Cp, 64)
    Co_pad = G.pad_out_channels(Cout)
    buf = K.ActBuf(N, H, W, Cp, pad, DEV)
    _fill_act(buf, x, L.PAD_REFLECT if reflect else L.PAD_ZERO)
    slab, _, _ = _wslab(w, False, 5 if pixel_row else 1, Cp, Co_pad, kwpad, k, k, Cout, Cin)
    Ho, Wo = G.conv_out(H, k, stride, pad), G.conv_out(W, k, stride, pad)
    y = torch.zeros(N, Ho, Wo, Co_pad, device=DEV)
    table = G.taps_conv_fwd_window(k, stride, 0)
    kw = dict(rw_pitch=2 * Cp, BN=64) if pixel_row else {}
    if pixel_row:
        assert kwpad == 8 * Cp
    a = K.conv_args(buf.view(interior=False) if pixel_row else buf.window_view(kwpad), None, table, kwpad, slab, None,
                    k * Co_pad, Co_pad, y.data_ptr(), True, (Ho * Wo * Co_pad, Wo * Co_pad, Co_pad), (0, 0), Ho, Wo, **kw)
    K.run_conv(a)
    torch.cuda.synchronize()
    xp = F.pad(x, (pad,) * 4, mode="reflect" if reflect else "constant")
    ref = F.conv2d(xp, w, None, stride=stride)
    return _result(y[..., :Cout].permute(0, 3, 1, 2), ref, 2e-4)


def case_convT_fwd(N=2, H=8, W=8, Cin=128, Cout=64):
    """ConvTranspose2d(k3, s2, p1, op1) forward as four gather phases."""
    _setup()
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    w = _bf(torch.randn(Cin, Cout, 3, 3, device=DEV) * 0.05)
    Co_pad = G.pad_out_channels(Cout)
    buf = K.ActBuf(N, H, W, Cin, 1, DEV)   # explicit (reflect) halo that must NOT be read
    _fill_act(buf, x, L.PAD_REFLECT)
    slab, _, _ = _wslab(w, True, 0, 0, Co_pad, Cin, 3, 3, Cout, Cin)
    Ho, Wo = 2 * H, 2 * W
    y = torch.zeros(N, Ho, Wo, Co_pad, device=DEV)
    table = G.taps_convT_fwd(3, 3, 2, 1)
    a = K.conv_args(buf.view(interior=True), None, table, Cin, slab, None, 9 * Co_pad, Co_pad, y.data_ptr(), True,
                    (Ho * Wo * Co_pad, Wo * Co_pad, Co_pad), (0, 0), Ho, Wo)
    K.run_conv(a)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=1)
    return _result(y[..., :Cout].permute(0, 3, 1, 2), ref, 2e-4)


def case_conv_dgrad(N=2, H=16, W=16, Cin=64, Cout=128, k=3, stride=1, pad=1, shift=0):
    """Gradient w.r.t. the (zero-padded, implicit halo) input of a Conv2d."""
    _setup()
    Ho, Wo = G.conv_out(H, k, stride, pad), G.conv_out(W, k, stride, pad)
    dy = _bf(torch.randn(N, Cout, Ho, Wo, device=DEV))
    w = _bf(torch.randn(Cout, Cin, k, k, device=DEV) * 0.05)
    Kc = G.round_up(Cout, 64)
    Ci_pad = G.pad_out_channels(Cin)
    buf = K.ActBuf(N, Ho, Wo, G.pad_in_channels(Cout), 0, DEV)
    _fill_act(buf, dy, L.PAD_ZERO)
    slab, _, _ = _wslab(w, False, 2, 0, Ci_pad, Kc, k, k, Cout, Cin)
    dx = torch.zeros(N, H, W, Ci_pad, device=DEV)
    table = G.taps_conv_dgrad(k, k, stride, -pad)
    kw = {}
    if shift:
        table = G.taps_rowshift_dgrad(k, k, -pad)
        kw = dict(shift_kw=k, shift_brow_step=-1, shift_base_mode=2, BN=min(Ci_pad, 64))
    a = K.conv_args(buf.view(interior=True), None, table, Kc, slab, None, k * k * Ci_pad, Ci_pad, dx.data_ptr(), True,
                    (H * W * Ci_pad, W * Ci_pad, Ci_pad), (0, 0), H, W, **kw)
    K.run_conv(a)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=stride, padding=pad)
    return _result(dx[..., :Cin].permute(0, 3, 1, 2), ref, 2e-4)


def case_conv_dgrad_flat(N=3, H=30, W=30, Cin=256, Cout=256, k=3):
    """Flattened tiling (SscgConvArgs.flat_*): gradient w.r.t. the halo-padded input of a stride-1 conv that reads an
    explicit halo (the residual-block convs) — dRaw sits in a buffer with a zero halo of k - 1, tiles are 128
    consecutive positions of the flattened padded output; vs conv_transpose2d (full correlation)."""
    _setup()
    dy = _bf(torch.randn(N, Cout, H, W, device=DEV))
    w = _bf(torch.randn(Cout, Cin, k, k, device=DEV) * 0.05)
    Kc = G.round_up(Cout, 64)
    Ci_pad = G.pad_out_channels(Cin)
    buf = K.ActBuf(N, H, W, G.pad_out_channels(Cout), k - 1, DEV)
    _fill_act(buf, dy, L.PAD_ZERO)
    slab, _, _ = _wslab(w, False, 2, 0, Ci_pad, Kc, k, k, Cout, Cin)
    Ho, Wo = H + k - 1, W + k - 1
    dx = torch.zeros(N, Ho, Wo, Ci_pad, device=DEV, dtype=torch.bfloat16)
    npx = N * buf.Hp * buf.Wp
    fview = L.make_view(buf.hi.data_ptr(), 1, 1, npx, buf.C, npx * buf.C, npx * buf.C, buf.C)
    table = G.taps_conv_dgrad(k, k, 1, -(k - 1))
    a = K.conv_args(fview, None, table, Kc, slab, None, k * k * Ci_pad, Ci_pad, dx.data_ptr(), False,
                    (Ho * Wo * Ci_pad, Wo * Ci_pad, Ci_pad), (0, 0), Ho, Wo, flat=(buf.Wp, buf.Hp * buf.Wp, N))
    K.run_conv(a)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy, w)
    err, scale, _ = _result(dx.float()[..., :Cin].permute(0, 3, 1, 2), ref, 0)
    return err, scale, scale * 2.0 ** -8


def case_conv_wgrad(N=2, H=16, W=16, Cin=64, Cout=128, k=3, stride=1, pad=1, ksplit=None):
    _setup()
    Ho, Wo = G.conv_out(H, k, stride, pad), G.conv_out(W, k, stride, pad)
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    dy = _bf(torch.randn(N, Cout, Ho, Wo, device=DEV))
    Cp = G.pad_in_channels(Cin)
    Kc = G.round_up(Cp, 64)
    Co_pad = G.round_up(Cout, 64)
    xb = K.ActBuf(N, H, W, Cp, 0, DEV)
    _fill_act(xb, x, L.PAD_ZERO)
    dyb = K.ActBuf(N, Ho, Wo, G.pad_in_channels(Cout), 0, DEV)
    _fill_act(dyb, dy, L.PAD_ZERO)
    dw = torch.zeros(k * k * Co_pad * Kc, device=DEV)
    table = G.taps_conv_fwd(k, k, stride, -pad)
    a = K.wgrad_args(dyb.view(interior=True), None, xb.view(interior=True), None, table, Kc, Co_pad, dw,
                     k * k * Co_pad, ksplit=ksplit)
    K.run_wgrad(a)
    torch.cuda.synchronize()
    dw1 = dw.clone()
    K.run_wgrad(a)                    # accumulates: second launch doubles every element exactly; fixed split order
    torch.cuda.synchronize()
    assert torch.equal(dw, 2 * dw1), "conv_wgrad is not reproducible / does not accumulate"
    dw.copy_(dw1)
    grad = torch.zeros(Cout, Cin, k, k, device=DEV)
    ua = K.wprep_args(grad, False, Cout, Cin, k, k, 0, 0, Co_pad, Kc, None)
    K.run_wgrad_unpack(ua, dw, grad)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, k, k), dy, stride=stride, padding=pad)
    return _result(grad, ref, 2e-4 * (N * Ho * Wo) ** 0.5)


def case_wgrad_window(N=2, H=16, W=16, Cin=3, Cout=64, k=7, pad=3, pixel_row=False):
    """Weight gradient of a small-Cin stem through the row-window view, or (pixel_row) through the plain view with
    overlapping MN-major rows (conv_wgrad.cu, RW; slab columns in sscg_wprep mode 5 order)."""
    _setup()
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    dy = _bf(torch.randn(N, Cout, H, W, device=DEV))
    Cp = G.pad_in_channels(Cin)
    kwpad = G.round_up(k * Cp, 64)
    Co_pad = G.round_up(Cout, 64)
    xb = K.ActBuf(N, H, W, Cp, pad, DEV)
    _fill_act(xb, x, L.PAD_REFLECT)
    dyb = K.ActBuf(N, H, W, Co_pad, 0, DEV)
    _fill_act(dyb, dy, L.PAD_ZERO)
    dw = torch.zeros(k * Co_pad * kwpad, device=DEV)
    table = G.taps_conv_fwd_window(k, 1, 0)
    if pixel_row:
        assert kwpad == 8 * Cp
        a = K.wgrad_args(dyb.view(interior=True), None, xb.view(interior=False), None, table, kwpad, Co_pad, dw, k * Co_pad,
                         rw_pitch=2 * Cp)
    else:
        a = K.wgrad_args(dyb.view(interior=True), None, xb.window_view(kwpad), None, table, kwpad, Co_pad, dw, k * Co_pad)
    K.run_wgrad(a)
    torch.cuda.synchronize()
    dw1 = dw.clone()
    K.run_wgrad(a)
    torch.cuda.synchronize()
    assert torch.equal(dw, 2 * dw1), "conv_wgrad (window) is not reproducible / does not accumulate"
    dw.copy_(dw1)
    grad = torch.zeros(Cout, Cin, k, k, device=DEV)
    ua = K.wprep_args(grad, False, Cout, Cin, k, k, 5 if pixel_row else 1, Cp, Co_pad, kwpad, None)
    K.run_wgrad_unpack(ua, dw, grad)
    torch.cuda.synchronize()
    xp = F.pad(x, (pad,) * 4, mode="reflect")
    ref = torch.nn.grad.conv2d_weight(xp, (Cout, Cin, k, k), dy, stride=1, padding=0)
    return _result(grad, ref, 2e-4 * (N * H * W) ** 0.5)


def case_wgrad7(N=2, H=16, W=40, Cout=21):
    """7x7 head weight gradient with the horizontal taps as GEMM columns (conv_wgrad7.cu): x = 64-channel activation with
    its reflect halo of 3, dY in a zero-haloed (6) buffer; accumulate semantics and bit reproducibility included."""
    _setup()
    Cin = 64
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    dy = _bf(torch.randn(N, Cout, H, W, device=DEV))
    xb = K.ActBuf(N, H, W, Cin, 3, DEV)
    _fill_act(xb, x, L.PAD_REFLECT)
    dyb = K.ActBuf(N, H, W, G.pad_out_channels(Cout), 6, DEV)
    _fill_act(dyb, dy, L.PAD_ZERO)
    dw = torch.zeros(7 * 64 * 448, device=DEV)
    a = K.wgrad7_args(xb, dyb, dw)
    K.run_wgrad(a)
    torch.cuda.synchronize()
    dw1 = dw.clone()
    K.run_wgrad(a)
    torch.cuda.synchronize()
    assert torch.equal(dw, 2 * dw1), "conv_wgrad7 is not reproducible / does not accumulate"
    dw.copy_(dw1)
    grad = torch.zeros(Cout, Cin, 7, 7, device=DEV)
    ua = K.wprep_args(grad, False, Cout, Cin, 7, 7, 1, 64, 64, 448, None)
    K.run_wgrad_unpack(ua, dw, grad)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(F.pad(x, (3,) * 4, mode="reflect"), (Cout, Cin, 7, 7), dy, stride=1, padding=0)
    return _result(grad, ref, 2e-4 * (N * H * W) ** 0.5)


def case_convT_bwd(N=2, H=8, W=8, Cin=128, Cout=64):
    """ConvTranspose2d dgrad (strided gather over dY) and wgrad (roles swapped)."""
    _setup()
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    w = _bf(torch.randn(Cin, Cout, 3, 3, device=DEV) * 0.05)
    dy = _bf(torch.randn(N, Cout, 2 * H, 2 * W, device=DEV))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2, padding=1, output_padding=1).backward(dy)
    Kc = G.round_up(Cout, 64)
    Ci_pad = G.pad_out_channels(Cin)
    dyb = K.ActBuf(N, 2 * H, 2 * W, G.pad_in_channels(Cout), 0, DEV)
    _fill_act(dyb, dy, L.PAD_ZERO)
    xb = K.ActBuf(N, H, W, Cin, 0, DEV)
    _fill_act(xb, x, L.PAD_ZERO)
    slab, _, _ = _wslab(w, True, 2, 0, Ci_pad, Kc, 3, 3, Cout, Cin)
    dx = torch.zeros(N, H, W, Ci_pad, device=DEV)
    table = G.taps_convT_dgrad(3, 3, 2, 1)
    a = K.conv_args(dyb.view(interior=True), None, table, Kc, slab, None, 9 * Ci_pad, Ci_pad, dx.data_ptr(), True,
                    (H * W * Ci_pad, W * Ci_pad, Ci_pad), (0, 0), H, W)
    K.run_conv(a)
    # wgrad: "dy" role = x (M = Cin), "x" role = dY (strided), slab [tap][Ci_pad][Kc = Cout]
    Ci_pad64 = G.round_up(Cin, 64)
    dw = torch.zeros(9 * Ci_pad64 * Kc, device=DEV)
    wa = K.wgrad_args(xb.view(interior=True), None, dyb.view(interior=True), None, table, Kc, Ci_pad64, dw,
                      9 * Ci_pad64)
    K.run_wgrad(wa)
    grad = torch.zeros(Cin, Cout, 3, 3, device=DEV)
    ua = K.wprep_args(grad, True, Cout, Cin, 3, 3, 2, 0, Ci_pad64, Kc, None)
    K.run_wgrad_unpack(ua, dw, grad)
    torch.cuda.synchronize()
    e1 = _result(dx[..., :Cin].permute(0, 3, 1, 2), xr.grad, 2e-4)
    e2 = _result(grad, wr.grad, 2e-4 * (N * H * W) ** 0.5)
    # normalise both to one verdict (ratio to tolerance)
    worst = max(e1[0] / e1[2], e2[0] / e2[2])
    return worst, 1.0, 1.0


def case_apply_fwd(N=2, H=12, W=12, Cc=64, pad=1, reflect=True, act=L.ACT_RELU, residual=False, norm=True):
    """InstanceNorm + activation (+ residual) + halo write vs torch."""
    _setup()
    raw = _bf(torch.randn(N, Cc, H, W, device=DEV) * 2 + 0.5)
    rawb = raw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    st = K.stats_encode(torch.stack([raw.sum(dim=(2, 3)), (raw * raw).sum(dim=(2, 3))], dim=-1))
    dst = K.ActBuf(N, H, W, Cc, pad, DEV)
    a = L.ApplyArgs()
    a.raw, a.raw_fp32 = rawb.data_ptr(), 0
    a.stats, a.eps = (st.data_ptr() if norm else None), 1e-5
    a.N, a.H, a.W, a.C = N, H, W, Cc
    a.act, a.slope, a.drop_seed = act, 0.2, 0
    resb = None
    if residual:
        res = _bf(torch.randn(N, Cc, H, W, device=DEV))
        resb = K.ActBuf(N, H, W, Cc, 1, DEV)
        _fill_act(resb, res, L.PAD_REFLECT)
        a.res = resb.view(interior=True)
    a.dst, a.dst_lo = dst.hi.data_ptr(), None
    a.pad, a.pad_mode = pad, (L.PAD_REFLECT if reflect else L.PAD_ZERO)
    K.run_apply(a)
    torch.cuda.synchronize()
    ref = F.instance_norm(raw, eps=1e-5) if norm else raw
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    elif act == L.ACT_LRELU:
        ref = F.leaky_relu(ref, 0.2)
    if residual:
        ref = ref + res
    if pad:
        ref = F.pad(ref, (pad,) * 4, mode="reflect" if reflect else "constant")
    got = dst.as_nhwc(interior=False).float().permute(0, 3, 1, 2)
    return _result(got, ref, 2e-2)


def case_apply_bwd(N=2, H=12, W=12, Cc=64, pad=1, act=L.ACT_RELU, skip=True, dz_bf16=False, stream_mode=2):
    """Backward of pad(act(IN(raw))) (+ skip gradient): dRaw vs autograd."""
    _setup()
    raw = _bf(torch.randn(N, Cc, H, W, device=DEV) * 2 + 0.5)
    dyp = _bf(torch.randn(N, Cc, H + 2 * pad, W + 2 * pad, device=DEV))
    sk = _bf(torch.randn(N, Cc, H, W, device=DEV)) if skip else None
    rr = raw.clone().requires_grad_(True)
    z = F.instance_norm(rr, eps=1e-5)
    y = F.relu(z) if act == L.ACT_RELU else (F.leaky_relu(z, 0.2) if act == L.ACT_LRELU else z)
    yp = F.pad(y, (pad,) * 4, mode="reflect") if pad else y
    loss = (yp * dyp).sum()
    if skip:
        loss = loss + (y * sk).sum()
    loss.backward()
    rawb = raw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    st = K.stats_encode(torch.stack([raw.sum(dim=(2, 3)), (raw * raw).sum(dim=(2, 3))], dim=-1))
    dypb = K.ActBuf(N, H, W, Cc, pad, DEV)
    t = dyp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dypb.hi[: t.numel()].copy_(t.reshape(-1))
    a = L.BwdArgs()
    a.raw, a.raw_fp32 = rawb.data_ptr(), 0
    a.stats, a.eps = st.data_ptr(), 1e-5
    a.N, a.H, a.W, a.C = N, H, W, Cc
    a.act, a.slope, a.drop_seed = act, 0.2, 0
    a.dyp, a.dyp_fp32 = dypb.view(interior=False), 0
    a.pad, a.pad_mode = pad, L.PAD_REFLECT if pad else L.PAD_NONE
    if skip:
        skb = sk.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        a.skip = L.make_view(skb.data_ptr(), N, H, W, Cc, H * W * Cc, W * Cc, Cc)
        a.skip_fp32 = 0
    dz = torch.zeros(N, H, W, Cc, device=DEV, dtype=torch.bfloat16 if dz_bf16 else torch.float32)
    bst = K.stats_buffer(N, Cc, DEV)
    a.dz, a.dz_fp32, a.dz_lo = dz.data_ptr(), 0 if dz_bf16 else 1, None
    a.bstats = bst.data_ptr()
    draw = torch.zeros(N, H, W, Cc, device=DEV, dtype=torch.bfloat16)
    L.lib().sscg_set_stream_norm(stream_mode)
    K.run_bwd_prep(a)
    K.run_bwd_apply(a, draw)
    torch.cuda.synchronize()
    first = (bst.clone(), draw.clone())
    bst.zero_()
    K.run_bwd_prep(a)                # order-independent integer accumulation: a second run reproduces every bit
    K.run_bwd_apply(a, draw)
    torch.cuda.synchronize()
    L.lib().sscg_set_stream_norm(2)
    assert torch.equal(first[0], bst) and torch.equal(first[1], draw), "in_bwd_prep / apply are not reproducible"
    # bf16 dZ (the fast-mode layout, served by the bulk-pipelined kernels) adds one bf16 rounding of values up to ~10
    return _result(draw.float().permute(0, 3, 1, 2), rr.grad, 5e-2 if dz_bf16 else 2e-2)


def case_stream_ab(N=3, H=16, W=64, Cc=256, pad=1, act=L.ACT_RELU, skip=True, residual=True, drop=True):
    """Bulk-pipelined InstanceNorm kernels (norm_stream.cuh) against the register-batched ones on the same
    inputs, all-bf16 layouts, dropout on: forward output, dZ and the folded total gradient must be bit-identical
    (same arithmetic; ReLU and the dropout scale on packed bf16 pairs are exact), dRaw within one bf16 ulp (FMA form);
    the plane sums may differ in summation order only."""
    _setup()
    lib = L.lib()
    raw = (torch.randn(N, H, W, Cc, device=DEV) * 2 + 0.5).to(torch.bfloat16)
    st = K.stats_encode(torch.stack([raw.float().sum(dim=(1, 2)), (raw.float() ** 2).sum(dim=(1, 2))], dim=-1))
    seed = 0x1234567 if drop else 0
    worst = 0.0
    # ---- forward ------------------------------------------------------------------------------
    outs = []
    resb = None
    if residual:
        resb = K.ActBuf(N, H, W, Cc, 1, DEV)
        resb.hi.copy_(torch.randn_like(resb.hi.float()).to(torch.bfloat16))
    for on in (0, 2):
        lib.sscg_set_stream_norm(on)
        dst = K.ActBuf(N, H, W, Cc, pad, DEV)
        a = L.ApplyArgs()
        a.raw, a.raw_fp32 = raw.data_ptr(), 0
        a.stats, a.eps = st.data_ptr(), 1e-5
        a.N, a.H, a.W, a.C = N, H, W, Cc
        a.act, a.slope, a.drop_seed = act, 0.2, seed
        if resb is not None:
            a.res = resb.view(interior=True)
        a.dst, a.dst_lo = dst.hi.data_ptr(), None
        a.pad, a.pad_mode = pad, (L.PAD_REFLECT if pad else L.PAD_NONE)
        K.run_apply(a)
        torch.cuda.synchronize()
        outs.append(dst.hi.clone())
    worst = max(worst, float((outs[0].float() - outs[1].float()).abs().max()))
    assert float(outs[1].float().abs().max()) > 0
    # ---- backward -----------------------------------------------------------------------------
    dypb = K.ActBuf(N, H, W, Cc, pad, DEV)
    dypb.hi.copy_(torch.randn_like(dypb.hi.float()).to(torch.bfloat16))
    skb = torch.randn(N, H, W, Cc, device=DEV).to(torch.bfloat16) if skip else None
    res = []
    for on in (0, 2):
        lib.sscg_set_stream_norm(on)
        a = L.BwdArgs()
        a.raw, a.raw_fp32 = raw.data_ptr(), 0
        a.stats, a.eps = st.data_ptr(), 1e-5
        a.N, a.H, a.W, a.C = N, H, W, Cc
        a.act, a.slope, a.drop_seed = act, 0.2, seed
        a.dyp, a.dyp_fp32 = dypb.view(interior=False), 0
        a.pad, a.pad_mode = pad, L.PAD_REFLECT if pad else L.PAD_NONE
        if skip:
            a.skip = L.make_view(skb.data_ptr(), N, H, W, Cc, H * W * Cc, W * Cc, Cc)
        gout = torch.zeros(N, H, W, Cc, device=DEV, dtype=torch.bfloat16)
        a.g_out, a.g_fp32 = gout.data_ptr(), 0
        dz = torch.zeros(N, H, W, Cc, device=DEV, dtype=torch.bfloat16)
        bst = K.stats_buffer(N, Cc, DEV)
        a.dz, a.dz_fp32, a.dz_lo = dz.data_ptr(), 0, None
        a.bstats = bst.data_ptr()
        K.run_bwd_prep(a)
        torch.cuda.synchronize()
        bst_own = bst.clone()
        if on == 2:
            bst.copy_(res[0][2])          # same plane sums for the second half, so dRaw can be compared bitwise
        draw = torch.zeros(N, H, W, Cc, device=DEV, dtype=torch.bfloat16)
        K.run_bwd_apply(a, draw)
        torch.cuda.synchronize()
        res.append((dz.clone(), gout.clone(), bst_own, draw.clone()))
    lib.sscg_set_stream_norm(2)
    for i in (0, 1):
        worst = max(worst, float((res[0][i].float() - res[1][i].float()).abs().max()))
    assert float(res[1][3].float().abs().max()) > 0
    # dRaw: the pipelined kernel evaluates rstd * (dZ - m1 - zhat * m2) as two FMAs with per-channel constants
    # (norm_stream.cuh) — same value up to fp32 rounding, i.e. at most one bf16 ulp apart after the final rounding,
    # and on a small fraction of the elements only
    d0, d1 = res[0][3].float(), res[1][3].float()
    ulp = d0.abs().clamp_min(1e-3) * 2.0 ** -7
    assert bool(((d0 - d1).abs() <= ulp).all()), float(((d0 - d1).abs() / ulp).max())
    assert float(((d0 - d1) != 0).float().mean()) < 0.02
    # plane sums: same terms, different summation order
    b0, b1 = K.stats_decode(res[0][2]), K.stats_decode(res[1][2])
    rel = float((b1 - b0).abs().max() / b0.abs().max().clamp_min(1e-6))
    assert rel < 1e-4, rel
    return worst, 1.0, 0.0


def case_pack_unpack(N=2, Cc=21, H=10, W=12, pad=3):
    _setup()
    x = torch.randn(N, Cc, H, W, device=DEV)
    Cp = G.pad_in_channels(Cc)
    buf = K.ActBuf(N, H, W, Cp, pad, DEV)
    K.pack_nchw(x.contiguous(), buf, L.PAD_REFLECT)
    torch.cuda.synchronize()
    ref = F.pad(_bf(x), (pad,) * 4, mode="reflect")
    got = buf.as_nhwc(interior=False).float()[..., :Cc].permute(0, 3, 1, 2)
    e1 = (got - ref).abs().max().item()
    lab = torch.randint(0, Cc, (N, 1, H, W), device=DEV)
    buf2 = K.ActBuf(N, H, W, Cp, pad, DEV)
    K.onehot_pack(lab, Cc, buf2, L.PAD_REFLECT)
    oh = torch.zeros(N, Cc, H, W, device=DEV).scatter_(1, lab, 1.0)
    ref2 = F.pad(oh, (pad,) * 4, mode="reflect")
    got2 = buf2.as_nhwc(interior=False).float()[..., :Cc].permute(0, 3, 1, 2)
    e2 = (got2 - ref2).abs().max().item()
    src = torch.randn(N, H, W, 32, device=DEV)
    dst = torch.zeros(N, Cc, H, W, device=DEV)
    K.unpack_nhwc(src, N, Cc, H, W, 32, dst)
    torch.cuda.synchronize()
    e3 = (dst - src[..., :Cc].permute(0, 3, 1, 2)).abs().max().item()
    return max(e1, e2, e3), 1.0, 1e-6


def case_seg_head(N=2, Cc=21, H=24, W=40):
    """Fused softmax + cross-entropy + argmax (forward and backward) vs torch."""
    _setup()
    from sscg_b200.losses import seg_head
    logits = (torch.randn(N, Cc, H, W, device=DEV) * 3).requires_grad_(True)
    labels = torch.randint(0, Cc, (N, 1, H, W), device=DEV)
    wprobe = torch.randn(N, Cc, H, W, device=DEV)
    loss, probs, am = seg_head(logits, labels)
    (loss * 1.7 + (probs * wprobe).sum()).backward()
    ref = logits.detach().clone().requires_grad_(True)
    rl = F.cross_entropy(ref, labels.squeeze(1))
    rp = F.softmax(ref, dim=1)
    (rl * 1.7 + (rp * wprobe).sum()).backward()
    e = max((probs - rp).abs().max().item(), abs(float(loss) - float(rl)),
            (logits.grad - ref.grad).abs().max().item() / max(1e-6, ref.grad.abs().max().item()) * 1e-1)
    exact = torch.equal(am, ref.detach().max(1)[1])
    # probabilities-only use (no labels): gradient must be the pure softmax Jacobian
    l2 = logits.detach().clone().requires_grad_(True)
    _, p2, _ = seg_head(l2, None)
    (p2 * wprobe).sum().backward()
    r2 = logits.detach().clone().requires_grad_(True)
    (F.softmax(r2, dim=1) * wprobe).sum().backward()
    e = max(e, (l2.grad - r2.grad).abs().max().item())
    # nn.CrossEntropyLoss label semantics: ignore_index pixels (-100 by default; 255 = VOC void) are left out of the mean
    # and of the gradient; int32 label maps are accepted; the loss itself is reproducible bit for bit
    for ign, dt in ((-100, torch.int64), (255, torch.int32)):
        lab_i = labels.clone()
        lab_i[torch.rand(lab_i.shape, device=DEV) < 0.3] = ign
        l3 = logits.detach().clone().requires_grad_(True)
        loss3, _, _ = seg_head(l3, lab_i.to(dt), ignore_index=ign)
        loss3b, _, _ = seg_head(l3.detach(), lab_i.to(dt), ignore_index=ign)
        assert torch.equal(loss3.detach(), loss3b), "seg_head loss is not reproducible"
        loss3.backward()
        r3 = logits.detach().clone().requires_grad_(True)
        rl3 = F.cross_entropy(r3, lab_i.squeeze(1), ignore_index=ign)
        rl3.backward()
        e = max(e, abs(float(loss3) - float(rl3)),
                (l3.grad - r3.grad).abs().max().item() / max(1e-9, r3.grad.abs().max().item()) * 1e-1)
    assert K.device_error() == 0
    # a label outside [0, C) that is not ignore_index raises the device error flag (torch device-asserts)
    bad = labels.clone()
    bad[0, 0, 0, 0] = Cc + 3
    seg_head(logits.detach(), bad)
    torch.cuda.synchronize()
    code = K.device_error()
    assert (code >> 16) & 0x7fff == 31, code
    return (e if exact else 1.0), 1.0, 2e-5


def _nexp_slab(w, mode, CoW):
    Co, Ci = w.shape[0], w.shape[1]
    NT = G.round_up(7 * CoW, 16)
    nn = ((Co if mode == 3 else Ci) + CoW - 1) // CoW
    dst = torch.zeros(7 * nn * NT * 64, dtype=torch.bfloat16, device=DEV)
    a = K.wprep_args(w, False, Co, Ci, 7, 7, mode, CoW, NT, 64, dst)
    K.run_wprep(a)
    return dst, nn


def case_conv7_nexp_fwd(N=2, H=20, W=40, Cout=21, act=L.ACT_NONE, bias=True):
    """N-expanded 7x7 head convolution (conv_nexp.cu) vs conv2d on the reflect-padded input."""
    _setup()
    Cin = 64
    x = _bf(torch.randn(N, Cin, H, W, device=DEV))
    w = _bf(torch.randn(Cout, Cin, 7, 7, device=DEV) * 0.05)
    b = torch.randn(Cout, device=DEV) if bias else None
    buf = K.ActBuf(N, H, W, Cin, 3, DEV)
    _fill_act(buf, x, L.PAD_REFLECT)
    CoW = 8 if Cout <= 8 else (16 if Cout <= 16 else (24 if Cout <= 24 else 32))
    slab, nn = _nexp_slab(w, 3, CoW)
    Cp = G.pad_out_channels(Cout)
    y = torch.zeros(N, H, W, Cp, device=DEV)
    bias_pad = None
    if bias:
        bias_pad = torch.zeros(max(Cp, CoW), device=DEV)
        bias_pad[:Cout] = b
    a = K.conv7_args(buf.hi.data_ptr(), Cin, N, buf.Hp, buf.Wp, slab, CoW, nn, 4, min(CoW, Cp), y.data_ptr(), True,
                     (H * W * Cp, W * Cp, Cp), bias=bias_pad, act=act)
    K.run_conv7(a)
    torch.cuda.synchronize()
    ref = F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w, b)
    if act == L.ACT_TANH:
        ref = torch.tanh(ref)
    return _result(y[..., :Cout].permute(0, 3, 1, 2), ref, 3e-4)


def case_conv7_nexp_dgrad(N=2, H=20, W=40, Cout=21):
    """Data gradient of the 7x7 head w.r.t. its halo-padded 64-channel input: N-expanded kernel on dRaw held in a
    zero-haloed (6) buffer vs conv_transpose2d."""
    _setup()
    Cin = 64
    dy = _bf(torch.randn(N, Cout, H, W, device=DEV))
    w = _bf(torch.randn(Cout, Cin, 7, 7, device=DEV) * 0.05)
    Cp = G.pad_out_channels(Cout)
    buf = K.ActBuf(N, H, W, Cp, 6, DEV)
    _fill_act(buf, dy, L.PAD_ZERO)
    slab, nn = _nexp_slab(w, 4, 32)
    Ho, Wo = H + 6, W + 6
    gx = torch.zeros(N, Ho, Wo, Cin, device=DEV, dtype=torch.bfloat16)
    a = K.conv7_args(buf.hi.data_ptr(), Cp, N, buf.Hp, buf.Wp, slab, 32, nn, Cp // 16, 32, gx.data_ptr(), False,
                     (Ho * Wo * Cin, Wo * Cin, Cin))
    K.run_conv7(a)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy, w)
    err, scale, _ = _result(gx.float().permute(0, 3, 1, 2), ref, 0)
    return err, scale, scale * 2.0 ** -8


def case_conv7_nexp_stem_dgrad(N=2, H=20, W=40, Cin=3):
    """Data gradient of the 7x7 STEM (Cin -> 64) w.r.t. its halo-padded network input, as the engine launches it
    (engine.py `nexp_stem`): dRaw (64 channels) in a zero-haloed (6) buffer, CoW = 8 / 16 / 24 / 32 output columns per
    horizontal tap for Cin <= 8 / 16 / 24 / 32, fp32 output with the channel pitch of gact[0]; vs conv_transpose2d."""
    _setup()
    Cout = 64
    dy = _bf(torch.randn(N, Cout, H, W, device=DEV))
    w = _bf(torch.randn(Cout, Cin, 7, 7, device=DEV) * 0.05)
    buf = K.ActBuf(N, H, W, Cout, 6, DEV)
    _fill_act(buf, dy, L.PAD_ZERO)
    CoW = 8 if Cin <= 8 else (16 if Cin <= 16 else (24 if Cin <= 24 else 32))
    slab, nn = _nexp_slab(w, 4, CoW)
    assert nn == 1
    Cpitch = G.pad_out_channels(G.pad_in_channels(Cin))
    Ho, Wo = H + 6, W + 6
    gx = torch.zeros(N, Ho, Wo, Cpitch, device=DEV)
    a = K.conv7_args(buf.hi.data_ptr(), Cout, N, buf.Hp, buf.Wp, slab, CoW, 1, 4, CoW, gx.data_ptr(), True,
                     (Ho * Wo * Cpitch, Wo * Cpitch, Cpitch), tag=5)
    K.run_conv7(a)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy, w)
    err, scale, _ = _result(gx[..., :Cin].permute(0, 3, 1, 2), ref, 0)
    return err, scale, 3e-4 * max(1.0, scale)


def case_lsgan(shape=(16, 1, 30, 30), target=1.0):
    """Fused LSGAN loss (forward mean + gradient) vs nn.MSELoss against a constant target."""
    _setup()
    from sscg_b200.losses import lsgan_loss
    x = torch.randn(*shape, device=DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = lsgan_loss(xa, target) * 0.37
    la.backward()
    lb = F.mse_loss(xb, torch.full_like(xb, target)) * 0.37
    lb.backward()
    torch.cuda.synchronize()
    e = max(abs(float(la) - float(lb)), float((xa.grad - xb.grad).abs().max()) * x.numel())
    return e, float(lb), 1e-5 * max(1.0, abs(float(lb)))


def case_l1(shape=(2, 3, 33, 35)):
    """Fused L1 loss vs nn.L1Loss (odd element count: exercises the scalar tail)."""
    _setup()
    from sscg_b200.losses import l1_loss
    x, y = torch.randn(*shape, device=DEV), torch.randn(*shape, device=DEV)
    x.view(-1)[:7] = y.view(-1)[:7]                        # exact ties: sign(0) = 0
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = l1_loss(xa, y) * 1.7
    la.backward()
    lb = F.l1_loss(xb, y) * 1.7
    lb.backward()
    torch.cuda.synchronize()
    e = max(abs(float(la) - float(lb)), float((xa.grad - xb.grad).abs().max()) * x.numel())
    return e, float(lb), 1e-5


def case_flat_adam(steps=6):
    """optim.FlatAdam (one sscg_adam_flat launch over a flat bucket) vs torch.optim.Adam(betas=(0.5, 0.999)) on the
    same gradients, through a LambdaLR decay and a state_dict round trip."""
    _setup()
    from sscg_b200.optim import FlatAdam
    from sscg_b200.step import FlatGrads
    shapes = [(7, 3, 3, 3), (7,), (5, 7, 1, 1), (1,), (33, 2)]
    pa = [torch.nn.Parameter(torch.randn(*sh, device=DEV)) for sh in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    fg = FlatGrads(pa)
    oa = FlatAdam(pa, fg, lr=2e-4, betas=(0.5, 0.999))
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    lam = lambda ep: 1.0 - max(0, ep - 2) / 6.0               # utils.py:434-441 with epochs=8, decay_epoch=2
    sa = torch.optim.lr_scheduler.LambdaLR(oa, lr_lambda=lam)
    sb = torch.optim.lr_scheduler.LambdaLR(ob, lr_lambda=lam)
    worst = 0.0
    for it in range(steps):
        for a, b in zip(pa, pb):
            g = torch.randn_like(b) * (0.1 + it)
            a.grad.copy_(g)
            b.grad = g.clone()
        oa.step()
        ob.step()
        sa.step()
        sb.step()
        if it == 2:                                            # checkpoint round trip (model.py:641-655)
            sd = oa.state_dict()
            oa.load_state_dict(sd)
        for a, b in zip(pa, pb):
            worst = max(worst, float((a.detach() - b.detach()).abs().max()))
    torch.cuda.synchronize()
    ma = oa.state[pa[0]]["exp_avg"]
    mb = ob.state[pb[0]]["exp_avg"]
    worst = max(worst, float((ma - mb).abs().max()))
    return worst, 1.0, 2e-6


CASES = {
    "seg_head_loss_c21": lambda: case_seg_head(),
    "seg_head_loss_c4": lambda: case_seg_head(N=3, Cc=4, H=17, W=9),
    # forward conv, regular mode
    "fwd_3x3_reflect_64": lambda: case_conv_fwd(),
    "fwd_3x3_reflect_256": lambda: case_conv_fwd(N=2, H=16, W=16, Cin=256, Cout=256),
    "fwd_3x3_zero_oob": lambda: case_conv_fwd(reflect=False, explicit=False),
    "fwd_3x3_s2": lambda: case_conv_fwd(Cin=64, Cout=128, stride=2, reflect=False, explicit=False),
    "fwd_4x4_s2": lambda: case_conv_fwd(Cin=64, Cout=128, k=4, stride=2, reflect=False, explicit=False),
    "fwd_4x4_s1_odd": lambda: case_conv_fwd(H=10, W=10, Cin=128, Cout=512, k=4, stride=1, reflect=False,
                                            explicit=False),
    "fwd_4x4_s1_cout1": lambda: case_conv_fwd(H=9, W=9, Cin=128, Cout=1, k=4, stride=1, reflect=False,
                                              explicit=False, bias=True),
    "fwd_7x7_head21": lambda: case_conv_fwd(Cin=64, Cout=21, k=7, pad=3, bias=True),
    "fwd_7x7_head3_tanh": lambda: case_conv_fwd(Cin=64, Cout=3, k=7, pad=3, bias=True, act=L.ACT_TANH),
    "fwd_bf16_stats": lambda: case_conv_fwd(Cin=128, Cout=256, out_fp32=False, stats=True),
    "fwd_bf16_lrelu_bias": lambda: case_conv_fwd(Cin=64, Cout=64, out_fp32=False, bias=True, act=L.ACT_LRELU),
    "fwd_wide_image": lambda: case_conv_fwd(N=1, H=8, W=256, Cin=64, Cout=64),
    "fwd_split3": lambda: case_conv_fwd(Cin=128, Cout=128, split=3),
    # row-shift mode (7x7, one row box shared by the 7 horizontal taps)
    "shift_fwd_head21": lambda: case_conv_fwd(N=2, H=12, W=200, Cin=64, Cout=21, k=7, pad=3, bias=True, shift=1),
    "shift_fwd_head3_tanh": lambda: case_conv_fwd(N=1, H=9, W=300, Cin=64, Cout=3, k=7, pad=3, bias=True,
                                                  act=L.ACT_TANH, shift=1),
    "shift_fwd_oob": lambda: case_conv_fwd(N=2, H=10, W=70, Cin=64, Cout=21, k=7, pad=3, reflect=False, explicit=False,
                                           shift=1),
    "shift_dgrad_head": lambda: case_conv_dgrad(N=2, H=12, W=150, Cin=64, Cout=21, k=7, pad=3, shift=1),
    "shift_dgrad_stem": lambda: case_conv_dgrad(N=2, H=10, W=140, Cin=3, Cout=64, k=7, pad=3, shift=1),
    # window mode
    "win_stem_c3": lambda: case_conv_window(),
    "win_stem_c21": lambda: case_conv_window(Cin=21),
    "win_stem_c1": lambda: case_conv_window(Cin=1),
    "pixrow_stem_c3": lambda: case_conv_window(pixel_row=True),
    "pixrow_stem_c1_wide": lambda: case_conv_window(N=3, H=20, W=200, Cin=1, pixel_row=True),     # ragged second tile
    "pixrow_stem_c3_256": lambda: case_conv_window(N=2, H=32, W=256, Cin=3, pixel_row=True),
    "pixrow_stem_c21": lambda: case_conv_window(N=2, H=20, W=200, Cin=21, pixel_row=True),       # 3 channel groups
    "pixrow_stem_c12": lambda: case_conv_window(N=1, H=16, W=128, Cin=12, pixel_row=True),       # 2 groups
    "pixrow_stem_c30": lambda: case_conv_window(N=2, H=16, W=140, Cin=30, pixel_row=True),       # 4 groups
    "win_d0_c3": lambda: case_conv_window(Cin=3, k=4, stride=2, pad=1, reflect=False),
    "win_d0_c21": lambda: case_conv_window(Cin=21, k=4, stride=2, pad=1, reflect=False),
    # transposed conv
    "convT_fwd": lambda: case_convT_fwd(),
    "convT_fwd_256": lambda: case_convT_fwd(Cin=256, Cout=128),
    "convT_bwd": lambda: case_convT_bwd(),
    # 160 tiles on 148 SMs: the last, partial wave runs as half-N tiles (tail_from in conv_igemm.cu)
    "conv_tail_split_stats": lambda: case_conv_fwd(N=5, H=64, W=64, Cin=64, Cout=256, out_fp32=False, stats=True),
    "conv_tail_split_bias_act": lambda: case_conv_fwd(N=5, H=64, W=64, Cin=64, Cout=256, bias=True, act=L.ACT_LRELU),
    # dgrad
    "dgrad_3x3_s1": lambda: case_conv_dgrad(),
    "dgrad_3x3_s2": lambda: case_conv_dgrad(stride=2),
    "dgrad_4x4_s2": lambda: case_conv_dgrad(k=4, stride=2),
    "dgrad_4x4_s1_odd": lambda: case_conv_dgrad(H=10, W=10, k=4, stride=1),
    "dgrad_7x7_small_cout": lambda: case_conv_dgrad(Cin=64, Cout=21, k=7, pad=3),
    "dgrad_to_c3": lambda: case_conv_dgrad(Cin=3, Cout=64, k=7, pad=3),
    "dgrad_flat_3x3_c256": lambda: case_conv_dgrad_flat(),
    "dgrad_flat_3x3_c64_odd": lambda: case_conv_dgrad_flat(N=2, H=13, W=21, Cin=64, Cout=128),
    "dgrad_flat_res_shape": lambda: case_conv_dgrad_flat(N=16, H=64, W=64),     # 576 tiles: four full waves
    # wgrad
    "wgrad_3x3_s1": lambda: case_conv_wgrad(),
    "wgrad_3x3_s1_256": lambda: case_conv_wgrad(Cin=256, Cout=256),
    "wgrad_3x3_s2": lambda: case_conv_wgrad(stride=2),
    "wgrad_4x4_s2": lambda: case_conv_wgrad(k=4, stride=2),
    "wgrad_4x4_s1_odd": lambda: case_conv_wgrad(H=10, W=10, k=4, stride=1),
    "wgrad_cout21": lambda: case_conv_wgrad(Cin=64, Cout=21, k=7, pad=3),
    "wgrad_ksplit1": lambda: case_conv_wgrad(ksplit=1),
    "wgrad_window_c3": lambda: case_wgrad_window(),
    "wgrad_window_c21": lambda: case_wgrad_window(Cin=21),            # Kc = 192 -> one 192-wide tile
    "wgrad_pixrow_c3": lambda: case_wgrad_window(pixel_row=True),
    "wgrad_pixrow_c21": lambda: case_wgrad_window(N=2, H=20, W=200, Cin=21, pixel_row=True),     # 3 groups, ragged rows
    "wgrad_pixrow_c12": lambda: case_wgrad_window(N=3, H=16, W=128, Cin=12, pixel_row=True),
    "wgrad_pixrow_c30": lambda: case_wgrad_window(N=1, H=24, W=70, Cin=30, pixel_row=True),
    "wgrad_window_c64_wide": lambda: case_wgrad_window(Cin=64, Cout=21),   # head conv: Kc = 448 -> 448-wide tile
    "wgrad_window_c64_wide_big": lambda: case_wgrad_window(N=2, H=24, W=40, Cin=64, Cout=3),
    "wgrad7_head_c21": lambda: case_wgrad7(),
    "wgrad7_head_c3": lambda: case_wgrad7(N=3, H=24, W=72, Cout=3),            # Cy = 16 (SWIZZLE_32B segments)
    "wgrad7_head_c19_wide": lambda: case_wgrad7(N=2, H=32, W=250, Cout=19),    # 4 column blocks, the last one partial
    "wgrad7_head_c4_many_units": lambda: case_wgrad7(N=20, H=64, W=64, Cout=4),
    # elementwise
    "apply_fwd_relu_reflect": lambda: case_apply_fwd(),
    "apply_fwd_res": lambda: case_apply_fwd(act=L.ACT_NONE, residual=True),
    "apply_fwd_lrelu_zero": lambda: case_apply_fwd(Cc=128, pad=1, reflect=False, act=L.ACT_LRELU),
    "apply_fwd_c512": lambda: case_apply_fwd(Cc=512, H=7, W=7, pad=0, act=L.ACT_LRELU),
    "apply_bwd_relu": lambda: case_apply_bwd(),
    "apply_bwd_nopad": lambda: case_apply_bwd(pad=0, act=L.ACT_LRELU, skip=False),
    "apply_bwd_pad3": lambda: case_apply_bwd(pad=3, act=L.ACT_RELU, skip=False),
    "apply_bwd_generic_big": lambda: case_apply_bwd(N=16, H=64, W=64, Cc=256, pad=1, stream_mode=0),
    "apply_bwd_generic_lrelu_nopad": lambda: case_apply_bwd(N=3, H=31, W=31, Cc=512, pad=0, act=L.ACT_LRELU, skip=False,
                                                            stream_mode=0),
    "apply_bwd_stream_res_shape": lambda: case_apply_bwd(N=16, H=64, W=64, Cc=256, pad=1, dz_bf16=True),
    # bulk-pipelined (cp.async.bulk ring) variants: eligible all-bf16 shapes
    "apply_fwd_stream_res": lambda: case_apply_fwd(N=3, H=16, W=64, Cc=256, pad=1, act=L.ACT_NONE, residual=True),
    "apply_fwd_stream_pad3": lambda: case_apply_fwd(N=2, H=20, W=128, Cc=64, pad=3),
    "apply_fwd_stream_c512": lambda: case_apply_fwd(N=2, H=31, W=31, Cc=512, pad=0, act=L.ACT_LRELU),
    "apply_bwd_stream_res": lambda: case_apply_bwd(N=3, H=16, W=64, Cc=256, pad=1, dz_bf16=True),
    "apply_bwd_stream_pad3": lambda: case_apply_bwd(N=2, H=20, W=128, Cc=64, pad=3, skip=False, dz_bf16=True),
    "apply_bwd_stream_c512": lambda: case_apply_bwd(N=2, H=31, W=31, Cc=512, pad=0, act=L.ACT_LRELU, skip=False,
                                                    dz_bf16=True),
    "apply_bwd_stream_many_samples": lambda: case_apply_bwd(N=40, H=8, W=64, Cc=128, pad=1, dz_bf16=True),
    "stream_ab_res": lambda: case_stream_ab(),
    "stream_ab_pad3": lambda: case_stream_ab(N=2, H=20, W=128, Cc=64, pad=3, skip=False, residual=False),
    "stream_ab_nopad_lrelu": lambda: case_stream_ab(N=2, H=31, W=31, Cc=512, pad=0, act=L.ACT_LRELU, skip=False,
                                                    residual=False, drop=False),
    "conv7_nexp_fwd_c21": lambda: case_conv7_nexp_fwd(),
    "conv7_nexp_fwd_c3_tanh": lambda: case_conv7_nexp_fwd(Cout=3, act=L.ACT_TANH),
    "conv7_nexp_fwd_c4_odd": lambda: case_conv7_nexp_fwd(N=3, H=13, W=23, Cout=4, bias=False),
    "conv7_nexp_fwd_big": lambda: case_conv7_nexp_fwd(N=2, H=128, W=256, Cout=19),
    "conv7_nexp_dgrad_c21": lambda: case_conv7_nexp_dgrad(),
    "conv7_nexp_dgrad_c3": lambda: case_conv7_nexp_dgrad(Cout=3),
    "conv7_nexp_dgrad_big": lambda: case_conv7_nexp_dgrad(N=2, H=128, W=256, Cout=20),
    "conv7_nexp_stem_dgrad_c3": lambda: case_conv7_nexp_stem_dgrad(),                       # CoW = 8
    "conv7_nexp_stem_dgrad_c1": lambda: case_conv7_nexp_stem_dgrad(N=3, H=13, W=23, Cin=1),  # CoW = 8, ACDC input
    "conv7_nexp_stem_dgrad_c21": lambda: case_conv7_nexp_stem_dgrad(Cin=21),                 # CoW = 24, label-map input
    "conv7_nexp_stem_dgrad_c19_big": lambda: case_conv7_nexp_stem_dgrad(N=2, H=128, W=256, Cin=19),
    "lsgan_real": lambda: case_lsgan(),
    "lsgan_fake_odd": lambda: case_lsgan(shape=(3, 1, 7, 5), target=0.0),
    "l1_loss": lambda: case_l1(),
    "l1_loss_big": lambda: case_l1(shape=(4, 3, 64, 64)),
    "flat_adam": lambda: case_flat_adam(),
    "pack_unpack": lambda: case_pack_unpack(),
}

"""CPU: tap tables of the implicit-GEMM convolutions against brute-force enumeration of
torch.nn.Conv2d / ConvTranspose2d index arithmetic (reference arch/ops.py:40-57 shapes)."""
import itertools

import pytest
import torch
import torch.nn.functional as F

import sscg_b200  # noqa: F401
from sscg_b200 import geometry as G


def _apply_table(table, x, w_taps, Hout, Wout):
    """Dense evaluation of a tap table on a single-channel image with per-tap scalar weights."""
    H, W = x.shape
    y = torch.zeros(Hout, Wout, dtype=torch.float64)
    os_ = 2 if table.n_phases == 4 else 1
    for p in range(table.n_phases):
        ph, pw = (p >> 1, p & 1) if table.n_phases == 4 else (0, 0)
        for (dh, dw, brow) in table.taps[table.phase_start[p]:table.phase_start[p + 1]]:
            for i in range((Hout - ph + os_ - 1) // os_):
                for j in range((Wout - pw + os_ - 1) // os_):
                    hi = i * table.stride + dh + table.org_h
                    wi = j * table.stride + dw + table.org_w
                    if 0 <= hi < H and 0 <= wi < W:
                        y[i * os_ + ph, j * os_ + pw] += x[hi, wi] * w_taps[brow]
    return y


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (3, 2, 1), (4, 2, 1), (4, 1, 1), (7, 1, 3)])
def test_conv_fwd_and_dgrad_tables(k, s, p):
    torch.manual_seed(0)
    H, W = 9 if s == 1 else 10, 11 if s == 1 else 12
    x = torch.randn(H, W, dtype=torch.float64)
    w = torch.randn(k, k, dtype=torch.float64)
    ref = F.conv2d(x[None, None], w[None, None], stride=s, padding=p)[0, 0]
    y = _apply_table(G.taps_conv_fwd(k, k, s, -p), x, w.reshape(-1), ref.shape[0], ref.shape[1])
    assert torch.allclose(y, ref, atol=1e-12)
    dy = torch.randn_like(ref)
    dx_ref = torch.nn.grad.conv2d_input((1, 1, H, W), w[None, None], dy[None, None], stride=s, padding=p)[0, 0]
    dx = _apply_table(G.taps_conv_dgrad(k, k, s, -p), dy, w.reshape(-1), H, W)
    assert torch.allclose(dx, dx_ref, atol=1e-12)


def test_convT_tables():
    torch.manual_seed(1)
    x = torch.randn(5, 6, dtype=torch.float64)
    w = torch.randn(3, 3, dtype=torch.float64)
    ref = F.conv_transpose2d(x[None, None], w[None, None], stride=2, padding=1, output_padding=1)[0, 0]
    assert ref.shape == (G.convT_out(5, 3, 2, 1, 1), G.convT_out(6, 3, 2, 1, 1))
    y = _apply_table(G.taps_convT_fwd(3, 3, 2, 1), x, w.reshape(-1), 10, 12)
    assert torch.allclose(y, ref, atol=1e-12)
    dy = torch.randn_like(ref)
    xr = x.clone().requires_grad_(True)
    F.conv_transpose2d(xr[None, None], w[None, None], stride=2, padding=1, output_padding=1).backward(dy[None, None])
    dx = _apply_table(G.taps_convT_dgrad(3, 3, 2, 1), dy, w.reshape(-1), 5, 6)
    assert torch.allclose(dx, xr.grad, atol=1e-12)


def test_explicit_halo_tables_and_padding_helpers():
    # explicit halo (org = 0): taps index the padded buffer directly
    t = G.taps_conv_fwd(3, 3, 1, 0)
    assert (t.org_h, t.org_w, len(t.taps)) == (0, 0, 9)
    t = G.taps_conv_fwd_window(7, 1, 0)
    assert [tp[2] for tp in t.taps] == list(range(7)) and all(tp[1] == 0 for tp in t.taps)
    assert [G.pad_out_channels(c) for c in (1, 3, 21, 64, 128, 256, 512)] == [16, 16, 32, 64, 128, 256, 512]
    assert [G.pad_in_channels(c) for c in (1, 3, 4, 19, 20, 21, 64)] == [8, 8, 8, 24, 24, 24, 64]
    assert G.pick_tile(64) == (2, 64) and G.pick_tile(31) == (4, 32) and G.pick_tile(256) == (1, 128)
    assert G.pick_tile(30, 64) == (2, 32)
    assert G.pick_tile(66, 128, 66) in ((16, 8), (8, 16))       # 45 tiles instead of 66
    assert G.pick_tile(64, 128, 64) == (2, 64) and G.pick_tile(256, 128, 256) == (1, 128)
    # every dgrad phase of the 4x4 stride-2 PatchGAN conv has exactly 4 taps; 3x3 stride-2 has 1/2/2/4
    assert G.taps_conv_dgrad(4, 4, 2, -1).phase_start == [0, 4, 8, 12, 16]
    ps = G.taps_conv_dgrad(3, 3, 2, -1).phase_start
    assert sorted(ps[i + 1] - ps[i] for i in range(4)) == [1, 2, 2, 4]

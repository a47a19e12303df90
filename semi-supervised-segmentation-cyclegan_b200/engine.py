"""Execution engine: turns a ResNet-block generator or an n-layer PatchGAN (module trees with the
reference's layout, reference arch/generators.py:65-95, arch/discriminators.py:42-63) into a list
of fused *stages* and runs their forward / backward through the C ABI (libsscg_b200.so).

A stage = one convolution-like GEMM + what follows it up to the next convolution's input buffer:

    act[i] --conv_igemm--> raw[i] (+ InstanceNorm sums in the epilogue)
           --in_apply----> act[i+1]   (normalise, ReLU/LeakyReLU, dropout, residual add, halo)

Stages without InstanceNorm (PatchGAN stem/tail, generator head) apply bias + activation in the
GEMM epilogue and write act[i+1] (or the fp32 output) directly.

Backward per stage:  in_bwd_prep (+ in_bwd_apply)  ->  dRaw;  conv_wgrad -> weight-gradient slab;
conv_igemm with the dgrad tap table -> gradient w.r.t. act[i].

Precision modes: "bf16" (fast; bf16 operands and storage, fp32 accumulate/statistics) and
"bf16x3" (parity; operands carried as hi+lo bf16 planes, raw outputs and gradients in fp32).
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib as L
from . import geometry as G
from . import kernels as K


@dataclass
class StageSpec:
    kind: str                 # 'conv' | 'window' | 'convT'
    k: int
    stride: int
    pad: int
    in_halo: int              # explicit halo carried by the input buffer (0 = implicit zero fill)
    in_reflect: bool          # halo content of the input buffer
    Cin: int
    Cout: int
    norm: bool
    act: int
    weight: torch.nn.Parameter
    bias: Optional[torch.nn.Parameter]
    dropout: bool = False
    residual_from: Optional[int] = None   # index of the act buffer added after normalisation
    final: bool = False                   # writes the fp32 NHWC network output
    name: str = ""


def _conv_geom(spec: StageSpec, Hin, Win):
    if spec.kind == "convT":
        return G.convT_out(Hin, spec.k, spec.stride, spec.pad, 1), G.convT_out(Win, spec.k, spec.stride, spec.pad, 1)
    return G.conv_out(Hin, spec.k, spec.stride, spec.pad), G.conv_out(Win, spec.k, spec.stride, spec.pad)


class StageWeights:
    """bf16 GEMM operand slabs + fp32 wgrad accumulator of one stage (shape independent)."""

    def __init__(self, spec: StageSpec, split: bool, need_dgrad: bool, device, first: bool = False):
        s = spec
        self.spec = s
        self.split = split
        k = s.k
        self.transposed = s.kind == "convT"
        # channel pitch of the stage input: the packed network input is padded to 8 channels, internal
        # activations carry the producing GEMM's N padding
        self.Cp_in = G.pad_in_channels(s.Cin) if first else G.pad_out_channels(s.Cin)
        # stride-1 stem over <= 32 input channels in bf16 mode: pixel-row forward (conv_igemm.cu, RW): the forward slab is
        # ordered (channel group, kw, channel in group); weight / data gradients keep the row-window order
        self.pixel_row = (s.kind == "window" and not split and s.stride == 1 and self.Cp_in in (8, 16, 24, 32) and k <= 8
                          and G.pad_out_channels(s.Cout) == 64 and os.environ.get("SSCG_PIXEL_ROW", "1") != "0")
        if s.kind == "window":
            self.Kc = G.round_up(k * self.Cp_in, 64)
            self.ntaps_fwd = k
            fmode = 1
        else:
            self.Kc = G.round_up(self.Cp_in, 64)
            self.ntaps_fwd = k * k
            fmode = 0
        self.Co_pad = G.pad_out_channels(s.Cout)          # forward N / channel pitch of raw
        self.Co_pitch = self.Co_pad
        bf = torch.bfloat16
        self.w_fwd = torch.zeros(self.ntaps_fwd * self.Co_pad * self.Kc, dtype=bf, device=device)
        self.w_fwd_lo = torch.zeros_like(self.w_fwd) if split else None
        if self.pixel_row:
            assert self.Kc == 8 * self.Cp_in          # 64 K-columns per group of 8 channels
        self.prep_fwd = K.wprep_args(s.weight, self.transposed, s.Cout, s.Cin, k, k, 5 if self.pixel_row else fmode,
                                     self.Cp_in, self.Co_pad, self.Kc, self.w_fwd, self.w_fwd_lo)
        # dgrad: rows = input channels (padded to a legal N tile), K = output-channel pitch
        self.need_dgrad = need_dgrad
        self.Ci_pad = G.pad_out_channels(self.Cp_in)
        self.Kc_d = G.round_up(self.Co_pitch, 64)
        if need_dgrad:
            self.w_dg = torch.zeros(k * k * self.Ci_pad * self.Kc_d, dtype=bf, device=device)
            self.w_dg_lo = torch.zeros_like(self.w_dg) if split else None
            self.prep_dg = K.wprep_args(s.weight, self.transposed, s.Cout, s.Cin, k, k, 2, 0, self.Ci_pad, self.Kc_d,
                                        self.w_dg, self.w_dg_lo)
        # wgrad accumulator (fp32) and its unpack descriptor
        if self.transposed:
            # roles swapped: rows = Cin (M), K axis = Cout
            self.wg_rows = G.round_up(s.Cin, 64)
            self.wg_Kc = G.round_up(self.Co_pitch, 64)
            self.wg_taps = k * k
            wmode = 2
        else:
            self.wg_rows = G.round_up(s.Cout, 64)
            self.wg_Kc = self.Kc
            self.wg_taps = self.ntaps_fwd
            wmode = 5 if (self.pixel_row and os.environ.get("SSCG_PIXEL_ROW_WG", "1") != "0") else fmode
        self.wg_pixel_row = (wmode == 5)       # weight-gradient slab in pixel-row column order (conv_wgrad.cu, RW)
        # 7x7 head conv (64 input channels, explicit halo): its weight gradient reads the activation as
        # a row window of k*64 contiguous elements, so one CTA handles a whole filter row and dY is
        # fetched once per row instead of once per tap
        self.wg_window = (s.kind == "conv" and s.in_halo > 0 and s.stride == 1 and self.Cp_in == 64
                          and k * 64 <= 448 and k > 3 and not split)
        if self.wg_window:
            self.wg_Kc = k * 64
            self.wg_taps = k
            wmode = 1
        self.dw = torch.zeros(self.wg_taps * self.wg_rows * self.wg_Kc, dtype=torch.float32, device=device)
        self.unpack_wg = K.wprep_args(None, self.transposed, s.Cout, s.Cin, k, k, wmode, self.Cp_in, self.wg_rows,
                                      self.wg_Kc, None)
        self.bias_pad = None
        if s.bias is not None and not s.norm:
            self.bias_pad = torch.zeros(self.Co_pad, dtype=torch.float32, device=device)
        # 7x7 head (64 -> <= 32 channels) in bf16 mode: N-expanded kernel (csrc/conv_nexp.cu) for the forward and the
        # data gradient; the seven horizontal taps are GEMM columns, shift-added in the epilogue
        self.nexp = (s.kind == "conv" and k == 7 and s.stride == 1 and s.in_halo == 3 and self.Cp_in == 64
                     and s.Cout <= 32 and not split and not s.norm)
        # 7x7 stem (small Cin -> 64): only its data gradient (64 -> Cin columns) takes the N-expanded kernel
        self.nexp_stem = (first and s.kind == "window" and k == 7 and s.stride == 1 and s.in_halo == 3 and s.Cout == 64
                          and self.Co_pitch == 64 and s.Cin <= 32 and not split and s.norm)
        if self.nexp_stem:
            self.nx_CoW = 8 if s.Cin <= 8 else (16 if s.Cin <= 16 else (24 if s.Cin <= 24 else 32))
            nt_d = G.round_up(7 * self.nx_CoW, 16)
            self.nx_dg_tiles = 1
            self.w_nx_dg = torch.zeros(7 * nt_d * 64, dtype=bf, device=device)
            self.prep_nx_dg = K.wprep_args(s.weight, False, s.Cout, s.Cin, 7, 7, 4, self.nx_CoW, nt_d, 64, self.w_nx_dg)
        if self.nexp:
            self.nx_CoW = 8 if s.Cout <= 8 else (16 if s.Cout <= 16 else (24 if s.Cout <= 24 else 32))
            nt_f = G.round_up(7 * self.nx_CoW, 16)
            self.w_nx = torch.zeros(7 * nt_f * 64, dtype=bf, device=device)
            self.prep_nx = K.wprep_args(s.weight, False, s.Cout, s.Cin, 7, 7, 3, self.nx_CoW, nt_f, 64, self.w_nx)
            nt_d = G.round_up(7 * 32, 16)
            self.nx_dg_tiles = s.Cin // 32
            self.w_nx_dg = torch.zeros(7 * self.nx_dg_tiles * nt_d * 64, dtype=bf, device=device)
            self.prep_nx_dg = K.wprep_args(s.weight, False, s.Cout, s.Cin, 7, 7, 4, 32, nt_d, 64, self.w_nx_dg)

    def prep_descs(self):
        """Slab descriptors of this stage with the source pointer refreshed (parameters may have been re-allocated)."""
        out = [self.prep_fwd]
        if self.need_dgrad:
            out.append(self.prep_dg)
        if self.nexp:
            out += [self.prep_nx, self.prep_nx_dg]
        if self.nexp_stem:
            out.append(self.prep_nx_dg)
        for d in out:
            d.w = self.spec.weight.data_ptr()
        return out

    def prepare(self):
        for d in self.prep_descs():
            K.run_wprep(d)
        if self.bias_pad is not None:
            self.bias_pad[: self.spec.Cout].copy_(self.spec.bias.detach())


class Ctx:
    """Per-call activation storage (saved for backward)."""

    def __init__(self):
        self.act: List[K.ActBuf] = []
        self.raw: List[Optional[K.ActBuf]] = []
        self.stats = None
        self.stat_off: List[int] = []
        self.out = None
        self.drop_seed = 0
        self.busy = False


class NetPlan:
    """Stage list + buffers of one network for a fixed input shape (N, H, W)."""

    def __init__(self, specs: List[StageSpec], weights: List[StageWeights], N, H, W, precision, device,
                 need_input_grad_capable=True):
        self.specs, self.weights = specs, weights
        self.N, self.H, self.W = N, H, W
        self.split = 3 if precision == "bf16x3" else 1
        self.device = device
        self.geom = []
        h, w = H, W
        for s in specs:
            ho, wo = _conv_geom(s, h, w)
            self.geom.append((h, w, ho, wo))
            h, w = ho, wo
        self.Hout, self.Wout = h, w
        self.Cout = specs[-1].Cout
        self.ctx_pool: List[Ctx] = []
        self.drop_ctr = None      # optional device int64 counter mixed into dropout seeds (CUDA-graph replays)
        # scratch of the fixed-order split-K reduction of the weight gradients (all of a plan's wgrad launches are
        # ordered on one stream)
        self.ws_wg, self.ws_wg7 = K.WsPool(device), K.WsPool(device)
        self.overlap_wgrad = False  # side-stream wgrad: measured no gain on B200 (power-capped, GEMMs contend); kept as an option
        self._scratch_ready = False
        self._args_cache = {}

    def _direct(self, s: StageSpec) -> bool:
        """Un-normalised stage whose GEMM epilogue writes its consumer's buffer directly (bias +
        activation fused).  In bf16x3 mode intermediate stages go through raw(fp32) + in_apply so the
        consumer gets hi/lo planes."""
        return (not s.norm) and (s.final or self.split == 1)

    def _rowshift_ok(self, s: StageSpec, n_cols: int) -> bool:
        """7x7 stride-1 convolutions in bf16 mode run in row-shift mode (see SscgConvArgs.shift_kw) when
        the GEMM's N fits the 16/32-column tiles of that mode (dgrad splits wider N into 32-column tiles)."""
        return (self.split == 1 and s.k == 7 and s.stride == 1 and s.kind in ("conv", "window")
                and n_cols in (16, 32))

    # ------------------------------------------------------------------ buffers
    def _new_ctx(self) -> Ctx:
        c = Ctx()
        N, sp = self.N, self.split == 3
        nstat = 0
        for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
            hin, win, ho, wo = self.geom[i]
            c.act.append(K.ActBuf(N, hin, win, wt.Cp_in, s.in_halo, self.device, split=sp))
            if s.norm:
                c.raw.append(K.ActBuf(N, ho, wo, wt.Co_pitch, 0, self.device, fp32=sp))
                c.stat_off.append(nstat)
                nstat += N * wt.Co_pitch * 2 * L.SSCG_STAT_WORDS
            elif not self._direct(s):
                c.raw.append(K.ActBuf(N, ho, wo, wt.Co_pitch, 0, self.device, fp32=True))
                c.stat_off.append(-1)
            else:
                c.raw.append(None)
                c.stat_off.append(-1)
        c.stats = torch.zeros(max(nstat, 2), dtype=torch.int64, device=self.device)      # binned plane sums
        last = self.weights[-1]
        c.out = K.ActBuf(N, self.Hout, self.Wout, last.Co_pitch, 0, self.device, fp32=True)
        return c

    def acquire_ctx(self) -> Ctx:
        for c in self.ctx_pool:
            if not c.busy:
                c.busy = True
                return c
        c = self._new_ctx()
        c.busy = True
        self.ctx_pool.append(c)
        return c

    def release_ctx(self, c: Ctx):
        c.busy = False

    def _ensure_scratch(self):
        """Backward scratch shared by all contexts of this plan (backward passes are serial)."""
        if self._scratch_ready:
            return
        N, sp = self.N, self.split == 3
        dev = self.device
        # gradient w.r.t. act[i] buffers (padded extents where the halo is explicit)
        self.gact: List[Optional[K.ActBuf]] = []
        for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
            hin, win, _, _ = self.geom[i]
            cp = wt.Ci_pad if i == 0 else wt.Cp_in
            fp32 = sp or i == 0
            self.gact.append(K.ActBuf(N, hin, win, cp, s.in_halo, dev, fp32=fp32))
        last = self.weights[-1]
        self.gout = K.ActBuf(N, self.Hout, self.Wout, last.Co_pitch, 0, dev, fp32=sp)
        mx = max(self.geom[i][2] * self.geom[i][3] * self.weights[i].Co_pitch for i in range(len(self.specs)))
        self.dz = torch.zeros(N * mx + K.SLACK, dtype=torch.float32 if sp else torch.bfloat16, device=dev)
        # dRaw scratch, double-buffered by stage parity: the weight-gradient GEMM of stage i (side stream)
        # reads buffer i%2 while the main stream already produces dRaw of stage i-1 in the other one
        self.draws = [torch.zeros(N * mx + K.SLACK, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        self.draws_lo = [torch.zeros_like(d) if sp else None for d in self.draws]
        self.draw, self.draw_lo = self.draws[0], self.draws_lo[0]     # (diagnostics: last written buffer)
        self.wstream = torch.cuda.Stream(device=dev) if torch.device(dev).type == "cuda" else None
        self.ev_draw = [torch.cuda.Event() for _ in range(2)] if self.wstream is not None else None
        self.ev_wg = [torch.cuda.Event() for _ in range(2)] if self.wstream is not None else None
        # dRaw of an N-expanded head stage lives in its own buffer with a zero halo of 6 (input layout of the
        # data-gradient kernel); nothing else writes it, so the halo stays zero
        self.draw_nx = {}
        for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
            if (wt.nexp and not sp and i > 0) or (wt.nexp_stem and not sp and i == 0):
                _, _, ho, wo = self.geom[i]
                self.draw_nx[i] = K.ActBuf(N, ho, wo, wt.Co_pitch, 6, dev)
        # 3x3 stride-1 stages behind an explicit halo (the residual-block convs): dRaw goes into a zero-haloed
        # (k - 1 = 2) ping-pong buffer so that the data gradient can tile the FLATTENED padded output: 36 tiles per
        # 66 x 66 sample instead of 45, i.e. 4 waves of the persistent grid instead of 5 at bs 16
        self.draw_flat = {}
        self._flat_bufs = {}
        if not sp and os.environ.get("SSCG_FLAT_DGRAD", "1") != "0":
            for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
                if (s.kind == "conv" and s.k == 3 and s.stride == 1 and s.in_halo == 1 and s.pad == 1 and s.norm and i > 0
                        and wt.need_dgrad and i not in self.draw_nx):
                    _, _, ho, wo = self.geom[i]
                    key = (ho, wo, wt.Co_pitch)
                    if key not in self._flat_bufs:
                        self._flat_bufs[key] = [K.ActBuf(N, ho, wo, wt.Co_pitch, 2, dev) for _ in range(2)]
                    self.draw_flat[i] = self._flat_bufs[key][i & 1]
        self.tbuf = [None, None]   # residual-path total gradients (ping-pong), allocated lazily
        nb = sum(N * wt.Co_pitch * 2 * L.SSCG_STAT_WORDS for wt in self.weights)
        self.bstats = torch.zeros(nb, dtype=torch.int64, device=dev)                      # binned plane sums
        self.bstat_off = []
        o = 0
        for wt in self.weights:
            self.bstat_off.append(o)
            o += N * wt.Co_pitch * 2 * L.SSCG_STAT_WORDS
        self._scratch_ready = True

    def _tbuf(self, which, like: K.ActBuf):
        if self.tbuf[which] is None:
            self.tbuf[which] = K.ActBuf(like.N, like.H, like.W, like.C, 0, self.device, fp32=self.split == 3)
        return self.tbuf[which]

    # ------------------------------------------------------------------ forward
    def _fwd_table(self, s: StageSpec):
        if s.kind == "window":
            return G.taps_conv_fwd_window(s.k, s.stride, 0)
        if s.kind == "convT":
            return G.taps_convT_fwd(s.k, s.k, s.stride, s.pad)
        return G.taps_conv_fwd(s.k, s.k, s.stride, 0 if s.in_halo else -s.pad)

    def _pixel_row_ok(self, s: StageSpec, wt: StageWeights, buf: K.ActBuf):
        return wt.pixel_row and buf.C == wt.Cp_in

    def _x_view(self, s: StageSpec, wt: StageWeights, buf: K.ActBuf):
        """(view, lo pointer) of the stage input as the forward GEMM reads it."""
        if s.kind == "window":
            return buf.window_view(wt.Kc), (buf.lo.data_ptr() if buf.lo is not None else None)
        if s.kind == "convT" or not s.in_halo:
            return buf.view(interior=True), buf.lo_ptr(interior=True)
        return buf.view(interior=False), buf.lo_ptr(interior=False)

    def forward(self, c: Ctx, x=None, labels=None, n_classes=None, training=True, drop_seed=0, parts=None):
        """x: NCHW fp32 (or labels int64 N x 1 x H x W to be one-hot encoded on the fly), or `parts`: a list of such
        tensors whose batches are packed one after the other (one batched pass instead of several, step.py).
        Leaves the fp32 NHWC result in c.out; returns c."""
        sp = self.split
        c.stats.zero_()          # plane-sum accumulators (the GEMM epilogues add into them)
        s0 = self.specs[0]
        mode0 = L.PAD_REFLECT if s0.in_reflect else L.PAD_ZERO
        if parts is None:
            parts = [labels if labels is not None else x]
        n_off = 0
        for t in parts:
            if t.dtype == torch.int64:
                K.onehot_pack(t, s0.Cin if n_classes is None else n_classes, c.act[0], mode0, n_off)
            else:
                K.pack_nchw(t, c.act[0], mode0, n_off)
            n_off += t.shape[0]
        assert n_off == self.N
        c.drop_seed = drop_seed if training else 0
        nst = len(self.specs)
        for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
            hin, win, ho, wo = self.geom[i]
            key = ("f", id(c), i)
            args = self._args_cache.get(key)
            if args is None:
                view, lo = self._x_view(s, wt, c.act[i])
                table = self._fwd_table(s)
                if not self._direct(s):
                    dst = c.raw[i]
                    kw = {}
                    if self._pixel_row_ok(s, wt, c.act[i]):
                        # stem over an 8-channel buffer: the image row itself is the K-major operand (conv_igemm.cu, RW)
                        view, lo = c.act[i].view(interior=False), None
                        kw = dict(rw_pitch=2 * wt.Cp_in, BN=64)
                    ca = K.conv_args(view, lo, table, wt.Kc, wt.w_fwd, wt.w_fwd_lo, wt.ntaps_fwd * wt.Co_pad, wt.Co_pad,
                                     dst.hi.data_ptr(), dst.fp32, (dst.sN, dst.sH, dst.sW), (0, 0), ho, wo,
                                     bias=None if s.norm else wt.bias_pad,
                                     stats=c.stats[c.stat_off[i]:] if s.norm else None, split=sp,
                                     tag=1 if s.name.startswith("res") else 4, **kw)
                    nxt = c.act[i + 1]
                    aa = L.ApplyArgs()
                    aa.raw, aa.raw_fp32 = dst.hi.data_ptr(), 1 if dst.fp32 else 0
                    aa.stats, aa.eps = (c.stats[c.stat_off[i]:].data_ptr() if s.norm else None), 1e-5
                    aa.N, aa.H, aa.W, aa.C = self.N, ho, wo, wt.Co_pitch
                    aa.act, aa.slope = s.act, 0.2
                    if s.residual_from is not None:
                        rb = c.act[s.residual_from]
                        aa.res = rb.view(interior=True)
                        aa.res_lo = rb.lo_ptr(interior=True)
                    aa.dst, aa.dst_lo = nxt.hi.data_ptr(), (nxt.lo.data_ptr() if nxt.lo is not None else None)
                    aa.pad = nxt.pad
                    aa.pad_mode = L.PAD_REFLECT if self.specs[i + 1].in_reflect else L.PAD_ZERO
                    args = (ca, aa)
                else:
                    dst = c.out if s.final else c.act[i + 1]
                    assert dst.pad == 0
                    kw = {}
                    if wt.nexp and self.split == 1 and dst.fp32:
                        src = c.act[i]
                        ca = K.conv7_args(src.hi.data_ptr(), src.C, self.N, src.Hp, src.Wp, wt.w_nx, wt.nx_CoW, 1, 4,
                                          min(wt.nx_CoW, dst.C), dst.hi.data_ptr(), True, (dst.sN, dst.sH, dst.sW),
                                          bias=wt.bias_pad, act=s.act, tag=4)
                        self._args_cache[key] = (ca, None)
                        K.run_conv7(ca)
                        continue
                    if self._rowshift_ok(s, wt.Co_pad):      # 7x7 head: one row box feeds the 7 horizontal taps
                        table = G.taps_rowshift_fwd(s.k, s.k, 0 if s.in_halo else -s.pad)
                        kw = dict(shift_kw=s.k, shift_brow_step=1, BN=min(wt.Co_pad, 32))
                    ca = K.conv_args(view, lo, table, wt.Kc, wt.w_fwd, wt.w_fwd_lo, wt.ntaps_fwd * wt.Co_pad, wt.Co_pad,
                                     dst.hi.data_ptr(), dst.fp32, (dst.sN, dst.sH, dst.sW), (0, 0), ho, wo,
                                     bias=wt.bias_pad, act=s.act, split=sp, tag=4, **kw)
                    args = (ca, None)
                self._args_cache[key] = args
            ca, aa = args
            if isinstance(ca, L.Conv7Args):
                K.run_conv7(ca)
                continue
            K.run_conv(ca)
            if aa is not None:
                aa.drop_seed = (c.drop_seed * 1000003 + i + 1) if (s.dropout and c.drop_seed) else 0
                aa.drop_ctr = self.drop_ctr.data_ptr() if self.drop_ctr is not None else None
                K.run_apply(aa)
            assert i < nst
        return c

    def output_nchw(self, c: Ctx):
        y = torch.empty(self.N, self.Cout, self.Hout, self.Wout, dtype=torch.float32, device=self.device)
        K.unpack_nhwc(c.out.hi, self.N, self.Cout, self.Hout, self.Wout, c.out.C, y)
        return y

    # ------------------------------------------------------------------ backward
    def backward(self, c: Ctx, grad_out=None, need_dx=True, need_dw=True, accumulate_dw=False, gout_ready=False,
                 debug_hook=None, dx_range=None):
        """grad_out: NCHW fp32 gradient of the network output (or gout_ready=True when self.gout was
        filled by a fused loss kernel).  Weight-gradient slabs are accumulated in self.weights[i].dw
        (zeroed first unless accumulate_dw).  Returns grad_in NCHW fp32 (or None); with dx_range = (n0, n1) the input
        gradient is computed and returned for those samples only (batched passes whose other parts are data)."""
        if dx_range is not None and tuple(dx_range) == (0, self.N):
            dx_range = None
        self._ensure_scratch()
        sp = self.split
        N = self.N
        if not gout_ready:
            K.pack_nchw(grad_out.contiguous(), self.gout, L.PAD_ZERO)
        self.bstats.zero_()
        if need_dw and not accumulate_dw:
            if getattr(self, "dw_flat", None) is not None:
                self.dw_flat.zero_()
            else:
                for wt in self.weights:
                    wt.dw.zero_()
        nst = len(self.specs)
        wg_pending = [False, False]
        for i in range(nst - 1, -1, -1):
            s, wt = self.specs[i], self.weights[i]
            hin, win, ho, wo = self.geom[i]
            key = ("b", id(c), i, need_dx, need_dw, dx_range if i == 0 else None)
            args = self._args_cache.get(key)
            if args is None:
                args = self._build_bwd_args(c, i, need_dx, need_dw, dx_range if i == 0 else None)
                self._args_cache[key] = args
            ba, use_apply, wa, da = args
            ba.drop_seed = (c.drop_seed * 1000003 + i + 1) if (s.dropout and c.drop_seed) else 0
            ba.drop_ctr = self.drop_ctr.data_ptr() if self.drop_ctr is not None else None
            par = i & 1
            overlap = self.overlap_wgrad and self.wstream is not None and debug_hook is None
            main = torch.cuda.current_stream()
            if overlap and wg_pending[par]:
                main.wait_event(self.ev_wg[par])          # the wgrad that last read this dRaw buffer is done
                wg_pending[par] = False
            K.run_bwd_prep(ba)
            if use_apply:
                if i in self.draw_nx:
                    K.run_bwd_apply(ba, self.draw_nx[i].hi, None)
                elif i in self.draw_flat:
                    K.run_bwd_apply(ba, self.draw_flat[i].hi, None)
                else:
                    K.run_bwd_apply(ba, self.draws[par], self.draws_lo[par])
            self.draw, self.draw_lo = self.draws[par], self.draws_lo[par]
            if wa is not None and not overlap:
                K.run_wgrad(wa)
            if da is not None:
                (K.run_conv7 if isinstance(da, L.Conv7Args) else K.run_conv)(da)
            if wa is not None and overlap:
                # fork AFTER the dgrad: the side-stream wgrad then runs next to the memory-bound
                # prep/apply kernels of stage i-1 (they co-reside with a GEMM CTA on an SM) instead of
                # fighting the dgrad GEMM on the critical path for whole SMs
                self.ev_draw[par].record(main)
                with torch.cuda.stream(self.wstream):
                    self.wstream.wait_event(self.ev_draw[par])
                    K.run_wgrad(wa)
                    self.ev_wg[par].record(self.wstream)
                wg_pending[par] = True
            if debug_hook is not None:
                debug_hook(i, self)
        for par in (0, 1):                                  # join the side stream before anyone reads dw
            if wg_pending[par]:
                torch.cuda.current_stream().wait_event(self.ev_wg[par])
        gx = None
        if need_dx:
            s0 = self.specs[0]
            n0, n1 = dx_range if dx_range is not None else (0, N)
            gx = torch.empty(n1 - n0, s0.Cin, self.H, self.W, dtype=torch.float32, device=self.device)
            K.unpack_fold(self.gact[0], s0.Cin, gx, L.PAD_REFLECT if s0.in_reflect else L.PAD_ZERO, n0, n1)
        return gx

    def _draw_view(self, i, lo=False):
        wt = self.weights[i]
        _, _, ho, wo = self.geom[i]
        if i in self.draw_nx and not lo:
            return self.draw_nx[i].view(interior=True)
        if i in self.draw_flat and not lo:
            return self.draw_flat[i].view(interior=True)
        t = self.draws_lo[i & 1] if lo else self.draws[i & 1]
        cp = wt.Co_pitch
        return L.make_view(t.data_ptr(), self.N, ho, wo, cp, ho * wo * cp, wo * cp, cp)

    def _build_bwd_args(self, c: Ctx, i, need_dx, need_dw, dx_range=None):
        s, wt = self.specs[i], self.weights[i]
        sp = self.split
        N = self.N
        hin, win, ho, wo = self.geom[i]
        nst = len(self.specs)
        cp = wt.Co_pitch
        # ---- 1. elementwise backward -> dRaw in self.draw ------------------------------------
        ba = L.BwdArgs()
        ba.N, ba.H, ba.W, ba.C = N, ho, wo, cp
        ba.eps, ba.slope = 1e-5, 0.2
        ba.act = s.act
        if i == nst - 1:
            g = self.gout
            ba.dyp, ba.dyp_fp32 = g.view(interior=False), 1 if g.fp32 else 0
            ba.pad, ba.pad_mode = 0, L.PAD_NONE
        else:
            g = self.gact[i + 1]
            nxt = self.specs[i + 1]
            ba.dyp, ba.dyp_fp32 = g.view(interior=False), 1 if g.fp32 else 0
            ba.pad = nxt.in_halo
            ba.pad_mode = (L.PAD_REFLECT if nxt.in_reflect else L.PAD_ZERO) if nxt.in_halo else L.PAD_NONE
        # residual-path gradients (generator): see _residual_plan
        skip, gout_t = self.res_bwd.get(i, (None, None)) if hasattr(self, "res_bwd") else (None, None)
        if skip is not None:
            sb = self._resolve_t(skip, c.act[i + 1])
            ba.skip, ba.skip_fp32 = sb.view(interior=False), 1 if sb.fp32 else 0
        if gout_t is not None:
            tb = self._resolve_t(gout_t, c.act[i + 1])
            ba.g_out, ba.g_fp32 = tb.hi.data_ptr(), 1 if tb.fp32 else 0
        ba.bstats = self.bstats[self.bstat_off[i]:].data_ptr()
        if s.norm:
            raw = c.raw[i]
            ba.raw, ba.raw_fp32 = raw.hi.data_ptr(), 1 if raw.fp32 else 0
            ba.stats = c.stats[c.stat_off[i]:].data_ptr()
            ba.dz, ba.dz_fp32, ba.dz_lo = self.dz.data_ptr(), 1 if sp == 3 else 0, None
            if (gout_t is not None and sp == 1 and s.act == L.ACT_NONE and not s.dropout and not ba.g_fp32
                    and os.environ.get("SSCG_DZ_ALIAS", "1") != "0"):
                # no activation, no dropout: dZ equals the folded total gradient — one buffer, one store
                ba.dz = ba.g_out
            use_apply = True
        else:
            # activation applied in the GEMM epilogue: its output carries the sign (LeakyReLU) / value (tanh);
            # in bf16x3 mode the pre-activation (bias included) is kept in raw[i] and has the same sign
            src = c.out if s.final else (c.act[i + 1] if self._direct(s) else c.raw[i])
            ba.raw, ba.raw_fp32 = src.hi.data_ptr(), 1 if src.fp32 else 0
            ba.stats = None
            ba.dz, ba.dz_fp32 = self.draws[i & 1].data_ptr(), 0
            ba.dz_lo = self.draws_lo[i & 1].data_ptr() if self.draws_lo[i & 1] is not None else None
            if i in self.draw_nx:
                ba.dz, ba.dz_pad = self.draw_nx[i].hi.data_ptr(), 6
            use_apply = False
        if use_apply and i in self.draw_nx:
            ba.draw_pad = 6
        if use_apply and i in self.draw_flat:
            ba.draw_pad = 2
        # ---- 2. wgrad ------------------------------------------------------------------------
        wa = None
        if need_dw:
            xview, xlo = self._x_view(s, wt, c.act[i])
            dview = self._draw_view(i)
            dlo = self.draws_lo[i & 1].data_ptr() if self.draws_lo[i & 1] is not None else None
            if s.kind == "convT":
                table = G.taps_convT_dgrad(s.k, s.k, s.stride, s.pad)
                wa = K.wgrad_args(xview, xlo, dview, dlo, table, wt.wg_Kc, wt.wg_rows, wt.dw, wt.wg_taps * wt.wg_rows,
                                  split=sp, tag=6, ws_pool=self.ws_wg)
            else:
                table = self._fwd_table(s)
                wkw = {}
                if wt.wg_pixel_row and c.act[i].C == wt.Cp_in:
                    xview, xlo = c.act[i].view(interior=False), None
                    wkw = dict(rw_pitch=2 * wt.Cp_in)
                if wt.wg_window:
                    table = G.taps_conv_fwd_window(s.k, 1, 0)
                    xview, xlo = c.act[i].window_view(wt.wg_Kc), None
                if (wt.wg_window and wt.nexp and i in self.draw_nx and sp == 1 and wt.wg_rows == 64 and wt.wg_Kc == 448
                        and self.draw_nx[i].C in (16, 32) and min(ho, wo) >= 8 and os.environ.get("SSCG_WGRAD7", "1") != "0"):
                    # 7x7 head: horizontal taps as GEMM columns (csrc/conv_wgrad7.cu), same slab layout
                    wa = K.wgrad7_args(c.act[i], self.draw_nx[i], wt.dw, tag=6, ws_pool=self.ws_wg7)
                else:
                    wa = K.wgrad_args(dview, dlo, xview, xlo, table, wt.wg_Kc, wt.wg_rows, wt.dw, wt.wg_taps * wt.wg_rows,
                                      split=sp, tag=3 if s.name.startswith("res") else 6, ws_pool=self.ws_wg, **wkw)
        # ---- 3. dgrad ------------------------------------------------------------------------
        da = None
        dkw = {}
        n0, n1 = dx_range if dx_range is not None else (0, N)      # stage 0 only: samples whose input gradient is wanted
        if i == 0 and i in self.draw_nx and need_dx and wt.need_dgrad:
            gin, src = self.gact[0], self.draw_nx[0]          # fp32 gradient w.r.t. the halo-padded network input
            assert gin.pad == 3 and gin.fp32 and gin.C >= wt.nx_CoW
            da = K.conv7_args(src.hi.data_ptr() + n0 * src.sN * src.esize, src.C, n1 - n0, src.Hp, src.Wp, wt.w_nx_dg,
                              wt.nx_CoW, 1, 4, wt.nx_CoW, gin.hi.data_ptr() + n0 * gin.sN * gin.esize, True,
                              (gin.sN, gin.sH, gin.sW), tag=5)
        elif i in self.draw_nx and i > 0 and wt.need_dgrad and not self.gact[i].fp32:
            gin, src = self.gact[i], self.draw_nx[i]
            assert gin.pad == 3 and gin.C == 32 * wt.nx_dg_tiles
            da = K.conv7_args(src.hi.data_ptr(), src.C, N, src.Hp, src.Wp, wt.w_nx_dg, 32, wt.nx_dg_tiles, src.C // 16, 32,
                              gin.hi.data_ptr(), False, (gin.sN, gin.sH, gin.sW), tag=5)
        elif i in self.draw_flat:
            gin, src = self.gact[i], self.draw_flat[i]
            assert gin.pad == 1 and gin.Hp == src.Hp - 2 and gin.Wp == src.Wp - 2
            npx = N * src.Hp * src.Wp
            fview = L.make_view(src.hi.data_ptr(), 1, 1, npx, src.C, npx * src.C, npx * src.C, src.C)
            table = G.taps_conv_dgrad(s.k, s.k, 1, -(s.k - 1))        # taps (k-1-a, k-1-b) >= 0 into the zero-haloed dRaw
            da = K.conv_args(fview, None, table, wt.Kc_d, wt.w_dg, None, s.k * s.k * wt.Ci_pad, wt.Ci_pad,
                             gin.hi.data_ptr(), gin.fp32, (gin.sN, gin.sH, gin.sW), (0, 0), gin.Hp, gin.Wp, split=sp,
                             tag=2 if s.name.startswith("res") else 5, flat=(src.Wp, src.Hp * src.Wp, N))
        elif (i > 0 or need_dx) and wt.need_dgrad:
            gin = self.gact[i]
            dview = self._draw_view(i)
            dlo = self.draws_lo[i & 1].data_ptr() if self.draws_lo[i & 1] is not None else None
            if s.kind == "convT":
                table = G.taps_convT_dgrad(s.k, s.k, s.stride, s.pad)
                Ho_d, Wo_d = hin, win
                yoff = (gin.pad, gin.pad)
            else:
                org = 0 if s.in_halo else -s.pad
                table = G.taps_conv_dgrad(s.k, s.k, s.stride, org)
                if self._rowshift_ok(s, min(wt.Ci_pad, 32)) and wt.Ci_pad <= 64:
                    table = G.taps_rowshift_dgrad(s.k, s.k, org)
                    dkw = dict(shift_kw=s.k, shift_brow_step=-1, BN=wt.Ci_pad)   # 16 / 32 (stems), 64 (head)
                if s.in_halo:
                    Ho_d, Wo_d, yoff = gin.Hp, gin.Wp, (0, 0)
                else:
                    Ho_d, Wo_d, yoff = hin, win, (0, 0)
            gptr = gin.hi.data_ptr()
            if i == 0 and dx_range is not None:
                dview = K.sub_view(dview, n0, n1)
                dlo = (dlo + n0 * dview.sN * 2) if dlo is not None else None
                gptr += n0 * gin.sN * gin.esize
            da = K.conv_args(dview, dlo, table, wt.Kc_d, wt.w_dg, wt.w_dg_lo, s.k * s.k * wt.Ci_pad, wt.Ci_pad,
                             gptr, gin.fp32, (gin.sN, gin.sH, gin.sW), yoff, Ho_d, Wo_d, split=sp,
                             tag=2 if s.name.startswith("res") else 5, **dkw)
        return ba, use_apply, wa, da

    def _resolve_t(self, tag, like):
        kind, idx = tag
        if kind == "t":
            return self._tbuf(idx, like)
        return self.gact[idx]   # ("g", stage index): a dgrad output used directly as a total gradient

    # ------------------------------------------------------------------ gradients -> parameters
    def param_grads(self, scale=1.0, into=None, weights=True):
        """Unpack the wgrad slabs into parameter-shaped fp32 gradients.  Returns a list aligned with
        [(weight, bias) for every stage]; biases cancelled by InstanceNorm get exact zeros."""
        out = []
        for i, (s, wt) in enumerate(zip(self.specs, self.weights)):
            # with `into`, a None entry is a parameter without a gradient buffer (frozen): nothing to add there
            gw = torch.zeros_like(s.weight, dtype=torch.float32) if into is None else into[i][0]
            if weights and gw is not None:
                K.run_wgrad_unpack(wt.unpack_wg, wt.dw, gw, scale)
            gb = None
            if s.bias is not None:
                gb = torch.zeros_like(s.bias, dtype=torch.float32) if into is None else into[i][1]
                if not s.norm and gb is not None:
                    K.bias_grad(self.bstats[self.bstat_off[i]:], self.N, s.Cout, wt.Co_pitch, gb, scale)
            out.append((gw, gb))
        return out

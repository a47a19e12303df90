"""Timing of the 7x7 head weight gradient at production size: conv_wgrad7 vs the window-mode conv_wgrad."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402,F401
from sscg_b200 import geometry as G  # noqa: E402
from sscg_b200 import kernels as K  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    cfgs = [(16, 21), (16, 3), (32, 21)] if len(sys.argv) < 2 else [(16, 21)]
    for N, Cout in cfgs:
        H = W = 256
        xb = K.ActBuf(N, H, W, 64, 3, DEV)
        xb.hi.normal_()
        dyb = K.ActBuf(N, H, W, G.pad_out_channels(Cout), 6, DEV)
        t = torch.zeros(N, dyb.Hp, dyb.Wp, dyb.C, device=DEV)
        t[:, 6:-6, 6:-6, :Cout].normal_()
        dyb.hi[:t.numel()].copy_(t.reshape(-1).to(torch.bfloat16))
        dw = torch.zeros(7 * 64 * 448, device=DEV)
        a7 = K.wgrad7_args(xb, dyb, dw)
        t7 = timeit(lambda: K.run_wgrad(a7), reps=20 if len(sys.argv) < 2 else 2)
        line = "head wgrad N=%d Cout=%d: wgrad7 %.1f us" % (N, Cout, t7)
        if len(sys.argv) < 2:
            table = G.taps_conv_fwd_window(7, 1, 0)
            aw = K.wgrad_args(dyb.view(interior=True), None, xb.window_view(448), None, table, 448, 64, dw, 7 * 64)
            line += ", window %.1f us" % timeit(lambda: K.run_wgrad(aw))
        print(line, "dev_err", K.device_error(), flush=True)


if __name__ == "__main__":
    main()

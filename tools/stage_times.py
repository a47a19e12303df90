"""GPU diagnostic: per-launch CUDA-event timing of one forward+backward of each of the four networks at
the bench workload (bs 16, 256x256), aggregated by (kernel kind, geometry).  Shows which layers sit
far from the tensor roofline.  Writes gpurun_out/stage_times.{json,md}."""
import argparse
import contextlib
import io
import json
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402,F401
from sscg_b200 import kernels as K  # noqa: E402
from sscg_b200.arch import define_Dis, define_Gen  # noqa: E402

REC = []          # (net, key, flops, bytes, ev0, ev1)
NET = ["?"]


def _ntaps_conv(a):
    return a.phase_start[a.n_phases]


def _timed(fn, key_fn):
    def w(a, *rest):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(a, *rest)
        e1.record()
        key, flops, byts = key_fn(a)
        REC.append((NET[0], key, flops, byts, e0, e1))
        return r
    return w


def key_conv(a):
    nt = _ntaps_conv(a)
    px = a.x.N * a.Ho * a.Wo
    if a.n_phases == 4:
        flops = 2.0 * a.x.N * (a.Ho // 2) * (a.Wo // 2) * nt * a.Kc * a.Co_pad     # all four phases, nt taps in total
    else:
        flops = 2.0 * px * nt * a.Kc * a.Co_pad * (a.shift_kw if a.shift_kw else 1)
    kind = {1: "fwd", 2: "dgrad", 4: "fwd", 5: "dgrad"}.get(a.tag, "conv")
    key = "%s %dx%d K=%d taps=%d N=%d BN=%d s=%d ph=%d%s%s" % (kind, a.Ho, a.Wo, a.Kc, nt, a.Co_pad, a.BN, a.stride,
                                                             a.n_phases, " shift" if a.shift_kw else "",
                                                             " stats" if a.stats else "")
    byts = px * a.Co_pad * (4 if a.y_fp32 else 2) + a.x.N * a.x.H * a.x.W * min(a.x.C, a.Kc) * 2
    return key, flops, byts


def key_conv7(a):
    nt = (7 * a.CoW + 15) // 16 * 16
    px = a.N * (a.Hp - 6) * (a.Wp - 6)
    flops = 2.0 * a.N * (a.Hp - 6) * a.Wp * 7 * 16 * a.ksteps * nt * a.n_ntiles
    key = "%s nexp %dx%d K=%dx7 N=%dx%d" % ({4: "fwd", 5: "dgrad"}.get(a.tag, "conv7"), a.Hp - 6, a.Wp - 6, 16 * a.ksteps, nt,
                                            a.n_ntiles)
    return key, flops, px * a.CoW * a.n_ntiles * (4 if a.y_fp32 else 2) + a.N * a.Hp * a.Wp * a.x_pitch * 2


def key_wgrad(a):
    from sscg_b200 import _lib as L
    if isinstance(a, L.Wgrad7Args):
        flops = 2.0 * a.N * a.H * a.W * 49 * 64 * a.Cy
        return "wgrad7 %dx%d Cy=%d (7x7 head, taps as columns)" % (a.H, a.W, a.Cy), flops, a.N * a.H * a.W * (64 + a.Cy) * 2
    px = a.dy.N * a.dy.H * a.dy.W
    flops = 2.0 * px * a.n_taps * a.Kc * a.Co_pad
    key = "wgrad %dx%d K=%d taps=%d M=%d BN=%d ks=%d" % (a.dy.H, a.dy.W, a.Kc, a.n_taps, a.Co_pad, a.BN, a.ksplit)
    return key, flops, px * a.dy.C * 2 + a.x.N * a.x.H * a.x.W * min(a.x.C, a.Kc) * 2


def key_apply(a):
    el = a.N * a.H * a.W * a.C
    return "in_apply %dx%dx%d res=%d pad=%d" % (a.H, a.W, a.C, 1 if a.res.ptr else 0, a.pad), 0.0, el * (4 + (2 if a.res.ptr else 0))


def key_prep(a):
    el = a.N * a.H * a.W * a.C
    return "bwd_prep %dx%dx%d norm=%d skip=%d pad=%d" % (a.H, a.W, a.C, 1 if a.stats else 0, 1 if a.skip.ptr else 0,
                                                         a.pad), 0.0, el * 6


def key_bapply(a):
    el = a.N * a.H * a.W * a.C
    return "bwd_apply %dx%dx%d" % (a.H, a.W, a.C), 0.0, el * 6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--classes", type=int, default=21)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    torch.manual_seed(0)
    K.run_conv = _timed(K.run_conv, key_conv)
    K.run_wgrad = _timed(K.run_wgrad, key_wgrad)
    K.run_conv7 = _timed(K.run_conv7, key_conv7)
    K.run_apply = _timed(K.run_apply, key_apply)
    K.run_bwd_prep = _timed(K.run_bwd_prep, key_prep)
    K.run_bwd_apply = _timed(K.run_bwd_apply, key_bapply)
    C = a.classes
    with contextlib.redirect_stdout(io.StringIO()):
        nets = OrderedDict(
            Gsi=define_Gen(3, C, 64, "resnet_9blocks_softmax", norm="instance", use_dropout=True, gpu_ids=[0]),
            Gis=define_Gen(C, 3, 64, "resnet_9blocks", norm="instance", use_dropout=True, gpu_ids=[0]),
            Di=define_Dis(3, 64, "n_layers", norm="instance", gpu_ids=[0]),
            Ds=define_Dis(C, 64, "n_layers", norm="instance", gpu_ids=[0]))
    cin = dict(Gsi=3, Gis=C, Di=3, Ds=C)
    for rep in range(a.reps + 1):
        if rep == 1:
            torch.cuda.synchronize()
            REC.clear()
        for name, net in nets.items():
            NET[0] = name
            x = (torch.rand(a.batch, cin[name], a.size, a.size, device="cuda") * 2 - 1).requires_grad_(True)
            y = net(x)
            y.square().mean().backward()
    torch.cuda.synchronize()
    agg = OrderedDict()
    for net, key, flops, byts, e0, e1 in REC:
        k = (net, key)
        d = agg.setdefault(k, dict(ms=0.0, n=0, flops=flops, bytes=byts))
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
    rows = []
    for (net, key), d in agg.items():
        us = 1e3 * d["ms"] / d["n"]
        rows.append(dict(net=net, key=key, launches_per_pass=d["n"] // a.reps, avg_us=us,
                         tflops=d["flops"] / (us * 1e-6) / 1e12 if d["flops"] else None,
                         gbs=d["bytes"] / (us * 1e-6) / 1e9, ms_per_pass=d["ms"] / a.reps))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stage_times.json"), "w") as f:
        json.dump(rows, f, indent=1)
    lines = ["| net | launch | n/pass | avg us | ms/pass | TFLOP/s (padded) | GB/s (min traffic) |", "|---|---|---:|---:|---:|---:|---:|"]
    for net in nets:
        tot = 0.0
        for r in sorted([r for r in rows if r["net"] == net], key=lambda r: -r["ms_per_pass"]):
            tot += r["ms_per_pass"]
            lines.append("| %s | %s | %d | %.1f | %.3f | %s | %.0f |" % (
                r["net"], r["key"], r["launches_per_pass"], r["avg_us"], r["ms_per_pass"],
                ("%.0f" % r["tflops"]) if r["tflops"] else "-", r["gbs"]))
        lines.append("| %s | **total fwd+bwd** | | | %.3f | | |" % (net, tot))
    with open(os.path.join(ROOT, "gpurun_out", "stage_times.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    print("device_error", K.device_error())


if __name__ == "__main__":
    main()

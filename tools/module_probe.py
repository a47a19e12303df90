"""GPU diagnostic: fused modules (arch.define_Gen / define_Dis on cuda:0) against the CPU oracle.
Prints error metrics for forward outputs and gradients in both precision modes; writes
gpurun_out/module_probe.json.  Not a test — used to calibrate the tolerances in tests/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402
from oracle import ref_arch as RA  # noqa: E402
from sscg_b200.arch import define_Dis, define_Gen  # noqa: E402

OUT = []


def metrics(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b)
    return {"rel_l2": float(d.norm() / max(b.norm().item(), 1e-30)),
            "max_rel": float(d.abs().max() / max(b.abs().max().item(), 1e-30))}


def run_net(kind, cfg, N, H, W, precision, seed=0):
    torch.manual_seed(seed)
    if kind == "gen":
        cin, cout, ngf, name = cfg
        net = define_Gen(cin, cout, ngf, name, norm="instance", use_dropout=False, gpu_ids=[0])
    else:
        cin, ndf = cfg
        net = define_Dis(cin, ndf, "n_layers", norm="instance", gpu_ids=[0])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.05)
    net.precision = precision
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x = (torch.rand(N, cin, H, W) * 2 - 1)
    xg = x.cuda().requires_grad_(True)
    t0 = time.time()
    y = net(xg)
    torch.cuda.synchronize()
    probe = torch.randn(y.shape)
    (y * probe.cuda()).sum().backward()
    torch.cuda.synchronize()
    t1 = time.time()
    # oracle fp32
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    if kind == "gen":
        tanh = not name.endswith("softmax")
        yr = RA.resnet_generator(sdr, xr, 9 if "9" in name else 6, tanh=tanh)
        with torch.no_grad():
            ye = RA.resnet_generator(sd, x, 9 if "9" in name else 6, tanh=tanh, emulate_bf16=True, live_norm_bias=False)
    else:
        yr = RA.nlayer_discriminator(sdr, xr, 3)
        with torch.no_grad():
            ye = RA.nlayer_discriminator(sd, x, 3, emulate_bf16=True, live_norm_bias=False)
    (yr * probe).sum().backward()
    rec = {"kind": kind, "cfg": list(cfg), "shape": [N, H, W], "precision": precision, "secs": t1 - t0}
    rec["y_vs_fp32"] = metrics(y, yr)
    rec["y_vs_emul"] = metrics(y, ye)
    rec["emul_vs_fp32"] = metrics(ye, yr)
    rec["gx"] = metrics(xg.grad, xr.grad)
    worst = {"rel_l2": 0.0}
    allg = {}
    for (k, p) in net.named_parameters():
        g = sdr[k].grad
        if k.endswith("bias") and g.abs().max() < 1e-4:
            continue
        m = metrics(p.grad, g)
        allg[k] = m["rel_l2"]
        if m["rel_l2"] > worst["rel_l2"]:
            worst = dict(m, name=k)
    rec["worst_param_grad"] = worst
    rec["param_grad_rel_l2"] = allg
    if kind == "gen" and not tanh:
        am, ar = y.argmax(1).cpu(), yr.argmax(1)
        top2 = yr.detach().topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])
        safe = margin > 1e-3 * yr.detach().abs().max()
        rec["argmax_mismatch_all"] = float((am != ar).float().mean())
        rec["argmax_mismatch_safe"] = float(((am != ar) & safe).float().sum())
    from sscg_b200 import kernels as K
    rec["dev_error"] = K.device_error()
    print(json.dumps({k: v for k, v in rec.items() if k != "param_grad_rel_l2"}), flush=True)
    OUT.append(rec)


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    jobs = [
        ("gen", (3, 5, 4, "resnet_9blocks_softmax"), 2, 16, 16),
        ("dis", (3, 4), 2, 32, 32),
        ("gen", (3, 21, 64, "resnet_9blocks_softmax"), 2, 64, 64),
        ("gen", (21, 3, 64, "resnet_9blocks"), 2, 64, 64),
        ("dis", (3, 64), 2, 64, 64),
        ("dis", (21, 64), 2, 128, 128),
        ("gen", (1, 4, 64, "resnet_9blocks_softmax"), 1, 32, 48),
    ]
    for prec in ("bf16x3", "bf16"):
        for j in jobs:
            if which != "all" and which != prec:
                continue
            try:
                run_net(j[0], j[1], j[2], j[3], j[4], prec)
            except Exception as e:  # keep going: one failure must not hide the rest
                import traceback
                traceback.print_exc()
                OUT.append({"kind": j[0], "cfg": list(j[1]), "precision": prec, "error": repr(e)})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "module_probe.json"), "w") as f:
        json.dump(OUT, f, indent=1)

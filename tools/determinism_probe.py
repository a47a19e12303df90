"""Run-to-run reproducibility of the fused path on cuda:0 (prints one JSON line per check):
forward with dropout under a fixed seed, one full bf16 training step from identical weights (losses, flat gradient
buckets, updated weights) and CUDA-graph replays against eager launches.  Every difference should be exactly 0:
the kernels contain no floating-point atomics (fixed-order reductions, csrc/sscg_ptx.cuh)."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402,F401
from sscg_b200.arch import define_Gen  # noqa: E402
from sscg_b200.step import GraphedStep, SemiSupCycleGAN  # noqa: E402


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def main():
    N, H = int(os.environ.get("N", "4")), int(os.environ.get("HW", "128"))
    torch.manual_seed(1)
    g = quiet(define_Gen, 3, 21, 64, "resnet_9blocks_softmax", norm="instance", use_dropout=True, gpu_ids=[0])
    x = torch.rand(N, 3, H, H, device="cuda") * 2 - 1
    for prec in ("bf16", "bf16x3"):
        g.precision = prec
        g.train()
        outs = []
        for _ in range(3):
            torch.manual_seed(5)
            xg = x.clone().requires_grad_(True)
            y = g(xg)
            g.zero_grad(set_to_none=True)
            y.square().mean().backward()
            outs.append((y.detach().clone(), xg.grad.clone(), [p.grad.clone() for p in g.parameters()]))
        d_y = max(float((o[0] - outs[0][0]).abs().max()) for o in outs[1:])
        d_gx = max(float((o[1] - outs[0][1]).abs().max()) for o in outs[1:])
        d_gw = max(float((a - b).abs().max()) for o in outs[1:] for a, b in zip(o[2], outs[0][2]))
        print(json.dumps({"check": "generator fwd+bwd x3, dropout, fixed seed", "precision": prec, "max_abs_diff_y": d_y,
                          "max_abs_diff_gx": d_gx, "max_abs_diff_gw": d_gw}), flush=True)
    # ---- full step, twice from identical weights ---------------------------------------------------
    l_img = torch.rand(N, 3, H, H, device="cuda") * 2 - 1
    unl = torch.rand(N, 3, H, H, device="cuda") * 2 - 1
    l_gt = torch.randint(0, 21, (N, 1, H, H), device="cuda")
    for prec in ("bf16", "bf16x3"):
        res = []
        for rep in range(2):
            torch.manual_seed(0)
            np.random.seed(0)
            m = quiet(SemiSupCycleGAN, n_classes=21, variant="classic", use_dropout=True, device="cuda:0", precision=prec)
            for _ in range(3):
                out = m.train_step(l_img, l_gt, unl)
            torch.cuda.synchronize()
            res.append(({k: float(v) for k, v in out.items()}, m.g_grads.flat.clone(), m.d_grads.flat.clone(),
                        torch.cat([p.detach().reshape(-1) for p in m.Gsi.parameters()])))
        print(json.dumps({"check": "3 training steps x2 from identical weights", "precision": prec,
                          "loss_max_abs_diff": max(abs(res[0][0][k] - res[1][0][k]) for k in res[0][0]),
                          "g_grad_max_abs_diff": float((res[0][1] - res[1][1]).abs().max()),
                          "d_grad_max_abs_diff": float((res[0][2] - res[1][2]).abs().max()),
                          "weights_max_abs_diff": float((res[0][3] - res[1][3]).abs().max())}), flush=True)
    # ---- graph replay vs eager -------------------------------------------------------------------
    for prec in ("bf16", "bf16x3"):
        outs = []
        for graph in (False, True):
            torch.manual_seed(0)
            np.random.seed(0)
            m = quiet(SemiSupCycleGAN, n_classes=21, variant="classic", use_dropout=False, device="cuda:0", precision=prec,
                      graph_safe=graph)
            if graph:
                gs = GraphedStep(m, l_img, l_gt, unl, warmup=3)
                for _ in range(2):
                    o = gs(l_img, l_gt, unl).clone()
                torch.cuda.synchronize()
                host = {k: float(v) for k, v in zip(gs.KEYS, o)}
            else:
                for _ in range(5):
                    o = m.train_step(l_img, l_gt, unl)
                host = {k: float(v) for k, v in o.items()}
            outs.append((host, torch.cat([p.detach().reshape(-1) for p in m.Gsi.parameters()])))
        print(json.dumps({"check": "5 steps: CUDA-graph replays vs eager launches (no dropout)", "precision": prec,
                          "loss_max_abs_diff": max(abs(outs[0][0][k] - outs[1][0][k]) for k in outs[0][0]),
                          "weights_max_abs_diff": float((outs[0][1] - outs[1][1]).abs().max())}), flush=True)


if __name__ == "__main__":
    main()

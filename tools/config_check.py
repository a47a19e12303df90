"""BASELINE.json configs #3 / #4 / #5 at FULL size on one GPU (per-GPU batch): parity-mode forward of Gsi and Ds against
the fp32 oracle on two samples of the full-size input, one bf16 training step at the full per-GPU batch (finite losses,
no device error), and the step repeated from the same seeds (bit-identical losses).  Prints one JSON line per config."""
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
from oracle import ref_arch as RA  # noqa: E402
from sscg_b200 import kernels as K  # noqa: E402
from sscg_b200.step import SemiSupCycleGAN  # noqa: E402

CONFIGS = {3: dict(cimg=3, ncls=19, h=256, w=512, batch=8), "3b": dict(cimg=3, ncls=20, h=256, w=512, batch=8),
           4: dict(cimg=1, ncls=4, h=256, w=256, batch=32), 5: dict(cimg=3, ncls=21, h=512, w=512, batch=4)}


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def main():
    import gc
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    want = sys.argv[1:]
    for idx, c in CONFIGS.items():
        if want and str(idx) not in want:
            continue
        torch.manual_seed(0)
        np.random.seed(0)
        m = quiet(SemiSupCycleGAN, n_classes=c["ncls"], img_channels=c["cimg"], variant="classic", use_dropout=True,
                  device="cuda:0", precision="bf16x3")
        g = torch.Generator().manual_seed(7)
        x = torch.rand(2, c["cimg"], c["h"], c["w"], generator=g) * 2 - 1
        lab = torch.randint(0, c["ncls"], (2, 1, c["h"], c["w"]), generator=g)
        m.Gsi.eval()
        with torch.no_grad():
            y = m.Gsi(x.cuda()).cpu()
            d = m.Ds.forward_onehot(lab.cuda()).cpu()
        yr = RA.resnet_generator({k: v.detach().cpu() for k, v in m.Gsi.state_dict().items()}, x, 9, tanh=False, use_dropout=True)
        dr = RA.nlayer_discriminator({k: v.detach().cpu() for k, v in m.Ds.state_dict().items()}, RA.make_one_hot(lab, c["ncls"]), 3)
        top2 = yr.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-3 * yr.abs().max()
        flips = int(((y.argmax(1) != yr.argmax(1)) & safe).sum())
        res = {"config": idx, **c, "gsi_fwd_max_rel": float((y - yr).abs().max() / yr.abs().max()),
               "argmax_flips_off_near_ties": flips, "ds_fwd_max_rel": float((d - dr).abs().max() / dr.abs().max())}
        del m, y, d
        gc.collect()
        torch.cuda.empty_cache()
        losses = []
        for rep in range(2):
            torch.manual_seed(0)
            np.random.seed(0)
            m = quiet(SemiSupCycleGAN, n_classes=c["ncls"], img_channels=c["cimg"], variant="classic", use_dropout=True,
                      device="cuda:0", precision="bf16")
            g = torch.Generator().manual_seed(9)
            N = c["batch"]
            l_img = (torch.rand(N, c["cimg"], c["h"], c["w"], generator=g) * 2 - 1).cuda()
            unl = (torch.rand(N, c["cimg"], c["h"], c["w"], generator=g) * 2 - 1).cuda()
            l_gt = torch.randint(0, c["ncls"], (N, 1, c["h"], c["w"]), generator=g).cuda()
            for _ in range(2):
                out = m.train_step(l_img, l_gt, unl)
            torch.cuda.synchronize()
            losses.append({k: float(v) for k, v in out.items()})
            del m, out, l_img, unl, l_gt
            gc.collect()
            torch.cuda.empty_cache()
        res["bf16_step_losses"] = losses[0]
        res["finite"] = all(np.isfinite(v) for v in losses[0].values())
        res["reproducible"] = losses[0] == losses[1]
        res["device_error"] = K.device_error()
        res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

"""Numbers behind the tolerances of tests/test_modules_gpu.py and tests/test_step_gpu.py (prints JSON lines):
  * golden toy nets, precision bf16: gradients vs the bf16-EMULATED oracle (autograd through oracle/ref_arch.py with
    emulate_bf16=True: forward roundings at the kernels' rounding points, gradients rounded to bf16 where the kernels
    store them in bf16);
  * full-width nets: fp32 oracle vs fp64 oracle gradients (the ReLU-kink floor of fp32 itself), and the kernels in both
    precisions against the fp64 oracle / the emulated oracle;
  * whole step, precision bf16: 9 losses and gradients vs ref_step.full_step(emulate_bf16=True)."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sscg_b200  # noqa: E402,F401
from oracle import ref_arch as RA  # noqa: E402
from oracle import ref_step as RS  # noqa: E402
from sscg_b200.arch import define_Dis, define_Gen  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def cancelled(name):
    if not name.endswith(".bias"):
        return False
    parts = name.split(".")
    if parts[0] == "res_model":
        return len(parts) != 3
    return name not in ("dis_model.0.bias", "dis_model.5.bias")


def oracle_grads(kind, sd, x, probe, tanh, dtype=torch.float32, emulate=False):
    sdr = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.detach().to(dtype).clone().requires_grad_(True)
    if kind == "gen":
        y = RA.resnet_generator(sdr, xr, 9, tanh=tanh, emulate_bf16=emulate, live_norm_bias=not emulate)
    else:
        y = RA.nlayer_discriminator(sdr, xr, 3, emulate_bf16=emulate, live_norm_bias=not emulate)
    (y * probe.to(dtype)).sum().backward()
    return y.detach(), xr.grad, {k: v.grad for k, v in sdr.items()}


def worst(gk, go):
    w = ("", 0.0)
    for k, g in go.items():
        if cancelled(k) or g is None:
            continue
        r = rel(gk[k], g)
        if r > w[1]:
            w = (k, r)
    return w


def main():
    T = lambda a: torch.from_numpy(np.asarray(a))          # noqa: E731
    # ---- golden toy nets, bf16 vs emulated oracle ----------------------------------------------------
    z = np.load(os.path.join(GOLD, "gen_tiny.npz"))
    for tag, name in (("softmax", "resnet_9blocks_softmax"), ("tanh", "resnet_9blocks")):
        sd = {k[len(tag) + 3:]: T(z[k]) for k in z.files if k.startswith(tag + ".w.")}
        net = quiet(define_Gen, 3, 5, 4, name, norm="instance", use_dropout=False, gpu_ids=[0])
        net.load_state_dict(sd)
        net.precision = "bf16"
        x = T(z[tag + ".x"]).cuda().requires_grad_(True)
        y = net(x)
        probe = T(z[tag + ".probe"])
        (y * probe.cuda()).sum().backward()
        ye, gxe, gwe = oracle_grads("gen", sd, T(z[tag + ".x"]), probe, tag == "tanh", emulate=True)
        gk = {k: p.grad for k, p in net.named_parameters()}
        print(json.dumps({"check": "toy generator, bf16 vs emulated oracle", "net": name, "y_max_rel": float(
            (y.detach().cpu() - ye).abs().max() / ye.abs().max()), "gx_rel_l2": rel(x.grad, gxe),
            "worst_weight_grad": worst(gk, gwe)}), flush=True)
    z = np.load(os.path.join(GOLD, "dis_tiny.npz"))
    sd = {k[2:]: T(z[k]) for k in z.files if k.startswith("w.")}
    net = quiet(define_Dis, 3, 4, "n_layers", n_layers_D=3, norm="instance", gpu_ids=[0])
    net.load_state_dict(sd)
    net.precision = "bf16"
    x = T(z["x"]).cuda().requires_grad_(True)
    y = net(x)
    (y * T(z["probe"]).cuda()).sum().backward()
    ye, gxe, gwe = oracle_grads("dis", sd, T(z["x"]), T(z["probe"]), False, emulate=True)
    print(json.dumps({"check": "toy discriminator, bf16 vs emulated oracle", "y_max_rel": float(
        (y.detach().cpu() - ye).abs().max() / ye.abs().max()), "gx_rel_l2": rel(x.grad, gxe),
        "worst_weight_grad": worst({k: p.grad for k, p in net.named_parameters()}, gwe)}), flush=True)
    # ---- full-width nets -------------------------------------------------------------------------
    FULL = [("gen", (3, 21, "resnet_9blocks_softmax"), 2, 64, 64), ("gen", (21, 3, "resnet_9blocks"), 2, 64, 64),
            ("gen", (1, 4, "resnet_9blocks_softmax"), 1, 32, 48), ("dis", (3,), 2, 64, 64), ("dis", (21,), 2, 128, 128)]
    for kind, cfg, N, H, W in FULL:
        torch.manual_seed(0)
        if kind == "gen":
            net = quiet(define_Gen, cfg[0], cfg[1], 64, cfg[2], norm="instance", use_dropout=False, gpu_ids=[0])
            tanh = not cfg[2].endswith("softmax")
        else:
            net = quiet(define_Dis, cfg[0], 64, "n_layers", norm="instance", gpu_ids=[0])
            tanh = False
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.05)
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        x = torch.rand(N, cfg[0], H, W) * 2 - 1
        res = {}
        probe = None
        for prec in ("bf16x3", "bf16"):
            net.precision = prec
            net.zero_grad(set_to_none=True)
            xg = x.cuda().requires_grad_(True)
            y = net(xg)
            if probe is None:
                probe = torch.randn(y.shape)
            (y * probe.cuda()).sum().backward()
            res[prec] = (y.detach().cpu(), xg.grad.cpu(), {k: p.grad.cpu().clone() for k, p in net.named_parameters()})
        y64, gx64, gw64 = oracle_grads(kind, sd, x, probe, tanh, torch.float64)
        y32, gx32, gw32 = oracle_grads(kind, sd, x, probe, tanh, torch.float32)
        ye, gxe, gwe = oracle_grads(kind, sd, x, probe, tanh, torch.float32, emulate=True)
        print(json.dumps({
            "check": "full-width %s %s %dx%dx%d" % (kind, cfg, N, H, W),
            "fp32_oracle_vs_fp64": {"y": rel(y32, y64), "gx": rel(gx32, gx64), "gw": worst(gw32, gw64)},
            "bf16x3_vs_fp64": {"y": rel(res["bf16x3"][0], y64), "gx": rel(res["bf16x3"][1], gx64),
                               "gw": worst(res["bf16x3"][2], gw64)},
            "bf16_vs_emulated": {"y": rel(res["bf16"][0], ye), "gx": rel(res["bf16"][1], gxe), "gw": worst(res["bf16"][2], gwe)},
            "emulated_vs_fp64": {"y": rel(ye, y64), "gx": rel(gxe, gx64), "gw": worst(gwe, gw64)},
            "bf16_vs_fp64": {"y": rel(res["bf16"][0], y64), "gx": rel(res["bf16"][1], gx64), "gw": worst(res["bf16"][2], gw64)},
        }), flush=True)
    # ---- whole step, bf16 vs emulated oracle --------------------------------------------------------
    from sscg_b200.step import SemiSupCycleGAN
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    for variant in ("classic", "head"):
        names = ["Gis", "Gsi", "Di", "Ds"] + (["old_Gis", "old_Gsi", "old_Di"] if variant == "head" else [])
        nets = {nm: {k[len(nm) + 1:]: T(z[k]) for k in z.files if k.startswith(nm + ".")} for nm in names}
        l_img, l_gt, unl = T(z["l_img"]), T(z["l_gt"]), T(z["unl_img"])
        losses, grads, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant=variant, emulate_bf16=True)
        l32, g32, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant=variant)
        m = quiet(SemiSupCycleGAN, n_classes=21, ngf=4, ndf=4, variant=variant, use_dropout=False, device="cuda:0",
                  precision="bf16")
        m.load_state(nets)
        out = m.train_step(l_img.cuda(), l_gt.cuda(), unl.cuda())
        dl = {k: abs(float(out[k]) - losses[k]) / max(1.0, abs(losses[k])) for k in losses}
        dl32 = {k: abs(losses[k] - l32[k]) / max(1.0, abs(l32[k])) for k in losses}
        w = ("", 0.0)
        w32 = ("", 0.0)
        for nm in ("Gis", "Gsi", "Di", "Ds"):
            for pname, p in m.nets[nm].named_parameters():
                if cancelled(pname):
                    continue
                r = rel(p.grad, grads[nm][pname])
                if r > w[1]:
                    w = (nm + "." + pname, r)
                r = rel(grads[nm][pname], g32[nm][pname])
                if r > w32[1]:
                    w32 = (nm + "." + pname, r)
        print(json.dumps({"check": "toy step, bf16 vs emulated oracle", "variant": variant, "loss_rel": dl,
                          "worst_grad": w, "emulated_vs_fp32_loss_rel": dl32, "emulated_vs_fp32_worst_grad": w32}),
              flush=True)


def step_probe():
    """Per-network flat-gradient agreement of a bf16 step with the emulated oracle, toy and mid-size nets."""
    from sscg_b200.step import SemiSupCycleGAN
    for ngf, hw in ((4, 32), (16, 64), (32, 64)):
        torch.manual_seed(3)
        m = quiet(SemiSupCycleGAN, n_classes=21, ngf=ngf, ndf=ngf, variant="classic", use_dropout=False, device="cuda:0",
                  precision="bf16")
        nets = {nm: {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for nm, net in m.nets.items()}
        g = torch.Generator().manual_seed(5)
        l_img = torch.rand(2, 3, hw, hw, generator=g) * 2 - 1
        unl = torch.rand(2, 3, hw, hw, generator=g) * 2 - 1
        l_gt = torch.randint(0, 21, (2, 1, hw, hw), generator=g)
        le, ge, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant="classic", emulate_bf16=True)
        l32, g32, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant="classic")
        res = {}
        for prec in ("bf16", "bf16x3"):
            for net in m.nets.values():
                net.precision = prec
            m.load_state(nets)
            out = m.train_step(l_img.cuda(), l_gt.cuda(), unl.cuda())
            per = {}
            for nm in ("Gis", "Gsi", "Di", "Ds"):
                names = [k for k, _ in m.nets[nm].named_parameters() if not cancelled(k)]
                fk = torch.cat([dict(m.nets[nm].named_parameters())[k].grad.detach().cpu().reshape(-1).double() for k in names])
                fe = torch.cat([ge[nm][k].reshape(-1).double() for k in names])
                f32 = torch.cat([g32[nm][k].reshape(-1).double() for k in names])
                cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))     # noqa: E731
                per[nm] = {"rel_vs_emulated": float((fk - fe).norm() / fe.norm()), "cos_vs_emulated": cos(fk, fe),
                           "rel_vs_fp32": float((fk - f32).norm() / f32.norm()), "cos_vs_fp32": cos(fk, f32),
                           "emulated_rel_vs_fp32": float((fe - f32).norm() / f32.norm())}
            res[prec] = {"loss_rel_vs_emulated": {k: abs(float(out[k]) - le[k]) / max(1.0, abs(le[k])) for k in le},
                         "loss_rel_vs_fp32": {k: abs(float(out[k]) - l32[k]) / max(1.0, abs(l32[k])) for k in le},
                         "grads": per}
            # Adam moved the weights: restore for the next precision
        print(json.dumps({"check": "step flat gradients", "ngf": ngf, "hw": hw, **res}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "step":
        step_probe()
    else:
        main()

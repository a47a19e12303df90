"""ORACLE tooling — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported read-only) on CPU.  Run in the build container only:

    python -m oracle.make_golden

Fixtures (all tiny so they can be committed):
  gen_tiny.npz   — define_Gen(3, 5, ngf=4, 'resnet_9blocks_softmax' / 'resnet_9blocks') weights, input,
                   forward output and input/weight gradients of sum(out * probe)
  dis_tiny.npz   — define_Dis(3, ndf=4, 'n_layers') likewise
  step_head.npz  — the 9 scalars the reference's literal semisuper_cycleGAN.train() logs at step 0
                   (model.py:548-550) under the stub harness of SURVEY.md §4.2 (fake tensorboardX,
                   synthetic VOCDataset, deeplab->resnet_9blocks*, pixel->n_layers), together with
                   the synthetic batch and every net's initial state_dict.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import arch  # noqa: F401  (reference package)
    return arch


def _np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def golden_modules():
    arch = _import_reference()
    torch.manual_seed(0)
    out = {}
    for name, tanh in (("resnet_9blocks_softmax", False), ("resnet_9blocks", True)):
        net = arch.define_Gen(3, 5, 4, name, norm="instance", use_dropout=False, gpu_ids=[])
        # non-zero biases so that the live biases (head conv) are exercised
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.1)
        net.eval()
        x = torch.rand(2, 3, 16, 16) * 2 - 1
        x.requires_grad_(True)
        y = net(x)
        probe = torch.randn_like(y)
        (y * probe).sum().backward()
        tag = "softmax" if not tanh else "tanh"
        out.update(_np(net.state_dict(), f"{tag}.w."))
        out.update({f"{tag}.g." + k: p.grad.numpy() for k, p in net.named_parameters()})
        out[f"{tag}.x"] = x.detach().numpy()
        out[f"{tag}.y"] = y.detach().numpy()
        out[f"{tag}.probe"] = probe.numpy()
        out[f"{tag}.gx"] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "gen_tiny.npz"), **out)

    out = {}
    net = arch.define_Dis(3, 4, "n_layers", n_layers_D=3, norm="instance", gpu_ids=[])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    x.requires_grad_(True)
    y = net(x)
    probe = torch.randn_like(y)
    (y * probe).sum().backward()
    out.update(_np(net.state_dict(), "w."))
    out.update({"g." + k: p.grad.numpy() for k, p in net.named_parameters()})
    out["x"], out["y"], out["probe"], out["gx"] = x.detach().numpy(), y.detach().numpy(), probe.numpy(), x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "dis_tiny.npz"), **out)


class _Recorder:
    def __init__(self):
        self.scalars = []

    def add_scalars(self, tag, d, step):
        self.scalars.append((tag, {k: float(v) for k, v in d.items()}, int(step)))

    def add_image(self, *a, **k):
        pass


def run_reference_train(N=2, H=32, W=32, C=21, ngf=4, ndf=4, seed=0, steps=1):
    """Drive the reference's literal train() for `steps` steps on synthetic data."""
    _import_reference()
    rec = _Recorder()
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = lambda *a, **k: rec
    sys.modules["tensorboardX"] = tb
    import model as ref_model  # noqa: E402  (reference model.py; needs the stub above)

    gen = torch.Generator().manual_seed(seed + 1)
    imgs = torch.rand(2 * N * steps, 3, H, W, generator=gen) * 2 - 1
    gts = torch.randint(0, C, (2 * N * steps, 1, H, W), generator=gen)

    class Synth(torch.utils.data.Dataset):
        def __init__(self, *a, name="label", **k):
            self.off = 0 if name == "label" else N * steps
            self.n = N * steps

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return imgs[self.off + i], gts[self.off + i], "s%d" % i

    ref_model.VOCDataset = Synth
    orig_gen, orig_dis = ref_model.define_Gen, ref_model.define_Dis
    created = {}

    def gen_wrap(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, gpu_ids=[0]):
        if netG == "deeplab":
            netG = "resnet_9blocks" if output_nc == 3 else "resnet_9blocks_softmax"
        net = orig_gen(input_nc, output_nc, ngf, netG, norm=norm, use_dropout=use_dropout, gpu_ids=gpu_ids)
        created.setdefault("gen", []).append(net)
        return net

    def dis_wrap(input_nc, ndf, netD, n_layers_D=3, norm="batch", gpu_ids=[0]):
        if netD == "pixel":
            netD = "n_layers"
        net = orig_dis(input_nc, ndf, netD, n_layers_D=n_layers_D, norm=norm, gpu_ids=gpu_ids)
        created.setdefault("dis", []).append(net)
        return net

    ref_model.define_Gen, ref_model.define_Dis = gen_wrap, dis_wrap
    args = argparse.Namespace(
        epochs=1, decay_epoch=0, batch_size=N, lr=2e-4, gpu_ids=[], crop_height=H, crop_width=W, lamda_img=0.5,
        lamda_gt=0.1, lab_CE_weight=1.0, lab_MSE_weight=1.0, adversarial_weight=1.0, discriminator_weight=1.0,
        checkpoint_dir="/tmp/sscg_golden_ckpt", dataset="voc2012", norm="instance", no_dropout=True, ngf=ngf, ndf=ndf)
    torch.manual_seed(seed)
    np.random.seed(seed)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        m = ref_model.semisuper_cycleGAN(args)
        # order of construction (model.py:215-230): Gis, Gsi, old_Gis, old_Gsi ; Di, Ds, old_Di
        names_g = ["Gis", "Gsi", "old_Gis", "old_Gsi"]
        names_d = ["Di", "Ds", "old_Di"]
        init = {}
        for nm, net in zip(names_g, created["gen"]):
            init[nm] = {k: v.detach().clone() for k, v in net.state_dict().items()}
        for nm, net in zip(names_d, created["dis"]):
            init[nm] = {k: v.detach().clone() for k, v in net.state_dict().items()}
        # shuffle=True loaders: make the batch order deterministic and known
        orig_loader = ref_model.DataLoader
        ref_model.DataLoader = lambda ds, batch_size, shuffle, drop_last: orig_loader(
            ds, batch_size=batch_size, shuffle=False, drop_last=drop_last)
        try:
            m.train(args)
        except (AttributeError, StopIteration, TypeError, IndexError):
            pass   # the loop dies after validation at `iter(val_loader).next()` (model.py:577)
    finally:
        os.chdir(cwd)
        ref_model.define_Gen, ref_model.define_Dis = orig_gen, orig_dis
    return rec.scalars, init, imgs, gts


def golden_step():
    N, H, W, C = 2, 32, 32, 21
    scalars, init, imgs, gts = run_reference_train(N, H, W, C, ngf=4, ndf=4, seed=0, steps=1)
    out = {}
    first = {}
    for tag, d, step in scalars:
        if step == 0:
            first.update(d)
    for k, v in first.items():
        out["loss." + k] = np.float64(v)
    for nm, sd in init.items():
        out.update(_np(sd, nm + "."))
    out["l_img"], out["l_gt"] = imgs[:N].numpy(), gts[:N].numpy()
    out["unl_img"] = imgs[N:2 * N].numpy()
    np.savez_compressed(os.path.join(OUT, "step_head.npz"), **out)
    return first


def golden_pool(steps=140, seed=7):
    """Decision sequence of the reference's history pool (utils.py:278-299 Sample_from_Pool) under a fixed numpy
    seed, in the call order of the training loop (model.py:490-493: recon_img, fake_img, fake_gt once per step, one
    whole batch per call).  Item k of pool p is identified by the integer 1000 * p + k; the fixture records which
    item every call returned."""
    _import_reference()
    import utils as ref_utils
    np.random.seed(seed)
    pools = [ref_utils.Sample_from_Pool() for _ in range(3)]
    ret = np.zeros((steps, 3), dtype=np.int64)
    for k in range(steps):
        for p in range(3):
            ret[k, p] = int(pools[p]([np.array([1000 * p + k])])[0][0])
    np.savez_compressed(os.path.join(OUT, "pool_decisions.npz"), returned=ret, seed=np.int64(seed))
    return ret


def golden_extra():
    """state_dict layout (keys, shapes, trainable flags) and a seeded output checksum of the reference's non-hot
    families that sscg_b200.arch.extra restates as stock torch modules: deeplab, unet_128, unet_256, fc_disc."""
    import contextlib
    import io
    import json
    arch = _import_reference()
    out = {}
    cases = [("deeplab", lambda: arch.define_Gen(3, 21, 64, "deeplab", norm="instance", use_dropout=True, gpu_ids=[]), (1, 3, 65, 65)),
             ("unet_128", lambda: arch.define_Gen(3, 5, 8, "unet_128", norm="instance", use_dropout=True, gpu_ids=[]), (1, 3, 128, 128)),
             ("unet_256", lambda: arch.define_Gen(3, 5, 8, "unet_256", norm="batch", use_dropout=False, gpu_ids=[]), (1, 3, 256, 256)),
             ("fc_disc", lambda: arch.define_Dis(21, 16, "fc_disc", gpu_ids=[]), (1, 21, 64, 64))]
    for name, make, shape in cases:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            net = make()
        net.eval()
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            y = net(x)
        out[name] = {"keys": [[k, list(v.shape)] for k, v in net.state_dict().items()],
                     "trainable": [k for k, p in net.named_parameters() if p.requires_grad],
                     "out_shape": list(y.shape), "out_sum": float(y.double().sum()), "out_abs_sum": float(y.double().abs().sum())}
    with open(os.path.join(OUT, "extra_modules.json"), "w") as f:
        json.dump(out, f)
    return {k: (len(v["keys"]), v["out_shape"]) for k, v in out.items()}


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        print(golden_extra())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pool":        # add the pool fixture without touching the others
        r = golden_pool()
        print("pool_decisions:", r.shape, "stored batches returned:", int((r != np.arange(len(r))[:, None] + 1000 * np.arange(3)).sum()))
        sys.exit(0)
    golden_modules()
    golden_pool()
    golden_extra()
    print(golden_step())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

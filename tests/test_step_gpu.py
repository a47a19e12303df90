"""-m gpu: the full training step (sscg_b200.step.SemiSupCycleGAN on cuda:0, fused kernels underneath)
against (1) the 9 scalars the reference's literal train() loop logged at step 0 (golden, made by
oracle/make_golden.py from the unmodified reference) and (2) the oracle's gradients.
Tolerance: 1e-3 relative on the loss scalars in parity mode (bf16x3); gradients rel-L2 <= 5e-2: a step
chains up to three networks (~70 ReLU/LeakyReLU layers), so the rare kink flips described in
tests/test_modules_gpu.py accumulate to the 1e-2 level even though every forward value agrees to 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_step as RS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("img_dis_loss", "gt_dis_loss", "cycle_img_dis_loss", "img_gen_loss", "gt_gen_loss", "img_cycle_loss",
        "gt_cycle_loss", "lab_loss_CE", "lab_loss_MSE")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _sd(z, prefix):
    return {k[len(prefix):]: _t(z[k]) for k in z.files if k.startswith(prefix)}


def _build(variant, precision, z):
    import sscg_b200  # noqa: F401
    from sscg_b200.step import SemiSupCycleGAN
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = SemiSupCycleGAN(n_classes=21, ngf=4, ndf=4, variant=variant, use_dropout=False, device="cuda:0",
                        precision=precision)
    names = ["Gis", "Gsi", "Di", "Ds"] + (["old_Gis", "old_Gsi", "old_Di"] if variant == "head" else [])
    m.load_state({nm: _sd(z, nm + ".") for nm in names})
    return m, names


def test_head_step_matches_reference_train_loop():
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    m, names = _build("head", "bf16x3", z)
    out = m.train_step(_t(z["l_img"]).cuda(), _t(z["l_gt"]).cuda(), _t(z["unl_img"]).cuda())
    for k in KEYS:
        ref = float(z["loss." + k])
        assert abs(float(out[k]) - ref) <= 1e-3 * max(1.0, abs(ref)), (k, float(out[k]), ref)


@pytest.mark.parametrize("variant", ["classic", "head"])
def test_step_gradients_match_oracle(variant):
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    m, names = _build(variant, "bf16x3", z)
    nets = {nm: _sd(z, nm + ".") for nm in names}
    l_img, l_gt, unl = _t(z["l_img"]), _t(z["l_gt"]), _t(z["unl_img"])
    losses, grads, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant=variant)
    out = m.train_step(l_img.cuda(), l_gt.cuda(), unl.cuda())
    for k in KEYS:
        assert abs(float(out[k]) - losses[k]) <= 1e-3 * max(1.0, abs(losses[k])), (k, float(out[k]), losses[k])
    for nm in ("Gis", "Gsi", "Di", "Ds"):
        for pname, p in m.nets[nm].named_parameters():
            g = grads[nm][pname]
            parts = pname.split(".")
            cancelled = pname.endswith(".bias") and (
                (parts[0] == "res_model" and len(parts) != 3) or
                (parts[0] == "dis_model" and pname not in ("dis_model.0.bias", "dis_model.5.bias")))
            if cancelled:
                assert float(p.grad.abs().max()) == 0.0
                continue     # cancelled by InstanceNorm: fp32 noise in the oracle, exact zero here
            rel = float((p.grad.cpu() - g).norm() / max(float(g.norm()), 1e-30))
            assert rel <= 5e-2, (nm, pname, rel)


def _flat(grads, names):
    return torch.cat([grads[k].detach().reshape(-1).double().cpu() for k in names])


@pytest.mark.parametrize("variant", ["classic", "head"])
def test_bf16_step_matches_emulated_oracle(variant):
    """The BENCHMARKED mode (bf16) against the bf16-emulated oracle step (oracle/ref_step.full_step(emulate_bf16=True):
    forward roundings at the kernels' rounding points, gradients rounded to bf16 where the kernels store them in bf16).
      * the losses that are smooth functions of one or two networks: within 2e-3 (measured <= 8e-4);
      * the losses behind argmax -> one-hot (model.py:435-438,509-512) or a three-network chain are discontinuous /
        chaotic in the logits: they must stay inside the envelope of the emulation's own distance from fp32;
      * gradients, per network (all parameters flattened): the fused step is as close to the fp32 oracle as the emulation
        is (<= 1.25x + 0.02; measured 0.85-1.06x) and the two bf16 gradients point the same way (cosine >= 0.95 classic,
        >= 0.8 head; measured 0.988-0.99999 / 0.85).  Every single kernel on this path is pinned tightly in kernel_cases.py; a wrong tap
        table / mask / scale anywhere in the step gives a cosine below 0.9 or a ratio far above 1."""
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    m, names = _build(variant, "bf16", z)
    nets = {nm: _sd(z, nm + ".") for nm in names}
    l_img, l_gt, unl = _t(z["l_img"]), _t(z["l_gt"]), _t(z["unl_img"])
    le, ge, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant=variant, emulate_bf16=True)
    l32, g32, _ = RS.full_step(nets, l_img, l_gt, unl, 21, variant=variant)
    out = m.train_step(l_img.cuda(), l_gt.cuda(), unl.cuda())
    smooth = ["lab_loss_CE", "lab_loss_MSE", "gt_cycle_loss", "img_gen_loss", "img_dis_loss"]
    if variant == "classic":
        smooth.append("img_cycle_loss")
    for k in KEYS:
        got, want = float(out[k]), le[k]
        if k in smooth:
            assert abs(got - want) <= 2e-3 * max(1.0, abs(want)), (k, got, want)
        else:
            env = abs(le[k] - l32[k])
            assert abs(got - l32[k]) <= 1.5 * env + 5e-3 * max(1.0, abs(l32[k])), (k, got, want, l32[k])
    for nm in ("Gis", "Gsi", "Di", "Ds"):
        params = dict(m.nets[nm].named_parameters())
        live = [k for k in params if not (k.endswith(".bias") and (
            (k.startswith("res_model") and len(k.split(".")) != 3) or
            (k.startswith("dis_model") and k not in ("dis_model.0.bias", "dis_model.5.bias"))))]
        fk = _flat({k: params[k].grad for k in live}, live)
        fe, f32 = _flat(ge[nm], live), _flat(g32[nm], live)
        rel_k = float((fk - f32).norm() / f32.norm())
        rel_e = float((fe - f32).norm() / f32.norm())
        cos = float((fk * fe).sum() / (fk.norm() * fe.norm()))
        assert rel_k <= 1.25 * rel_e + 0.02, (nm, rel_k, rel_e)
        # head variant: the generator loss holds MSE(old_Di(Gis(softmax(Gsi(x))))) at weight 1 (model.py:432,452,466), a
        # three-network chain whose gradient is chaotic in bf16 on these 4-channel toy nets (measured cosine 0.85)
        assert cos >= (0.95 if variant == "classic" else 0.8), (nm, cos)


def test_bf16_step_runs_and_is_finite_with_dropout():
    import sscg_b200  # noqa: F401
    from sscg_b200.step import SemiSupCycleGAN
    torch.manual_seed(0)
    m = SemiSupCycleGAN(n_classes=21, variant="classic", use_dropout=True, device="cuda:0", precision="bf16")
    l_img = (torch.rand(2, 3, 64, 64) * 2 - 1).cuda()
    unl = (torch.rand(2, 3, 64, 64) * 2 - 1).cuda()
    l_gt = torch.randint(0, 21, (2, 1, 64, 64)).cuda()
    w0 = m.Gsi.res_model[1][0].weight.detach().clone()
    for _ in range(3):
        out = m.train_step(l_img, l_gt, unl)
        assert all(bool(torch.isfinite(v)) for v in out.values())
    assert float((m.Gsi.res_model[1][0].weight - w0).abs().max()) > 0      # Adam moved the weights
    # random init, 21 classes: CE near ln(21); L1 between two U(-1,1)-like images near 2/3
    assert abs(float(out["lab_loss_CE"]) - 3.04) < 0.6 and abs(float(out["lab_loss_MSE"]) - 0.66) < 0.2
    host = m.train_step_host(l_img.cpu().pin_memory(), l_gt.cpu().pin_memory(), unl.cpu().pin_memory())
    assert set(host) == set(KEYS)


def test_graphed_step_matches_eager_losses():
    """CUDA-graph replay of the whole step == eager execution of the same step (parity mode, no dropout)."""
    import sscg_b200  # noqa: F401
    from sscg_b200.step import GraphedStep, SemiSupCycleGAN
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    names = ["Gis", "Gsi", "Di", "Ds"]
    l_img, l_gt, unl = _t(z["l_img"]).cuda(), _t(z["l_gt"]).cuda(), _t(z["unl_img"]).cuda()
    outs = []
    for graph in (False, True):
        np.random.seed(0)
        torch.manual_seed(0)
        m = SemiSupCycleGAN(n_classes=21, ngf=4, ndf=4, variant="classic", use_dropout=False, device="cuda:0",
                            precision="bf16x3", graph_safe=graph)
        m.load_state({nm: _sd(z, nm + ".") for nm in names})
        if graph:
            gs = GraphedStep(m, l_img, l_gt, unl, warmup=3)      # 3 eager steps + capture (not replayed yet)
            for _ in range(2):
                before = m.Gsi.state_dict()["res_model.10.res_block.1.0.weight"].detach().clone()
                host = gs.step_host(l_img.cpu().pin_memory(), l_gt.cpu().pin_memory(), unl.cpu().pin_memory())
                # every replay re-derives the bf16 operand slabs from the weights it starts with (a stale slab
                # would freeze the network while Adam keeps moving the fp32 master copy)
                wt = [w for w in m.Gsi._runner.weights if w.spec.name == "res6.conv1"][0]
                slab = wt.w_fwd.float().view(9, wt.Co_pad, wt.Kc)[:, :before.shape[0], :before.shape[1]]
                slab = slab + wt.w_fwd_lo.float().view(9, wt.Co_pad, wt.Kc)[:, :before.shape[0], :before.shape[1]]
                want = before.permute(2, 3, 0, 1).reshape(9, before.shape[0], before.shape[1])
                assert float((slab - want).abs().max()) <= 1e-4 * float(want.abs().max())
                after = m.Gsi.state_dict()["res_model.10.res_block.1.0.weight"]
                assert float((after - before).abs().max()) > 0          # ... and the replay did update them
            assert gs.launches_per_step > 100
        else:
            for _ in range(5):
                o = m.train_step(l_img, l_gt, unl)
            host = {k: float(v) for k, v in o.items()}
        outs.append(host)
    # step 5 of training from identical weights on a fixed batch: graph replays == eager launches, bit for bit.  (Round 1
    # accumulated statistics, split-K partials and loss sums with float atomics; Adam's sign-like first updates turned
    # that last-bit noise into 1e-2-level loss differences between ANY two runs.  The reductions are order-independent
    # now — binned integer accumulators and fixed-order split-K / loss sums, csrc/sscg_ptx.cuh.)
    for k in KEYS:
        assert outs[0][k] == outs[1][k], (k, outs[0][k], outs[1][k])


def test_step_host_prefetch_overlaps_without_changing_results():
    """GraphedStep.step_host(prefetch=next batch): the next batch's host-to-device copy runs on a copy stream during the
    step; results are bit-identical to the plain path."""
    import sscg_b200  # noqa: F401
    from sscg_b200.step import GraphedStep, SemiSupCycleGAN
    g = torch.Generator().manual_seed(4)
    batches = [((torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).pin_memory(),
                torch.randint(0, 5, (2, 1, 32, 32), generator=g).pin_memory(),
                (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).pin_memory()) for _ in range(3)]
    runs = []
    for prefetch in (False, True):
        torch.manual_seed(0)
        np.random.seed(0)
        m = SemiSupCycleGAN(n_classes=5, ngf=4, ndf=4, variant="classic", use_dropout=True, device="cuda:0", precision="bf16",
                            graph_safe=True)
        gs = GraphedStep(m, *[t.cuda() for t in batches[0]], warmup=2)
        out = []
        for i in range(5):
            kw = {"prefetch": batches[(i + 1) % 3]} if prefetch else {}
            out.append(gs.step_host(*batches[i % 3], **kw))
        runs.append(out)
    assert runs[0] == runs[1]
    assert len({round(o["lab_loss_CE"], 6) for o in runs[0]}) > 1      # different batches were really consumed


def test_bf16_step_is_reproducible_with_dropout():
    """Two runs of three bf16 training steps (dropout on) from the same seeds: identical losses, gradients, weights."""
    import sscg_b200  # noqa: F401
    from sscg_b200.step import SemiSupCycleGAN
    l_img = (torch.rand(2, 3, 64, 64) * 2 - 1).cuda()
    unl = (torch.rand(2, 3, 64, 64) * 2 - 1).cuda()
    l_gt = torch.randint(0, 21, (2, 1, 64, 64)).cuda()
    res = []
    for _ in range(2):
        torch.manual_seed(0)
        np.random.seed(0)
        m = SemiSupCycleGAN(n_classes=21, variant="classic", use_dropout=True, device="cuda:0", precision="bf16")
        for _ in range(3):
            out = m.train_step(l_img, l_gt, unl)
        torch.cuda.synchronize()
        res.append(({k: float(v) for k, v in out.items()}, m.g_grads.flat.clone(), m.d_grads.flat.clone(),
                    torch.cat([p.detach().reshape(-1) for p in m.Gsi.parameters()])))
    assert res[0][0] == res[1][0]
    for i in (1, 2, 3):
        assert torch.equal(res[0][i], res[1][i])


@pytest.mark.parametrize("name,C,cimg,H,W", [("cityscapes_19", 19, 3, 128, 256), ("cityscapes_20", 20, 3, 128, 256),
                                             ("acdc", 4, 1, 128, 128), ("voc_large", 21, 3, 256, 256)])
def test_other_baseline_configs_run(name, C, cimg, H, W):
    """BASELINE.json configs #3-#5 (class counts 19/20/4, 1-channel input, non-square and larger crops) at
    reduced batch/size: forward parity of Gsi in parity mode + one finite bf16 training step."""
    import sscg_b200  # noqa: F401
    from oracle import ref_arch as RA
    from sscg_b200.step import SemiSupCycleGAN
    torch.manual_seed(0)
    m = SemiSupCycleGAN(n_classes=C, img_channels=cimg, variant="classic", use_dropout=False, device="cuda:0",
                        precision="bf16x3")
    x = torch.rand(1, cimg, H // 2, W // 2) * 2 - 1
    m.Gsi.eval()
    with torch.no_grad():
        y = m.Gsi(x.cuda()).cpu()
    yr = RA.resnet_generator({k: v.detach().cpu() for k, v in m.Gsi.state_dict().items()}, x, 9, tanh=False)
    assert float((y - yr).abs().max() / yr.abs().max()) <= 1e-3
    for net in m.nets.values():
        net.precision = "bf16"
    m.Gsi.train()
    l_img = (torch.rand(2, cimg, H, W) * 2 - 1).cuda()
    unl = (torch.rand(2, cimg, H, W) * 2 - 1).cuda()
    l_gt = torch.randint(0, C, (2, 1, H, W)).cuda()
    out = m.train_step(l_img, l_gt, unl)
    assert all(bool(torch.isfinite(v)) for v in out.values()), out

"""ORACLE — test infrastructure only.  CPU restatement of one semi-supervised CycleGAN training step
(reference model.py:370-552) on top of oracle.ref_arch, for the north-star configuration
Gis = resnet_9blocks, Gsi = resnet_9blocks_softmax, Di = Ds = n_layers(3).

Two variants:
  * "head"    — literal HEAD semantics including the frozen auxiliary nets old_Gis / old_Gsi / old_Di
                (model.py:225-230, 418-423, 432, 501-502).
  * "classic" — 2 generators + 2 discriminators with the L1 image-cycle loss that HEAD still carries
                as a comment (model.py:453), weighted by lamda_img (main.py:21).

The function returns the 9 logged scalars (model.py:548-550) plus the gradients autograd produces
for every parameter of the trained nets in both phases; optimizer updates are applied by the
caller (tests compare gradients, bench's CPU leg calls torch.optim.Adam like model.py:286-287).
"""
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import ref_arch as RA


@dataclass
class StepWeights:
    lamda_img: float = 0.5          # main.py:21 (used by the classic variant only)
    lamda_gt: float = 0.1           # main.py:22
    lab_CE_weight: float = 1.0      # main.py:25
    lab_MSE_weight: float = 1.0     # main.py:26
    adversarial_weight: float = 1.0     # main.py:29
    discriminator_weight: float = 1.0   # main.py:30


def _gen(sd, x, tanh, e):
    # bf16 emulation drops the conv biases that InstanceNorm cancels, like the kernels (they would shift the bf16
    # rounding of the raw conv outputs)
    return RA.resnet_generator(sd, x, 9, tanh=tanh, use_dropout=False, emulate_bf16=e, live_norm_bias=not e)


def _dis(sd, x, e):
    return RA.nlayer_discriminator(sd, x, 3, emulate_bf16=e, live_norm_bias=not e)


def _mse_to(x, target):
    # nn.MSELoss against torch.ones/zeros targets, model.py:441-446,514-528
    return ((x - target) ** 2).mean()


def _req(sd, flag):
    return {k: v.detach().clone().requires_grad_(flag) for k, v in sd.items()}


def generator_phase(nets, l_img, l_gt, unl_img, C, variant="classic", w=StepWeights(), emulate_bf16=False,
                    dead_forwards=False):
    """model.py:379-474 up to (not including) g_optimizer.step().
    nets: dict of state_dicts {"Gis","Gsi","Di","Ds"[,"old_Gis","old_Gsi","old_Di"]}.
    Returns (losses dict, grads dict {"Gis": {...}, "Gsi": {...}}, tensors needed by the D phase)."""
    e = emulate_bf16
    Gis, Gsi = _req(nets["Gis"], True), _req(nets["Gsi"], True)
    Di, Ds = _req(nets["Di"], False), _req(nets["Ds"], False)          # set_grad(False), model.py:379
    fake_img = _gen(Gis, RA.make_one_hot(l_gt, C), True, e)               # model.py:385
    fake_gt = _gen(Gsi, unl_img, False, e)                                # :386
    lab_gt = _gen(Gsi, l_img, False, e)                                   # :387
    lab_loss_CE = F.cross_entropy(lab_gt, l_gt.squeeze(1))                # :398
    lab_gt_p = F.softmax(lab_gt, dim=1)                                   # :401
    fake_gt_p = F.softmax(fake_gt, dim=1)                                 # :402
    recon_img = _gen(Gis, fake_gt_p, True, e)                             # :408
    # recon_lab_img = Gis(lab_gt_p) (model.py:409) feeds no loss: it cannot change any
    # result; `dead_forwards=True` still executes it so that CPU-baseline timings do the reference's work
    if dead_forwards:
        with torch.no_grad():
            _gen(Gis, lab_gt_p.detach(), True, e)
    recon_gt = _gen(Gsi, fake_img, False, e)                              # :410
    fake_img_dis = _dis(Di, fake_img, e)                                  # :431
    fake_gt_oh = RA.argmax_one_hot(fake_gt_p, C)                          # :435-437
    fake_gt_dis = _dis(Ds, fake_gt_oh, e)                                 # :438
    img_gen_loss = _mse_to(fake_img_dis, 1.0)                             # :445
    gt_gen_loss = _mse_to(fake_gt_dis, 1.0)                               # :446
    gt_cycle_loss = F.cross_entropy(recon_gt, l_gt.squeeze(1))            # :455
    lab_loss_MSE = (fake_img - l_img).abs().mean()                        # :461 (named MSE, is L1)
    extra = {}
    if variant == "head":
        old_Gis, old_Gsi = _req(nets["old_Gis"], False), _req(nets["old_Gsi"], False)
        old_Di = _req(nets["old_Di"], False)
        resnet_fake_gt = F.softmax(_gen(old_Gsi, unl_img, False, e), dim=1)       # :418,421
        resnet_recon_img = _gen(old_Gis, resnet_fake_gt, True, e)                 # :422
        if dead_forwards:                                                         # :419,423 (unused results)
            with torch.no_grad():
                _gen(old_Gis, F.softmax(_gen(old_Gsi, l_img, False, e), dim=1), True, e)
        img_cycle_loss = _mse_to(_dis(old_Di, recon_img, e), 1.0)                 # :432,452
        unsup = w.adversarial_weight * (img_gen_loss + gt_gen_loss) + img_cycle_loss + gt_cycle_loss * w.lamda_gt  # :466
        extra["resnet_recon_img"] = resnet_recon_img.detach()
    else:
        img_cycle_loss = (recon_img - unl_img).abs().mean()                       # :453 (commented at HEAD)
        unsup = (w.adversarial_weight * (img_gen_loss + gt_gen_loss) + img_cycle_loss * w.lamda_img
                 + gt_cycle_loss * w.lamda_gt)
    full = w.lab_CE_weight * lab_loss_CE + w.lab_MSE_weight * lab_loss_MSE       # :464
    gen_loss = full + unsup                                                     # :468
    gen_loss.backward()                                                         # :472
    losses = {"lab_loss_CE": lab_loss_CE, "lab_loss_MSE": lab_loss_MSE, "img_gen_loss": img_gen_loss,
              "gt_gen_loss": gt_gen_loss, "img_cycle_loss": img_cycle_loss, "gt_cycle_loss": gt_cycle_loss}
    losses = {k: float(v.detach()) for k, v in losses.items()}
    grads = {"Gis": {k: v.grad for k, v in Gis.items()}, "Gsi": {k: v.grad for k, v in Gsi.items()}}
    tensors = {"fake_img": fake_img.detach(), "fake_gt": fake_gt_p.detach(), "recon_img": recon_img.detach(),
               "lab_gt": lab_gt.detach(), "recon_gt": recon_gt.detach(), **extra}
    return losses, grads, tensors


def discriminator_phase(nets, l_gt, unl_img, fake_img, fake_gt, recon_img, C, variant="classic", w=StepWeights(),
                        resnet_recon_img=None, emulate_bf16=False):
    """model.py:481-539 up to (not including) d_optimizer.step().  fake_img / fake_gt / recon_img are
    the history-pool outputs (model.py:490-495); with a pool that is not yet full they are the
    current batch (utils.py:286-289)."""
    e = emulate_bf16
    Di, Ds = _req(nets["Di"], True), _req(nets["Ds"], True)              # set_grad(True), model.py:481
    unl_img_dis = _dis(Di, unl_img, e)                                     # :499
    fake_img_dis = _dis(Di, fake_img, e)                                   # :500
    real_gt_dis = _dis(Ds, RA.make_one_hot(l_gt, C), e)                    # :506-507
    fake_gt_dis = _dis(Ds, RA.argmax_one_hot(fake_gt, C), e)               # :509-512
    img_dis_loss = (_mse_to(unl_img_dis, 1.0) + _mse_to(fake_img_dis, 0.0)) * 0.5     # :521-522,531
    gt_dis_loss = (_mse_to(real_gt_dis, 1.0) + _mse_to(fake_gt_dis, 0.0)) * 0.5       # :523-524,532
    grads = {}
    if variant == "head":
        old_Di = _req(nets["old_Di"], True)
        cycle = _mse_to(_dis(old_Di, resnet_recon_img, e), 1.0) + _mse_to(_dis(old_Di, recon_img, e), 0.0)  # :501-502,527-528,534
        total = w.discriminator_weight * (img_dis_loss + gt_dis_loss) + cycle           # :538
    else:
        cycle = torch.zeros(())
        total = w.discriminator_weight * (img_dis_loss + gt_dis_loss)
    total.backward()                                                        # :539
    grads["Di"] = {k: v.grad for k, v in Di.items()}
    grads["Ds"] = {k: v.grad for k, v in Ds.items()}
    if variant == "head":
        grads["old_Di"] = {k: v.grad for k, v in old_Di.items()}
    losses = {"img_dis_loss": float(img_dis_loss.detach()), "gt_dis_loss": float(gt_dis_loss.detach()),
              "cycle_img_dis_loss": float(cycle.detach())}
    return losses, grads


def full_step(nets, l_img, l_gt, unl_img, C, variant="classic", w=StepWeights(), emulate_bf16=False,
              dead_forwards=False):
    """One step with an empty history pool (pool passes the current batch through, utils.py:286-289).
    NOTE: the reference updates the generators (g_optimizer.step(), model.py:474) before the D phase,
    but the D phase only consumes tensors produced before that update, so gradients of both phases
    are functions of the pre-step weights."""
    gl, gg, t = generator_phase(nets, l_img, l_gt, unl_img, C, variant, w, emulate_bf16, dead_forwards)
    dl, dg = discriminator_phase(nets, l_gt, unl_img, t["fake_img"], t["fake_gt"], t["recon_img"], C, variant, w,
                                 t.get("resnet_recon_img"), emulate_bf16)
    losses = dict(gl)
    losses.update(dl)
    grads = dict(gg)
    grads.update(dg)
    return losses, grads, t

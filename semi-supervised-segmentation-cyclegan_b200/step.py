"""The semi-supervised CycleGAN training step (reference model.py:370-552) on top of the drop-in
modules — the caller of the hot path.

`SemiSupCycleGAN.train_step` follows the reference's step line by line (citations inline) with the
north-star network choice (Gis = resnet_9blocks, Gsi = resnet_9blocks_softmax, Di = Ds = n_layers(3),
SURVEY.md §3.2) and two variants:
  * 'head'    — literal HEAD semantics with the frozen auxiliary nets old_Gis / old_Gsi / old_Di;
  * 'classic' — 2 generators + 2 discriminators and the L1 image-cycle loss (model.py:453).
Differences from the reference, all outside the arithmetic of the step:
  * the history pool keeps whole batches ON THE DEVICE (utils.py:278-299 keeps numpy copies and
    round-trips 113 MB over PCIe per step, model.py:490-495); the swap decisions consume
    numpy's RNG exactly like the reference;
  * data parallelism: one process per GPU, gradients of each optimizer live in one flat fp32
    bucket that is all-reduced (AVG) over NCCL after gen_loss.backward() / discriminator_loss.backward();
  * `interp` (model.py:268) is the identity for the ResNet generators (same output size) and is elided
    after checking the shapes.
"""
import copy
import itertools
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from .arch import define_Dis, define_Gen, set_grad
from .losses import l1_loss, lsgan_loss, seg_head


@dataclass
class StepWeights:
    lamda_img: float = 0.5              # main.py:21 (classic variant only; unused at HEAD)
    lamda_gt: float = 0.1               # main.py:22
    lab_CE_weight: float = 1.0          # main.py:25
    lab_MSE_weight: float = 1.0         # main.py:26
    adversarial_weight: float = 1.0     # main.py:29
    discriminator_weight: float = 1.0   # main.py:30


def make_one_hot(labels, C):
    # reference utils.py:314-350 — scatter_ of ones along the channel axis
    one_hot = torch.zeros(labels.size(0), C, labels.size(2), labels.size(3), dtype=torch.float32, device=labels.device)
    return one_hot.scatter_(1, labels.long(), 1)


class DevicePool:
    """Device-resident Sample_from_Pool (reference utils.py:278-299): stores whole batches; once 50
    are held, returns a stored batch with probability 0.5 (and replaces it with the new one)."""

    def __init__(self, max_elements=50):
        self.max_elements = max_elements
        self.cur_elements = 0
        self.items = []

    def __call__(self, in_items):
        return_items = []
        for in_item in in_items:
            if self.cur_elements < self.max_elements:
                self.items.append(in_item)
                self.cur_elements = self.cur_elements + 1
                return_items.append(in_item)
            else:
                if np.random.ranf() > 0.5:
                    idx = np.random.randint(0, self.max_elements)
                    tmp = self.items[idx]
                    self.items[idx] = in_item
                    return_items.append(tmp)
                else:
                    return_items.append(in_item)
        return return_items


class GraphPool:
    """CUDA-graph-safe history pool with the same decisions as Sample_from_Pool (utils.py:278-299).
    The host draws from numpy's RNG exactly like the reference (`host_decide`, called once per step
    and pool, in the reference's call order) and hands (use_stored, store, idx) to the device through
    a small int64 tensor; `device_apply` is made of capturable torch ops on a static [max, ...] store."""

    def __init__(self, max_elements=50):
        self.max_elements = max_elements
        self.cur_elements = 0
        self.storage = None
        self.dec = None          # device int64 [3] = (use_stored, store, idx)

    def host_decide(self):
        if self.cur_elements < self.max_elements:
            idx = self.cur_elements
            self.cur_elements += 1
            return (0, 1, idx)
        if np.random.ranf() > 0.5:
            return (1, 1, int(np.random.randint(0, self.max_elements)))
        return (0, 0, 0)

    def device_apply(self, x):
        if self.storage is None:
            self.storage = torch.zeros((self.max_elements,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        idx = self.dec[2:3]
        old = self.storage.index_select(0, idx)[0]
        out = torch.where(self.dec[0] != 0, old, x)
        self.storage.index_copy_(0, idx, torch.where(self.dec[1] != 0, x, old).unsqueeze(0))
        return out


class FlatGrads:
    """All gradients of one optimizer in a single fp32 bucket (p.grad are views into it)."""

    def __init__(self, params):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat.is_cuda:        # NCCL averages in the collective (one launch less per bucket)
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:                        # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))


class _StockGrads:
    """Gradient bookkeeping of the stock-torch arm: plain .grad tensors owned by autograd."""

    def __init__(self, params):
        self.params = list(params)

    def zero(self):
        for p in self.params:
            p.grad = None

    def allreduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in self.params:
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
                    p.grad.div_(dist.get_world_size(group))


class SemiSupCycleGAN:
    def __init__(self, n_classes=21, img_channels=3, ngf=64, ndf=64, variant="classic", use_dropout=True, lr=2e-4,
                 device="cuda", precision=None, weights=None, keep_dead_forward=True, fused_adam=True,
                 graph_safe=False, stock=None, batch_passes=None):
        """stock: None = this repo's fused path on CUDA.  "fp32" | "tf32" | "bf16_autocast" = the SAME step on the
        stock torch.nn module trees (cuDNN convolutions, ATen InstanceNorm / losses, torch.optim.Adam): the on-box GPU
        baseline of bench.py (the reference's own modules executed by stock PyTorch, arch/ops.py:40-74)."""
        assert variant in ("classic", "head")
        assert stock in (None, "fp32", "tf32", "bf16_autocast")
        self.stock = stock
        # batch_passes: run the pairs of passes that share a network and have independent inputs — Gsi(unl_img) with
        # Gsi(l_img) (model.py:386-387), Gis(one_hot(l_gt)) with Gis(fake_gt) (:385,408), Di(unl_img) with Di(fake_img)
        # (:499-500), Ds(real) with Ds(fake) (:506-512) — as ONE pass over the concatenated batch.  InstanceNorm is per
        # sample and every loss is a batch mean, so no result changes; default on for the fused CUDA path.
        self.batch_passes = batch_passes
        self.C, self.variant = n_classes, variant
        self.w = weights or StepWeights()
        self.keep_dead_forward = keep_dead_forward
        gpu_ids = [torch.device(device).index or 0] if torch.device(device).type == "cuda" else []
        # same construction order as the reference (model.py:215-230)
        self.Gis = define_Gen(n_classes, img_channels, ngf, "resnet_9blocks", norm="instance", use_dropout=use_dropout,
                              gpu_ids=gpu_ids)
        self.Gsi = define_Gen(img_channels, n_classes, ngf, "resnet_9blocks_softmax", norm="instance",
                              use_dropout=use_dropout, gpu_ids=gpu_ids)
        self.Di = define_Dis(img_channels, ndf, "n_layers", n_layers_D=3, norm="instance", gpu_ids=gpu_ids)
        self.Ds = define_Dis(n_classes, ndf, "n_layers", n_layers_D=3, norm="instance", gpu_ids=gpu_ids)
        self.nets = {"Gis": self.Gis, "Gsi": self.Gsi, "Di": self.Di, "Ds": self.Ds}
        if variant == "head":
            self.old_Gis = define_Gen(n_classes, img_channels, ngf, "resnet_9blocks", norm="instance",
                                      use_dropout=use_dropout, gpu_ids=gpu_ids)
            self.old_Gsi = define_Gen(img_channels, n_classes, ngf, "resnet_9blocks_softmax", norm="instance",
                                      use_dropout=use_dropout, gpu_ids=gpu_ids)
            self.old_Di = define_Dis(img_channels, ndf, "n_layers", n_layers_D=3, norm="instance", gpu_ids=gpu_ids)
            self.nets.update({"old_Gis": self.old_Gis, "old_Gsi": self.old_Gsi, "old_Di": self.old_Di})
        for n in self.nets.values():
            n.precision = precision
            if stock is not None:
                n.fusable = False                      # forward() takes the stock nn.Sequential path
                if stock == "bf16_autocast":
                    n.to(memory_format=torch.channels_last)
        self.MSE, self.L1, self.CE = nn.MSELoss(), nn.L1Loss(), nn.CrossEntropyLoss()     # model.py:270-272
        self.softmax = nn.Softmax2d()                                                     # model.py:273
        g_params = list(itertools.chain(self.Gis.parameters(), self.Gsi.parameters()))
        d_params = list(itertools.chain(self.Di.parameters(), self.Ds.parameters()))
        cuda = torch.device(device).type == "cuda"
        self.graph_safe = graph_safe
        if graph_safe:
            # device-side step counter: mixed into the dropout seeds so that replays draw fresh masks
            self.step_counter = torch.zeros(1, dtype=torch.int64, device=device)
            for n in self.nets.values():
                if getattr(n, "_runner", None) is not None:
                    n._runner.drop_ctr = self.step_counter
        if stock is not None:
            self.g_grads, self.d_grads = _StockGrads(g_params), _StockGrads(d_params)
        else:
            self.g_grads, self.d_grads = FlatGrads(g_params), FlatGrads(d_params)
        if stock is not None:
            kw = {"capturable": True} if graph_safe else {}
            self.g_optimizer = torch.optim.Adam(g_params, lr=lr, betas=(0.5, 0.999), fused=cuda or None, **kw)
            self.d_optimizer = torch.optim.Adam(d_params, lr=lr, betas=(0.5, 0.999), fused=cuda or None, **kw)
        elif fused_adam and cuda:
            # one flat-bucket Adam launch per optimizer (optim.FlatAdam, sscg_adam_flat); graph capturable
            from .optim import FlatAdam
            self.g_optimizer = FlatAdam(g_params, self.g_grads, lr=lr, betas=(0.5, 0.999),
                                        post_step=[lambda: self._invalidate_weights((self.Gis, self.Gsi))])   # model.py:286
            self.d_optimizer = FlatAdam(d_params, self.d_grads, lr=lr, betas=(0.5, 0.999),
                                        post_step=[lambda: self._invalidate_weights((self.Di, self.Ds))])     # model.py:287
        else:
            kw = {"capturable": True} if graph_safe else {}
            self.g_optimizer = torch.optim.Adam(g_params, lr=lr, betas=(0.5, 0.999), **kw)       # model.py:286
            self.d_optimizer = torch.optim.Adam(d_params, lr=lr, betas=(0.5, 0.999), **kw)       # model.py:287
        for n in (self.Gis, self.Gsi, self.Di, self.Ds):          # p.grad are persistent views: accumulate in place
            if getattr(n, "_runner", None) is not None:
                n._runner.direct_grad = True
                n._runner.defer_unpack = True
        P = GraphPool if graph_safe else DevicePool
        self.pool_recon, self.pool_fake_img, self.pool_fake_gt = P(), P(), P()           # model.py:350-352
        self._dec_fed = False
        if graph_safe:
            self.pool_dec = torch.zeros(3, 3, dtype=torch.int64, device=device)
            self.pool_dec_host = torch.zeros(3, 3, dtype=torch.int64).pin_memory()
            for i, p in enumerate((self.pool_recon, self.pool_fake_img, self.pool_fake_gt)):
                p.dec = self.pool_dec[i]
        self.Gsi.train()
        self.Gis.train()                                                                 # model.py:363-364

    @staticmethod
    def _invalidate_weights(nets):
        """The flat Adam kernel writes the parameters behind autograd's back: tell the runners to re-derive
        their bf16 weight slabs at the next forward."""
        for n in nets:
            r = getattr(n, "_runner", None)
            if r is not None:
                r.invalidate_weights()

    @staticmethod
    def _flush_wgrad(nets):
        for n in nets:
            r = getattr(n, "_runner", None)
            if r is not None and r.specs is not None:
                r.flush_wgrad()

    def load_state(self, sds):
        for k, sd in sds.items():
            self.nets[k].load_state_dict(sd)

    def _ones(self, t):
        return torch.ones_like(t)

    def train_step(self, l_img, l_gt, unl_img):
        """One optimisation step on device tensors: l_img, unl_img N x Cimg x H x W fp32 in [-1, 1],
        l_gt N x 1 x H x W int64.  Returns the 9 logged scalars (model.py:548-550) as 0-d device tensors."""
        import contextlib
        ctx = contextlib.nullcontext()
        if self.stock == "bf16_autocast":
            ctx = torch.autocast("cuda", dtype=torch.bfloat16)
            l_img = l_img.contiguous(memory_format=torch.channels_last)
            unl_img = unl_img.contiguous(memory_format=torch.channels_last)
        with ctx:
            for point in self.step_segments(l_img, l_gt, unl_img):
                (self.g_grads if point == "g" else self.d_grads).allreduce_mean()
        return self.last_losses

    def step_segments(self, l_img, l_gt, unl_img):
        """The step as a generator that yields at the two points where gradients must be all-reduced
        ("g" after gen_loss.backward(), "d" after discriminator_loss.backward()).  The eager driver
        (train_step) and the CUDA-graph driver (GraphedStep: one graph per segment, collectives between
        them stay eager) share this single body.  Results land in self.last_losses."""
        C, w = self.C, self.w
        head = self.variant == "head"
        frozen_d = [self.Di, self.Ds] + ([self.old_Di] if head else [])
        # ---- generator phase (model.py:379-474) --------------------------------------------
        set_grad(frozen_d, False)                                                        # :379
        if head:
            set_grad([self.old_Gsi, self.old_Gis], False)                                # :380
        self.g_grads.zero()                                                              # :381
        onehot_in = l_img.is_cuda and self.stock is None     # label glue: one-hot inputs go in as label maps (arch.*.forward_onehot)
        fused = l_img.is_cuda and self.stock is None         # fused softmax + cross-entropy + argmax kernel (losses.py)
        batched = fused and (self.batch_passes if self.batch_passes is not None else True)
        N = l_img.shape[0]
        if batched:
            # Gsi(unl_img) and Gsi(l_img) as one pass (:386-387); softmax / CE / argmax of both halves in one launch: the
            # unlabeled half carries ignore_index labels, so the mean cross-entropy is over the labeled half (:398)
            logits2 = self.Gsi.forward_parts([unl_img.float(), l_img])
            lab2 = torch.cat([torch.full_like(l_gt, -100), l_gt])
            lab_loss_CE, probs2, arg2 = seg_head(logits2, lab2)                          # :398,401-402,435
            fake_gt, lab_gt = probs2[:N], probs2[N:]
            fake_gt_arg = arg2[:N]
            # Gis(one_hot(l_gt)) and Gis(fake_gt) as one pass (:385,408); the input gradient is computed for the second
            # half only (the first is data)
            img2 = self.Gis.forward_parts([l_gt, fake_gt])
            fake_img, recon_img = img2[:N], img2[N:]
        else:
            fake_img = (self.Gis.forward_onehot(l_gt) if onehot_in
                        else self.Gis(make_one_hot(l_gt, C).float()))                    # :385
            fake_gt = self.Gsi(unl_img.float())                                          # :386
            lab_gt = self.Gsi(l_img)                                                     # :387
            if fused:
                lab_loss_CE, lab_gt, _ = seg_head(lab_gt, l_gt)                          # :398,401
                _, fake_gt, fake_gt_arg = seg_head(fake_gt, None)                        # :402,435
            else:
                lab_loss_CE = self.CE(lab_gt, l_gt.squeeze(1))                           # :398
                lab_gt = self.softmax(lab_gt)                                            # :401
                fake_gt = self.softmax(fake_gt)                                          # :402
                fake_gt_arg = fake_gt.data.max(1)[1]                                     # :435
            recon_img = self.Gis(fake_gt.float())                                        # :408
        assert fake_img.shape[2:] == l_img.shape[2:] and fake_gt.shape[2:] == l_img.shape[2:]   # interp == identity
        if self.keep_dead_forward:
            with torch.no_grad():
                self.Gis(lab_gt.float())      # recon_lab_img (:409) feeds no loss: forward only
        recon_gt = self.Gsi(fake_img.float())                                            # :410
        if head:
            with torch.no_grad():             # frozen aux nets, inputs without grad (:418-423)
                resnet_fake_gt = self.softmax(self.old_Gsi(unl_img.float()))
                resnet_lab_gt = self.softmax(self.old_Gsi(l_img))
                resnet_recon_img = self.old_Gis(resnet_fake_gt.float())
                self.old_Gis(resnet_lab_gt.float())                                      # resnet_recon_lab_img: unused
        fake_img_dis = self.Di(fake_img)                                                 # :431
        if onehot_in:
            fake_gt_dis = self.Ds.forward_onehot(fake_gt_arg.unsqueeze(1))               # :435-438
        else:
            fake_gt_disc = make_one_hot(fake_gt_arg.unsqueeze(1), C)                     # :435-437
            fake_gt_dis = self.Ds(fake_gt_disc.float())                                  # :438
        if fused:     # fused LSGAN / L1 kernels (losses.py): scalar target, one pass forward, one backward
            MSE1 = lambda t: lsgan_loss(t, 1.0)
            MSE0 = lambda t: lsgan_loss(t, 0.0)
            L1 = l1_loss
        else:
            MSE1 = lambda t: self.MSE(t, torch.ones_like(t))
            MSE0 = lambda t: self.MSE(t, torch.zeros_like(t))
            L1 = self.L1
        img_gen_loss = MSE1(fake_img_dis)                                                # :445
        gt_gen_loss = MSE1(fake_gt_dis)                                                  # :446
        if fused:
            gt_cycle_loss, _, _ = seg_head(recon_gt, l_gt)                               # :455
        else:
            gt_cycle_loss = self.CE(recon_gt, l_gt.squeeze(1))                           # :455
        lab_loss_MSE = L1(fake_img, l_img)                                               # :461
        fullsupervisedloss = w.lab_CE_weight * lab_loss_CE + w.lab_MSE_weight * lab_loss_MSE      # :464
        if head:
            resnet_fake_img_dis = self.old_Di(recon_img)                                 # :432
            img_cycle_loss = MSE1(resnet_fake_img_dis)                                   # :452
            unsupervisedloss = (w.adversarial_weight * (img_gen_loss + gt_gen_loss) + img_cycle_loss
                                + gt_cycle_loss * w.lamda_gt)                            # :466
        else:
            img_cycle_loss = L1(recon_img, unl_img)                                      # :453
            unsupervisedloss = (w.adversarial_weight * (img_gen_loss + gt_gen_loss) + img_cycle_loss * w.lamda_img
                                + gt_cycle_loss * w.lamda_gt)
        gen_loss = fullsupervisedloss + unsupervisedloss                                 # :468
        gen_loss.backward()                                                              # :472
        self._flush_wgrad((self.Gis, self.Gsi))
        yield "g"
        self.g_optimizer.step()                                                          # :474
        # ---- discriminator phase (model.py:481-542) ----------------------------------------
        set_grad(frozen_d, True)                                                         # :481
        self.d_grads.zero()                                                              # :482
        if self.graph_safe:
            # the decisions of this step come from feed_pool_decisions(): GraphedStep / train_step_host call it before
            # the step; a direct train_step() draws them here (never during capture: a replay reads whatever was fed)
            if not torch.cuda.is_current_stream_capturing():
                if not self._dec_fed:
                    self.feed_pool_decisions()
                self._dec_fed = False
            recon_img = self.pool_recon.device_apply(recon_img.detach())                 # :490
            fake_img = self.pool_fake_img.device_apply(fake_img.detach())                # :491
            fake_gt = self.pool_fake_gt.device_apply(fake_gt.detach())                   # :493
            self.step_counter.add_(1)
        else:
            recon_img = self.pool_recon([recon_img.detach()])[0]                         # :490
            fake_img = self.pool_fake_img([fake_img.detach()])[0]                        # :491
            fake_gt = self.pool_fake_gt([fake_gt.detach()])[0]                           # :493
        if batched:
            d2 = self.Di.forward_parts([unl_img, fake_img])                              # :499-500 as one pass
            unl_img_dis, fake_img_dis = d2[:N], d2[N:]
            s2 = self.Ds.forward_parts([l_gt, fake_gt.data.max(1)[1].unsqueeze(1)])      # :506-512 as one pass
            real_gt_dis, fake_gt_dis = s2[:N], s2[N:]
        else:
            unl_img_dis = self.Di(unl_img)                                               # :499
            fake_img_dis = self.Di(fake_img)                                             # :500
        if batched:
            pass
        elif onehot_in:
            real_gt_dis = self.Ds.forward_onehot(l_gt)                                   # :506-507
            fake_gt_dis = self.Ds.forward_onehot(fake_gt.data.max(1)[1].unsqueeze(1))    # :509-512
        else:
            real_gt_dis = self.Ds(make_one_hot(l_gt, C).float())                         # :506-507
            fake_gt_disc = make_one_hot(fake_gt.data.max(1)[1].unsqueeze(1), C)          # :509-511
            fake_gt_dis = self.Ds(fake_gt_disc.float())                                  # :512
        img_dis_loss = (MSE1(unl_img_dis) + MSE0(fake_img_dis)) * 0.5                    # :521-522,531
        gt_dis_loss = (MSE1(real_gt_dis) + MSE0(fake_gt_dis)) * 0.5                      # :523-524,532
        if head:
            resnet_recon_img_dis = self.old_Di(resnet_recon_img)                         # :501
            resnet_fake_img_dis = self.old_Di(recon_img)                                 # :502
            cycle_img_dis_loss = MSE1(resnet_recon_img_dis) + MSE0(resnet_fake_img_dis)   # :527-534
            discriminator_loss = w.discriminator_weight * (img_dis_loss + gt_dis_loss) + cycle_img_dis_loss  # :538
        else:
            cycle_img_dis_loss = torch.zeros((), device=l_img.device)
            discriminator_loss = w.discriminator_weight * (img_dis_loss + gt_dis_loss)
        discriminator_loss.backward()                                                    # :539
        self._flush_wgrad((self.Di, self.Ds))
        yield "d"
        self.d_optimizer.step()                                                          # :542
        self.last_losses = {"img_dis_loss": img_dis_loss.detach(), "gt_dis_loss": gt_dis_loss.detach(),
                "cycle_img_dis_loss": cycle_img_dis_loss.detach(), "img_gen_loss": img_gen_loss.detach(),
                "gt_gen_loss": gt_gen_loss.detach(), "img_cycle_loss": img_cycle_loss.detach(),
                "gt_cycle_loss": gt_cycle_loss.detach(), "lab_loss_CE": lab_loss_CE.detach(),
                "lab_loss_MSE": lab_loss_MSE.detach()}

    def feed_pool_decisions(self):
        """graph_safe mode: draw this step's pool decisions on the host (numpy RNG, reference order:
        recon_img, fake_img, fake_gt — model.py:490-493) and ship them to the device asynchronously."""
        for i, p in enumerate((self.pool_recon, self.pool_fake_img, self.pool_fake_gt)):
            d = p.host_decide()
            self.pool_dec_host[i, 0], self.pool_dec_host[i, 1], self.pool_dec_host[i, 2] = d
        self.pool_dec.copy_(self.pool_dec_host, non_blocking=True)
        self._dec_fed = True

    def train_step_host(self, l_img_host, l_gt_host, unl_img_host):
        """End-to-end entry: pinned host buffers in, the 9 scalars out on the host."""
        dev = next(self.Gis.parameters()).device
        if self.graph_safe:
            self.feed_pool_decisions()
        l_img = l_img_host.to(dev, non_blocking=True)
        l_gt = l_gt_host.to(dev, non_blocking=True)
        unl_img = unl_img_host.to(dev, non_blocking=True)
        out = self.train_step(l_img, l_gt, unl_img)
        stacked = torch.stack([out[k] for k in sorted(out)])
        host = stacked.cpu()
        return {k: float(v) for k, v in zip(sorted(out), host)}


class GraphedStep:
    """The whole training step (both phases, both optimizer updates, NCCL all-reduces) captured once
    into a CUDA graph and replayed: the ~2000 kernel launches of a step cost one graph launch on the
    host.  Inputs are copied into static device buffers; the 9 loss scalars come back in a static
    tensor (`losses`, ordered like sorted(KEYS))."""

    KEYS = ("cycle_img_dis_loss", "gt_cycle_loss", "gt_dis_loss", "gt_gen_loss", "img_cycle_loss", "img_dis_loss",
            "img_gen_loss", "lab_loss_CE", "lab_loss_MSE")

    def __init__(self, model: SemiSupCycleGAN, l_img, l_gt, unl_img, warmup=3):
        assert model.graph_safe, "build the model with graph_safe=True"
        from . import kernels as K
        self.m = model
        self.l_img, self.l_gt, self.unl_img = l_img.clone(), l_gt.clone(), unl_img.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # real steps: allocate every buffer, reach steady state
                model.feed_pool_decisions()
                model.train_step(self.l_img, self.l_gt, self.unl_img)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        import torch.distributed as dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # capture records the kernels without running them: no pool decision is consumed here (the
        # captured kernels read whatever decision is in model.pool_dec at replay time)
        l0 = K.launch_count()
        if self.world == 1:
            self.graphs = [torch.cuda.CUDAGraph()]
            with torch.cuda.graph(self.graphs[0]):
                out = model.train_step(self.l_img, self.l_gt, self.unl_img)
                self.losses = torch.stack([out[k] for k in self.KEYS])
            self.points = []
        else:
            # NCCL collectives stay outside the graphs: segment | all-reduce G | segment | all-reduce D | segment
            pool = torch.cuda.graph_pool_handle()
            gen = model.step_segments(self.l_img, self.l_gt, self.unl_img)
            self.graphs, self.points = [], []
            done = False
            while not done:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    try:
                        self.points.append(next(gen))
                    except StopIteration:
                        done = True
                        out = model.last_losses
                        self.losses = torch.stack([out[k] for k in self.KEYS])
                self.graphs.append(g)
        self.launches_per_step = K.launch_count() - l0
        torch.cuda.synchronize()

    def __call__(self, l_img, l_gt, unl_img):
        """Device or pinned-host inputs; returns the static loss tensor (valid after the replay)."""
        self.l_img.copy_(l_img, non_blocking=True)
        self.l_gt.copy_(l_gt, non_blocking=True)
        self.unl_img.copy_(unl_img, non_blocking=True)
        self.m.feed_pool_decisions()
        for opt in (self.m.g_optimizer, self.m.d_optimizer):      # LambdaLR rewrites param_groups: push to the device scalar
            if hasattr(opt, "sync_lr"):
                opt.sync_lr()
        for i, g in enumerate(self.graphs):
            g.replay()
            if i < len(self.points):
                (self.m.g_grads if self.points[i] == "g" else self.m.d_grads).allreduce_mean()
        self.m._dec_fed = False           # consumed by the replay
        return self.losses

    def step_host(self, l_img_host, l_gt_host, unl_img_host, prefetch=None):
        """End-to-end: pinned host buffers in, the 9 scalars on the host out.
        prefetch = the NEXT step's (l_img, l_gt, unl_img) pinned host batch (optional): its host-to-device copy is
        enqueued on a copy stream right after this step's graph has been launched and overlaps the step (a 33 MB batch
        is 1.3 ms of PCIe time at bs 16); the next call finds the batch in device staging buffers."""
        batch = (l_img_host, l_gt_host, unl_img_host)
        staged = getattr(self, "_staged", None)
        if staged is not None and all(a is b for a, b in zip(staged[0], batch)):
            torch.cuda.current_stream().wait_event(staged[2])
            batch = staged[1]                                     # device copies made during the previous step
        losses = self(*batch)
        self._staged = None
        if prefetch is not None:
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream()
                self._stage_sets = [[torch.empty_like(self.l_img), torch.empty_like(self.l_gt), torch.empty_like(self.unl_img)]
                                    for _ in range(2)]
                self._stage_k = 0
            # ping-pong staging: the set written now was last read two steps ago, and every step_host() ends with a
            # host-side wait on its losses, so that read has completed
            self._stage_k ^= 1
            bufs = self._stage_sets[self._stage_k]
            with torch.cuda.stream(self._copy_stream):
                for dst, src in zip(bufs, prefetch):
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._staged = (tuple(prefetch), tuple(bufs), ev)
        host = losses.cpu()
        return {k: float(v) for k, v in zip(self.KEYS, host)}

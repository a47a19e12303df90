"""CPU: pins the oracle (oracle/ref_arch.py, oracle/ref_step.py) against
 (1) golden vectors generated from the unmodified reference (tests/golden/*.npz, made by
     oracle/make_golden.py), and
 (2) the reference itself when /root/reference is present (build container only).
The oracle is fp32 PyTorch-functional code; agreement with the reference's nn.Module path is
required to 1e-6 relative (same ATen kernels, so normally bit-identical)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_arch as RA
from oracle import ref_step as RS

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _sd(z, prefix):
    return {k[len(prefix):]: _t(z[k]) for k in z.files if k.startswith(prefix)}


def _close(a, b, rtol=1e-6):
    scale = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() <= rtol * scale


@pytest.mark.parametrize("tag,tanh", [("softmax", False), ("tanh", True)])
def test_generator_matches_golden(tag, tanh):
    z = np.load(os.path.join(GOLD, "gen_tiny.npz"))
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(z, tag + ".w.").items()}
    x = _t(z[tag + ".x"]).clone().requires_grad_(True)
    y = RA.resnet_generator(sd, x, 9, tanh=tanh, use_dropout=False)
    assert _close(y.detach(), _t(z[tag + ".y"]))
    (y * _t(z[tag + ".probe"])).sum().backward()
    assert _close(x.grad, _t(z[tag + ".gx"]), 1e-5)
    for k, p in sd.items():
        g = _t(z[tag + ".g." + k])
        if k.endswith(".bias") and g.abs().max() < 1e-4:
            continue   # biases cancelled by InstanceNorm: reference grads are fp32 noise
        assert _close(p.grad, g, 1e-4), k


def test_discriminator_matches_golden():
    z = np.load(os.path.join(GOLD, "dis_tiny.npz"))
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(z, "w.").items()}
    x = _t(z["x"]).clone().requires_grad_(True)
    y = RA.nlayer_discriminator(sd, x, 3)
    assert _close(y.detach(), _t(z["y"]))
    (y * _t(z["probe"])).sum().backward()
    assert _close(x.grad, _t(z["gx"]), 1e-5)
    for k, p in sd.items():
        g = _t(z["g." + k])
        if k.endswith(".bias") and g.abs().max() < 1e-4:
            continue
        assert _close(p.grad, g, 1e-4), k


def test_step_head_matches_reference_train_loop_golden():
    """The 9 scalars logged by the reference's literal train() at step 0 (model.py:548-550)."""
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    nets = {nm: _sd(z, nm + ".") for nm in ("Gis", "Gsi", "Di", "Ds", "old_Gis", "old_Gsi", "old_Di")}
    losses, grads, _ = RS.full_step(nets, _t(z["l_img"]), _t(z["l_gt"]), _t(z["unl_img"]), 21, variant="head")
    for k in ("img_dis_loss", "gt_dis_loss", "cycle_img_dis_loss", "img_gen_loss", "gt_gen_loss", "img_cycle_loss",
              "gt_cycle_loss", "lab_loss_CE", "lab_loss_MSE"):
        ref = float(z["loss." + k])
        assert abs(losses[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (k, losses[k], ref)
    assert set(grads) == {"Gis", "Gsi", "Di", "Ds", "old_Di"}


def test_emulated_bf16_oracle_is_close_to_fp32_oracle():
    z = np.load(os.path.join(GOLD, "gen_tiny.npz"))
    sd = _sd(z, "softmax.w.")
    x = _t(z["softmax.x"])
    y = RA.resnet_generator(sd, x, 9, tanh=False)
    ye = RA.resnet_generator(sd, x, 9, tanh=False, emulate_bf16=True, live_norm_bias=False)
    rel = (y - ye).norm() / y.norm()
    assert 1e-4 < rel < 0.1


def test_one_hot_and_argmax_tie_rule():
    lab = torch.tensor([[[[0, 2], [1, 2]]]])
    oh = RA.make_one_hot(lab, 3)
    assert oh.shape == (1, 3, 2, 2) and oh.sum().item() == 4 and oh[0, 2, 0, 1] == 1
    p = torch.zeros(1, 3, 1, 1)   # all equal -> first max wins (model.py:435)
    assert RA.argmax_one_hot(p, 3)[0, :, 0, 0].tolist() == [1.0, 0.0, 0.0]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_oracle_equals_reference_modules():
    sys.path.insert(0, REF)
    try:
        import arch as ref_arch
        torch.manual_seed(3)
        g = ref_arch.define_Gen(21, 3, 8, "resnet_9blocks", norm="instance", use_dropout=True, gpu_ids=[]).eval()
        x = torch.rand(1, 21, 32, 32)
        assert _close(RA.resnet_generator(g.state_dict(), x, 9, tanh=True, use_dropout=True), g(x).detach())
        g6 = ref_arch.define_Gen(3, 4, 8, "resnet_6blocks_softmax", norm="instance", use_dropout=False, gpu_ids=[]).eval()
        x = torch.rand(1, 3, 24, 40)
        assert _close(RA.resnet_generator(g6.state_dict(), x, 6, tanh=False), g6(x).detach())
        d = ref_arch.define_Dis(21, 8, "n_layers", n_layers_D=3, norm="instance", gpu_ids=[])
        x = torch.rand(2, 21, 64, 64)
        assert _close(RA.nlayer_discriminator(d.state_dict(), x, 3), d(x).detach())
    finally:
        sys.path.remove(REF)
        for m in [m for m in sys.modules if m == "arch" or m.startswith("arch.")]:
            del sys.modules[m]

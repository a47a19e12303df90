"""-m gpu: the drop-in modules (arch.define_Gen / define_Dis on cuda:0, fused sm_100a path) against
the CPU oracle and the golden vectors generated from the unmodified reference.

Tolerances (north_star: 1e-3 relative to the fp32 reference, exact argmax):
  * precision 'bf16x3' (parity mode): forward max|a-b| <= 1e-3 * max|b| and argmax identical off the
    near-tie set (top-2 margin < 1e-3 * max|logit|); gradients of the tiny golden nets rel-L2 <= 1e-3.
    For full-width nets the gradient error is set by ReLU / LeakyReLU kinks, not by arithmetic: a forward deviation
    d flips the sign of ~d * (density of pre-activations at 0) of the units, each flipped unit is an O(1) error of its
    gradient, so rel-L2 ~ sqrt(fraction flipped) (3e-5 forward -> ~5e-3 gradient).  test_parity_mode_gradients_sit_
    at_the_kink_floor DEMONSTRATES this: the fp64 oracle itself, evaluated at an input perturbed so that its output
    moves as much as the kernels' output deviates, shows the same gradient change; the kernels must stay within 3x
    of that floor (measured 0.6-1.5x).  The fixed bound 3e-2 is kept as a backstop.
  * precision 'bf16' (fast mode, the benchmarked one): on the tiny golden nets the forward agrees with the
    bf16-emulated oracle to <= 1e-5 (proves the forward rounding points are exactly the documented ones) and every
    gradient agrees with autograd through the emulated oracle (gradients rounded to bf16 where the kernels store them in
    bf16) to rel-L2 <= 3e-2 — measured 0.6-1.5e-2, which is the noise of ~70 independent bf16 roundings of
    gradient tensors along the chain (each 2^-9 / sqrt(3) = 1.1e-3 rel-L2; the emulation rounds at the same tensors
    but not always before / after the same fp32 additions); a wrong tap table, stride, halo or mask gives >= 0.5.
    For full-width nets (K = 2304 contractions) two bf16 implementations diverge chaotically at the 1e-2 level
    (different fp32 summation order -> different bf16 roundings of raw outputs), so outputs AND gradients are checked
    against the envelope: deviation from the fp32 oracle <= 1.5x (outputs: 2x) the deviation of the emulated oracle
    itself (measured ratio 0.97-1.02): the kernels are as accurate as a plain bf16 PyTorch implementation of the same
    rounding points.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_arch as RA

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _sd(z, prefix):
    return {k[len(prefix):]: _t(z[k]) for k in z.files if k.startswith(prefix)}


def _max_rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / max(b.abs().max().item(), 1e-30))


def _rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / max(b.norm().item(), 1e-30))


def _norm_cancelled_bias(name):
    """Conv biases followed by InstanceNorm: mathematically zero gradient (the reference produces fp32
    noise there); the fused path returns exact zeros."""
    if not name.endswith(".bias"):
        return False
    parts = name.split(".")
    if parts[0] == "res_model":
        return len(parts) != 3          # `res_model.<idx>.bias` is the head conv (no norm after it)
    return name not in ("dis_model.0.bias", "dis_model.5.bias")


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import sscg_b200  # noqa: F401


@pytest.mark.parametrize("tag,name", [("softmax", "resnet_9blocks_softmax"), ("tanh", "resnet_9blocks")])
@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_generator_golden(tag, name, precision):
    _setup()
    from sscg_b200.arch import define_Gen
    z = np.load(os.path.join(GOLD, "gen_tiny.npz"))
    net = define_Gen(3, 5, 4, name, norm="instance", use_dropout=False, gpu_ids=[0])
    net.load_state_dict(_sd(z, tag + ".w."))
    net.precision = precision
    x = _t(z[tag + ".x"]).cuda().requires_grad_(True)
    y = net(x)
    (y * _t(z[tag + ".probe"]).cuda()).sum().backward()
    if precision == "bf16x3":
        assert _max_rel(y, _t(z[tag + ".y"])) <= 1e-3
        assert _rel_l2(x.grad, _t(z[tag + ".gx"])) <= 1e-3
        for k, p in net.named_parameters():
            if _norm_cancelled_bias(k):
                assert float(p.grad.abs().max()) == 0.0
                continue
            assert _rel_l2(p.grad, _t(z[tag + ".g." + k])) <= 1e-3, k
        if tag == "softmax":
            assert torch.equal(y.argmax(1).cpu(), _t(z[tag + ".y"]).argmax(1))
    else:
        sd = {k: v.clone().requires_grad_(True) for k, v in _sd(z, tag + ".w.").items()}
        xe = _t(z[tag + ".x"]).clone().requires_grad_(True)
        ye = RA.resnet_generator(sd, xe, 9, tanh=(tag == "tanh"), emulate_bf16=True, live_norm_bias=False)
        assert _max_rel(y, ye) <= 1e-5
        (ye * _t(z[tag + ".probe"])).sum().backward()          # autograd through the emulation: bf16-rounded gradients
        assert _rel_l2(x.grad, xe.grad) <= 3e-2, _rel_l2(x.grad, xe.grad)
        for k, p in net.named_parameters():
            if _norm_cancelled_bias(k):
                assert float(p.grad.abs().max()) == 0.0
                continue
            assert _rel_l2(p.grad, sd[k].grad) <= 3e-2, (k, _rel_l2(p.grad, sd[k].grad))


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_discriminator_golden(precision):
    _setup()
    from sscg_b200.arch import define_Dis
    z = np.load(os.path.join(GOLD, "dis_tiny.npz"))
    net = define_Dis(3, 4, "n_layers", n_layers_D=3, norm="instance", gpu_ids=[0])
    net.load_state_dict(_sd(z, "w."))
    net.precision = precision
    x = _t(z["x"]).cuda().requires_grad_(True)
    y = net(x)
    (y * _t(z["probe"]).cuda()).sum().backward()
    if precision == "bf16x3":
        assert _max_rel(y, _t(z["y"])) <= 1e-3
        assert _rel_l2(x.grad, _t(z["gx"])) <= 1e-3
        for k, p in net.named_parameters():
            if _norm_cancelled_bias(k):
                continue
            assert _rel_l2(p.grad, _t(z["g." + k])) <= 1e-3, k
    else:
        sd = {k: v.clone().requires_grad_(True) for k, v in _sd(z, "w.").items()}
        xe = _t(z["x"]).clone().requires_grad_(True)
        ye = RA.nlayer_discriminator(sd, xe, 3, emulate_bf16=True, live_norm_bias=False)
        assert _max_rel(y, ye) <= 1e-5
        (ye * _t(z["probe"])).sum().backward()
        assert _rel_l2(x.grad, xe.grad) <= 3e-2, _rel_l2(x.grad, xe.grad)
        for k, p in net.named_parameters():
            if _norm_cancelled_bias(k):
                continue
            assert _rel_l2(p.grad, sd[k].grad) <= 3e-2, (k, _rel_l2(p.grad, sd[k].grad))


def _full_width(kind, cfg, N, H, W, precision, want_emulated_grads=False):
    from sscg_b200.arch import define_Dis, define_Gen
    torch.manual_seed(0)
    if kind == "gen":
        cin, cout, name = cfg
        net = define_Gen(cin, cout, 64, name, norm="instance", use_dropout=False, gpu_ids=[0])
    else:
        cin, = cfg
        net = define_Dis(cin, 64, "n_layers", norm="instance", gpu_ids=[0])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.05)
    net.precision = precision
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x = torch.rand(N, cin, H, W) * 2 - 1
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    probe = torch.randn(y.shape)
    (y * probe.cuda()).sum().backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    sde = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xe = x.clone().requires_grad_(True)
    if kind == "gen":
        tanh = not name.endswith("softmax")
        yr = RA.resnet_generator(sdr, xr, 9, tanh=tanh)
        ye = RA.resnet_generator(sde, xe, 9, tanh=tanh, emulate_bf16=True, live_norm_bias=False)
    else:
        yr = RA.nlayer_discriminator(sdr, xr, 3)
        ye = RA.nlayer_discriminator(sde, xe, 3, emulate_bf16=True, live_norm_bias=False)
    (yr * probe).sum().backward()
    if want_emulated_grads:
        (ye * probe).sum().backward()
    _full_width.extra = dict(sd=sd, x=x, probe=probe, xe=xe, sde=sde)
    return net, y, yr, ye.detach(), xg.grad, xr.grad, sdr


FULL = [("gen", (3, 21, "resnet_9blocks_softmax"), 2, 64, 64),
        ("gen", (21, 3, "resnet_9blocks"), 2, 64, 64),
        ("gen", (1, 4, "resnet_9blocks_softmax"), 1, 32, 48),      # ACDC-like: 1-channel input, 4 classes
        ("dis", (3,), 2, 64, 64),
        ("dis", (21,), 2, 128, 128)]


@pytest.mark.parametrize("kind,cfg,N,H,W", FULL)
def test_full_width_parity_mode(kind, cfg, N, H, W):
    """bf16x3: outputs within 1e-3 of the fp32 oracle, exact argmax off the near-tie set."""
    _setup()
    net, y, yr, ye, gx, gxr, sdr = _full_width(kind, cfg, N, H, W, "bf16x3")
    assert _max_rel(y, yr) <= 1e-3
    if kind == "gen" and cfg[2].endswith("softmax"):
        yc, yrc = y.detach().cpu(), yr.detach()
        top2 = yrc.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-3 * yrc.abs().max()
        assert int(((yc.argmax(1) != yrc.argmax(1)) & safe).sum()) == 0
    assert _rel_l2(gx, gxr) <= 3e-2
    for k, p in net.named_parameters():
        if _norm_cancelled_bias(k):
            continue
        assert _rel_l2(p.grad, sdr[k].grad) <= 3e-2, k


@pytest.mark.parametrize("kind,cfg,N,H,W", FULL)
def test_full_width_fast_mode(kind, cfg, N, H, W):
    """bf16: outputs and gradients deviate from the fp32 oracle no more than a plain bf16 implementation does
    (the bf16-emulated oracle and autograd through it): outputs <= 2x, gradients <= 1.5x of its deviation."""
    _setup()
    net, y, yr, ye, gx, gxr, sdr = _full_width(kind, cfg, N, H, W, "bf16", want_emulated_grads=True)
    ex = _full_width.extra
    envelope = _rel_l2(ye, yr)
    assert _rel_l2(y, yr) <= 2.0 * envelope + 1e-4, (_rel_l2(y, yr), envelope)
    env_gx = _rel_l2(ex["xe"].grad, gxr)
    assert _rel_l2(gx, gxr) <= 1.5 * env_gx + 1e-3, (_rel_l2(gx, gxr), env_gx)
    for k, p in net.named_parameters():
        if _norm_cancelled_bias(k):
            continue
        env = _rel_l2(ex["sde"][k].grad, sdr[k].grad)
        # + 5e-3: the bias of the last conv sums the OUTPUT gradient, which the kernels hold in bf16 (one rounding,
        # measured 1.3-2.8e-3) while the emulation keeps that one tensor in fp32 (its envelope there is exactly 0)
        assert _rel_l2(p.grad, sdr[k].grad) <= 1.5 * env + 5e-3, (k, _rel_l2(p.grad, sdr[k].grad), env)


@pytest.mark.parametrize("kind,cfg,N,H,W", [FULL[0], FULL[4]])
def test_parity_mode_gradients_sit_at_the_kink_floor(kind, cfg, N, H, W):
    """The 3e-2 gradient tolerance of the parity mode is a property of ReLU networks, not of the kernels: evaluate the
    fp64 oracle at an input perturbed so that ITS output moves by as much as the kernels' output deviates from it
    (~3e-5); its gradients then change by the 'kink floor' (units whose pre-activation changed sign).  The kernels'
    gradient errors must stay within 3x of that floor."""
    _setup()
    net, y, yr, ye, gx, gxr, sdr = _full_width(kind, cfg, N, H, W, "bf16x3")
    ex = _full_width.extra
    tanh = kind == "gen" and not cfg[2].endswith("softmax")

    def oracle64(xin):
        sd64 = {k: v.double().clone().requires_grad_(True) for k, v in ex["sd"].items()}
        x64 = xin.double().clone().requires_grad_(True)
        y64 = (RA.resnet_generator(sd64, x64, 9, tanh=tanh) if kind == "gen" else RA.nlayer_discriminator(sd64, x64, 3))
        (y64 * ex["probe"].double()).sum().backward()
        return y64.detach(), x64.grad, {k: v.grad for k, v in sd64.items()}

    y0, gx0, gw0 = oracle64(ex["x"])
    d_kernel = _rel_l2(y, y0)
    assert d_kernel <= 2e-4
    noise = torch.randn(ex["x"].shape, generator=torch.Generator().manual_seed(11), dtype=torch.float64)
    y1, _, _ = oracle64(ex["x"].double() + 1e-6 * noise)
    amp = 1e-6 * d_kernel / max(_rel_l2(y1, y0), 1e-30)          # forward is linear in a perturbation this small
    y2, gx2, gw2 = oracle64(ex["x"].double() + amp * noise)
    assert 0.5 * d_kernel <= _rel_l2(y2, y0) <= 2.0 * d_kernel
    floor_gx = _rel_l2(gx2, gx0)
    worst_floor = max(_rel_l2(gw2[k], gw0[k]) for k in gw0 if not _norm_cancelled_bias(k))
    worst_kernel = max(_rel_l2(p.grad, gw0[k]) for k, p in net.named_parameters() if not _norm_cancelled_bias(k))
    assert _rel_l2(gx, gx0) <= 3.0 * floor_gx + 1e-4, (_rel_l2(gx, gx0), floor_gx)
    assert worst_kernel <= 3.0 * worst_floor + 1e-4, (worst_kernel, worst_floor)


def test_state_dict_roundtrip_and_dropout_determinism():
    _setup()
    from sscg_b200.arch import define_Gen, set_grad
    torch.manual_seed(1)
    g = define_Gen(3, 21, 64, "resnet_9blocks_softmax", norm="instance", use_dropout=True, gpu_ids=[0])
    assert "res_model.4.res_block.4.weight" in g.state_dict()      # dropout shifts the 2nd conv's index
    x = torch.rand(2, 3, 32, 32, device="cuda")
    for prec in ("bf16", "bf16x3"):        # no floating-point atomics anywhere: every run reproduces every bit
        g.precision = prec
        g.train()
        torch.manual_seed(5)
        a = g(x)
        torch.manual_seed(5)
        b = g(x)
        torch.manual_seed(6)
        c = g(x)
        assert torch.equal(a, b)                               # same seed -> same mask, same bits
        assert _max_rel(c, a) > 1e-3                           # different seed -> different mask
        g.eval()
        e1, e2 = g(x), g(x)
        assert torch.equal(e1, e2)
    g.precision = "bf16x3"
    set_grad([g], False)
    y = g(x.requires_grad_(True))
    y.sum().backward()
    assert all(p.grad is None for p in g.parameters()) and x.grad is not None


@pytest.mark.parametrize("N,H,W", [(3, 72, 88), (1, 200, 200), (2, 36, 132)])
def test_generator_odd_shapes_and_inference_path(N, H, W):
    """Sizes that are not multiples of the 128-pixel tile (reference examples are 200x200; any multiple of
    4 is legal for the two stride-2 stages), batch 1 and 3, eval + no_grad (testing.py:61 style call)."""
    _setup()
    from sscg_b200.arch import define_Gen
    torch.manual_seed(2)
    g = define_Gen(3, 21, 64, "resnet_9blocks_softmax", norm="instance", use_dropout=True, gpu_ids=[0])
    g.precision = "bf16x3"
    g.eval()
    x = torch.rand(N, 3, H, W) * 2 - 1
    with torch.no_grad():
        y = g(x.cuda())
    sd = {k: v.detach().cpu() for k, v in g.state_dict().items()}
    yr = RA.resnet_generator(sd, x, 9, tanh=False, use_dropout=True)
    assert y.shape == (N, 21, H, W)
    assert _max_rel(y, yr) <= 1e-3

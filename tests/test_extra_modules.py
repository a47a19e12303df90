"""Non-hot generator / discriminator names route to stock torch modules with the reference's module tree
(sscg_b200/arch/extra.py): state_dict keys / shapes / trainable sets and a seeded forward equal
the reference's (fixture tests/golden/extra_modules.json made by oracle/make_golden.py extra; live comparison where
/root/reference exists).  `Interp` (the reference's bilinear `interp`, model.py:268) against torch on CPU, and its CUDA
kernels against torch under -m gpu."""
import contextlib
import io
import json
import os
import sys

import pytest
import torch
import torch.nn.functional as F

import sscg_b200  # noqa: F401
from sscg_b200.arch import define_Dis, define_Gen
from sscg_b200.interp import Interp

GOLD = os.path.join(os.path.dirname(__file__), "golden", "extra_modules.json")

CASES = {"deeplab": (lambda: define_Gen(3, 21, 64, "deeplab", norm="instance", use_dropout=True, gpu_ids=[]), (1, 3, 65, 65)),
         "unet_128": (lambda: define_Gen(3, 5, 8, "unet_128", norm="instance", use_dropout=True, gpu_ids=[]), (1, 3, 128, 128)),
         "unet_256": (lambda: define_Gen(3, 5, 8, "unet_256", norm="batch", use_dropout=False, gpu_ids=[]), (1, 3, 256, 256)),
         "fc_disc": (lambda: define_Dis(21, 16, "fc_disc", gpu_ids=[]), (1, 21, 64, 64))}


@pytest.mark.parametrize("name", sorted(CASES))
def test_stock_module_matches_reference_fixture(name):
    want = json.load(open(GOLD))[name]
    make, shape = CASES[name]
    torch.manual_seed(0)                              # same construction + init order -> same weights as the reference
    with contextlib.redirect_stdout(io.StringIO()):
        net = make()
    assert [[k, list(v.shape)] for k, v in net.state_dict().items()] == want["keys"]
    assert [k for k, p in net.named_parameters() if p.requires_grad] == want["trainable"]
    net.eval()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = net(x)
    assert list(y.shape) == want["out_shape"]
    assert abs(float(y.double().sum()) - want["out_sum"]) <= 1e-4 * max(1.0, want["out_abs_sum"])
    assert abs(float(y.double().abs().sum()) - want["out_abs_sum"]) <= 1e-4 * max(1.0, want["out_abs_sum"])


def test_enet_lednet_delegate_to_the_reference_package_or_refuse():
    have_ref = os.path.isdir("/root/reference")
    if have_ref and "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    if have_ref:
        with contextlib.redirect_stdout(io.StringIO()):
            net = define_Gen(3, 5, 8, "lednet_128", norm="instance", gpu_ids=[])
        assert type(net).__name__ == "LEDNet" and type(net).__module__ == "arch.generators"
    else:
        with pytest.raises(NotImplementedError, match="outside the B200 hot path"):
            define_Gen(3, 5, 8, "enet", norm="instance", gpu_ids=[])


def test_interp_cpu_path_and_identity():
    x = torch.rand(2, 3, 9, 11)
    assert Interp((9, 11))(x) is x
    y = Interp((33, 41))(x)
    assert torch.allclose(y, F.interpolate(x, size=(33, 41), mode="bilinear", align_corners=True))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size", [((2, 21, 41, 41), (321, 321)), ((1, 3, 33, 65), (256, 512)), ((2, 4, 64, 48), (17, 23)),
                                        ((1, 2, 1, 7), (5, 1))])
def test_interp_cuda_kernels_match_torch(shape, size):
    x = torch.randn(*shape, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = Interp(size)(xa)
    yb = F.interpolate(xb, size=size, mode="bilinear", align_corners=True)
    probe = torch.randn_like(yb)
    (ya * probe).sum().backward()
    (yb * probe).sum().backward()
    assert float((ya - yb).abs().max()) <= 1e-5 * max(1.0, float(yb.abs().max()))
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-4 * max(1.0, float(xb.grad.abs().max()))
    # the gather backward is reproducible
    xc = x.clone().requires_grad_(True)
    (Interp(size)(xc) * probe).sum().backward()
    assert torch.equal(xc.grad, xa.grad)

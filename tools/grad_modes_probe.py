"""GPU diagnostic: gradients of a full-width generator in fast (bf16) vs parity (bf16x3) mode on the same weights
and input; prints the five parameters with the largest relative L2 difference.  The fast mode takes backward kernel
paths the parity mode does not (N-expanded 7x7, flattened residual dgrad, pipelined norm passes)."""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402,F401
from sscg_b200.arch import define_Gen  # noqa: E402

for name, cin, cout in (("resnet_9blocks_softmax", 3, 21), ("resnet_9blocks", 21, 3)):
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        net = define_Gen(cin, cout, 64, name, norm="instance", use_dropout=False, gpu_ids=[0])
    x0 = (torch.rand(2, cin, 64, 64) * 2 - 1).cuda()
    probe = torch.randn(2, cout, 64, 64).cuda()
    grads = {}
    for precision in ("bf16x3", "bf16"):
        net.precision = precision
        net.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        (net(x) * probe).sum().backward()
        grads[precision] = {"x": x.grad.clone(), **{k: p.grad.clone() for k, p in net.named_parameters()}}
    rows = []
    for k, g in grads["bf16x3"].items():
        n = float(g.norm())
        if n < 1e-12:
            continue
        rows.append((float((grads["bf16"][k] - g).norm()) / n, k))
    rows.sort(reverse=True)
    print(name, ["%s %.3f" % (k, r) for r, k in rows[:5]], "x %.3f" % [r for r, k in rows if k == "x"][0])

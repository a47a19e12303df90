"""CPU: the C-ABI library loads and exports every symbol include/sscg_b200.h declares (no compute
calls without a GPU); host-side step logic (stock-torch CPU path of the drop-in modules — the
reference's own gpu_ids=[] behaviour) equals the oracle; state_dict layout equals the reference's."""
import os
import re

import numpy as np
import pytest
import torch

import sscg_b200  # noqa: F401
from oracle import ref_step as RS
from sscg_b200 import _lib
from sscg_b200.arch import define_Dis, define_Gen, set_grad
from sscg_b200.step import DevicePool, SemiSupCycleGAN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sscg_b200.h")).read()
    declared = set(re.findall(r"\b(sscg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())
    assert lib.sscg_version() >= 100


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in sscg_b200/_lib.py have the size and field offsets a C compiler gives the structs of
    include/sscg_b200.h (an argument block that drifts from the header corrupts kernel arguments silently)."""
    import ctypes as C
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    pairs = [("SscgTap", _lib.Tap), ("SscgView", _lib.View), ("SscgConvArgs", _lib.ConvArgs),
             ("SscgWgradArgs", _lib.WgradArgs), ("SscgWgrad7Args", _lib.Wgrad7Args), ("SscgApplyArgs", _lib.ApplyArgs), ("SscgBwdArgs", _lib.BwdArgs),
             ("SscgWprepArgs", _lib.WprepArgs), ("SscgWbatchEntry", _lib.WbatchEntry), ("SscgConv7Args", _lib.Conv7Args)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sscg_b200.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0; }"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = {}
    for line in out.splitlines():
        c, f, v = line.split()
        want[(c, f)] = int(v)
    for cname, ct in pairs:
        assert C.sizeof(ct) == want[(cname, "sizeof")], cname
        for fname, _ in ct._fields_:
            assert getattr(ct, fname).offset == want[(cname, fname)], (cname, fname)


def test_define_errors_match_reference_strings():
    with pytest.raises(NotImplementedError, match=r"Generator model name \[foo\] is not recognized"):
        define_Gen(3, 3, 8, "foo", norm="instance", gpu_ids=[])
    with pytest.raises(NotImplementedError, match=r"Discriminator model name \[bar\] is not recognized"):
        define_Dis(3, 8, "bar", norm="instance", gpu_ids=[])
    with pytest.raises(NotImplementedError, match=r"normalization layer \[group\] is not found"):
        define_Gen(3, 3, 8, "resnet_9blocks", norm="group", gpu_ids=[])


def test_state_dict_layout_matches_golden_reference_keys():
    z = np.load(os.path.join(GOLD, "gen_tiny.npz"))
    ref_keys = sorted(k[len("softmax.w."):] for k in z.files if k.startswith("softmax.w."))
    g = define_Gen(3, 5, 4, "resnet_9blocks_softmax", norm="instance", use_dropout=False, gpu_ids=[])
    assert sorted(g.state_dict().keys()) == ref_keys
    for k, v in g.state_dict().items():
        assert tuple(v.shape) == z["softmax.w." + k].shape
    gd = define_Gen(3, 5, 4, "resnet_9blocks", norm="instance", use_dropout=True, gpu_ids=[])
    assert "res_model.4.res_block.4.weight" in gd.state_dict() and "res_model.4.res_block.3.weight" not in gd.state_dict()
    z = np.load(os.path.join(GOLD, "dis_tiny.npz"))
    d = define_Dis(3, 4, "n_layers", norm="instance", gpu_ids=[])
    assert sorted(d.state_dict().keys()) == sorted(k[2:] for k in z.files if k.startswith("w."))


def test_set_grad_and_pool():
    d = define_Dis(3, 4, "n_layers", norm="instance", gpu_ids=[])
    set_grad([d], False)
    assert not any(p.requires_grad for p in d.parameters())
    set_grad([d], True)
    assert all(p.requires_grad for p in d.parameters())
    np.random.seed(0)
    pool = DevicePool(max_elements=3)
    items = [torch.full((1,), float(i)) for i in range(10)]
    out = [pool([it])[0] for it in items]
    assert [float(o) for o in out[:3]] == [0.0, 1.0, 2.0]          # pass-through until full (utils.py:286-289)
    assert len(pool.items) == 3
    assert any(float(o) != float(i) for o, i in zip(out[3:], items[3:]))    # swaps happen afterwards


@pytest.mark.parametrize("variant", ["classic", "head"])
def test_cpu_step_equals_oracle(variant):
    z = np.load(os.path.join(GOLD, "step_head.npz"))
    t = lambda a: torch.from_numpy(np.asarray(a))
    sd = lambda p: {k[len(p):]: t(z[k]) for k in z.files if k.startswith(p)}
    names = ["Gis", "Gsi", "Di", "Ds"] + (["old_Gis", "old_Gsi", "old_Di"] if variant == "head" else [])
    m = SemiSupCycleGAN(n_classes=21, ngf=4, ndf=4, variant=variant, use_dropout=False, device="cpu")
    m.load_state({nm: sd(nm + ".") for nm in names})
    nets = {nm: sd(nm + ".") for nm in names}
    losses, grads, _ = RS.full_step(nets, t(z["l_img"]), t(z["l_gt"]), t(z["unl_img"]), 21, variant=variant)
    out = m.train_step(t(z["l_img"]), t(z["l_gt"]), t(z["unl_img"]))
    for k, v in losses.items():
        assert abs(float(out[k]) - v) <= 1e-5 * max(1.0, abs(v)), k
    for nm in ("Gis", "Gsi", "Di", "Ds"):
        for pname, p in m.nets[nm].named_parameters():
            g = grads[nm][pname]
            assert float((p.grad - g).abs().max()) <= 1e-5 * max(1e-3, float(g.abs().max())), (nm, pname)
    if variant == "head":   # the 9 scalars of the reference's literal loop
        for k in losses:
            assert abs(float(out[k]) - float(z["loss." + k])) <= 2e-5 * max(1.0, abs(float(z["loss." + k])))

"""RunningScore (sscg_b200.metrics) against the reference's runningScore (utils.py:357-412): CPU path here
(pinned to the reference import when /root/reference is present, and to a hand-checked 3-class example
otherwise); the CUDA confusion-matrix kernel against the CPU path in the -m gpu test."""
import os
import sys

import numpy as np
import pytest
import torch

import sscg_b200  # noqa: F401
from sscg_b200.metrics import RunningScore


def _maps(n_classes, seed=0, shape=(3, 17, 23)):
    rng = np.random.RandomState(seed)
    lt = rng.randint(0, n_classes, size=shape)
    lp = rng.randint(0, n_classes, size=shape)
    lt[0, :2, :] = 255                     # "ignore" label outside the class range (VOC border)
    return lt, lp


def test_hand_checked_example():
    rs = RunningScore(3, "acdc")
    lt = np.array([[[0, 0, 1, 1, 2, 2]]])
    lp = np.array([[[0, 1, 1, 1, 2, 0]]])
    rs.update(lt, lp)
    assert rs.confusion_matrix.tolist() == [[1, 1, 0], [0, 2, 0], [1, 0, 1]]
    score, cls_iu = rs.get_scores()
    assert abs(score["Overall Acc: \t"] - 4 / 6) < 1e-12
    assert abs(score["Mean IoU : \t"] - np.mean([1 / 3, 2 / 3, 1 / 2])) < 1e-12
    assert abs(cls_iu[1] - 2 / 3) < 1e-12


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present")
@pytest.mark.parametrize("dataset,C", [("voc2012", 21), ("cityscapes", 20), ("acdc", 4)])
def test_matches_reference_running_score(dataset, C):
    sys.path.insert(0, "/root/reference")
    try:
        import utils as ref_utils
    finally:
        sys.path.remove("/root/reference")
    lt, lp = _maps(C)
    a, b = RunningScore(C, dataset), ref_utils.runningScore(C, dataset)
    for _ in range(2):
        a.update(torch.from_numpy(lt), torch.from_numpy(lp))
        b.update(lt, lp)
    assert np.array_equal(a.confusion_matrix, b.confusion_matrix)
    (sa, ia), (sb, ib) = a.get_scores(), b.get_scores()
    for k in sb:
        assert abs(sa[k] - sb[k]) < 1e-12
    assert set(ia) == set(ib) and all(abs(ia[k] - ib[k]) < 1e-12 or (np.isnan(ia[k]) and np.isnan(ib[k])) for k in ib)
    a.reset()
    assert a.confusion_matrix.sum() == 0


def test_forward_onehot_cpu_equals_one_hot_forward():
    import contextlib
    import io
    from sscg_b200.arch import define_Dis, define_Gen
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        g = define_Gen(5, 3, 4, "resnet_9blocks", norm="instance", use_dropout=False, gpu_ids=[])
        d = define_Dis(5, 4, "n_layers", norm="instance", gpu_ids=[])
    lab = torch.randint(0, 5, (2, 1, 32, 32))
    oh = torch.zeros(2, 5, 32, 32).scatter_(1, lab, 1)
    assert torch.equal(g.forward_onehot(lab), g(oh))
    assert torch.equal(d.forward_onehot(lab), d(oh))


@pytest.mark.gpu
def test_confusion_kernel_matches_cpu():
    lt, lp = _maps(21, seed=3, shape=(4, 64, 96))
    a, b = RunningScore(21, "voc2012"), RunningScore(21, "voc2012")
    for _ in range(3):
        a.update(torch.from_numpy(lt).cuda(), torch.from_numpy(lp).cuda())
        b.update(lt, lp)
    (sa, ia), (sb, ib) = a.get_scores(), b.get_scores()
    assert np.array_equal(a.confusion_matrix, b.confusion_matrix)
    assert all(abs(sa[k] - sb[k]) < 1e-12 for k in sb)


@pytest.mark.gpu
def test_forward_onehot_gpu_equals_one_hot_forward():
    import contextlib
    import io
    from sscg_b200.arch import define_Dis, define_Gen
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        g = define_Gen(21, 3, 8, "resnet_9blocks", norm="instance", use_dropout=False, gpu_ids=[0])
        d = define_Dis(21, 8, "n_layers", norm="instance", gpu_ids=[0])
    for net in (g, d):
        net.precision = "bf16x3"      # parity mode: run-to-run noise of the atomically accumulated sums stays ~1e-5
    lab = torch.randint(0, 21, (2, 1, 64, 64)).cuda()
    oh = torch.zeros(2, 21, 64, 64, device="cuda").scatter_(1, lab, 1)
    w = None
    for net in (g, d):
        ya = net.forward_onehot(lab)
        yb = net(oh)
        # the packed operand is bit-identical; the InstanceNorm sums are accumulated with atomics, so two runs of the
        # same network differ in the last bits of the statistics (in bf16 mode that flips roundings downstream: ~2e-2)
        assert float((ya - yb).abs().max()) <= 1e-3 * float(yb.abs().max())
        ya.square().mean().backward()                        # weight gradients flow with a label-map input
        w = next(net.parameters())
        assert w.grad is not None and bool(torch.isfinite(w.grad).all()) and float(w.grad.abs().max()) > 0

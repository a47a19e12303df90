"""History pool (reference utils.py:278-299 Sample_from_Pool, used at model.py:350-352,490-495): the device-resident
pools of sscg_b200.step return exactly the batches the reference's pool returns under the same numpy seed, past the
50-batch fill point.  Golden decision sequence: tests/golden/pool_decisions.npz (oracle/make_golden.py pool, generated
from the unmodified reference); a live comparison against /root/reference runs where that directory exists."""
import os
import sys

import numpy as np
import pytest
import torch

import sscg_b200  # noqa: F401
from sscg_b200.step import DevicePool, GraphPool

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pool_decisions.npz")


def _run_graph_pools(steps, seed, device):
    """Three GraphPools driven like SemiSupCycleGAN.feed_pool_decisions + step_segments (reference call order)."""
    np.random.seed(seed)
    pools = [GraphPool() for _ in range(3)]
    dec = torch.zeros(3, 3, dtype=torch.int64, device=device)
    for i, p in enumerate(pools):
        p.dec = dec[i]
    ret = np.zeros((steps, 3), dtype=np.int64)
    for k in range(steps):
        host = torch.tensor([p.host_decide() for p in pools], dtype=torch.int64)
        dec.copy_(host)
        for i, p in enumerate(pools):
            x = torch.full((2, 3, 4, 4), float(1000 * i + k), device=device)
            ret[k, i] = int(p.device_apply(x)[0, 0, 0, 0].item())
    return ret, pools


def _run_device_pools(steps, seed):
    np.random.seed(seed)
    pools = [DevicePool() for _ in range(3)]
    ret = np.zeros((steps, 3), dtype=np.int64)
    for k in range(steps):
        for i, p in enumerate(pools):
            ret[k, i] = int(p([torch.full((1,), float(1000 * i + k))])[0].item())
    return ret


def test_pools_reproduce_reference_decisions_past_the_fill_point():
    z = np.load(GOLD)
    want, seed = z["returned"], int(z["seed"])
    steps = want.shape[0]
    assert steps > 100 and (want[:50] == np.arange(50)[:, None] + 1000 * np.arange(3)).all()   # filling: pass-through
    assert (want[50:] != np.arange(50, steps)[:, None] + 1000 * np.arange(3)).any()              # later: stored batches
    got_g, pools = _run_graph_pools(steps, seed, "cpu")
    assert (got_g == want).all()
    assert (_run_device_pools(steps, seed) == want).all()
    # contents of the stores equal what the reference's pool holds: replay its decisions on the host
    np.random.seed(seed)
    items = [[None] * 50 for _ in range(3)]
    for k in range(steps):
        for i in range(3):
            if k < 50:
                items[i][k] = 1000 * i + k
            elif np.random.ranf() > 0.5:
                items[i][np.random.randint(0, 50)] = 1000 * i + k
    for i, p in enumerate(pools):
        assert [int(v) for v in p.storage[:, 0, 0, 0, 0].tolist()] == items[i]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present on this machine")
def test_golden_equals_live_reference_pool():
    sys.path.insert(0, "/root/reference")
    import utils as ref_utils
    z = np.load(GOLD)
    np.random.seed(int(z["seed"]))
    pools = [ref_utils.Sample_from_Pool() for _ in range(3)]
    for k in range(z["returned"].shape[0]):
        for p in range(3):
            assert int(pools[p]([np.array([1000 * p + k])])[0][0]) == z["returned"][k, p]


@pytest.mark.gpu
def test_graph_pool_on_device_matches_reference_decisions():
    z = np.load(GOLD)
    got, _ = _run_graph_pools(z["returned"].shape[0], int(z["seed"]), "cuda")
    assert (got == z["returned"]).all()


@pytest.mark.gpu
def test_graph_safe_train_step_draws_its_own_pool_decisions():
    """A direct train_step() in graph_safe mode (no GraphedStep / train_step_host) must advance the pools itself."""
    from sscg_b200.step import SemiSupCycleGAN
    torch.manual_seed(0)
    np.random.seed(0)
    m = SemiSupCycleGAN(n_classes=5, ngf=4, ndf=4, variant="classic", use_dropout=False, device="cuda:0",
                        precision="bf16", graph_safe=True)
    l_img = (torch.rand(2, 3, 32, 32) * 2 - 1).cuda()
    unl = (torch.rand(2, 3, 32, 32) * 2 - 1).cuda()
    l_gt = torch.randint(0, 5, (2, 1, 32, 32)).cuda()
    for k in range(3):
        m.train_step(l_img, l_gt, unl)
        assert m.pool_recon.cur_elements == k + 1 and m.pool_fake_gt.cur_elements == k + 1
    assert float(m.pool_fake_img.storage[2].abs().sum()) > 0 and float(m.pool_fake_img.storage[3].abs().sum()) == 0

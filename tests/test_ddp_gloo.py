"""CPU, world_size 2 over gloo: batch sharding reproduces the single-process gradients (InstanceNorm
has no cross-sample coupling and every loss is a batch mean).  Exercises the flat
gradient buckets + all-reduce plumbing of sscg_b200.step with the stock-torch CPU module path."""
import os
import socket
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, contextlib, io
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["SSCG_ROOT"])
sys.path.insert(0, os.path.join(os.environ["SSCG_ROOT"], "tests"))
from test_ddp_gloo import _make, _data
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.set_num_threads(2)
m = _make()
l_img, l_gt, unl = _data()
per = l_img.shape[0] // world
sl = slice(rank * per, rank * per + per)
out = m.train_step(l_img[sl], l_gt[sl], unl[sl])
if rank == 0:
    torch.save({"g": m.g_grads.flat, "d": m.d_grads.flat, "loss": {k: float(v) for k, v in out.items()}},
               os.environ["SSCG_OUT"])
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed=0):
    import contextlib
    import io
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import sscg_b200  # noqa: F401
    from sscg_b200.step import SemiSupCycleGAN
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = SemiSupCycleGAN(n_classes=5, ngf=4, ndf=4, variant="classic", use_dropout=False, device="cpu")
    return m


def _data():
    g = torch.Generator().manual_seed(3)
    l_img = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    unl = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    l_gt = torch.randint(0, 5, (4, 1, 32, 32), generator=g)
    return l_img, l_gt, unl


def test_two_rank_gloo_matches_single_process():
    m = _make()
    l_img, l_gt, unl = _data()
    ref_out = m.train_step(l_img, l_gt, unl)
    g_ref, d_ref = m.g_grads.flat.clone(), m.d_grads.flat.clone()
    port = _free_port()
    with tempfile.TemporaryDirectory() as td:
        out_path = os.path.join(td, "rank0.pt")
        wpath = os.path.join(td, "worker.py")
        with open(wpath, "w") as f:
            f.write(WORKER)
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                       SSCG_ROOT=ROOT, SSCG_OUT=out_path, OMP_NUM_THREADS="2")
            procs.append(subprocess.Popen([sys.executable, wpath], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT))
        for p in procs:
            out, _ = p.communicate(timeout=600)
            assert p.returncode == 0, out.decode()[-2000:]
        res = torch.load(out_path)
    # exact up to summation order: per-rank batch means average to the global batch mean
    assert float((res["g"] - g_ref).abs().max()) <= 1e-5 * float(g_ref.abs().max())
    assert float((res["d"] - d_ref).abs().max()) <= 1e-5 * float(d_ref.abs().max())
    # rank-local losses are means over the local half batch; their scale must match the global ones
    assert abs(res["loss"]["lab_loss_CE"] - float(ref_out["lab_loss_CE"])) < 0.2

"""Runs `--warm` un-profiled training steps, then ONE step between cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`).  Same workload as bench.py (configs[1], bs 16, 256x256)."""
import argparse
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sscg_b200  # noqa: E402,F401
from sscg_b200.step import SemiSupCycleGAN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--variant", default="classic")
a = ap.parse_args()
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = SemiSupCycleGAN(n_classes=21, variant=a.variant, use_dropout=True, device="cuda:0", precision="bf16")
batches = bench.make_batches(2, a.batch, "cuda:0", 100)
for i in range(a.warm):
    m.train_step(*batches[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.train_step(*batches[0])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")

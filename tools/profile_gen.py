"""One forward + backward of a generator (Gsi by default) between cudaProfilerStart/Stop, un-batched (bs 16):
the pass tools/stage_times.py times.  Use with `ncu --profile-from-start off`."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200  # noqa: E402,F401
from sscg_b200.arch import define_Gen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="gsi")
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
torch.manual_seed(0)
if a.net == "gsi":
    net = define_Gen(3, 21, 64, "resnet_9blocks_softmax", norm="instance", use_dropout=True, gpu_ids=[0])
    x = torch.randn(a.batch, 3, 256, 256, device="cuda:0", requires_grad=True)
else:
    net = define_Gen(21, 3, 64, "resnet_9blocks", norm="instance", use_dropout=True, gpu_ids=[0])
    x = torch.softmax(torch.randn(a.batch, 21, 256, 256, device="cuda:0"), 1).requires_grad_(True)
for _ in range(2):
    net(x).square().mean().backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net(x).square().mean().backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one pass")

"""Micro-benchmark of the fused segmentation head (softmax + CE + argmax, and its backward) at the step's shape."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200
from sscg_b200 import losses

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for N in (16, 32):
    x = torch.randn(N, 21, 256, 256, device="cuda", requires_grad=True)
    lab = torch.randint(0, 21, (N, 256, 256), device="cuda")
    lab[N // 2:] = -100
    gp = torch.randn(N, 21, 256, 256, device="cuda")
    def fwd():
        return losses.seg_head(x, lab)
    out = fwd()
    print(N, "fwd us", round(timeit(fwd), 1))
    def both():
        o = losses.seg_head(x, lab)
        loss, probs = o[0], o[1]
        torch.autograd.backward([loss, probs], [torch.ones_like(loss), gp])
        x.grad = None
    print(N, "fwd+bwd us", round(timeit(both), 1))

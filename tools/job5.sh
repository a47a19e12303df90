#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "wgrad7" > gpurun_out/pytest_w7.log 2>&1; echo "w7 rc=$?"; grep -E "passed|failed|Error|error|assert|max abs" gpurun_out/pytest_w7.log | head -20
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, kernel_cases as kc
from sscg_b200 import kernels as K
for name in ("wgrad7_head_c21", "wgrad7_head_c3", "wgrad7_head_c19_wide", "wgrad7_head_c4_many_units"):
    try:
        err, scale, tol = kc.CASES[name]()
        print(name, "err %.3e scale %.3e tol %.3e" % (err, scale, tol), "dev_err", K.device_error())
    except Exception as e:
        print(name, "EXC", repr(e)[:300], "dev_err", K.device_error())
PY
# timing at production size: head wgrad old (window) vs new
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, torch.nn.functional as F
import kernel_cases as kc
from sscg_b200 import kernels as K, _lib as L, geometry as G
DEV="cuda"
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
for N in (16, 32):
  for Cout in (21, 3):
    H=W=256
    xb = K.ActBuf(N, H, W, 64, 3, DEV); xb.hi.normal_()
    dyb = K.ActBuf(N, H, W, G.pad_out_channels(Cout), 6, DEV)
    t = torch.zeros(N, dyb.Hp, dyb.Wp, dyb.C, device=DEV); t[:, 6:-6, 6:-6, :Cout].normal_(); dyb.hi[:t.numel()].copy_(t.reshape(-1).to(torch.bfloat16))
    dw = torch.zeros(7*64*448, device=DEV)
    a7 = K.wgrad7_args(xb, dyb, dw)
    t7 = timeit(lambda: K.run_wgrad(a7))
    table = G.taps_conv_fwd_window(7, 1, 0)
    aw = K.wgrad_args(dyb.view(interior=True), None, xb.window_view(448), None, table, 448, 64, dw, 7*64)
    tw = timeit(lambda: K.run_wgrad(aw))
    print("head wgrad N=%d Cout=%d: wgrad7 %.1f us, window %.1f us, dev_err %d" % (N, Cout, t7, tw, K.device_error()))
PY

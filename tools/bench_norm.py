"""Micro-benchmark of the InstanceNorm apply / backward kernels at production shapes (bf16 mode).
Usage: SSCG_LIB=<variant .so> python tools/bench_norm.py"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscg_b200
from sscg_b200 import _lib as L, kernels as K

def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us

def run(N, H, W, C, pad, residual, skip, drop=12345, act=None, norm=True):
    dev = "cuda"
    raw = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
    st = K.stats_encode(torch.stack([raw.float().sum(dim=(1, 2)), (raw.float() ** 2).sum(dim=(1, 2))], dim=-1))
    dst = K.ActBuf(N, H, W, C, pad, dev)
    a = L.ApplyArgs()
    a.raw, a.raw_fp32, a.stats, a.eps = raw.data_ptr(), 0, st.data_ptr(), 1e-5
    a.N, a.H, a.W, a.C, a.act, a.slope, a.drop_seed = N, H, W, C, (L.ACT_RELU if act is None else act), 0.2, drop
    if not norm: a.stats = None
    resb = K.ActBuf(N, H, W, C, 1, dev)
    if residual:
        a.res = resb.view(interior=True)
    a.dst, a.dst_lo, a.pad, a.pad_mode = dst.hi.data_ptr(), None, pad, L.PAD_REFLECT
    t_apply = timeit(lambda: K.run_apply(a))
    b_apply = (N * H * W * C * 2 * (2 if residual else 1) + N * (H + 2 * pad) * (W + 2 * pad) * C * 2)
    dyp = K.ActBuf(N, H, W, C, pad, dev)
    dyp.hi.normal_()
    ba = L.BwdArgs()
    ba.raw, ba.raw_fp32, ba.stats, ba.eps = raw.data_ptr(), 0, st.data_ptr(), 1e-5
    ba.N, ba.H, ba.W, ba.C, ba.act, ba.slope, ba.drop_seed = N, H, W, C, (L.ACT_RELU if act is None else act), 0.2, drop
    ba.dyp, ba.dyp_fp32, ba.pad, ba.pad_mode = dyp.view(interior=False), 0, pad, L.PAD_REFLECT if pad else L.PAD_NONE
    sk = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
    gout = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    if skip:
        ba.skip = L.make_view(sk.data_ptr(), N, H, W, C, H * W * C, W * C, C)
        ba.g_out, ba.g_fp32 = gout.data_ptr(), 0
    dz = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    bst = K.stats_buffer(N, C, dev)
    ba.dz, ba.dz_fp32, ba.dz_lo, ba.bstats = dz.data_ptr(), 0, None, bst.data_ptr()
    t_prep = timeit(lambda: K.run_bwd_prep(ba))
    b_prep = N * H * W * C * 2 * (3 + (2 if skip else 0))
    draw = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    t_bapply = timeit(lambda: K.run_bwd_apply(ba, draw))
    b_bapply = N * H * W * C * 2 * 3
    return {"shape": [N, H, W, C, pad, residual, skip, drop, act, norm],
            "apply_us": round(t_apply, 1), "apply_GBs": round(b_apply / t_apply / 1e3),
            "prep_us": round(t_prep, 1), "prep_GBs": round(b_prep / t_prep / 1e3),
            "bapply_us": round(t_bapply, 1), "bapply_GBs": round(b_bapply / t_bapply / 1e3)}

if __name__ == "__main__":
    print(os.environ.get("SSCG_LIB", "default"))
    for cfg in [(16, 64, 64, 256, 1, False, False), (16, 64, 64, 256, 1, False, False, 0), (16, 64, 64, 256, 0, False, False, 0),
                (16, 64, 64, 256, 0, False, False, 0, 0), (16, 64, 64, 256, 0, False, False, 0, 0, False),
                (16, 64, 64, 256, 1, True, True), (16, 64, 64, 256, 1, True, True, 0, 0), (16, 256, 256, 64, 3, False, False),
                (16, 128, 128, 128, 0, False, False)][slice(*[int(x) for x in os.environ.get("BN_CFG", "0:99").split(":")])]:
        print(json.dumps(run(*cfg)), flush=True)

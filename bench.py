#!/usr/bin/env python
"""bench.py — training-step throughput of the semi-supervised CycleGAN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variant classic|head]

One "step" = one full optimisation step of model.py:370-552 (generator phase + discriminator phase,
both optimizer updates) on one synthetic batch per rank.  Workload at every N: BASELINE.json
configs[1] — synthetic VOC 3x256x256 / 21 classes, batch 16 labeled + 16 unlabeled images per GPU,
bf16 compute.  metric = labeled images per second summed over ranks (weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the
same metric through the public host API (pinned host buffers -> H2D -> step -> D2H of the 9 losses).
`--impl reference` times the reference's algorithm on the host CPU (the oracle port, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_step_images_per_sec_256x256_bf16"
UNIT = "img/s"
H = W = 256
NCLS = 21
BATCH = 16
CIMG = 3
# BASELINE.json configs (index = position in `configs`, 1-based as SURVEY.md §8d numbers them): per-GPU shapes
CONFIGS = {
    2: dict(name="configs[1]: full semisupervised_cycleGAN step, synthetic VOC 3x256x256 / 21-class, bs=16 per GPU",
            cimg=3, ncls=21, h=256, w=256, batch=16, gpus=1),
    3: dict(name="configs[2]: Cityscapes 3x256x512 / 19-class, bs=8 per GPU (8 GPUs in BASELINE.json)",
            cimg=3, ncls=19, h=256, w=512, batch=8, gpus=8),
    4: dict(name="configs[3]: ACDC 1x256x256 / 4-class, bs=32, 1 GPU", cimg=1, ncls=4, h=256, w=256, batch=32, gpus=1),
    5: dict(name="configs[4]: VOC 3x512x512 / 21-class, bs=4 per GPU (4 GPUs in BASELINE.json)",
            cimg=3, ncls=21, h=512, w=512, batch=4, gpus=4),
}


def apply_config(idx, batch=None):
    """Select the workload (module-level shape constants used by the data / FLOP helpers)."""
    global H, W, NCLS, BATCH, CIMG
    c = CONFIGS[idx]
    H, W, NCLS, CIMG = c["h"], c["w"], c["ncls"], c["cimg"]
    BATCH = batch or c["batch"]
    return c


# ------------------------------------------------------------------------------------------------
# algorithmic FLOPs (SURVEY.md §8d / BASELINE.md §3): 1 MAC = 2 FLOP, transposed conv at 9 taps per
# input pixel, backward of a conv = dgrad (+ wgrad when the net is trained)
# ------------------------------------------------------------------------------------------------
def f_gen(h, w, ci, co, ngf=64):
    hw = h * w
    return 2 * (hw * 49 * ngf * ci + (hw // 4) * 9 * ngf * 2 * ngf + (hw // 16) * 9 * 2 * ngf * 4 * ngf
                + 18 * (hw // 16) * 9 * (4 * ngf) ** 2 + (hw // 16) * 9 * 4 * ngf * 2 * ngf
                + (hw // 4) * 9 * 2 * ngf * ngf + hw * 49 * ngf * co)


def f_dis(h, w, ci, ndf=64):
    return 2 * ((h * w // 4) * 16 * ci * ndf + (h * w // 16) * 16 * ndf * 2 * ndf + (h * w // 64) * 16 * 2 * ndf * 4 * ndf
                + (h // 8 - 1) * (w // 8 - 1) * 16 * 4 * ndf * 8 * ndf + (h // 8 - 2) * (w // 8 - 2) * 16 * 8 * ndf)


def step_flops_per_sample(variant, h=None, w=None, c=None, cimg=None):
    """Per labeled sample per step.  Routes follow SURVEY.md §3.2: Gsi 3 fwd + 3 bwd, Gis 3 fwd + 2 bwd
    (first-layer dgrad skipped where the input needs no gradient is ignored here: < 1 %)."""
    h, w, c, cimg = h or H, w or W, c or NCLS, cimg or CIMG
    gis, gsi = f_gen(h, w, c, cimg), f_gen(h, w, cimg, c)
    di, ds = f_dis(h, w, cimg), f_dis(h, w, c)
    fwd = 3 * gis + 3 * gsi
    bwd = 2 * (2 * gis) + 3 * (2 * gsi)                    # dgrad + wgrad
    # G phase discriminators (frozen): Di fwd + dgrad, Ds fwd only
    fwd += di + ds
    bwd += di
    # D phase: 2 Di + 2 Ds forward, dgrad + wgrad each
    fwd += 2 * di + 2 * ds
    bwd += 2 * (2 * di) + 2 * (2 * ds)
    if variant == "head":
        fwd += 2 * gsi + 2 * gis + di                      # old_Gsi x2, old_Gis x2, old_Di(recon_img) in G phase
        bwd += di                                          # dgrad through old_Di
        fwd += 2 * di                                      # D phase: old_Di x2
        bwd += 2 * (2 * di)
        fwd -= 0
    return fwd, bwd


def res_conv_flops(n, h=None, w=None, ngf=64):
    return 2 * n * ((h or H) // 4) * ((w or W) // 4) * 9 * (4 * ngf) * (4 * ngf)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def make_batches(n_batches, batch, device, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        l_img = torch.rand(batch, CIMG, H, W, generator=g) * 2 - 1
        unl_img = torch.rand(batch, CIMG, H, W, generator=g) * 2 - 1
        coarse = torch.randint(0, NCLS, (batch, 1, H // 16, W // 16), generator=g)      # blocky masks
        l_gt = coarse.repeat_interleave(16, 2).repeat_interleave(16, 3).contiguous()
        if pin:
            out.append((l_img.pin_memory(), l_gt.pin_memory(), unl_img.pin_memory()))
        else:
            out.append((l_img.to(device), l_gt.to(device), unl_img.to(device)))
    return out


# ------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(n, variant, threads=None):
    """The reference's algorithm on the host CPU: oracle port (oracle/ref_step.py) + torch Adam
    (model.py:286-287).  Returns (step callable, threads used)."""
    from oracle import ref_step as RS
    import sscg_b200  # noqa: F401
    from sscg_b200.arch import define_Dis, define_Gen
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        nets = {"Gis": define_Gen(NCLS, CIMG, 64, "resnet_9blocks", "instance", False, []).state_dict(),
                "Gsi": define_Gen(CIMG, NCLS, 64, "resnet_9blocks_softmax", "instance", False, []).state_dict(),
                "Di": define_Dis(CIMG, 64, "n_layers", 3, "instance", []).state_dict(),
                "Ds": define_Dis(NCLS, 64, "n_layers", 3, "instance", []).state_dict()}
        if variant == "head":
            nets["old_Gis"] = define_Gen(NCLS, CIMG, 64, "resnet_9blocks", "instance", False, []).state_dict()
            nets["old_Gsi"] = define_Gen(CIMG, NCLS, 64, "resnet_9blocks_softmax", "instance", False, []).state_dict()
            nets["old_Di"] = define_Dis(CIMG, 64, "n_layers", 3, "instance", []).state_dict()
    nets = {k: {kk: vv.detach().clone() for kk, vv in sd.items()} for k, sd in nets.items()}
    g_params = [p.requires_grad_(True) for k in ("Gis", "Gsi") for p in nets[k].values()]
    d_params = [p.requires_grad_(True) for k in ("Di", "Ds") for p in nets[k].values()]
    g_opt = torch.optim.Adam(g_params, lr=2e-4, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(d_params, lr=2e-4, betas=(0.5, 0.999))
    batches = make_batches(2, n, "cpu", 7)
    state = {"i": 0}

    def step():
        l_img, l_gt, unl_img = batches[state["i"] % len(batches)]
        state["i"] += 1
        losses, grads, _ = RS.full_step(nets, l_img, l_gt, unl_img, NCLS, variant=variant, dead_forwards=True)
        for k in ("Gis", "Gsi"):
            for name, p in nets[k].items():
                p.grad = grads[k][name]
        g_opt.step()
        for k in ("Di", "Ds"):
            for name, p in nets[k].items():
                p.grad = grads[k][name]
        d_opt.step()
        return losses

    return step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_batch
    # all host threads the process may use (torchrun exports OMP_NUM_THREADS=1 for N > 1: override it)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    step, threads = cpu_reference_step_factory(n, args.variant, threads=avail)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "%d-image (of %d) %dx%d batch per step, %s variant, fp32, oracle port + torch Adam" % (n, BATCH, H, W, args.variant)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.cfg["name"], "variant": args.variant, "per_step_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def gpu_baseline(args, dev, batches, steps=4, warmup=2):
    """The on-box GPU baseline SURVEY.md §8d asks for: the SAME step (step.SemiSupCycleGAN, model.py:379-542) on the
    stock torch.nn module trees — the reference's own modules (arch/ops.py:40-74) executed by cuDNN / ATen with
    torch.optim.Adam(fused) and a device-resident history pool — at the same batch and shapes, in three precisions:
    fp32 (TF32 off), TF32 (PyTorch's default for cuDNN convolutions) and bf16 autocast + channels_last.  Eager
    launches (the reference's step cannot be graph-captured: its pool round-trips through numpy); the bf16 arm is also
    timed as CUDA-graph replays of the same stock step, the most generous form of the baseline.  Returns img/s per arm."""
    import contextlib
    import gc
    import io
    from sscg_b200.step import GraphedStep, SemiSupCycleGAN
    out = {"what": "same training step on stock torch.nn modules (cuDNN convolutions, ATen InstanceNorm / losses, "
                   "torch.optim.Adam(fused=True), device pool), batch %d, %dx%d, eager launches" % (args.batch, H, W),
           "steps_timed": steps}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    arms = [("fp32", "fp32", False, False), ("tf32", "tf32", True, False), ("bf16_autocast", "bf16_autocast", True, False),
            ("bf16_autocast_cudagraph", "bf16_autocast", True, True)]
    for name, stock, tf32, graph in arms:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            torch.manual_seed(0)
            with contextlib.redirect_stdout(io.StringIO()):
                m = SemiSupCycleGAN(n_classes=NCLS, img_channels=CIMG, variant=args.variant,
                                    use_dropout=not args.no_dropout, device=dev, stock=stock, graph_safe=graph)
            if graph:
                gs = GraphedStep(m, *batches[0], warmup=warmup + 1)
                step = lambda b: gs(*b)                              # noqa: E731
            else:
                step = lambda b: m.train_step(*b)                    # noqa: E731
            for i in range(warmup):
                step(batches[i % len(batches)])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                last = step(batches[i % len(batches)])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": args.batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
            if not graph:
                out[name]["lab_loss_CE"] = float(last["lab_loss_CE"])
            del m, step, last
            if graph:
                del gs
        except Exception as e:                                       # noqa: BLE001 — report, keep the other arms
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    vals = {k: v["value"] for k, v in out.items() if isinstance(v, dict) and "value" in v}
    if vals:
        best = max(vals, key=vals.get)
        out["best"] = {"arm": best, "value": vals[best], "unit": UNIT}
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT at debug levels VERSION and WARN, in front of the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        # ... and whatever the environment says, keep fd 1 clean while the communicator comes up (created eagerly here)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import sscg_b200  # noqa: F401
    from sscg_b200 import kernels as K
    from sscg_b200.step import GraphedStep, SemiSupCycleGAN

    import contextlib
    import io
    use_graph = not args.no_graph
    torch.manual_seed(0)          # identical initial weights on every rank
    with contextlib.redirect_stdout(io.StringIO()):
        model = SemiSupCycleGAN(n_classes=NCLS, img_channels=CIMG, variant=args.variant, use_dropout=not args.no_dropout,
                                device=dev, precision=args.precision, graph_safe=use_graph)
    if world > 1:
        for p in list(model.g_grads.params) + list(model.d_grads.params):
            dist.broadcast(p.data, 0)
    dev_batches = make_batches(3, args.batch, dev, 100 + rank)
    host_batches = make_batches(3, args.batch, dev, 100 + rank, pin=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if use_graph:
        gs = GraphedStep(model, *dev_batches[0], warmup=3)      # 3 real steps + capture
        step_dev = lambda b: gs(*b)                             # noqa: E731
        step_host = gs.step_host
    else:
        step_dev = lambda b: model.train_step(*b)               # noqa: E731
        step_host = model.train_step_host
    for i in range(args.warmup):
        step_dev(dev_batches[i % 3])
    barrier()
    # ---- timed region: device-resident inputs -----------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_dev(dev_batches[i % 3])
    e1.record()
    barrier()
    launches = (gs.launches_per_step * args.steps) if use_graph else (K.launch_count() - l0)
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # ---- end-to-end region: pinned host buffers, H2D + D2H inside ---------------------------
    # every step: H2D of its pinned batch + graph replay + D2H of the nine losses; the H2D of step i + 1 is issued on a
    # copy stream while step i runs (GraphedStep.step_host(prefetch=...)), as an input pipeline would
    pf = (lambda i: {"prefetch": host_batches[(i + 1) % 3]}) if use_graph else (lambda i: {})
    for i in range(2):
        step_host(*host_batches[i % 3], **pf(i))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = step_host(*host_batches[i % 3], **pf(i))
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    # ---- per-kernel attribution: same steps, eager launches bracketed by CUDA events ------------
    # (graph replays cannot carry per-launch events; the kernels and their arguments are identical)
    prof_steps = 2
    if use_graph:
        model.feed_pool_decisions()
    model.train_step(*dev_batches[0])
    torch.cuda.synchronize()
    K.prof_begin()
    for i in range(prof_steps):
        if use_graph:
            model.feed_pool_decisions()
        model.train_step(*dev_batches[i % 3])
    prof, prof_complete = K.prof_end()
    # ---- north-star sub-metric: one fused generator forward (Gsi, train mode) -------------------------
    # eager launches and, so that ~50 ctypes launches cannot be host-bound, replays of a CUDA graph of the same pass
    gen_ms = gen_graph_ms = None
    with torch.no_grad():
        x_img = dev_batches[0][0]
        for _ in range(3):
            model.Gsi(x_img)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(10):
            model.Gsi(x_img)
        g1.record()
        torch.cuda.synchronize()
        gen_ms = g0.elapsed_time(g1) / 10
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model.Gsi(x_img)
            torch.cuda.current_stream().wait_stream(side)
            gg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gg):
                model.Gsi(x_img)
            for _ in range(3):
                gg.replay()
            torch.cuda.synchronize()
            g0.record()
            for _ in range(20):
                gg.replay()
            g1.record()
            torch.cuda.synchronize()
            gen_graph_ms = g0.elapsed_time(g1) / 20
            del gg
        except Exception:                                            # noqa: BLE001
            gen_graph_ms = None
    dev_err = K.device_error()
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        del model
        if use_graph:
            del gs, step_dev, step_host
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        gpu_base = gpu_baseline(args, dev, dev_batches)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- derived numbers ------------------------------------------------------------------------
    peaks, peak_src = load_peaks()
    ms_step = ms_total / args.steps
    value = world * args.batch / (ms_step / 1e3)
    fwd, bwd = step_flops_per_sample(args.variant)
    tf_step = (fwd + bwd) * args.batch / 1e12
    res_ms, res_n = prof.get("res_conv_fwd", (0.0, 0))
    roof = None
    traffic, traffic_src = None, None
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_ncu_traffic.json")))
    tpath = tfiles[-1] if tfiles else ""                               # dram bytes of the same kernel from `ncu --set full` (latest round)
    if os.path.exists(tpath) and args.batch == 16 and args.config == 2:
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    if res_n:
        # 18 residual-block convs per generator pass, 6 passes per step at `batch` samples each (some passes run batched
        # two at a time, so launches differ in size): total algorithmic FLOPs of the class / total time of the class
        res_flops_step = 108 * res_conv_flops(args.batch)
        achieved = res_flops_step * prof_steps / 1e12 / (res_ms / 1e3)
        peak = peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel<256,1>: 3x3 256->256 @%dx%d residual-block conv, forward "
                                             "(108 sample-pass launches per step; passes that share a network run batched)"
                                             % (H // 4, W // 4),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src + " (sustained cuBLAS bf16: kernel timed inside a long step)",
                "avg_launch_ms": res_ms / res_n, "launches_timed": res_n,
                "algorithmic_flops_per_launch": res_flops_step * prof_steps / res_n,
                "algorithmic_flops_per_sample_pass_launch": res_conv_flops(args.batch),
                "timed_in": "eager pass of %d identical steps right after the timed region, CUDA events on the "
                            "launching stream around every launch" % prof_steps}
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3", "data": "synthetic",
            "config": {"workload": "%s; %d labeled + %d unlabeled %dx%dx%d images per GPU and step, %d classes, "
                                   "Gis/Gsi = resnet_9blocks[_softmax], Di/Ds = n_layers(3), dropout %s"
                                   % (args.cfg["name"], args.batch, args.batch, CIMG, H, W, NCLS,
                                      "off" if args.no_dropout else "on"),
                       "baseline_config_index": args.config,
                       "variant": args.variant, "batch_per_gpu": args.batch, "global_batch": world * args.batch,
                       "parallelism": "dp%d" % world, "precision": args.precision,
                       "cuda_graph": use_graph,
                       "l2": "working set per step (several GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "algorithmic_tflop_per_step_per_gpu": tf_step,
                       "step_tflops_achieved": tf_step / (ms_step / 1e3),
                       "step_frac_of_sustained_peak": tf_step / (ms_step / 1e3) / peaks["bf16_tflops_sustained"]},
            "roofline": roof,
            "generator_forward": {"what": "Gsi = resnet_9blocks_softmax forward incl. NCHW<->NHWC boundary, bs %d, %dx%d"
                                          % (args.batch, H, W),
                                  "ms": gen_ms, "ms_cuda_graph": gen_graph_ms,
                                  "tflops": f_gen(H, W, CIMG, NCLS) * args.batch / 1e12 / (gen_ms / 1e3),
                                  "frac_of_sustained_peak": f_gen(H, W, CIMG, NCLS) * args.batch / 1e12 / (gen_ms / 1e3)
                                  / peaks["bf16_tflops_sustained"],
                                  "frac_of_sustained_peak_cuda_graph":
                                      (f_gen(H, W, CIMG, NCLS) * args.batch / 1e12 / (gen_graph_ms / 1e3)
                                       / peaks["bf16_tflops_sustained"]) if gen_graph_ms else None},
            "kernel_time_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items()},
            "kernel_launches_per_step": {k: v[1] / prof_steps for k, v in prof.items()},
            "kernel_profile_complete": prof_complete,
            "e2e": {"value": world * args.batch * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 9 * 4, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks, "device_error": dev_err,
            "losses_last_step": last}
    if gpu_base is not None:
        line["gpu_baseline"] = gpu_base
        if "best" in gpu_base:
            line["gpu_baseline"]["ratio_vs_best"] = value / gpu_base["best"]["value"]
            line["gpu_baseline"]["e2e_ratio_vs_best"] = line["e2e"]["value"] / gpu_base["best"]["value"]
    if world == 1 and not args.no_cpu_baseline:
        step, threads = cpu_reference_step_factory(args.ref_batch, args.variant)
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "1 step on a %d-image (of %d) %dx%d batch, %s variant, fp32, oracle port + "
                                          "torch Adam, %.1f s" % (args.ref_batch, BATCH, H, W, args.variant, dt)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="classic", choices=["classic", "head"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config (1-based index as in SURVEY.md 8d): 2 = the headline workload")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--ref-batch", type=int, default=2,
                    help="images per CPU-reference step (bounded sample; 2 is the reference's minimum, model.py:435)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-torch/cuDNN arm on the same GPU")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    args = ap.parse_args()
    args.cfg = apply_config(args.config, args.batch)
    args.batch = BATCH
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

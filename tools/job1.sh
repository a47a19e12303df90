#!/bin/bash
# GPU job: tests, determinism probe, bench (default + BN128 variant)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/determinism_probe.py > gpurun_out/determinism.log 2>&1; echo "det rc=$?"; cat gpurun_out/determinism.log | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_a.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], json.dumps(d["kernel_time_ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY

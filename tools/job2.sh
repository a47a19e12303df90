#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_probe.py > gpurun_out/parity_probe.log 2>&1; echo "parity rc=$?"; grep -v Warn gpurun_out/parity_probe.log | tail -20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
print(json.dumps(d.get("gpu_baseline"), indent=1))
print(json.dumps(d.get("generator_forward")))
print(json.dumps(d.get("cpu_baseline")))
PY

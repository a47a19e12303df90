#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -q -x > gpurun_out/pytest_step.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_step.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
PY

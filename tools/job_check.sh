#!/bin/bash
# full GPU test suite + reproducibility probe + one bench line (no baselines): the quick check after a kernel change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 300 python tools/determinism_probe.py > gpurun_out/det.log 2>&1; echo "det rc=$?"; grep -c ': 0.0' gpurun_out/det.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
k=d.get("kernel_time_ms_per_step"); print({a: round(b,2) for a,b in k.items()}, round(sum(k.values()),2))
print(d["generator_forward"])
PY

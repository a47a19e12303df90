#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 600 python tools/stage_times.py > gpurun_out/st_tr.log 2>&1; echo "rc=$?"; grep -E "fwd |total" gpurun_out/st_tr.log | grep -v nexp | cut -c1-120
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
k=d.get("kernel_time_ms_per_step"); print({a: round(b,2) for a,b in k.items()}, round(sum(k.values()),2))
PY

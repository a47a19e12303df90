#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in old new2; do
SSCG_LIB=$PWD/variants/lib_$v.so timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"; tail -3 gpurun_out/bench_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
k=d.get("kernel_time_ms_per_step"); print({a: round(b,2) for a,b in k.items()}, round(sum(k.values()),2))
PY
done
done

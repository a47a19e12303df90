#!/bin/bash
# bench.py at the other BASELINE shapes (per-GPU batch), each with its own on-box stock-torch baseline
mkdir -p gpurun_out
for c in 3 4 5; do
timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg $c rc=$?"; tail -2 gpurun_out/bench_cfg$c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg$c.json").read().strip().splitlines()[-1])
g=d.get("gpu_baseline") or {}
print("cfg $c", d["config"].get("workload"), round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3) if d.get("roofline") else None, "best stock", g.get("best"), "ratio", g.get("ratio_vs_best"))
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/config_check.py > gpurun_out/config_check.log 2>&1; echo "check rc=$?"; grep config gpurun_out/config_check.log | cut -c1-400
for c in 3 4 5; do
  timeout 1500 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "bench cfg $c rc=$?"; tail -2 gpurun_out/bench_cfg$c.err
done
python - <<'PY'
import json
for c in (3,4,5):
    try:
        d=json.loads(open("gpurun_out/bench_cfg%d.json"%c).read().strip().splitlines()[-1])
        gb=d.get("gpu_baseline") or {}
        print(c, "img/s", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3),
              "step_frac", round(d["config"]["step_frac_of_sustained_peak"],3), "baseline best", gb.get("best"), "ratio", gb.get("ratio_vs_best"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(c, "ERR", e)
PY

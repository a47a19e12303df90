#!/bin/bash
# Round-2 evidence on ONE B200: tests, reproducibility / parity probes, per-layer timings, ncu launch list and --set full
# captures of the kernels DESIGN.md discusses, full-size config checks, final bench records.  Outputs: gpurun_out/r02_*
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke.log
timeout 600 python tools/determinism_probe.py 2>&1 | grep '"check"' > $O/r02_determinism.log; echo "det rc=$?"
timeout 900 python tools/parity_probe.py 2>&1 | grep '"check"' > $O/r02_parity_probe.log
timeout 900 python tools/parity_probe.py step 2>&1 | grep '"check"' >> $O/r02_parity_probe.log; echo "parity done"
timeout 600 python tools/stage_times.py > $O/r02_stage_times.log 2>&1; cp $O/stage_times.md $O/r02_stage_times.md
for c in 3 3b 4 5; do timeout 600 python tools/config_check.py $c 2>&1 | grep '"config"'; done > $O/r02_config_check.log; echo "config checks: $(wc -l < $O/r02_config_check.log)"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches.csv python tools/profile_step.py > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
cap() {  # name regex skip count script   (demangled names: template arguments read "(int)256")
  timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o $O/r02_full_$1 python $5 > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
G="tools/profile_gen.py --net gsi"
cap res_fwd 'conv_igemm_kernel<.int.256,' 6 2 "$G"
cap res_wgrad 'conv_wgrad_kernel<.int.256,' 4 2 "$G"
cap wgrad7 'conv_wgrad7_kernel' 0 1 "$G"
cap nexp 'conv_nexp_kernel' 0 3 "$G"
cap stem 'conv_igemm_kernel<.int.64,' 0 2 "$G"
cap stem_wgrad 'conv_wgrad_kernel<.int.(64|192),' 0 2 "tools/profile_gen.py --net gis"
cap norm 'in_(apply|bwd_prep|bwd_apply)_stream' 24 10 "$G"
cap seg 'seg_head' 0 2 tools/profile_step.py
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
# summarise on the box (gpurun_out is capped at 64 MiB): keep the summaries + metric dumps, and three of the reports
python tools/summarize_ncu_full.py $O/r02_ncu_full_summary.md $O/r02_full_*.ncu-rep > $O/summarize.log 2>&1; echo "summary rc=$?"
for r in $O/r02_full_*.ncu-rep; do n=$(basename $r .ncu-rep); python tools/ncu_metrics.py $r 2>/dev/null | grep -v "per_second\|_elapsed  " > $O/${n}_metrics.txt; done
rm -f $O/r02_full_nexp.ncu-rep $O/r02_full_norm.ncu-rep $O/r02_full_res_wgrad.ncu-rep $O/r02_full_seg.ncu-rep $O/r02_full_stem_wgrad.ncu-rep
du -sh $O
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench rc=$?"; tail -2 $O/r02_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "clocks", d["clocks"])
print("gpu_baseline", {k: (v.get("value") if isinstance(v, dict) else v) for k, v in d["gpu_baseline"].items() if k != "what"})
print("gen fwd", d["generator_forward"])
print("cpu", d.get("cpu_baseline"))
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"])
PY

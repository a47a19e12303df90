#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"; tail -2 gpurun_out/bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${n}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
done

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/pytest_nccl.log 2>&1; echo "nccl test rc=$?"; grep -E "2-rank|passed|failed|skipped|Error" gpurun_out/pytest_nccl.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY

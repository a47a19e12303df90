#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "win_ or pixrow" 2>&1 | tail -8
timeout 600 python tools/stage_times.py > gpurun_out/st_rw.log 2>&1; echo "rc=$?"; grep -E "fwd 256x256 K=64|total" gpurun_out/st_rw.log | cut -c1-120

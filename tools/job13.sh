#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "nexp" > gpurun_out/pytest_nexp.log 2>&1; echo "nexp rc=$?"; tail -3 gpurun_out/pytest_nexp.log
SSCG_DEBUG=1 timeout 600 python tools/stage_times.py > gpurun_out/st2.log 2>&1; grep -E "conv7_nexp NT" gpurun_out/st2.log | sort | uniq -c; grep -E "nexp|total" gpurun_out/st2.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stage_times.py > gpurun_out/st_wstat.log 2>&1; grep -E "nexp|wgrad7|total" gpurun_out/st_wstat.log
echo "---- without weight-stationary"
SSCG_NEXP_NO_WSTAT=1 timeout 600 python tools/stage_times.py > gpurun_out/st_nowstat.log 2>&1; grep -E "nexp|total" gpurun_out/st_nowstat.log
cp gpurun_out/stage_times.md gpurun_out/stage_times_nowstat.md

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad7 -s 3 -c 1 -o gpurun_out/prof_w7 -f python tools/w7_time.py ncu > gpurun_out/ncu_w7.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_w7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:in_bwd_prep_stream -s 6 -c 1 -o gpurun_out/prof_prep -f python tools/bench_norm.py > gpurun_out/ncu_prep.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_prep.log
ls -la gpurun_out/*.ncu-rep

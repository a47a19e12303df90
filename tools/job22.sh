#!/bin/bash
mkdir -p gpurun_out
for v in wg3 wg4; do
SSCG_LIB=$PWD/variants/lib_$v.so timeout 600 python tools/stage_times.py > gpurun_out/st_$v.log 2>&1; echo "$v rc=$?"; grep -E "wgrad |total" gpurun_out/st_$v.log | grep -v wgrad7 | cut -c1-110
SSCG_LIB=$PWD/variants/lib_$v.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k wgrad 2>&1 | tail -2
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stage_times.py > gpurun_out/st_base.log 2>&1; echo "base rc=$?"; grep -E "wgrad 64x64 K=256|fwd 64x64 K=256|total" gpurun_out/st_base.log | head -8
SSCG_WG_TWO=1 SSCG_LIB=$PWD/variants/lib_wg2.so timeout 600 python tools/stage_times.py > gpurun_out/st_wg2.log 2>&1; echo "wg2 rc=$?"; grep -E "wgrad 64x64 K=256|fwd 64x64 K=256|total" gpurun_out/st_wg2.log | head -8
SSCG_WG_TWO=1 SSCG_LIB=$PWD/variants/lib_wg2.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k wgrad 2>&1 | tail -3

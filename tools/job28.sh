#!/bin/bash
SSCG_LIB=$PWD/variants/lib_nored.so timeout 600 python tools/stage_times.py > gpurun_out/st_nored.log 2>&1; echo "rc=$?"; grep -E "fwd |total" gpurun_out/st_nored.log | grep -v nexp | cut -c1-120

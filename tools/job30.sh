#!/bin/bash
for v in 4b; do echo "== vec $v"; SSCG_LIB=$PWD/variants/lib_seg$v.so timeout 300 python tools/bench_seg.py 2>&1 | tail -5; done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -q -x 2>&1 | tail -3

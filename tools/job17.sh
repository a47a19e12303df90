#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k.log
for v in new2; do
  SSCG_LIB=$PWD/variants/lib_$v.so timeout 300 python tools/bench_norm.py > gpurun_out/bn_$v.log 2>&1; echo "== $v rc=$?"; cat gpurun_out/bn_$v.log | cut -c1-250
done
for v in new2; do
SSCG_LIB=$PWD/variants/lib_$v.so BN_CFG=0:1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:in_bwd_prep_stream --launch-skip 10 --launch-count 1 -o gpurun_out/prep_$v -f python tools/bench_norm.py > gpurun_out/ncu_prep_$v.log 2>&1; echo "ncu $v rc=$?"
SSCG_LIB=$PWD/variants/lib_$v.so BN_CFG=0:1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:in_apply_stream --launch-skip 10 --launch-count 1 -o gpurun_out/apply_$v -f python tools/bench_norm.py > gpurun_out/ncu_apply_$v.log 2>&1; echo "ncu $v rc=$?"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/parity_probe.py step > gpurun_out/parity_step.log 2>&1; echo "parity rc=$?"; grep -v Warn gpurun_out/parity_step.log | tail -8

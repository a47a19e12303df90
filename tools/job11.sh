#!/bin/bash
mkdir -p gpurun_out
for c in 3 3b 4 5; do timeout 600 python tools/config_check.py $c 2>&1 | grep '"config"' ; done | tee gpurun_out/config_check.log

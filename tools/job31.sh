#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_nexp_kernel -c 3 -f -o gpurun_out/nexp3 python tools/profile_gen.py --net gsi > gpurun_out/ncu_nexp3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep

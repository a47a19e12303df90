#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:conv_igemm_kernel<.int.(64|128),' -o gpurun_out/gsi_igemm -f python tools/profile_gen.py --net gsi > gpurun_out/ncu_gsi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_gsi.log
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:conv_igemm_kernel<.int.(64|128),' -o gpurun_out/gis_igemm -f python tools/profile_gen.py --net gis > gpurun_out/ncu_gis.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_gis.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 500 ncu --set full --clock-control none --import-source on -k regex:'dense_direct_pre_kernel|dense_direct_staged_kernel|dense_dmma_kernel' \
   --launch-skip 7 -c 3 -f -o gpurun_out/r01n_full_L30 python tools/ncu_target.py --reps 2 ) > gpurun_out/ncu_full.log 2>&1
tail -n 12 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep

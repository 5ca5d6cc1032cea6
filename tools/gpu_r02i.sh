#!/bin/bash
# one-GPU: ncu --set full of the tile launches of a scheduled QFT-30 (the real op mix)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_program' -c 5 -f -o $O/r02i_tile_qft30 \
   python bench.py --circuit qft --qubits 30 --steps 1 --warmup 1 --no-e2e --no-parity --no-cpu-baseline ) > $O/ncu_tile.log 2>&1
tail -n 5 $O/ncu_tile.log
ls -la $O/*.ncu-rep
echo done

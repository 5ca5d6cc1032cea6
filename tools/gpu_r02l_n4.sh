#!/bin/bash
# R = 4: exchanges with changing partners under rank skew, every transport on one process group, then the same
# script with the entry barrier of the packed transport switched off (negative control, temporary debug switch)
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4 > $O/gpu_n4.txt
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
    tests/mp_worker.py swapskew:13:5:auto,packed,packed-pieces,p2p,staged,packed-nobarrier,packed-pieces-nobarrier gpu ) > $O/swapskew_r4.log 2>&1
grep -E "SWAP_SKEW|MP_WORKER_OK|Error|error|real" $O/swapskew_r4.log | head -20
echo done

#!/bin/bash
# R = 8 on HEAD: the exchanges-with-changing-partners stress on every transport, then the four golden cases that failed in
# profiles/r02e_pytest_gpu_r8.log (automatic transport choice and forced packed transport), all on ONE process group
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    tests/mp_worker.py swapskew:14:5:auto,packed,packed-pieces,p2p,staged+r8_q11_c3+r8_q12+r8_q11_c3@packed+r8_q12@packed gpu ) > $O/parity_r8.log 2>&1
grep -E "SWAP_SKEW|MP_WORKER_OK|Error|rror:|real" $O/parity_r8.log | head -24
echo done

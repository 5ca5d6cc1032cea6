#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "multi_gpu and (r2 or ops or shor)" ) > gpurun_out/pytest_gpu_n2.log 2>&1
tail -n 6 gpurun_out/pytest_gpu_n2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 3 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -n 3 gpurun_out/bench_n2.err
python - <<'P'
import json
for l in open("gpurun_out/bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"], d["circuit_seconds"], d["swap_nvlink_gbs_per_gpu"], d.get("swap_wait_for_peers_ms_per_step"), d["swap_transport"]); print(json.dumps(d["kernel_breakdown"]))
P

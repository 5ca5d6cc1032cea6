#!/bin/bash
# Two-GPU validation of the swap transports: forced-transport parity, swap sweep per
# transport, bench with the packed transport enabled for low slots.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_validate_swap_n2.sh'
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== forced transports against the golden runs"
( HIQ_TEST_SWAP_TRANSPORTS=1 timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider -k swap_transports ) > gpurun_out/swap_transports_parity.log 2>&1
tail -n 3 gpurun_out/swap_transports_parity.log
echo "== swap sweep per transport (L=30)"
for mode in p2p packed; do
  ( HIQ_SWAP_MODE=$mode timeout 300 $TR --master-port 29541 tools/bench_swap.py --L 30 ) > gpurun_out/swap_n2_$mode.jsonl 2> gpurun_out/swap_n2_$mode.err
  cut -c1-160 gpurun_out/swap_n2_$mode.jsonl
done
echo "== random-34 bench, packed transport for low slots"
( HIQ_SWAP_PACKED=1 timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 1 --warmup 3 ) > gpurun_out/bench_n2_packed.json 2> gpurun_out/bench_n2_packed.err
python - <<'P'
import json
for l in open("gpurun_out/bench_n2_packed.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["swap_nvlink_gbs_per_gpu"], d["swap_transport"], [(k["kernel"], k["launches"], k["mean_ms"]) for k in d["kernel_breakdown"] if "swap" in k["kernel"]])
P

#!/bin/bash
# Round 2, two-GPU check: multi-GPU parity on HEAD (golden runs on every swap transport, operator / time-evolution / Shor
# cases, full-size properties at L = 32, direct diff against the compiled reference), swap sweep per transport, bench N=2.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu_n2.txt 2>&1; nproc >> $O/gpu_n2.txt; free -g >> $O/gpu_n2.txt
echo "== multi-GPU parity (R = 2 cases)"
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "multi_gpu or fullsize_properties or direct_diff or tile_program" ) > $O/pytest_gpu_n2.log 2>&1
tail -n 12 $O/pytest_gpu_n2.log
echo "== swap sweep per transport (L=32)"
for mode in p2p packed pull; do
  case $mode in
    p2p) envs="HIQ_SWAP_MODE=p2p" ;;
    packed) envs="HIQ_SWAP_MODE=packed" ;;
    pull) envs="HIQ_SWAP_MODE=packed HIQ_SWAP_PACKED_PULL=1" ;;
  esac
  ( env $envs timeout 300 $TR --master-port 29541 tools/bench_swap.py --L 32 --reps 2 ) > $O/swap_n2_$mode.jsonl 2> $O/swap_n2_$mode.err
  cut -c1-175 $O/swap_n2_$mode.jsonl
  tail -n 2 $O/swap_n2_$mode.err
done
echo "== bench N=2: in-place only vs packed for low slots"
for tag in inplace packed; do
  if [ $tag = packed ]; then export HIQ_SWAP_PACKED=1; else unset HIQ_SWAP_PACKED; fi
  ( time timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline ) > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
  tail -n 3 $O/bench_n2_$tag.err
  python - $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/bench_n2_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"]["seconds_per_step"], "swap GB/s", d["swap_nvlink_gbs_per_gpu"],
          "wait ms", d.get("swap_wait_for_peers_ms_per_step"), d["swap_transport"])
    print("   parity", d["parity"])
    print("   e2e_breakdown", d["e2e_breakdown"])
    for k in d["kernel_breakdown"]:
        print("     ", k)
except Exception as e:
    print(sys.argv[1], "ERR", e)
P
done
echo done

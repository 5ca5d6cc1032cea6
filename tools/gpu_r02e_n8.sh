#!/bin/bash
# Round 2, eight-GPU check: multi-GPU parity at R = 4 and R = 8 on HEAD (golden runs on every swap transport, operator /
# time-evolution / Shor cases, full-size properties at L = 32, direct diff against the compiled reference at 27 / 28 qubits),
# swap sweeps at N = 4 and N = 8, the bench line at N = 8 (35q@8) and the Shor-32 line.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu_n8.txt 2>&1; nproc >> $O/gpu_n8.txt; free -g >> $O/gpu_n8.txt
SEL="multi_gpu or fullsize_properties or direct_diff"
echo "== phase A: R = 4 parity on GPUs 0-3  ||  swap sweep N = 4 on GPUs 4-7"
( CUDA_VISIBLE_DEVICES=0,1,2,3 HIQ_TEST_R=4 timeout 1200 python -m pytest tests -m gpu -v -p no:cacheprovider -k "$SEL" ) > $O/pytest_gpu_r4.log 2>&1 &
PA=$!
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
    tools/bench_swap.py --L 32 --reps 2 ) > $O/swap_n4_auto.jsonl 2> $O/swap_n4_auto.err
grep '^{' $O/swap_n4_auto.jsonl | cut -c1-200
wait $PA
grep -E "PASSED|FAILED|ERROR|passed|failed" $O/pytest_gpu_r4.log | sed 's/.*:://' | tail -n 20
echo "== phase B: R = 8 parity"
( HIQ_TEST_R=8 timeout 1500 python -m pytest tests -m gpu -v -p no:cacheprovider -k "$SEL" ) > $O/pytest_gpu_r8.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed" $O/pytest_gpu_r8.log | sed 's/.*:://' | tail -n 20
echo "== phase C: swap sweep N = 8"
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/bench_swap.py --L 32 --reps 2 ) \
    > $O/swap_n8_auto.jsonl 2> $O/swap_n8_auto.err
grep '^{' $O/swap_n8_auto.jsonl | cut -c1-200
echo "== phase D: bench N = 8 (random-35) and Shor-32"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 3 --warmup 3 ) \
    > $O/bench_n8.json 2> $O/bench_n8.err
tail -n 4 $O/bench_n8.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n8.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"]["seconds_per_step"], "swap GB/s", d["swap_nvlink_gbs_per_gpu"],
          "wait ms", d.get("swap_wait_for_peers_ms_per_step"), d["swap_transport"])
    print("   parity", d["parity"])
    print("   cpu", d["cpu_baseline"])
    print("   e2e_breakdown", d["e2e_breakdown"])
    for k in d["kernel_breakdown"]:
        print("     ", k)
except Exception as e:
    print("ERR", e)
P
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29554 bench.py --circuit shor --gpus 8 --steps 2 --warmup 1 ) \
    > $O/bench_shor32_n8.json 2> $O/bench_shor32_n8.err
tail -c 900 $O/bench_shor32_n8.json; tail -n 3 $O/bench_shor32_n8.err
echo done

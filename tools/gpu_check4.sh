#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log
( time timeout 300 python tools/bench_sustained.py --L 33 --tag r01l --reps 25 --targets "5,9,17,25;20,25,29,32;29,30,31,32;2,3,4,5" ) > gpurun_out/sustained_r01l.log 2>&1
cut -c1-170 gpurun_out/sustained_r01l.log | tail -n 12
( time timeout 300 python bench.py --circuit random --qubits 30 --no-cpu-baseline ) > gpurun_out/bench_random30.json 2> gpurun_out/bench_random30.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_random30.json").read().strip().splitlines()[0]); print(d["value"], d["ms_per_step"], d["e2e"]["seconds_per_step"], d["clocks"]); print(json.dumps(d["kernel_breakdown"]))
P

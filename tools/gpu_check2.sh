#!/bin/bash
# block-structure A/B: GPU parity suite, prediag micro-benchmark at L=30, QFT-33 bench with and without the reduced product
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log
( time timeout 300 python tools/bench_prediag.py --L 30 --tag r01j ) > gpurun_out/prediag_r01j.log 2>&1
grep "mix\|dense_plain\|qft-like" gpurun_out/prediag_r01j.log | cut -c1-160
( time timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_n1_blocks.json 2> gpurun_out/bench_n1_blocks.err
python - <<'P'
import json
for f in ("gpurun_out/bench_n1_blocks.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[0]); print(f, d["value"], d["ms_per_step"], d["e2e"], d["clocks"]); print(json.dumps(d["kernel_breakdown"]))
    except Exception as e: print(f, "ERR", e)
P
( time HIQ_DENSE_BLOCKS=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e ) > gpurun_out/bench_n1_noblocks.json 2> gpurun_out/bench_n1_noblocks.err
python - <<'P'
import json
for f in ("gpurun_out/bench_n1_noblocks.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[0]); print(f, d["value"], d["ms_per_step"], d["clocks"]); print(json.dumps(d["kernel_breakdown"]))
    except Exception as e: print(f, "ERR", e)
P
echo done

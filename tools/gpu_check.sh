#!/bin/bash
# One-GPU round check: GPU parity suite, smoke, the bench line (+ reference arm, random-30 line) and the
# ncu launch list of the bench command.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
tail -n 15 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -n 3 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/bench_n1.json
( time timeout 400 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
( time timeout 300 python bench.py --circuit random --qubits 30 --no-cpu-baseline ) > gpurun_out/bench_random30.json 2> gpurun_out/bench_random30.err
tail -c 600 gpurun_out/bench_random30.json
( time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_qft33.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/ncu_bench.log
echo done

#!/bin/bash
# three-multiplication k=4 product: parity (default on), random-30 bench on/off, kernel parity with it off
set -u
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/pytest_gpu_3m.log 2>&1
tail -n 4 gpurun_out/pytest_gpu_3m.log
timeout 60 python bench.py --circuit random --qubits 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_random30_3m_on.json 2> gpurun_out/bench_random30_3m_on.err
HIQ_DENSE_3M=0 timeout 60 python bench.py --circuit random --qubits 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_random30_3m_off.json 2> gpurun_out/bench_random30_3m_off.err
python - <<'P'
import json
for f in ("on","off"):
    try:
        d=json.loads(open("gpurun_out/bench_random30_3m_%s.json"%f).read().strip().splitlines()[0]); print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], [(k["kernel"],k["launches"],k["mean_ms"]) for k in d["kernel_breakdown"][:3]])
    except Exception as e: print(f, "ERR", e)
P
( HIQ_DENSE_3M=0 timeout 60 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "dense" ) > gpurun_out/pytest_gpu_kernels_3m_off.log 2>&1
tail -n 2 gpurun_out/pytest_gpu_kernels_3m_off.log

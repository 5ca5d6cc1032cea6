#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 5 $O/pytest_gpu.log
( timeout 300 python bench.py --circuit qft --qubits 33 --steps 3 --warmup 2 --no-cpu-baseline --no-parity --no-e2e ) > $O/qft33_mono.json 2> $O/qft33_mono.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/qft33_mono.json") if l.startswith("{")][-1])
    print("qft33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"])
    for k in d["kernel_breakdown"]:
        print("    ", k)
except Exception as e:
    print("ERR", e)
P
tail -n 2 $O/qft33_mono.err
echo done

#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for ms in 4 3 2; do
  ( HIQ_TILE_MAX_STEPS=$ms timeout 300 python bench.py --circuit qft --qubits 33 --steps 3 --warmup 2 --no-cpu-baseline --no-parity --no-e2e ) > $O/qft33_ms$ms.json 2> $O/qft33_ms$ms.err
  python - $ms <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/qft33_ms%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print("max steps", sys.argv[1], "qft33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"])
    for k in d["kernel_breakdown"]:
        print("    ", k)
except Exception as e:
    print("ERR", e)
P
done
echo done

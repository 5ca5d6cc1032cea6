#!/bin/bash
# one-GPU: double-buffered tile kernel (parity, micro-benchmark with and without the second buffer, QFT-33 / random-33)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "tile or engine_tile or batched or fullsize_gpu" ) > $O/pytest_gpu.log 2>&1
tail -n 5 $O/pytest_gpu.log
for tag in nodb; do
  if [ $tag = nodb ]; then export HIQ_TILE_DOUBLE_BUFFER=0; else unset HIQ_TILE_DOUBLE_BUFFER; fi
  ( timeout 300 python tools/bench_tile.py --L 30 --tag $tag ) > $O/bench_tile_$tag.log 2>&1
  python - $tag <<'P'
import json, sys
for l in open("gpurun_out/bench_tile_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], "%-40s %8.3f ms  %6.3f ms/gate  %7.0f GB/s eff  T=%s" % (d["name"], d["ms"], d["ms_per_gate"], d["effective_gbs"], d.get("tile_bits")))
    elif "rror" in l:
        print(l.strip()[:200])
P
  ( timeout 300 python bench.py --circuit qft --qubits 33 --steps 3 --warmup 2 --no-cpu-baseline --no-parity --no-e2e ) > $O/qft33_$tag.json 2> $O/qft33_$tag.err
  python - $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/qft33_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "qft33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"])
    for k in d["kernel_breakdown"]:
        print("    ", k)
except Exception as e:
    print("ERR", e)
P
  tail -n 2 $O/qft33_$tag.err
done
unset HIQ_TILE_DOUBLE_BUFFER
echo done

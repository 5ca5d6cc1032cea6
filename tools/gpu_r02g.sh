#!/bin/bash
# one-GPU: block-form tile steps (parity, micro-benchmark, QFT-33 / random-33 lines)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 6 $O/pytest_gpu.log
( timeout 300 python tools/bench_tile.py --L 30 --tag blocks ) > $O/bench_tile_blocks.log 2>&1
python - <<'P'
import json
for l in open("gpurun_out/bench_tile_blocks.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-40s %8.3f ms  %6.3f ms/gate  %7.0f GB/s eff  T=%s" % (d["name"], d["ms"], d["ms_per_gate"], d["effective_gbs"], d.get("tile_bits")))
    elif "rror" in l:
        print(l.strip()[:200])
P
( timeout 400 python bench.py --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline ) > $O/bench_blocks.json 2> $O/bench_blocks.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_blocks.json") if l.startswith("{")][-1])
    q = d.get("qft33") or {}
    print("random33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"], "| qft33 ms/step", q.get("ms_per_step"), "passes", q.get("hbm_passes_per_step"))
    for k in d["kernel_breakdown"][:5]:
        print("    ", k)
    for k in (q.get("kernel_breakdown") or [])[:6]:
        print("  qft", k)
except Exception as e:
    print("ERR", e)
P
tail -n 3 $O/bench_blocks.err
echo done

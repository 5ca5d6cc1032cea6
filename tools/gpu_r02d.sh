#!/bin/bash
# Round 2, one-GPU check D: parity after the measurement / scratch-slab / time-evolution / tile changes, tile experiments.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -n 12 $O/pytest_gpu.log
echo "== tile micro-benchmark L=30 (min lo 4 / 3)"
( timeout 300 python tools/bench_tile.py --L 30 --tag lo4 ) > $O/bench_tile_lo4.log 2>&1
( HIQ_TILE_MIN_LO=3 timeout 300 python tools/bench_tile.py --L 30 --tag lo3 ) > $O/bench_tile_lo3.log 2>&1
python - <<'P'
import json
for tag in ("lo4", "lo3"):
    for l in open("gpurun_out/bench_tile_%s.log" % tag):
        if l.startswith("{"):
            d = json.loads(l)
            print(tag, "%-40s %8.3f ms  %6.3f ms/gate  %7.0f GB/s eff" % (d["name"], d["ms"], d["ms_per_gate"], d["effective_gbs"]))
        elif "rror" in l:
            print(tag, l.strip()[:200])
P
echo "== qft-33 (min lo 4 / 3)"
for lo in 4 3; do
  ( HIQ_TILE_MIN_LO=$lo timeout 300 python bench.py --circuit qft --qubits 33 --steps 3 --warmup 2 --no-cpu-baseline --no-parity --e2e-steps 2 ) > $O/qft33_lo$lo.json 2> $O/qft33_lo$lo.err
  python - $lo <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/qft33_lo%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print("lo", sys.argv[1], "qft33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"], "e2e", d["e2e"]["seconds_per_step"])
    for k in d["kernel_breakdown"]:
        print("    ", k)
except Exception as e:
    print("ERR", e)
P
done
echo "== shor-30 / shor-32 on one GPU"
( timeout 300 python bench.py --circuit shor --qubits 30 --steps 2 --warmup 1 ) > $O/bench_shor30.json 2> $O/bench_shor30.err
tail -c 500 $O/bench_shor30.json; tail -n 2 $O/bench_shor30.err
( timeout 300 python bench.py --circuit shor --qubits 32 --steps 1 --warmup 1 ) > $O/bench_shor32.json 2> $O/bench_shor32.err
tail -c 500 $O/bench_shor32.json; tail -n 2 $O/bench_shor32.err
echo done

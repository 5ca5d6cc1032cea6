#!/bin/bash
# Round 2, one-GPU check B: tile-resident gate programs (parity, micro-benchmark, bench A/B), the repaired bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 12 $O/pytest_gpu.log
echo "== tile micro-benchmark L=30"
( timeout 300 python tools/bench_tile.py --L 30 ) > $O/bench_tile.log 2>&1
cut -c1-190 $O/bench_tile.log | tail -n 20
echo "== bench A/B (random-33, steps 2)"
for tag in default maxfull2 single tileoff; do
  case $tag in
    default) envs="HIQ_TILE=1" ;;
    maxfull2) envs="HIQ_TILE_MAX_FULL=2" ;;
    single) envs="HIQ_TILE_SINGLE=1" ;;
    tileoff) envs="HIQ_TILE=0" ;;
  esac
  ( env $envs timeout 400 python bench.py --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline ) > $O/ab_random33_$tag.json 2> $O/ab_random33_$tag.err
  python - $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/ab_random33_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    q = d.get("qft33") or {}
    print(sys.argv[1], "random33 ms/step", round(d["ms_per_step"], 1), "passes", d["config"]["hbm_passes_per_step"], "| qft33 ms/step", q.get("ms_per_step"), "passes", q.get("hbm_passes_per_step"))
    for k in d["kernel_breakdown"][:5]:
        print("    ", k)
    for k in (q.get("kernel_breakdown") or [])[:6]:
        print("  qft", k)
except Exception as e:
    print(sys.argv[1], "ERR", e)
P
done
echo "== bench N=1 (default command)"
( time timeout 900 python bench.py ) > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 4 $O/bench_n1.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "phys", round(d["physical_hbm_gbs"]), "e2e", d["e2e"]["seconds_per_step"])
    print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "parity", d["parity"])
    print("cpu", d["cpu_baseline"])
    print("e2e_breakdown", d["e2e_breakdown"])
except Exception as e:
    print("ERR", e)
P
echo "== grover-20"
( timeout 300 python bench.py --circuit grover --steps 3 --warmup 3 ) > $O/bench_grover20.json 2> $O/bench_grover20.err
tail -c 700 $O/bench_grover20.json; tail -n 2 $O/bench_grover20.err
echo "== ncu --set full (L=29)"
( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_program|dense_direct|dense_dmma' \
   --launch-skip 6 -c 6 -f -o $O/r02b_full_L29 python tools/ncu_target_r02.py --reps 2 ) > $O/ncu_full.log 2>&1
tail -n 4 $O/ncu_full.log
ls -la $O/*.ncu-rep
echo done

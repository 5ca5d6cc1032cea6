#!/bin/bash
# one-GPU: tile experiments (tile size vs high slots), Shor lines after the measurement changes
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python tools/bench_tile.py --L 30 --tag t11 ) > $O/bench_tile_t11.log 2>&1
( HIQ_TILE_MIN_T=12 timeout 300 python tools/bench_tile.py --L 30 --tag t12 ) > $O/bench_tile_t12.log 2>&1
python - <<'P'
import json
for tag in ("t11", "t12"):
    for l in open("gpurun_out/bench_tile_%s.log" % tag):
        if l.startswith("{"):
            d = json.loads(l)
            print(tag, "%-40s %8.3f ms  %6.3f ms/gate  %7.0f GB/s eff  T=%s" % (d["name"], d["ms"], d["ms_per_gate"], d["effective_gbs"], d.get("tile_bits")))
        elif "rror" in l:
            print(tag, l.strip()[:200])
P
( timeout 300 python bench.py --circuit shor --qubits 30 --steps 2 --warmup 1 ) > $O/bench_shor30.json 2> $O/bench_shor30.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_shor30.json') if l.startswith('{')][-1]); print(d['value'], d['per_round_ms'], d['period_found_in'], d['candidate_divides_the_order_bound_in'])"
tail -n 2 $O/bench_shor30.err
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 4 $O/pytest_gpu.log
echo done

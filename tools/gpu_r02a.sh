#!/bin/bash
# Round 2, one-GPU check A: GPU parity suite (incl. the new full-size / direct-diff cases), the opt-in paths of round 1
# (block-loop kernel, slab pool), the new bench line + reference arm, Shor / Grover lines, ncu launch list, compute-sanitizer.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest -m gpu"
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 8 $O/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
tail -n 3 $O/smoke.log
echo "== block-loop kernel (HIQ_DENSE_BLOCKLOOP=1)"
( HIQ_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k blockloop ) > $O/exp_blockloop_parity.log 2>&1
tail -n 2 $O/exp_blockloop_parity.log
( timeout 200 python tools/bench_prediag.py --L 30 --tag staged ) > $O/exp_prediag_staged.log 2>&1
( HIQ_DENSE_BLOCKLOOP=1 timeout 200 python tools/bench_prediag.py --L 30 --tag blockloop ) > $O/exp_prediag_blockloop.log 2>&1
grep "mix" $O/exp_prediag_staged.log | cut -c1-120
grep "mix" $O/exp_prediag_blockloop.log | cut -c1-120
echo "== slab pool (HIQ_SLAB_POOL=1)"
( HIQ_SLAB_POOL=1 timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider -x ) > $O/exp_pytest_pool.log 2>&1
tail -n 2 $O/exp_pytest_pool.log
for tag in nopool pool; do
  if [ $tag = pool ]; then export HIQ_SLAB_POOL=1; else unset HIQ_SLAB_POOL; fi
  ( timeout 300 python bench.py --circuit qft --qubits 33 --steps 2 --warmup 2 --no-cpu-baseline --no-parity ) > $O/exp_bench_qft33_$tag.json 2> $O/exp_bench_qft33_$tag.err
  python - $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/exp_bench_qft33_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["seconds_per_step"], 3), d["e2e_breakdown"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
P
done
unset HIQ_SLAB_POOL
echo "== bench N=1 (random-33 + qft33 + e2e + parity + cpu baseline)"
( time timeout 900 python bench.py ) > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 4 $O/bench_n1.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "phys", round(d["physical_hbm_gbs"]), "e2e", d["e2e"]["seconds_per_step"])
    print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "parity", d["parity"])
    print("cpu", d["cpu_baseline"])
    print("qft33", {k: v for k, v in d["qft33"].items() if k not in ("roofline", "kernel_breakdown", "note")})
    print("e2e_breakdown", d["e2e_breakdown"])
    for k in d["kernel_breakdown"]:
        print("   ", k)
    for k in d["qft33"]["kernel_breakdown"]:
        print("   qft", k)
except Exception as e:
    print("ERR", e)
P
echo "== reference arm"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -c 1500 $O/bench_ref.json
echo "== shor-30 / grover-20"
( timeout 300 python bench.py --circuit shor --qubits 30 --steps 1 --warmup 1 ) > $O/bench_shor30.json 2> $O/bench_shor30.err
tail -c 900 $O/bench_shor30.json; tail -n 2 $O/bench_shor30.err
( timeout 300 python bench.py --circuit grover --steps 3 --warmup 2 --no-qft-line ) > $O/bench_grover20.json 2> $O/bench_grover20.err
tail -c 600 $O/bench_grover20.json; tail -n 2 $O/bench_grover20.err
( timeout 300 python bench.py --impl reference --circuit grover --steps 3 --warmup 1 ) > $O/bench_grover20_ref.json 2> $O/bench_grover20_ref.err
tail -c 600 $O/bench_grover20_ref.json
echo "== ncu launch list of the bench command"
( time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_random33.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-qft-line ) > $O/ncu_bench.log 2>&1
tail -n 2 $O/ncu_bench.log
echo "== compute-sanitizer"
( time timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x \
    -k "prediag or swap or tiled or diag_batch" ) > $O/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 6 $O/sanitizer_memcheck.log
( time timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x \
    -k "tiled or prediag" ) > $O/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -n 6 $O/sanitizer_racecheck.log
echo done

#!/bin/bash
# One-GPU validation of the opt-in code paths: parity first, then A/B timings.
#   gpurun --timeout 900 -- 'bash tools/gpu_validate_experiments.sh'
set -u
mkdir -p gpurun_out
echo "== block-loop kernel parity (HIQ_DENSE_BLOCKLOOP=1)"
( HIQ_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k blockloop ) > gpurun_out/exp_blockloop_parity.log 2>&1
tail -n 3 gpurun_out/exp_blockloop_parity.log
echo "== folded-diagonal micro-benchmark, staged vs block-loop"
( timeout 200 python tools/bench_prediag.py --L 30 --tag exp_staged ) > gpurun_out/exp_prediag_staged.log 2>&1
( HIQ_DENSE_BLOCKLOOP=1 timeout 200 python tools/bench_prediag.py --L 30 --tag exp_blockloop ) > gpurun_out/exp_prediag_blockloop.log 2>&1
grep "mix" gpurun_out/exp_prediag_staged.log | cut -c1-140
grep "mix" gpurun_out/exp_prediag_blockloop.log | cut -c1-140
echo "== QFT-33 bench: default, block-loop, slab pool"
for tag in default blockloop pool; do
  case $tag in
    default) envs="" ;;
    blockloop) envs="HIQ_DENSE_BLOCKLOOP=1" ;;
    pool) envs="HIQ_SLAB_POOL=1" ;;
  esac
  ( env $envs timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/exp_bench_qft33_$tag.json 2> gpurun_out/exp_bench_qft33_$tag.err
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/exp_bench_qft33_%s.json" % tag).read().strip().splitlines()[0])
    print(tag, "ms/step", round(d["ms_per_step"], 1), "e2e s", round(d["e2e"]["seconds_per_step"], 3), [(k["kernel"], k["launches"], k["mean_ms"]) for k in d["kernel_breakdown"][:3]])
except Exception as e:
    print(tag, "ERR", e)
P
done
echo "== whole GPU suite with the slab pool on"
( HIQ_SLAB_POOL=1 timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/exp_pytest_pool.log 2>&1
tail -n 3 gpurun_out/exp_pytest_pool.log

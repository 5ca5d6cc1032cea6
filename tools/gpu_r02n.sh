#!/bin/bash
# final one-GPU check of this tree: GPU parity suite, smoke, a fresh random-33 line (value, e2e, parity, roofline), Shor-32
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
tail -n 4 $O/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
( timeout 300 python bench.py --circuit shor --qubits 32 --steps 1 --warmup 0 ) > $O/bench_shor32_n1.json 2> $O/bench_shor32_n1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_shor32_n1.json') if l.startswith('{')][-1]); print('shor32', d['value'], d['per_round_ms'], d['period_found_in'], d.get('candidate_divides_the_order_bound_in'), d['last_run'])"
tail -n 2 $O/bench_shor32_n1.err
( timeout 300 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-qft-line ) > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1]); print('random33 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel'], 'parity', d['parity'] and d['parity']['ok'], d['clocks'])
for k in d['kernel_breakdown']: print('    ', k)"
tail -n 2 $O/bench_n1.err
echo done

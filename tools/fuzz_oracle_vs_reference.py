"""Differential campaign for the CHECKER: the numpy oracle (oracle/statevec.py) against the unmodified compiled reference
(oracle/_ref, R = 1, 2, 4, 8 OS processes) on random scripts — gates with controls, diagonal gates on global qubits, swaps,
probabilities, entropy, measurements (outcomes bit-exact), collapse, release of qubits; cluster sizes 3-5.
    python tools/fuzz_oracle_vs_reference.py <first seed> <last seed>"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scripts
from oracle import ref
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
t0 = time.time()
for seed in range(lo, hi):
    R = [1, 2, 4, 8][seed % 4]
    nq = 6 + seed % 4 + (R.bit_length() - 1)
    script = scripts.random_script(nq, R, 1000 + seed, ngates=60, queries=True, dealloc=(seed % 3 == 0), max_cluster=[3, 4, 5][seed % 3])
    try:
        exp = scripts.merge_rank_outputs(ref.run_script(script, R, 1, timeout=120))
        got = scripts.run_on_oracle(script, R)
        scripts.assert_outputs_match(script, got, exp)
    except Exception as e:
        bad += 1
        print("MISMATCH seed", seed, R, nq, type(e).__name__, str(e)[:300], flush=True)
print("seeds", lo, hi, "mismatches", bad, round(time.time() - t0, 1), "s")

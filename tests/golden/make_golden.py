"""Generates the golden fixtures in tests/golden from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference compiled by `make -C oracle`):
    python tests/golden/make_golden.py
Each fixture = the script (JSON) + every output of the reference run (npz): final/intermediate
slabs concatenated over ranks, slot maps, probabilities, measurement outcomes.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scripts  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [
    # name, nq, R, seed, ngates, max_local, max_cluster, queries, dealloc
    ("r1_q8", 8, 1, 100, 60, None, 4, True, True),
    ("r1_q11_c5", 11, 1, 101, 80, None, 5, True, False),
    ("r2_q9", 9, 2, 102, 60, None, 4, True, True),
    ("r4_q10", 10, 4, 103, 70, None, 4, True, False),
    ("r8_q12", 12, 8, 104, 80, None, 4, True, True),
    ("r8_q11_c3", 11, 8, 105, 60, None, 3, True, False),
    ("r4_q12_gates", 12, 4, 106, 120, None, 4, False, False),
    ("r2_q10_gates", 10, 2, 107, 120, None, 4, False, False),
]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, nq, R, seed, ng, ml, mc, queries, dealloc in CASES:
        script = scripts.random_script(nq, R, seed, ng, ml, mc, queries, dealloc)
        out = scripts.merge_rank_outputs(ref.run_script(script, R))
        arrays = {}
        meta = []
        for j, v in enumerate(out):
            if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], dict):
                arrays["vec%d" % j] = v[1]
                meta.append({"id2pos": {str(k): int(p) for k, p in v[0].items()}})
            elif isinstance(v, tuple) and v and v[0] == "error":
                meta.append({"error": v[1]})
            elif isinstance(v, list):
                meta.append({"list": [int(x) for x in v]})
            elif isinstance(v, float):
                meta.append({"float": v})
            else:
                meta.append(None)
        with open(os.path.join(here, name + ".json"), "w") as f:
            json.dump({"R": R, "script": json.loads(scripts.script_to_json(script)), "outputs": meta}, f)
        np.savez_compressed(os.path.join(here, name + ".npz"), **arrays)
        print(name, "ops", len(script), "arrays", len(arrays))


if __name__ == "__main__":
    main()

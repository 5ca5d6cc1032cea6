"""Writes tests/golden/greedy_*.json: what the UNMODIFIED reference scheduling engine (hiq/projectq/cengines/_greedyscheduler.py,
run by oracle/run_reference_greedy.py with the compiled reference schedulers) emits for the cases of
tests/test_scheduler.py::GREEDY_GOLDEN — relabelling, clusters, swaps, final controlled-Z roles, final slot maps.
Needs /root/reference and oracle/_ref.    python tests/make_golden_greedy.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import test_scheduler as T  # noqa: E402

if __name__ == "__main__":
    for kind, n, R, ml, cluster in T.GREEDY_GOLDEN:
        nq, cmds = T.greedy_case_circuit(kind, n, R)
        out = T._reference_python_engine(nq, cmds, R, ml, cluster, kind == "supremacy")
        path = os.path.join(HERE, "golden", "greedy_%s_%d_r%d_l%d_c%d.json" % (kind, n, R, ml, cluster))
        with open(path, "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(path, len(out["log"]), "log entries")

"""Golden GreedyScheduler command streams produced with the UNMODIFIED reference scheduler
(oracle/_ref/_sched_cpp) driving the same ProjectQ-free greedy loop.  Run in the build container:
    python tests/golden/make_golden_sched.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_scheduler as T  # noqa: E402
from oracle import ref  # noqa: E402

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in T.CIRCUITS:
        n, cmds, R, ml = T._circuit(name)
        log, _ = T.greedy_log(n, cmds, R, ml, ref.load_ref_sched())
        with open(os.path.join(here, "sched_%s.json" % name), "w") as f:
            json.dump(log, f)
        print(name, sum(1 for k, _ in log if k == "cluster"), "clusters", sum(1 for k, _ in log if k == "swap"), "swaps")

"""Dry-run analysis behind DESIGN.md section 10, item 0 (tile-resident multi-cluster pass): the sequence of dense fused
gates a circuit produces (slot space), their mixing bits, how many diagonal fused gates sit between two neighbours and
how many of those touch the earlier gate's mixing bits, and the size of the union of consecutive target sets.  CPU only
(the engine runs with HIQ_FLAG_DRY_RUN).   python tools/analyze_dense_chain.py qft 33 [ranks]"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import scripts  # noqa: E402
from hiqsimulator_b200 import _cppsim_mpi as M  # noqa: E402
from hiqsimulator_b200 import backends, cengines, ops  # noqa: E402
from hiqsimulator_b200 import kernels as K  # noqa: E402


def trace_of(kind, n, ranks=1, rank=0):
    g = ranks.bit_length() - 1
    cmds = bench.build_circuit(kind, n)
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=n - g, max_fused_qubits=4,
                               backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, rank, ranks, M.FLAG_DRY_RUN))
    eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
    eng.receive([ops.AllocateQureg(list(range(n)), 0)])
    eng.receive(copy.deepcopy(cmds))
    eng.flush()
    return be._simulator.trace()


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "qft"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 33
    ranks = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    prev, between = None, []
    chain, longest = 0, 0
    for d in trace_of(kind, n, ranks, ranks - 1):
        if d["kind"] == scripts.KIND["diag"]:
            between.append(sorted(int(s) for s in d["slots"]))
            continue
        if d["kind"] != scripts.KIND["dense"]:
            continue
        slots = [int(s) for s in d["slots"]]
        k = len(slots)
        ks, order = K.dense_block_shape(np.asarray(d["payload"]).reshape(1 << k, 1 << k))
        mix = sorted(slots[o] for o in order[:ks])
        line = "dense %-18s mixing %-14s" % (sorted(slots), mix)
        if prev is not None:
            union = sorted(set(prev[0]) | set(slots))
            touching = [p for p in between if set(p) & set(prev[1])]
            line += " | union with previous: %d slots | diagonals in between: %2d, touching its mixing bits: %d" % (
                len(union), len(between), len(touching))
            chain = chain + 1 if len(union) <= 6 else 0
            longest = max(longest, chain)
        print(line)
        prev, between = (slots, mix), []
    print("longest run of neighbours whose union is <= 6 slots: %d launches" % (longest + 1))


if __name__ == "__main__":
    main()

"""HBM passes a circuit will take, predicted on the CPU: the scheduled stream runs on a dry-run engine, whose launch
accounting (stats) follows the same holding / grouping logic as a device engine.
    python tools/plan_passes.py qft 33 [ranks]          env: HIQ_TILE=0|1, HIQ_TILE_MAX_FULL=n, HIQ_TILE_MAX_STEPS=n"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hiqsimulator_b200 import _cppsim_mpi as M  # noqa: E402
from hiqsimulator_b200 import backends, cengines, ops  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "qft"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 33
    ranks = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    g = ranks.bit_length() - 1
    cmds = bench.build_circuit(kind, n)
    for rank in sorted({0, ranks - 1}):
        be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=n - g, max_fused_qubits=4,
                                   backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, rank, ranks, M.FLAG_DRY_RUN))
        eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
        eng.receive([ops.AllocateQureg(list(range(n)), 0)])
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        be._simulator.synchronize()
        st = be._simulator.stats()
        plan = st["dense_passes"] + st["diag_passes"] + st["scale_passes"]
        print("%s-%d rank %d/%d: plan passes %d (dense %d, diag %d) -> launches %d (tile launches %d carrying %d dense gates), swaps %d"
              % (kind, n, rank, ranks, plan, st["dense_passes"], st["diag_passes"], st["gate_launches"], st["tile_launches"], st["tile_steps"],
                 st["total_swaps"]))


if __name__ == "__main__":
    main()

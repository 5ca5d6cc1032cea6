"""Rank body of the multi-process tests (launched with torch.distributed.run).

usage: mp_worker.py <golden name> <mode>      mode = dry (CPU, gloo) | gpu (NCCL)
Rank 0 compares the merged outputs with the golden fixture and exits non-zero on mismatch."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import scripts  # noqa: E402
from golden_util import load_golden  # noqa: E402


def main():
    name, mode = sys.argv[1], sys.argv[2]
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import world
    R, script, exp = load_golden(name)
    flags = M.FLAG_DRY_RUN if mode == "dry" else 0
    rank, size = world.init_world(flags)
    assert size == R, (size, R)
    if mode == "dry":
        e = M.SimulatorMPI(*script[0][1:])
        maps = []
        for op in script[1:]:
            if op[0] == "cheat_local":
                break
            got = getattr(e, op[0])(*op[1:])
            if op[0] == "get_qubits_ids":
                maps.append(list(got))
        gathered = world.gather_objects((e.trace(), maps))
        if rank == 0:
            state = scripts.replay_traces([g[0] for g in gathered], R)
            j = next(i for i, op in enumerate(script) if op[0] == "cheat_local")
            assert np.abs(state - exp[j][1]).max() <= 1e-12
            assert all(g[1] == gathered[0][1] for g in gathered)
    else:
        out = scripts.run_on_sim(M.SimulatorMPI, script)
        gathered = world.gather_objects(out)
        if rank == 0:
            merged = scripts.merge_rank_outputs(gathered)
            scripts.assert_outputs_match(script, merged, exp)
    world.barrier()
    if rank == 0:
        print("MP_WORKER_OK", name, mode)


if __name__ == "__main__":
    main()

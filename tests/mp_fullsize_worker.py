"""Rank body of tests/test_fullsize_multigpu.py (launched with torch.distributed.run, one process per GPU).

usage: mp_fullsize_worker.py props <qubits>          size-independent parity properties at full size (bench.parity_checks)
       mp_fullsize_worker.py diff <dir>              run <dir>/script.pkl on the CUDA engine and diff every rank's slab and
                                                     slot maps against <dir>/ref<rank>.npy / ref_ids.json, which the parent
                                                     obtained from the compiled reference (oracle/_ref, one process per rank)
Rank 0 prints FULLSIZE_OK <json> when every rank passed."""
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

TOL = 1e-12


def main():
    mode = sys.argv[1]
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, world
    rank, size = world.init_world(0)
    if mode == "props":
        import bench
        n = int(sys.argv[2])
        L = n - (size.bit_length() - 1)
        res = bench.parity_checks(n, L, lambda: backends.SimulatorMPI(gate_fusion=True, rnd_seed=12345, num_local_qubits=L, max_fused_qubits=4))
        ok = res["ok"]
        report = res
    else:
        import scripts
        d = sys.argv[2]
        with open(os.path.join(d, "script.pkl"), "rb") as f:
            script = pickle.load(f)
        with open(os.path.join(d, "ref_ids.json")) as f:
            ref_ids = json.load(f)
        out = scripts.run_on_sim(M.SimulatorMPI, script)
        errors = [(j, script[j][0], o) for j, o in enumerate(out) if isinstance(o, tuple) and len(o) == 2 and o[0] == "error"]
        assert not errors, errors[:3]
        ids = [list(o) for op, o in zip(script, out) if op[0] == "get_qubits_ids"][-1]
        id2pos, slab = out[-1]
        ref = np.load(os.path.join(d, "ref%d.npy" % rank), mmap_mode="r")
        err = 0.0
        step = 1 << 22
        for b in range(0, slab.shape[0], step):
            err = max(err, float(np.abs(slab[b:b + step] - ref[b:b + step]).max()))
        ok = err <= TOL and ids == ref_ids["ids"] and {int(k): v for k, v in ref_ids["id2pos"].items()} == dict(id2pos)
        report = {"max_abs_err": err, "maps_equal": ids == ref_ids["ids"], "amplitudes_per_rank": int(slab.shape[0])}
    gathered = world.gather_objects((ok, report))
    world.barrier()
    if rank == 0:
        assert all(g[0] for g in gathered), [g[1] for g in gathered]
        worst = gathered[0][1] if mode == "props" else {"max_abs_err": max(g[1]["max_abs_err"] for g in gathered),
                                                         "amplitudes_per_rank": gathered[0][1]["amplitudes_per_rank"]}
        print("FULLSIZE_OK", json.dumps(worst))


if __name__ == "__main__":
    main()

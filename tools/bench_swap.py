"""Swap sweep on N GPUs (mirrors the reference's opyt-mpi.cpp all_to_all loop): global<->local swaps of
q qubits for several local-slot classes, peer-mapped and staged transports, NVLink GB/s per GPU per
direction = 16 B * 2^L * (1 - 2^-q) / t.  Launch with torchrun; JSON lines on rank 0.
    HIQ_SWAP_MODE=p2p|packed|staged torchrun --nproc-per-node N tools/bench_swap.py --L 30     (default: automatic choice)
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=30)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import world
    rank, size = world.init_world(M.FLAG_TIMING)
    g = size.bit_length() - 1
    L = args.L
    n = L + g
    sim = M.SimulatorMPI(1, L, 4)
    sim.allocate_qureg(list(range(n)), 2.0 ** (-n / 2))
    sim.synchronize()
    mode = os.environ.get("HIQ_SWAP_MODE", "auto")
    cases = []
    for q in range(1, g + 1):
        cases += [(q, "top", list(range(L - q, L))), (q, "mid", list(range(12, 12 + q))), (q, "slot3+", list(range(3, 3 + q))),
                  (q, "bottom", list(range(0, q)))]
        # one low slot, the others in the middle: where the in-place exchange starts to lose (crossover for the packed one)
        cases += [(q, "slot%d+mid" % low, [low] + list(range(12, 12 + q - 1))) for low in (1, 2)]
        if q > 1:
            cases.append((q, "slot0+mid", [0] + list(range(12, 12 + q - 1))))
    for q, name, slots in cases:
        times = []
        for rep in range(args.reps + 1):
            loc = sim.get_local_qubits_ids()
            glo = sim.get_global_qubits_ids()
            pairs = []
            for j in range(q):
                pairs += [glo[j], loc[slots[j]]]
            sim.swap_qubits(pairs)
            t = [x for x in sim.collect_timings() if x[0] == 4]
            if rep >= 1:
                times.append(sum(x[3] for x in t))
        world.barrier()
        ms = float(np.median(times))
        nbytes = 16.0 * (1 << L) * (1 - 2.0 ** -q)
        st = sim.stats()
        all_ms = world.gather_objects(ms)
        if rank == 0:
            worst = max(all_ms)
            print(json.dumps({"mode": mode, "n_gpus": size, "L": L, "q": q, "slots": name, "ms": worst,
                              "nvlink_gbs_per_gpu_per_dir": nbytes / worst / 1e6, "frac_of_900": nbytes / worst / 1e6 / 900.0,
                              "swaps_p2p": st["swaps_p2p"], "swaps_staged": st["swaps_staged"], "swaps_packed": st.get("swaps_packed", 0)}), flush=True)
    p = sim.get_probability([False], [0])
    if rank == 0:
        print(json.dumps({"check_prob_q0_is_half": p}), flush=True)


if __name__ == "__main__":
    main()

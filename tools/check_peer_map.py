"""One in-place exchange at full slab size: the slab allocate_qureg maps in ONE virtual-memory piece is handed to the
peers and swapped through NVLink.  Prints the transport taken, the peer-mapping time and P(q0 = 0) (must stay 1/2).
    torchrun --nproc-per-node N tools/check_peer_map.py --L 33"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=33)
    args = ap.parse_args()
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import world
    rank, size = world.init_world(M.FLAG_TIMING)
    g = size.bit_length() - 1
    L = args.L
    n = L + g
    t0 = time.perf_counter()
    sim = M.SimulatorMPI(1, L, 4)
    sim.allocate_qureg(list(range(n)), 2.0 ** (-n / 2))
    sim.synchronize()
    t1 = time.perf_counter()
    loc, glo = sim.get_local_qubits_ids(), sim.get_global_qubits_ids()
    pairs = []
    for j in range(g):
        pairs += [glo[j], loc[L - 1 - j]]
    sim.swap_qubits(pairs)
    sim.synchronize()
    t2 = time.perf_counter()
    st = sim.stats()
    p = sim.get_probability([False], [0])
    if rank == 0:
        print(json.dumps({"n_gpus": size, "L": L, "allocate_s": t1 - t0, "first_swap_s": t2 - t1, "peer_map_s": st["peer_map_s"],
                          "slab_map_s": st["slab_grow_s"], "swaps_p2p": st["swaps_p2p"], "swaps_staged": st["swaps_staged"],
                          "swaps_packed": st.get("swaps_packed", 0), "prob_q0_is_0": p}), flush=True)


if __name__ == "__main__":
    main()

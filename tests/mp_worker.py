"""Rank body of the multi-process tests (launched with torch.distributed.run).

usage: mp_worker.py <golden name> <mode>      mode = dry (CPU, gloo) | gpu (NCCL)
       mp_worker.py ops:<qubits>:<seed> <mode>   operator-level script (scripts.operator_script) checked against the
                                                 numpy oracle instead of a golden fixture
       mp_worker.py swapskew:<qubits>:<seed> <mode>   exchanges with changing peer sets while half of the ranks are held back
Rank 0 compares the merged outputs with the expectation and exits non-zero on mismatch."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import scripts  # noqa: E402
from golden_util import load_golden  # noqa: E402


def main():
    names, mode = sys.argv[1], sys.argv[2]
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import world
    flags = M.FLAG_DRY_RUN if mode == "dry" else 0
    rank, size = world.init_world(flags)
    # several cases on one process group: name+name+...; name@transport forces the exchange transport for that case
    for name in names.split("+"):
        forced = None
        if "@" in name:
            name, forced = name.split("@")
        saved = {k: os.environ.get(k) for k in ("HIQ_SWAP_MODE", "HIQ_SWAP_PACKED_PIECE")}
        if forced:
            os.environ["HIQ_SWAP_MODE"] = forced
            if forced == "packed":
                os.environ["HIQ_SWAP_PACKED_PIECE"] = "16"
        try:
            one_case(name, mode, M, world, rank, size)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        if rank == 0:
            print("MP_WORKER_OK", name + ("@" + forced if forced else ""), mode, flush=True)


def one_case(name, mode, M, world, rank, size):
    if name.startswith("ops:") or name.startswith("tevo:"):
        return operator_case(name, mode, M, world, rank, size)
    if name.startswith("shor:"):
        return shor_case(name, mode, M, world, rank, size)
    if name.startswith("swapskew:"):
        return swap_skew_case(name, mode, M, world, rank, size)
    R, script, exp = load_golden(name)
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


def swap_skew_case(name, mode, M, world, rank, R):
    """Consecutive exchanges whose peer sets differ (scripts.swap_skew_script), with the ranks that meet a NEW partner in
    the next exchange held back on the host: the partner reaches that exchange while they are still inside the current
    one.  Whatever the transport shares between exchanges (the packed transport's process-wide staging buffers) has to
    survive this; the final slabs must equal the numpy oracle's.  name = swapskew:<qubits>:<seed>[:<transport>,...] —
    the transports (auto, packed, packed-pieces, p2p, staged) are run one after the other on the same process group
    (the engine reads HIQ_SWAP_MODE / HIQ_SWAP_PACKED_PIECE when it is constructed)."""
    import time
    parts = name.split(":")
    nq, seed = int(parts[1]), int(parts[2])
    transports = parts[3].split(",") if len(parts) > 3 else ["auto"]
    script, holds = scripts.swap_skew_script(nq, R, seed)
    hold_s = float(os.environ.get("HIQ_TEST_SKEW_SECONDS", "0.05"))

    def before_op(j, op):
        if rank in holds.get(j, ()):
            time.sleep(hold_s)
    exp = scripts.run_on_oracle(script, R) if rank == 0 else None
    if mode == "dry":
        e = M.SimulatorMPI(*script[0][1:])
        for j, op in enumerate(script[1:], 1):
            if op[0] == "cheat_local":
                break
            before_op(j, op)
            scripts._dispatch(e, op)
        gathered = world.gather_objects(e.trace())
        if rank == 0:
            state = scripts.replay_traces(gathered, R)
            assert np.abs(state - exp[-1][1]).max() <= 1e-12
    else:
        for transport in transports:
            for k in ("HIQ_SWAP_MODE", "HIQ_SWAP_PACKED_PIECE"):
                os.environ.pop(k, None)
            if transport != "auto":
                os.environ["HIQ_SWAP_MODE"] = transport.split("-")[0]
            if "pieces" in transport:
                os.environ["HIQ_SWAP_PACKED_PIECE"] = "16"  # many pieces: the two-buffer pipeline is exercised too
            keep = []
            out = scripts.run_on_sim(M.SimulatorMPI, script, keep, before_op)
            st = keep[0].stats()
            gathered = world.gather_objects((out, {k: int(st[k]) for k in ("swaps_p2p", "swaps_staged", "swaps_packed")}))
            if rank == 0:
                merged = scripts.merge_rank_outputs([g[0] for g in gathered])
                scripts.assert_outputs_match(script, merged, exp)
                print("SWAP_SKEW_OK", transport, gathered[0][1], flush=True)
            del keep, out
            world.barrier()
    world.barrier()


def shor_case(name, mode, M, world, rank, R):
    """C5 at test size through the whole host pipeline (GreedyScheduler + backend): the measured bits and the
    final state must equal the numpy oracle's driven by the same pipeline and seed."""
    from hiqsimulator_b200 import backends, cengines, circuits
    from oracle import statevec
    assert mode == "gpu"
    _, N, a, n, seed = name.split(":")
    N, a, n, seed = int(N), int(a), int(n), int(seed)
    L = n + 1 - (R.bit_length() - 1)

    def run(backend_class):
        be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=seed, num_local_qubits=L, max_fused_qubits=3, backend_class=backend_class)
        eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=3)])
        r, bits = circuits.run_shor(eng, N, a, n)
        return r, bits, [eng.measurements[q] for q in range(n + 1)], be
    r, bits, final, be = run(None)
    id2pos, full = be.cheat()
    gathered = world.gather_objects((r, bits, final, np.asarray(full).copy(), dict(id2pos)))
    if rank == 0:
        r0, bits0, final0, be0 = run(lambda s, ml, mc: statevec.SimulatorMPI(s, ml, mc, R))
        id2pos0, full0 = be0.cheat()
        for g in gathered:
            assert (g[0], g[1], g[2]) == (r0, bits0, final0), (g[:3], (r0, bits0, final0))
            assert g[4] == id2pos0 and np.abs(g[3] - full0).max() <= 1e-12
    world.barrier()


def operator_case(name, mode, M, world, rank, R):
    kind, nq, seed = name.split(":")
    script = (scripts.time_evolution_script if kind == "tevo" else scripts.operator_script)(int(nq), R, int(seed))
    if mode == "dry":
        stop = next(j for j, op in enumerate(script) if op[0] == "measure_qubits")
        script = script[:stop]
        e = M.SimulatorMPI(*script[0][1:])
        loads = []
        for op in script[1:]:
            if op[0] in ("cheat_local", "get_probability"):
                continue
            if op[0] == "set_wavefunction":
                loads.append(op[1])
            scripts._dispatch(e, op)
        gathered = world.gather_objects(e.trace())
        if rank == 0:
            exp = scripts.run_on_oracle(script, R)
            state = scripts.replay_traces(gathered, R, {"loads": loads})
            last = max(j for j, op in enumerate(script) if op[0] == "cheat_local")
            assert np.abs(state - exp[last][1]).max() <= 1e-12
    else:
        keep = []
        out = scripts.run_on_sim(M.SimulatorMPI, script, keep)
        errors = [(j, script[j][0], o) for j, o in enumerate(out) if isinstance(o, tuple) and len(o) == 2 and o[0] == "error"]
        assert not errors, errors[:3]
        # cheat(): every rank holds the concatenation of all slabs
        id2pos, full = keep[0].cheat()
        last = max(j for j, op in enumerate(script) if op[0] == "cheat_local")
        gathered = world.gather_objects((out, np.asarray(full).copy(), dict(id2pos)))
        if rank == 0:
            exp = scripts.run_on_oracle(script, R)
            merged = scripts.merge_rank_outputs([g[0] for g in gathered])
            scripts.assert_outputs_match(script, merged, exp, tol=1e-11 if kind == "tevo" else 1e-12)
            for g in gathered:
                assert np.array_equal(g[1], merged[last][1]) and g[2] == merged[last][0]
    world.barrier()


if __name__ == "__main__":
    main()

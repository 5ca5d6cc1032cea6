"""Runs the reference's UNMODIFIED Python scheduling engine (TEST INFRASTRUCTURE).

`/root/reference/hiq/projectq/cengines/_greedyscheduler.py` (class GreedyScheduler: the stage / cluster loop that calls
ClusterScheduler / SwapScheduler) and `/root/reference/hiq/projectq/ops/_gates.py` (MetaSwap,
AllocateQuregGate) are loaded from where they lie, byte for byte; what they import from ProjectQ (third party, absent here:
`projectq>=0.4.0`, requirements.txt:2) is supplied by the stand-ins of oracle/projectq_stand_ins.py — BasicEngine.send,
Command with qubits / control_qubits / all_qubits, BasicQubit, the gate base classes — restating the published ProjectQ
behaviour these two files rely on.  `hiq.projectq.cengines.SwapScheduler / ClusterScheduler` are the unmodified compiled reference
schedulers (oracle/_ref/_sched_cpp).  The engine drives a numpy-oracle backend (slot maps and swaps as the reference engine
does them) and everything it emits is logged.

    python -m oracle.run_reference_greedy job.json out.json      (own process: it installs fake top-level modules)

job  = {"n": qubits, "R": ranks, "max_local": .., "cluster": .., "supremacy": bool, "gates": [[uid, [targets], [controls], is_z], ..]}
out  = {"log": [["perm", ids] | ["cluster", uids] | ["swap", pair ids] ...], "gates": {uid: [[targets], [controls]]}}
Only tests use this (tests/test_scheduler.py); it needs /root/reference, so it never runs on the GPU box.
"""
from __future__ import annotations

import json
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REF, "hiq/projectq/cengines/_greedyscheduler.py"))


def main(job_path, out_path):
    with open(job_path) as f:
        job = json.load(f)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import ref, statevec
    from oracle import projectq_stand_ins as stand_ins
    op, ty = stand_ins.install()
    # the packages the two reference files import from: schedulers = the compiled reference, ops = the reference's own file
    gates_mod = stand_ins.install_hiq_packages(ref.load_ref_sched())
    gs_mod = stand_ins.load_unmodified("hiq_reference_greedyscheduler", "hiq/projectq/cengines/_greedyscheduler.py")

    n, R = job["n"], job["R"]
    sim = statevec.SimulatorMPI(1, job["max_local"], job["cluster"], R)
    log = []

    class Backend:  # what GreedyScheduler asks main_engine.backend (reference _greedyscheduler.py:139-149, 222)
        def get_qubits_ids(self):
            return list(sim.get_qubits_ids())

        def get_local_qubits_ids(self):
            return list(sim.get_local_qubits_ids())

        def get_global_qubits_ids(self):
            return list(sim.get_global_qubits_ids())

        def set_qubits_perm(self, ids):
            log.append(["perm", [int(x) for x in ids]])
            sim.set_qubits_perm(list(ids))

    cluster = []

    class Recorder:  # the engine after the scheduler: applies what changes the slot maps, logs everything
        def receive(self, command_list):
            for cmd in command_list:
                g = cmd.gate
                if isinstance(g, gates_mod.AllocateQuregGate):
                    sim.allocate_qureg([q.id for q in cmd.qubits[0]], 0)
                elif isinstance(g, op.AllocateQubitGate):
                    sim.allocate_qubit(cmd.qubits[0][0].id)
                elif isinstance(g, op.DeallocateQubitGate):
                    log.append(["dealloc", [cmd.qubits[0][0].id]])
                elif isinstance(g, gates_mod.MetaSwapGate):
                    pairs = [int(q.id) for q in cmd.qubits[0]]
                    log.append(["swap", pairs])
                    sim.swap_qubits(pairs)
                elif isinstance(g, op.FlushGate):
                    if cluster:
                        log.append(["cluster", list(cluster)])
                        del cluster[:]
                else:
                    cluster.append(g.uid)

    main_engine = types.SimpleNamespace(backend=Backend())
    gs = gs_mod.GreedyScheduler(supremacy_circuit=bool(job.get("supremacy")), cluster_size=job["cluster"])
    gs.main_engine = main_engine
    gs.next_engine = Recorder()

    def qubit(i):
        return ty.BasicQubit(main_engine, int(i))

    cmds = []
    for uid, targets, controls, is_z in job["gates"]:
        gate = op.ZGate() if is_z else op.BasicGate()
        gate.uid = uid
        cmds.append(op.Command(main_engine, gate, ([qubit(t) for t in targets],), [qubit(c) for c in controls]))
    alloc = op.Command(main_engine, gates_mod.AllocateQuregGate(0), ([qubit(i) for i in range(n)],))
    gs.receive([alloc])
    gs.receive(cmds)
    gs.receive([op.Command(main_engine, op.FlushGate(), ([ty.WeakQubitRef(main_engine, -1)],))])
    final = {str(c.gate.uid): [[q.id for reg in c.qubits for q in reg], [q.id for q in c.control_qubits]] for c in cmds}
    with open(out_path, "w") as f:
        json.dump({"log": log, "gates": final, "maps": Backend().get_qubits_ids()}, f)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

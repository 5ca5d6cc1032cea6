"""Runs the reference's UNMODIFIED Python scheduling engine (TEST INFRASTRUCTURE).

`/root/reference/hiq/projectq/cengines/_greedyscheduler.py` (class GreedyScheduler: the stage / cluster loop that calls
ClusterScheduler / SwapScheduler) and `/root/reference/hiq/projectq/ops/_gates.py` (MetaSwap,
AllocateQuregGate) are loaded from where they lie, byte for byte; what they import from ProjectQ (third party, absent here:
`projectq>=0.4.0`, requirements.txt:2) is supplied by ~60 lines of stand-ins below — BasicEngine.send, Command with
qubits / control_qubits / all_qubits, BasicQubit, the gate base classes — restating the published ProjectQ behaviour these
two files rely on.  `hiq.projectq.cengines.SwapScheduler / ClusterScheduler` are the unmodified compiled reference
schedulers (oracle/_ref/_sched_cpp).  The engine drives a numpy-oracle backend (slot maps and swaps as the reference engine
does them) and everything it emits is logged.

    python -m oracle.run_reference_greedy job.json out.json      (own process: it installs fake top-level modules)

job  = {"n": qubits, "R": ranks, "max_local": .., "cluster": .., "supremacy": bool, "gates": [[uid, [targets], [controls], is_z], ..]}
out  = {"log": [["perm", ids] | ["cluster", uids] | ["swap", pair ids] ...], "gates": {uid: [[targets], [controls]]}}
Only tests use this (tests/test_scheduler.py); it needs /root/reference, so it never runs on the GPU box.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REF, "hiq/projectq/cengines/_greedyscheduler.py"))


def _install_projectq_stand_ins():
    pq = types.ModuleType("projectq")
    ce = types.ModuleType("projectq.cengines")
    op = types.ModuleType("projectq.ops")
    ty = types.ModuleType("projectq.types")
    be = types.ModuleType("projectq.backends")
    me = types.ModuleType("projectq.meta")

    class BasicEngine:  # projectq/cengines/_basics.py: engines form a chain; send() hands commands to the next one
        def __init__(self):
            self.main_engine = None
            self.next_engine = None
            self.is_last_engine = False

        def send(self, command_list):
            self.next_engine.receive(command_list)

    class BasicQubit:  # projectq/types/_qubit.py
        def __init__(self, engine, idx):
            self.engine = engine
            self.id = idx

    class WeakQubitRef(BasicQubit):
        pass

    class BasicGate:  # projectq/ops/_basics.py (generate_command / make_tuple_of_qureg)
        @staticmethod
        def make_tuple_of_qureg(qubits):
            if not isinstance(qubits, tuple):
                qubits = (qubits,)
            qubits = list(qubits)
            for i in range(len(qubits)):
                if isinstance(qubits[i], BasicQubit):
                    qubits[i] = [qubits[i]]
            return tuple(qubits)

        def generate_command(self, qubits):
            qubits = self.make_tuple_of_qureg(qubits)
            engines = [q.engine for reg in qubits for q in reg]
            return Command(engines[0], self, qubits)

    class ClassicalInstructionGate(BasicGate):
        pass

    class FastForwardingGate(ClassicalInstructionGate):
        pass

    class FlushGate(FastForwardingGate):
        pass

    class AllocateQubitGate(ClassicalInstructionGate):
        pass

    class DeallocateQubitGate(FastForwardingGate):
        pass

    class ZGate(BasicGate):
        pass

    class Command:  # projectq/ops/_command.py
        def __init__(self, engine, gate, qubits, controls=(), tags=()):
            self.engine = engine
            self.gate = gate
            self.qubits = tuple(list(q) for q in qubits)
            self._control_qubits = list(controls)
            self.tags = list(tags)

        @property
        def control_qubits(self):
            return self._control_qubits

        @control_qubits.setter
        def control_qubits(self, qubits):
            self._control_qubits = list(qubits)

        @property
        def all_qubits(self):
            return (self._control_qubits,) + self.qubits

    class ResourceCounter:  # only patched by hiq/projectq/ops/_gates.py, never used here
        def _add_cmd(self, cmd):
            pass

    ce.BasicEngine = BasicEngine
    for cls in (BasicGate, ClassicalInstructionGate, FastForwardingGate, FlushGate, AllocateQubitGate, DeallocateQubitGate, ZGate, Command):
        setattr(op, cls.__name__, cls)
    ty.BasicQubit, ty.WeakQubitRef = BasicQubit, WeakQubitRef
    be.ResourceCounter = ResourceCounter
    me.get_control_count = lambda cmd: len(cmd.control_qubits)
    pq.cengines, pq.ops, pq.types, pq.backends, pq.meta = ce, op, ty, be, me
    for m in (pq, ce, op, ty, be, me):
        sys.modules[m.__name__] = m
    return op, ty


def _load_unmodified(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main(job_path, out_path):
    with open(job_path) as f:
        job = json.load(f)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import ref, statevec
    op, ty = _install_projectq_stand_ins()
    # the packages the two reference files import from: schedulers = the compiled reference, ops = the reference's own file
    hiq = types.ModuleType("hiq")
    hpq = types.ModuleType("hiq.projectq")
    hce = types.ModuleType("hiq.projectq.cengines")
    sched = ref.load_ref_sched()
    hce.SwapScheduler, hce.ClusterScheduler = sched.SwapScheduler, sched.ClusterScheduler
    for m in (hiq, hpq, hce):
        sys.modules[m.__name__] = m
    gates_mod = _load_unmodified("hiq.projectq.ops._gates", "hiq/projectq/ops/_gates.py")
    hop = types.ModuleType("hiq.projectq.ops")
    hop.MetaSwap, hop.AllocateQuregGate = gates_mod.MetaSwap, gates_mod.AllocateQuregGate
    sys.modules["hiq.projectq.ops"] = hop
    gs_mod = _load_unmodified("hiq_reference_greedyscheduler", "hiq/projectq/cengines/_greedyscheduler.py")

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

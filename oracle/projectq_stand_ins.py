"""Stand-ins for the few ProjectQ classes the reference's Python layer imports (TEST INFRASTRUCTURE).

ProjectQ (`projectq>=0.4.0`, reference requirements.txt:2) is third party and absent here.  The reference's own files
`hiq/projectq/cengines/_greedyscheduler.py`, `hiq/projectq/backends/_sim/_simulator_mpi.py` and `hiq/projectq/ops/_gates.py`
are loaded UNMODIFIED from /root/reference by oracle/run_reference_greedy.py and oracle/run_reference_pipeline.py; what
they import from ProjectQ is supplied here, restating the published behaviour they rely on and nothing more:
BasicEngine.send hands commands to the next engine; Command carries gate / qubits (tuple of registers) / control_qubits /
tags and all_qubits = (controls,) + qubits; BasicGate.generate_command wraps qubits into a one-register tuple;
Measure / Deallocate / Flush are fast-forwarding gates, Allocate is a classical instruction; Allocate, Deallocate and Measure
are singletons compared with ==.  mpi4py is replaced by a one-process stand-in (the engine's own communication goes
through the shared-memory Boost.MPI stand-in of oracle/shim)."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = "/root/reference"


def install():
    pq = types.ModuleType("projectq")
    ce = types.ModuleType("projectq.cengines")
    op = types.ModuleType("projectq.ops")
    ty = types.ModuleType("projectq.types")
    be = types.ModuleType("projectq.backends")
    me = types.ModuleType("projectq.meta")

    class BasicEngine:  # projectq/cengines/_basics.py
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

    class BasicGate:  # projectq/ops/_basics.py
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

    class MeasureGate(FastForwardingGate):
        pass

    class ZGate(BasicGate):
        pass

    class BasicMathGate(BasicGate):
        pass

    class TimeEvolution(BasicGate):
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

    class ResourceCounter:  # only patched by hiq/projectq/ops/_gates.py
        def _add_cmd(self, cmd):
            pass

    class LogicalQubitIDTag:  # projectq/meta/_logicalqubit.py
        def __init__(self, logical_qubit_id):
            self.logical_qubit_id = logical_qubit_id

    ce.BasicEngine = BasicEngine
    for cls in (BasicGate, ClassicalInstructionGate, FastForwardingGate, FlushGate, AllocateQubitGate, DeallocateQubitGate, MeasureGate, ZGate,
                BasicMathGate, TimeEvolution, Command):
        setattr(op, cls.__name__, cls)
    op.Allocate, op.Deallocate, op.Measure = AllocateQubitGate(), DeallocateQubitGate(), MeasureGate()
    op.NOT, op.H, op.R = BasicGate(), BasicGate(), BasicGate  # imported by the wrapper, not used on this path
    ty.BasicQubit, ty.WeakQubitRef = BasicQubit, WeakQubitRef
    be.ResourceCounter = ResourceCounter
    me.get_control_count = lambda cmd: len(cmd.control_qubits)
    me.LogicalQubitIDTag = LogicalQubitIDTag
    pq.cengines, pq.ops, pq.types, pq.backends, pq.meta = ce, op, ty, be, me
    for m in (pq, ce, op, ty, be, me):
        sys.modules[m.__name__] = m

    # mpi4py as the wrapper uses it (rc flags, thread checks; cheat()'s Allgather is not used by the runners)
    mpi4py = types.ModuleType("mpi4py")
    mpi4py.rc = types.SimpleNamespace()
    mpi = types.ModuleType("mpi4py.MPI")
    mpi.THREAD_FUNNELED = 1
    mpi.Is_thread_main = lambda: True
    mpi.Query_thread = lambda: 1
    mpi4py.MPI = mpi
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = mpi
    return op, ty


def load_unmodified(name, rel, package=None):
    """import a reference source file from where it lies, under module name `name`"""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install_hiq_packages(sched_module, cppsim_module=None):
    """the packages the reference files import from: hiq.projectq.cengines (schedulers), hiq.projectq.ops (the reference's
    own _gates.py), and — for the backend wrapper's relative import — hiq.projectq.backends._sim._cppsim_mpi"""
    mods = {}
    for name in ("hiq", "hiq.projectq", "hiq.projectq.cengines", "hiq.projectq.backends", "hiq.projectq.backends._sim"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        mods[name] = m
    mods["hiq.projectq.cengines"].SwapScheduler = sched_module.SwapScheduler
    mods["hiq.projectq.cengines"].ClusterScheduler = sched_module.ClusterScheduler
    gates_mod = load_unmodified("hiq.projectq.ops._gates", "hiq/projectq/ops/_gates.py")
    hop = types.ModuleType("hiq.projectq.ops")
    hop.MetaSwap, hop.AllocateQuregGate = gates_mod.MetaSwap, gates_mod.AllocateQuregGate
    sys.modules["hiq.projectq.ops"] = hop
    if cppsim_module is not None:
        sys.modules["hiq.projectq.backends._sim._cppsim_mpi"] = cppsim_module
    return gates_mod

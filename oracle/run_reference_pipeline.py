"""The reference's whole Python-to-engine pipeline, UNMODIFIED, on one rank of R (TEST INFRASTRUCTURE).

    GreedyScheduler   /root/reference/hiq/projectq/cengines/_greedyscheduler.py      loaded byte for byte
    SimulatorMPI      /root/reference/hiq/projectq/backends/_sim/_simulator_mpi.py   loaded byte for byte (the backend wrapper:
                      receive / _handle, the caller side of the drop-in boundary, SURVEY §8b "who calls it")
    _cppsim_mpi       oracle/_ref: the unmodified compiled engine          _sched_cpp: the unmodified compiled schedulers

ProjectQ and mpi4py (third party, absent) are replaced by oracle/projectq_stand_ins.py.  Every call the wrapper makes on the
engine object is recorded (method, arguments), so that the product's mirror of the wrapper (hiqsimulator_b200/backends.py)
can be compared call for call; the state, the slot maps and the measured bits come back too.

    python -m oracle.run_reference_pipeline job.pkl out.pkl    one process per rank (oracle.ref.run_module_on_ranks sets the
                                                               rank environment of the shared-memory Boost.MPI stand-in)
job = {"n", "max_local", "cluster", "seed", "gate_fusion", "gates": [(targets, controls, is_z, matrix)], "measure": [ids]}
out = {"calls": [(method, args...)], "maps", "id2pos", "state" (this rank's slab before the measurement), "bits": {id: bool}}
"""
from __future__ import annotations

import os
import pickle
import sys
import types

import numpy as np


def available() -> bool:
    from oracle import projectq_stand_ins as s
    return os.path.exists(os.path.join(s.REF, "hiq/projectq/backends/_sim/_simulator_mpi.py"))


def _summary(x):
    """arguments as plain comparable data (matrices rounded far below the parity tolerance)"""
    if isinstance(x, (list, tuple)):
        return [_summary(v) for v in x]
    if isinstance(x, (bool, np.bool_)):
        return bool(x)
    if isinstance(x, (int, np.integer)):
        return int(x)
    if isinstance(x, (float, complex, np.floating, np.complexfloating)):
        c = complex(x)
        return [round(c.real, 13), round(c.imag, 13)]
    return x


def main(job_path, out_path):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    with open(job_path, "rb") as f:
        job = pickle.load(f)
    from oracle import projectq_stand_ins as stand_ins
    from oracle import ref
    op, ty = stand_ins.install()
    refsim = ref.load_ref_sim()
    calls = []

    class RecordingEngine:
        """the compiled reference engine, every call logged"""

        def __init__(self, *args):
            calls.append(("ctor",) + tuple(_summary(a) for a in args))
            self._e = refsim.SimulatorMPI(*args)

        def __getattr__(self, name):
            fn = getattr(self._e, name)

            def logged(*args):
                if name not in ("get_qubits_ids", "get_local_qubits_ids", "get_global_qubits_ids", "cheat_local"):
                    calls.append((name,) + tuple(_summary(a) for a in args))
                return fn(*args)
            return logged

    cpp = types.ModuleType("hiq.projectq.backends._sim._cppsim_mpi")
    cpp.SimulatorMPI = RecordingEngine
    gates_mod = stand_ins.install_hiq_packages(ref.load_ref_sched(), cpp)
    gs_mod = stand_ins.load_unmodified("hiq_reference_greedyscheduler", "hiq/projectq/cengines/_greedyscheduler.py")
    be_mod = stand_ins.load_unmodified("hiq.projectq.backends._sim._simulator_mpi", "hiq/projectq/backends/_sim/_simulator_mpi.py",
                                       package="hiq.projectq.backends._sim")

    backend = be_mod.SimulatorMPI(gate_fusion=job["gate_fusion"], rnd_seed=job["seed"], num_local_qubits=job["max_local"],
                                  max_fused_qubits=job["cluster"])
    bits = {}
    main_engine = types.SimpleNamespace(backend=backend, mapper=None,
                                        set_measurement_result=lambda qb, value: bits.__setitem__(int(qb.id), bool(value)))
    backend.main_engine = main_engine
    backend.is_last_engine = True
    gs = gs_mod.GreedyScheduler(cluster_size=job["cluster"])
    gs.main_engine = main_engine
    gs.next_engine = backend

    def qubit(i):
        return ty.BasicQubit(main_engine, int(i))

    def flush():
        gs.receive([op.Command(main_engine, op.FlushGate(), ([ty.WeakQubitRef(main_engine, -1)],))])

    gs.receive([op.Command(main_engine, gates_mod.AllocateQuregGate(0), ([qubit(i) for i in range(job["n"])],))])
    cmds = []
    for targets, controls, is_z, matrix in job["gates"]:
        gate = op.ZGate() if is_z else op.BasicGate()
        gate.matrix = np.asarray(matrix, dtype=complex)
        cmds.append(op.Command(main_engine, gate, ([qubit(t) for t in targets],), [qubit(c) for c in controls]))
    gs.receive(cmds)
    flush()
    id2pos, slab = backend.cheat_local()
    out = {"maps": list(backend.get_qubits_ids()), "id2pos": dict(id2pos), "state": np.asarray(slab, dtype=np.complex128).copy()}
    if job.get("measure"):
        gs.receive([op.Command(main_engine, op.Measure, ([qubit(i) for i in job["measure"]],))])
        flush()
    out["calls"] = calls
    out["bits"] = bits
    with open(out_path, "wb") as f:
        pickle.dump(out, f)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

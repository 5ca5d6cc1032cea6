"""The product's Python layer against the reference's own, end to end, on the CPU.

oracle/run_reference_pipeline.py loads the reference's GreedyScheduler (_greedyscheduler.py) and its backend wrapper
(_simulator_mpi.py: receive / _handle — the caller side of the drop-in boundary) UNMODIFIED, on top of the unmodified compiled
engine and schedulers, one OS process per rank; ProjectQ / mpi4py are stand-ins (oracle/projectq_stand_ins.py).  The same
circuit goes through this repository's mirror (hiqsimulator_b200/cengines.py + backends.py) on the numpy oracle.  Compared:
every call the wrapper makes on the engine object — method, order, arguments: relabelling, gates with their matrices and
target / control roles, run() after every cluster, swaps, the measurement —, the state before the measurement, the slot maps
and the measured bits."""
import copy
import os
import pickle
import tempfile

import numpy as np
import pytest

from oracle import ref, statevec


def _summary(x):
    if isinstance(x, np.ndarray):
        x = x.tolist()
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


def _reference_pipeline(job, R):
    with tempfile.TemporaryDirectory() as tmp:
        jp = os.path.join(tmp, "job.pkl")
        with open(jp, "wb") as f:
            pickle.dump(job, f)
        return ref.run_module_on_ranks("oracle.run_reference_pipeline", jp, R, 1, timeout=600)


class _Recorder:
    """the numpy oracle behind the pybind surface of the reference engine, every call logged"""
    R = 1
    calls = None

    def __init__(self, seed, max_local, max_cluster):
        type(self).calls.append(("ctor", int(seed), int(max_local), int(max_cluster)))
        self._e = statevec.SimulatorMPI(seed, max_local, max_cluster, type(self).R)

    def __getattr__(self, name):
        if name == "apply_controlled_matrix":  # the reference binding takes nested lists only: the mirror must fall back to them
            raise AttributeError(name)
        fn = getattr(self._e, name)

        def logged(*args):
            if name not in ("get_qubits_ids", "get_local_qubits_ids", "get_global_qubits_ids", "cheat_local", "cheat"):
                type(self).calls.append((name,) + tuple(copy.deepcopy(a) for a in args))
            return fn(*args)
        return logged


@pytest.mark.parametrize("kind,n,R,ml,cluster,fusion", [("random", 10, 1, 10, 4, True), ("random", 11, 2, 10, 4, True), ("random", 12, 4, 10, 3, True),
                                                        ("qft", 11, 2, 10, 4, True), ("grover", 8, 2, 8, 4, True), ("random", 9, 1, 9, 4, False),
                                                        ("random", 12, 8, 9, 4, True)])
def test_python_layer_equals_the_unmodified_reference_pipeline(kind, n, R, ml, cluster, fusion):
    from oracle import run_reference_pipeline
    if not run_reference_pipeline.available() or not ref.have_ref():
        pytest.skip("needs /root/reference and oracle/_ref")
    from hiqsimulator_b200 import backends, cengines, circuits, ops
    if kind == "random":
        nq, cmds = circuits.random_circuit(n, 5, seed=n + R)
    elif kind == "qft":
        nq, cmds = circuits.qft_circuit(n)
    else:
        nq, cmds = circuits.grover_circuit(n, 2)
    seed = 11
    job = {"n": nq, "max_local": ml, "cluster": cluster, "seed": seed, "gate_fusion": fusion, "measure": list(range(nq)),
           "gates": [([int(q) for q in c.qubits], [int(q) for q in c.controls], bool(c.is_z), np.asarray(c.matrix)) for c in cmds]}
    want = _reference_pipeline(job, R)

    _Recorder.R, _Recorder.calls = R, []
    be = backends.SimulatorMPI(gate_fusion=fusion, rnd_seed=seed, num_local_qubits=ml, max_fused_qubits=cluster, backend_class=_Recorder)
    eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=cluster)])
    eng.allocate_qureg(nq)
    eng.receive(copy.deepcopy(cmds))
    eng.flush()
    id2pos, state = be._simulator._e.cheat()
    maps = list(be.get_qubits_ids())
    eng.receive([ops.Measure(list(range(nq)))])
    eng.flush()

    # call for call (every rank of the reference makes the same calls: compare with rank 0 and check the others agree)
    def canon(x):
        """one form for both logs: numbers (0, 0.0, 0j, and the runner's [re, im] pairs) -> (re, im) tuples"""
        if isinstance(x, np.ndarray):
            x = x.tolist()
        if isinstance(x, (bool, np.bool_)):
            return bool(x)
        if isinstance(x, (int, float, complex, np.integer, np.floating, np.complexfloating)):
            c = complex(x)
            return (round(c.real, 13) + 0.0, round(c.imag, 13) + 0.0)
        if isinstance(x, (list, tuple)):
            if len(x) == 2 and all(type(v) is float for v in x):
                return (x[0] + 0.0, x[1] + 0.0)
            return [canon(v) for v in x]
        return x
    got = [[c[0]] + [canon(a) for a in c[1:]] for c in _Recorder.calls]
    exp = [[c[0]] + [canon(a) for a in c[1:]] for c in want[0]["calls"]]
    assert len(got) == len(exp), (len(got), len(exp), [c[0] for c in got][:30], [c[0] for c in exp][:30])
    for j, (g, e) in enumerate(zip(got, exp)):
        assert g[0] == e[0], (j, g[0], e[0])
        assert g == e, (j, g[0], g[1:], e[1:])
    for r in range(1, R):
        assert [c[0] for c in want[r]["calls"]] == [c[0] for c in want[0]["calls"]]
    # state before the measurement, slot maps, measured bits
    assert maps == want[0]["maps"] and dict(id2pos) == want[0]["id2pos"]
    full = np.concatenate([want[r]["state"] for r in range(R)])
    assert np.abs(np.asarray(state) - full).max() <= 1e-12
    assert {q: bool(eng.measurements[q]) for q in range(nq)} == want[0]["bits"]
    if R >= 4:
        assert any(c[0] == "swap_qubits" for c in exp)  # the stage changes are part of what is compared

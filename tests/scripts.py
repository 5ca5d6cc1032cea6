"""Call-sequence ("script") helpers shared by the parity tests.

A script is a list of ops ``(method, *args)`` on the `_cppsim_mpi.SimulatorMPI` surface, starting
with ``("ctor", seed, max_local, max_cluster)``.  The same script is executed on
  * the compiled reference (oracle/_ref, one process per rank),
  * the numpy oracle (virtual ranks),
  * the CUDA engine (one process per GPU), or its dry-run trace,
so the tests read like the reference's own differential tests.
"""
from __future__ import annotations

import json

import numpy as np

from hiqsimulator_b200.gates import haar_unitary
from oracle import statevec


def random_script(nq, R, seed, ngates=60, max_local=None, max_cluster=4, queries=True, dealloc=False):
    """Random gate/swap/query sequence that respects the API contract (non-diagonal targets local,
    flush before a cluster would exceed 5 qubits)."""
    rng = np.random.default_rng(seed)
    ctor = ("ctor", 7 + seed, max_local or nq, max_cluster)
    script = [ctor, ("allocate_qureg", list(range(nq)), 0)]
    o = statevec.SimulatorMPI(*ctor[1:], R)
    o.allocate_qureg(list(range(nq)), 0)
    pending = set()

    def emit(op):
        nonlocal pending
        if op[0] == "apply_controlled_gate":
            need = set(op[2]) | set(op[3])
            if len(pending | need) > 5:
                script.append(("run",))
                o.run()
                pending = set()
            pending |= need
        else:
            pending = set()
        script.append(op)
        getattr(o, op[0])(*op[1:])
        if op[0] == "apply_controlled_gate":
            # the reference relies on its scheduler flushing right after a "huge" gate
            # (targets + local controls > max cluster; cluster_scheduler.h:122-136)
            n_local_ctrl = sum(1 for c in op[3] if c in o.get_local_qubits_ids())
            if len(op[2]) + n_local_ctrl > max_cluster:
                script.append(("run",))
                o.run()
                pending = set()

    for _ in range(ngates):
        kind = rng.integers(0, 10)
        loc = o.get_local_qubits_ids()
        glo = [q for q in o.get_global_qubits_ids() if q >= 0]
        if kind < 5:
            k = int(rng.integers(1, 4))
            ids = [int(x) for x in rng.choice(loc, size=k, replace=False)]
            rest = [q for q in loc + glo if q not in ids]
            nc = int(rng.integers(0, 3))
            ctrls = [int(x) for x in rng.choice(rest, size=min(nc, len(rest)), replace=False)]
            emit(("apply_controlled_gate", haar_unitary(1 << k, rng).tolist(), ids, ctrls))
        elif kind < 7:
            # diagonal gates may sit on global qubits, but only while the gate is not "huge"
            # (the reference cannot run a huge gate with a global target either)
            allq = loc + glo
            k = int(rng.integers(1, min(3, max_cluster) + 1))
            ids = [int(x) for x in rng.choice(allq, size=k, replace=False)]
            rest = [q for q in allq if q not in ids]
            nc = min(int(rng.integers(0, 2)), max_cluster - k)
            ctrls = [int(x) for x in rng.choice(rest, size=min(nc, len(rest)), replace=False)]
            d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k))
            emit(("apply_controlled_gate", np.diag(d).tolist(), ids, ctrls))
        elif kind < 8:
            emit(("run",))
        elif kind < 9 and glo:
            q = int(rng.integers(1, len(glo) + 1))
            gs = [int(x) for x in rng.choice(glo, size=q, replace=False)]
            ls = [int(x) for x in rng.choice(loc, size=q, replace=False)]
            pairs = []
            for a, b in zip(gs, ls):
                pairs += [a, b]
            emit(("run",))
            emit(("swap_qubits", pairs))
        else:
            ids = [int(rng.choice(loc))]
            rest = [q for q in loc + glo if q not in ids]
            ctrls = [int(x) for x in rng.choice(rest, size=min(5, len(rest)), replace=False)]
            emit(("apply_controlled_gate", haar_unitary(2, rng).tolist(), ids, ctrls))
    emit(("run",))
    script.append(("get_qubits_ids",))
    script.append(("cheat_local",))
    if queries:
        allq = o.get_local_qubits_ids() + [q for q in o.get_global_qubits_ids() if q >= 0]
        script.append(("get_probability", [True, False, True], allq[:3]))
        script.append(("get_probability", [False], [allq[-1]]))
        script.append(("entropy",))
        script.append(("measure_qubits", allq[:4]))
        script.append(("cheat_local",))
        script.append(("measure_qubits", allq[2:7]))
        script.append(("cheat_local",))
        script.append(("collapse_wavefunction", [allq[-1]], [_likely_bit(script, R, allq[-1])]))
        script.append(("cheat_local",))
        if dealloc:
            # measure everything, then release qubits one by one (they are classical now)
            script.append(("measure_qubits", list(allq)))
            for q in allq[: max(1, len(allq) // 2)]:
                script.append(("deallocate_qubit", int(q)))
                script.append(("get_qubits_ids",))
            script.append(("cheat_local",))
    return script


def _likely_bit(script, R, qid):
    """Outcome of qubit `qid` with probability >= 1/2 after the script so far (on the oracle)."""
    o = statevec.SimulatorMPI(*script[0][1:], R)
    for op in script[1:]:
        if op[0] == "cheat_local":
            continue
        getattr(o, op[0])(*op[1:])
    return o.get_probability([True], [qid]) >= 0.5


def run_on_oracle(script, R):
    o = statevec.SimulatorMPI(*script[0][1:], R)
    out = [None]
    for op in script[1:]:
        try:
            if op[0] == "cheat_local":
                id2pos, full = o.cheat()
                out.append((id2pos, full.copy()))
            else:
                out.append(getattr(o, op[0])(*op[1:]))
        except RuntimeError as e:
            out.append(("error", str(e)))
    return out


def merge_rank_outputs(per_rank):
    """results[rank][op] -> one list where cheat_local slabs are concatenated over ranks."""
    R = len(per_rank)
    out = []
    for j in range(len(per_rank[0])):
        v = per_rank[0][j]
        if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], dict):
            out.append((dict(v[0]), np.concatenate([np.asarray(per_rank[r][j][1]) for r in range(R)])))
        else:
            for r in range(1, R):
                w = per_rank[r][j]
                same = (list(w) == list(v)) if isinstance(v, (list, tuple)) else (w == v or (isinstance(v, float) and abs(w - v) < 1e-13))
                assert same, ("ranks disagree on op %d" % j, v, w)
            out.append(v)
    return out


def run_on_sim(make_sim, script):
    """Execute on one rank of an object with the pybind surface (reference module or ours)."""
    sim = None
    out = []
    for op in script:
        try:
            if op[0] == "ctor":
                sim = make_sim(*op[1:])
                out.append(None)
            elif op[0] == "cheat_local":
                d, v = sim.cheat_local()
                out.append((dict(d), np.asarray(v, dtype=np.complex128).copy()))
            else:
                out.append(getattr(sim, op[0])(*op[1:]))
        except RuntimeError as e:
            out.append(("error", str(e)))
    return out


def assert_outputs_match(script, got, exp, tol=1e-12):
    assert len(got) == len(exp) == len(script)
    for j, op in enumerate(script):
        g, e = got[j], exp[j]
        if isinstance(e, tuple) and len(e) == 2 and e[0] == "error":
            assert isinstance(g, tuple) and g[0] == "error", (j, op[0], g)
        elif op[0] == "cheat_local":
            assert dict(g[0]) == dict(e[0]), (j, g[0], e[0])
            assert g[1].shape == e[1].shape
            assert np.abs(g[1] - e[1]).max() <= tol, (j, float(np.abs(g[1] - e[1]).max()))
        elif op[0] in ("get_qubits_ids", "get_local_qubits_ids", "get_global_qubits_ids", "measure_qubits"):
            assert list(g) == list(e), (j, op[0], g, e)  # bit-exact
        elif op[0] in ("get_probability", "entropy"):
            assert abs(g - e) <= (tol if op[0] == "get_probability" else 1e-10), (j, op[0], g, e)
        elif op[0] == "get_amplitude":
            assert abs(g - e) <= tol
        else:
            assert g is None or g == e, (j, op[0], g)


# ------------------------------------------------------------------ JSON (golden fixtures)
def script_to_json(script):
    def enc(x):
        if isinstance(x, complex):
            return {"c": [x.real, x.imag]}
        if isinstance(x, (list, tuple)):
            return [enc(y) for y in x]
        if isinstance(x, (np.integer,)):
            return int(x)
        if isinstance(x, (np.bool_, bool)):
            return bool(x)
        return x
    return json.dumps([enc(list(op)) for op in script])


def script_from_json(s):
    def dec(x):
        if isinstance(x, dict) and "c" in x:
            return complex(x["c"][0], x["c"][1])
        if isinstance(x, list):
            return [dec(y) for y in x]
        return x
    return [tuple(dec(op)) for op in json.loads(s)]


# ------------------------------------------------------------------ dry-run trace replay
def replay_traces(traces, R):
    """Apply per-rank descriptor traces (engine dry-run) with the oracle kernels.
    Swaps are collective: the i-th SWAP descriptor of every rank is executed together."""
    KIND = {"none": 0, "dense": 1, "diag": 2, "scale": 3, "swap": 4, "grow": 5, "fill": 6}
    vec = [np.zeros(1, dtype=np.complex128) for _ in range(R)]
    vec[0][0] = 1.0
    cursors = [0] * R
    while True:
        swaps = []
        for r in range(R):
            t = traces[r]
            while cursors[r] < len(t) and t[cursors[r]]["kind"] != KIND["swap"]:
                d = t[cursors[r]]
                cursors[r] += 1
                if d["kind"] == KIND["grow"]:
                    vec[r] = np.concatenate([vec[r], np.zeros_like(vec[r])])
                elif d["kind"] == KIND["fill"]:
                    vec[r][:] = d["payload"][0]
                elif d["kind"] == KIND["dense"]:
                    statevec.apply_dense(vec[r], list(d["slots"]), np.asarray(d["payload"]), int(d["ctrl_mask"]))
                elif d["kind"] == KIND["diag"]:
                    statevec.apply_diag(vec[r], list(d["slots"]), np.asarray(d["payload"]), int(d["ctrl_mask"]))
                elif d["kind"] == KIND["scale"]:
                    vec[r] *= d["payload"][0]
            if cursors[r] < len(t):
                swaps.append(tuple(int(x) for x in t[cursors[r]]["aux"]))
                cursors[r] += 1
        if not swaps:
            break
        assert len(swaps) == R and len(set(swaps)) == 1, "ranks disagree on the swap plan"
        aux = swaps[0]
        L = int(np.log2(vec[0].shape[0]))
        g = int(np.log2(R))
        full = np.concatenate(vec).reshape((2,) * (g + L))
        nb = g + L
        for i in range(0, len(aux), 2):
            full = np.swapaxes(full, nb - 1 - (L + aux[i]), nb - 1 - aux[i + 1])
        full = np.ascontiguousarray(full).reshape(R, 1 << L)
        vec = [full[r].copy() for r in range(R)]
    return np.concatenate(vec)

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
                out.append(_dispatch(o, op))
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


def run_on_sim(make_sim, script, keep=None, before_op=None):
    """Execute on one rank of an object with the pybind surface (reference module or ours).
    keep: optional list that receives the simulator object (so the caller can go on using it);
    before_op(j, op): optional hook called ahead of every op (the skew tests hold ranks back with it)."""
    sim = None
    out = []
    for j, op in enumerate(script):
        if before_op is not None:
            before_op(j, op)
        try:
            if op[0] == "ctor":
                sim = make_sim(*op[1:])
                if keep is not None:
                    keep.append(sim)
                out.append(None)
            elif op[0] == "cheat_local":
                d, v = sim.cheat_local()
                out.append((dict(d), np.asarray(v, dtype=np.complex128).copy()))
            else:
                out.append(_dispatch(sim, op))
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
        elif op[0] in ("get_probability", "entropy", "get_expectation_value"):
            assert abs(g - e) <= (1e-10 if op[0] == "entropy" else tol), (j, op[0], g, e)
        elif op[0] == "get_amplitude":
            assert abs(g - e) <= tol
        else:
            assert g is None or g == e, (j, op[0], g)


# ------------------------------------------------------------------ operator-level scripts
# Python callables for emulate_math, by name (scripts stay JSON-able)
MATH_FUNCS = {
    "plus2": lambda v: [v[0] + 2],                     # the reference's commented-out Plus2Gate (_simulator_mpi_test.py:223-226)
    "minus3": lambda v: [v[0] - 3],
    "swap_regs": lambda v: [v[1], v[0]],
    "xor_into": lambda v: [v[0], v[0] ^ v[1]],
    "add_regs": lambda v: [v[0], v[0] + v[1]],
}


def random_terms(rng, n_ids, n_terms, real_coefs, max_factors=4):
    terms = []
    for _ in range(n_terms):
        nf = int(rng.integers(0, min(max_factors, n_ids) + 1))
        idx = sorted(int(x) for x in rng.choice(n_ids, size=nf, replace=False))
        term = [(i, "XYZ"[int(rng.integers(0, 3))]) for i in idx]
        c = float(rng.normal()) if real_coefs else complex(rng.normal(), rng.normal())
        terms.append((term, c))
    return terms


def operator_script(nq, R, seed, max_cluster=3, ngates=25):
    """Random gates, then the operator-level calls of the wrapper (get_expectation_value, apply_qubit_operator,
    emulate_math in its closed and tabulated forms, set_wavefunction), each followed by a look at the state."""
    rng = np.random.default_rng(1000 + seed)
    script = random_script(nq, R, seed, ngates=ngates, max_cluster=max_cluster, queries=False)
    o = statevec.SimulatorMPI(*script[0][1:], R)
    for op in script[1:]:
        if op[0] not in ("cheat_local", "get_qubits_ids"):
            getattr(o, op[0])(*op[1:])
    allq = [q for q in o.get_qubits_ids() if q >= 0]
    glo = [q for q in o.get_global_qubits_ids() if q >= 0]
    loc = o.get_local_qubits_ids()

    def look():
        script.append(("cheat_local",))

    ids = [int(x) for x in rng.permutation(allq)]
    # expectation values: mixed terms, the identity, a term with a repeated qubit, a long diagonal group
    script.append(("get_expectation_value", random_terms(rng, len(ids), 6, True) + [([], 0.4)], ids))
    script.append(("get_expectation_value", [([(0, "Y"), (0, "X"), (1, "Z")], 0.7), ([(2, "Z"), (2, "Y")], -0.2)], ids))
    zterms = []
    for _ in range(70):
        sub = sorted(int(x) for x in rng.choice(len(ids), size=int(rng.integers(1, 4)), replace=False))
        zterms.append(([(i, "Z") for i in sub], float(rng.normal())))
    script.append(("get_expectation_value", zterms, ids))
    if glo:
        gi = ids.index(glo[0])
        li = ids.index(loc[0])
        script.append(("get_expectation_value", [([(gi, "X")], 1.0), (sorted([(gi, "Y"), (li, "Z")]), 0.5), ([(gi, "Z")], -1.5)], ids))
    # one flip mask for every term -> in place; then a general operator -> accumulator
    a, b, c = 0, 1, 2
    script.append(("apply_qubit_operator", [([(a, "X"), (b, "Z")], 0.6 + 0.1j), ([(a, "Y")], -0.3j), ([(a, "X"), (c, "Z")], 0.2)], ids))
    look()
    script.append(("apply_qubit_operator", zterms[:5], ids))
    look()
    script.append(("apply_qubit_operator", random_terms(rng, len(ids), 7, False) + [([], 0.25)], ids))
    look()
    if glo:
        script.append(("apply_qubit_operator", [([(gi, "Y"), (li, "X")], 1.0)], ids))
        look()
    script.append(("get_probability", [True], [ids[0]]))
    # emulate_math: closed forms and tabulated Python functions, registers anywhere (local and global qubits)
    perm = [int(x) for x in rng.permutation(allq)]
    reg, rest = perm[:4], perm[4:]
    script.append(("emulate_math_add_constant", 5, reg, rest[:1]))
    look()
    script.append(("emulate_math_add_constant", -3, reg[:3], []))
    look()
    script.append(("emulate_math_add_constant_modN", 4, 11, reg, rest[:2]))
    look()
    script.append(("emulate_math_multiply_by_constant_modN", 7, 15, reg, rest[1:2]))
    look()
    if glo:
        reg_g = glo + [q for q in loc if q not in glo][:3]
        script.append(("emulate_math_multiply_by_constant_modN", 3, 1 << len(reg_g), reg_g, [q for q in allq if q not in reg_g][:1]))
        look()
        script.append(("emulate_math_add_constant", 1, sorted(allq), []))
        look()
    script.append(("emulate_math_fn", "plus2", [perm[:2]], perm[2:3]))
    look()
    script.append(("emulate_math_fn", "swap_regs", [perm[:2], perm[3:5]], []))
    look()
    script.append(("emulate_math_fn", "xor_into", [perm[1:3], perm[4:6]], perm[:1]))
    look()
    # set_wavefunction adopts the ordering it is given
    wf = rng.normal(size=1 << len(allq)) + 1j * rng.normal(size=1 << len(allq))
    wf /= np.linalg.norm(wf)
    order = [int(x) for x in rng.permutation(allq)]
    script.append(("set_wavefunction", [complex(x) for x in wf], order))
    script.append(("get_qubits_ids",))
    look()
    script.append(("get_expectation_value", random_terms(rng, len(order), 4, True), order))
    script.append(("measure_qubits", order[:3]))
    look()
    return script


def time_evolution_script(nq, R, seed):
    """random gates, then controlled and uncontrolled TimeEvolution calls (Hamiltonians with X/Y/Z strings on local and
    global qubits, identity terms, negative times), each followed by a look at the state"""
    rng = np.random.default_rng(2000 + seed)
    script = random_script(nq, R, seed, ngates=20, max_cluster=4, queries=False)
    o = statevec.SimulatorMPI(*script[0][1:], R)
    for op in script[1:]:
        if op[0] not in ("cheat_local", "get_qubits_ids"):
            getattr(o, op[0])(*op[1:])
    allq = [q for q in o.get_qubits_ids() if q >= 0]
    ids = [int(x) for x in rng.permutation(allq)]
    reg, rest = ids[:-2], ids[-2:]
    h1 = random_terms(rng, len(reg), 4, True, max_factors=3) + [([], 0.7)]
    script.append(("emulate_time_evolution", h1, 0.8, reg, [rest[0]]))
    script.append(("cheat_local",))
    h2 = random_terms(rng, len(ids), 3, True, max_factors=4)
    script.append(("emulate_time_evolution", h2, -0.45, ids, []))
    script.append(("cheat_local",))
    script.append(("emulate_time_evolution", [([], 1.3)], 0.5, reg, rest))  # a pure phase on the controlled part
    script.append(("cheat_local",))
    script.append(("get_probability", [True], [ids[0]]))
    return script


def _dispatch(sim, op):
    if op[0] == "emulate_math_fn":
        return sim.emulate_math(MATH_FUNCS[op[1]], op[2], op[3])
    return getattr(sim, op[0])(*op[1:])


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
KIND = {"none": 0, "dense": 1, "diag": 2, "scale": 3, "swap": 4, "grow": 5, "fill": 6,
        "pauli_expect": 7, "pauli_apply": 8, "pauli_commit": 9, "permute": 10, "load": 11, "launch": 13}
LAUNCH_DIAG_BATCH, LAUNCH_DENSE_PREDIAG, LAUNCH_TILE = 0, 1, 2


def decode_launch(d):
    """HIQ_DESC_LAUNCH record (include/hiq_b200.h) -> [(slots or None, matrix or None, [(slots, table), ...]), ...]"""
    aux = [int(x) for x in d["aux"]]
    pay = np.asarray(d["payload"])
    a = p = 0
    steps = []
    for _ in range(int(d["k"])):
        k = aux[a]
        a += 1
        slots = m = None
        if k >= 0:
            slots = tuple(aux[a:a + k])
            a += k
        n_ops = aux[a]
        a += 1
        if k >= 0:
            m = pay[p:p + (1 << (2 * k))].reshape(1 << k, 1 << k)
            p += 1 << (2 * k)
        ops = []
        for _ in range(n_ops):
            ko = aux[a]
            a += 1
            ops.append((list(aux[a:a + ko]), pay[p:p + (1 << ko)]))
            a += ko
            p += 1 << ko
        steps.append((slots, m, ops))
    assert a == len(aux) and p == len(pay)
    return steps


def _apply_launch(vec, d, info):
    """one multi-pass launch of a launch trace with the oracle kernels; when info["tile_emulator"] is set it additionally goes
    through its launcher's parameter image and the emulator of its kernel (tests/tile_emulator.py, tests/diag_emulator.py)"""
    steps = decode_launch(d)
    emu = None
    if info.get("tile_emulator"):
        from hiqsimulator_b200 import kernels as K
        L = int(np.log2(vec.shape[0]))
        emu = vec.copy()
        if d["form"] == LAUNCH_TILE:
            import tile_emulator
            tile_emulator.run_image(K.tile_program_image(L, steps), emu)
        elif d["form"] == LAUNCH_DIAG_BATCH:
            import diag_emulator
            diag_emulator.run_diag_batch_image(K.diag_batch_image(L, steps[0][2]), emu)
        else:
            import diag_emulator
            slots, m, ops = steps[0]
            diag_emulator.run_dense_prediag_image(K.dense_prediag_image(L, list(slots), m, ops), emu)
    for slots, m, ops in steps:
        for sl, table in ops:
            if sl:
                statevec.apply_diag(vec, sl, table, 0)
            else:
                vec *= table[0]
        if m is not None:
            statevec.apply_dense(vec, list(slots), m, 0)
    key = {LAUNCH_DIAG_BATCH: "diag_batch", LAUNCH_DENSE_PREDIAG: "dense_prediag", LAUNCH_TILE: "tile"}[d["form"]]
    info.setdefault("launch_forms", {}).setdefault(key, 0)
    info["launch_forms"][key] += 1
    if emu is not None:
        err = float(np.abs(emu - vec).max())
        assert err <= 1e-12, "%s image of a scheduled launch differs from the oracle by %g" % (key, err)
        if d["form"] == LAUNCH_TILE:
            info["tile_images_emulated"] = info.get("tile_images_emulated", 0) + 1
        info["images_emulated"] = info.get("images_emulated", 0) + 1
_COLLECTIVE = {KIND["swap"], KIND["pauli_expect"], KIND["pauli_apply"], KIND["pauli_commit"], KIND["permute"], KIND["load"]}


def replay_traces(traces, R, info=None):
    """Apply per-rank descriptor traces (engine dry-run) with the oracle kernels.
    Swaps and the operator-level passes are collective: the i-th such descriptor of every rank is executed
    together.  `info` (optional dict): "loads" = host vectors consumed by LOAD descriptors in order;
    on return "expect" = the value (summed over ranks) of every PAULI_EXPECT descriptor."""
    info = info if info is not None else {}
    loads = list(info.get("loads", []))
    info["expect"] = []
    vec = [np.zeros(1, dtype=np.complex128) for _ in range(R)]
    vec[0][0] = 1.0
    acc = [None] * R
    cursors = [0] * R
    while True:
        stops = []
        for r in range(R):
            t = traces[r]
            while cursors[r] < len(t) and t[cursors[r]]["kind"] not in _COLLECTIVE:
                d = t[cursors[r]]
                cursors[r] += 1
                if d["kind"] == KIND["grow"]:
                    vec[r] = np.concatenate([vec[r], np.zeros_like(vec[r])])
                elif d["kind"] == KIND["fill"]:
                    vec[r][:] = d["payload"][0]
                elif d["kind"] == KIND["dense"]:
                    emu = None
                    if info.get("tile_emulator"):  # a single-gate launch: through hiqk_apply_dense's image and its kernel emulator
                        import dense_emulator
                        from hiqsimulator_b200 import kernels as K
                        k = len(d["slots"])
                        emu = vec[r].copy()
                        dense_emulator.run_dense_image(K.dense_image(int(np.log2(emu.shape[0])), list(d["slots"]),
                                                                     np.asarray(d["payload"]).reshape(1 << k, 1 << k), int(d["ctrl_mask"])), emu)
                    statevec.apply_dense(vec[r], list(d["slots"]), np.asarray(d["payload"]), int(d["ctrl_mask"]))
                    if emu is not None:
                        err = float(np.abs(emu - vec[r]).max())
                        assert err <= 1e-12, "dense image of a scheduled launch differs from the oracle by %g" % err
                        info["images_emulated"] = info.get("images_emulated", 0) + 1
                elif d["kind"] == KIND["diag"]:
                    statevec.apply_diag(vec[r], list(d["slots"]), np.asarray(d["payload"]), int(d["ctrl_mask"]))
                elif d["kind"] == KIND["scale"]:
                    vec[r] *= d["payload"][0]
                elif d["kind"] == KIND["launch"]:
                    _apply_launch(vec[r], d, info)
            if cursors[r] < len(t):
                stops.append(t[cursors[r]])
                cursors[r] += 1
        if not stops:
            break
        assert len(stops) == R and len({d["kind"] for d in stops}) == 1, "ranks disagree on the collective sequence"
        kind = stops[0]["kind"]
        L = int(np.log2(vec[0].shape[0]))
        if kind == KIND["swap"]:
            plans = {tuple(int(x) for x in d["aux"]) for d in stops}
            assert len(plans) == 1, "ranks disagree on the swap plan"
            aux = plans.pop()
            g = int(np.log2(R))
            full = np.concatenate(vec).reshape((2,) * (g + L))
            nb = g + L
            for i in range(0, len(aux), 2):
                full = np.swapaxes(full, nb - 1 - (L + aux[i]), nb - 1 - aux[i + 1])
            full = np.ascontiguousarray(full).reshape(R, 1 << L)
            vec = [full[r].copy() for r in range(R)]
        elif kind in (KIND["pauli_expect"], KIND["pauli_apply"]):
            total = 0j
            new = [None] * R
            for r, d in enumerate(stops):
                aux = [int(x) for x in d["aux"]]
                mode, lx, src_rank = aux[0], aux[1], aux[2]
                terms = list(zip(aux[3:], [complex(c) for c in d["payload"]]))
                src = None if src_rank == r else vec[src_rank]
                if kind == KIND["pauli_expect"]:
                    total += statevec.pauli_expect(vec[r], lx, terms, src=src)
                elif mode == 0:
                    assert src is None
                    new[r] = vec[r].copy()
                    statevec.pauli_apply(new[r], lx, terms)
                else:
                    if mode == 1:
                        acc[r] = np.full(1 << L, np.nan + 0j)  # every element must be written by the first pass
                    statevec.pauli_apply(vec[r], lx, terms, acc=acc[r], accumulate=(mode == 2), src=src)
            if kind == KIND["pauli_expect"]:
                info["expect"].append(total)
            for r in range(R):
                if new[r] is not None:
                    vec[r] = new[r]
        elif kind == KIND["pauli_commit"]:
            vec = [a for a in acc]
            acc = [None] * R
        elif kind == KIND["permute"]:
            out = []
            for r, d in enumerate(stops):
                aux = [int(x) for x in d["aux"]]
                pk, a, N, cmask, takes_part = aux[:5]
                k = int(d["k"])
                pos = aux[5:5 + k]
                table = aux[5 + k:]
                gather = statevec.permute_gather(vec, r, L, pk, pos, cmask, a, N, table)
                if not takes_part:  # the rank filter is only a shortcut: the gather must be the identity there
                    assert np.array_equal(gather, vec[r])
                out.append(gather)
            vec = out
        elif kind == KIND["load"]:
            wf = np.asarray(loads.pop(0), dtype=np.complex128)
            for r, d in enumerate(stops):
                sl = int(d["aux"][0])
                vec[r] = np.zeros(1 << L, dtype=np.complex128) if sl < 0 else wf[sl << L:(sl + 1) << L].copy()
    return np.concatenate(vec)


# ------------------------------------------------------------------ scheduled circuits as scripts
def swap_skew_script(nq, R, seed, exchanges=12):
    """Consecutive single-pair exchanges on ALTERNATING global bits (the peer sets of two consecutive exchanges differ),
    low local slots (the packed transport in automatic mode), an all-pairs exchange now and then, a dense gate between
    some of them.  Returns (script, holds): holds[j] = the ranks to delay on the host before op j — the ranks whose
    partner in the NEXT exchange is a rank that was not delayed, so that the partner reaches the next exchange while
    the delayed rank is still inside this one (the case a process-wide staging buffer has to survive)."""
    rng = np.random.default_rng(seed)
    G = R.bit_length() - 1
    assert G >= 2, "needs at least 4 ranks"
    L = nq - G
    ctor = ("ctor", 11 + seed, L, 3)
    script = [ctor, ("allocate_qureg", list(range(nq)), 0)]
    o = statevec.SimulatorMPI(*ctor[1:], R)
    o.allocate_qureg(list(range(nq)), 0)
    holds = {}

    def emit(op):
        script.append(op)
        getattr(o, op[0])(*op[1:])

    def mix():
        loc = o.get_local_qubits_ids()
        for _ in range(3):
            ids = [int(x) for x in rng.choice(loc, size=2, replace=False)]
            emit(("apply_controlled_gate", haar_unitary(4, rng).tolist(), ids, []))
            emit(("run",))

    def rotate(ids):
        for q in ids:
            emit(("apply_controlled_gate", haar_unitary(2, rng).tolist(), [int(q)], []))
            emit(("run",))

    # every amplitude non-zero and different on every rank: rotate the local qubits, trade G of them for the global ones
    # (this first exchange also sets up the staging buffers, a collective step), rotate what came in
    rotate(o.get_local_qubits_ids())
    glo, loc = o.get_global_qubits_ids(), o.get_local_qubits_ids()
    pairs = []
    for gp in range(G):
        pairs += [int(glo[gp]), int(loc[gp])]
    emit(("swap_qubits", pairs))
    rotate(glo)
    mix()
    gpos_seq = [e % G for e in range(exchanges)]
    for e, g in enumerate(gpos_seq):
        glo = o.get_global_qubits_ids()
        loc = o.get_local_qubits_ids()
        if e % 5 == 4:  # all global bits at once: the peer set is the whole world
            slots = [int(x) for x in rng.choice(min(L, 4), size=G, replace=False)]
            pairs = []
            for gp, s in zip(range(G), slots):
                pairs += [int(glo[gp]), int(loc[s])]
        else:
            pairs = [int(glo[g]), int(loc[int(rng.integers(0, 3))])]
        if e % 2 == 0 and e + 1 < exchanges:
            nxt = gpos_seq[e + 1]
            holds[len(script)] = [r for r in range(R) if not (r >> nxt) & 1]
        emit(("swap_qubits", pairs))
        if e % 3 == 2:
            mix()
    emit(("get_qubits_ids",))
    script.append(("cheat_local",))
    return script, holds


def scheduled_script(kind, n, R, cluster=4, seed=1, depth=20):
    """The bench's pipeline (circuit generator -> GreedyScheduler -> backend) recorded as a script: the post-scheduler
    command stream (initial relabelling, gates, flushes, swaps) for `R` ranks, planned against a dry-run engine, ending
    with the slot maps and a look at the state.  Identical on every rank; runs on the compiled reference (one process
    per rank), the numpy oracle and the CUDA engine."""
    import bench
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import circuits, ops
    g = R.bit_length() - 1
    L = n - g
    cmds = circuits.qft_circuit(n)[1] if kind == "qft" else circuits.random_circuit(n, depth=depth)[1]
    stream, shape, _ = bench.schedule_circuit(n, L, cmds, lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, 0, R, M.FLAG_DRY_RUN), cluster=cluster)
    script = [("ctor", seed, L, cluster), ("allocate_qureg", list(range(n)), 0)]
    for c in stream:
        if isinstance(c, tuple):
            script.append(("set_qubits_perm", list(c[1])))
        elif c.kind == ops.FLUSH:
            script.append(("run",))
        elif c.kind == ops.METASWAP:
            script.append(("run",))
            script.append(("swap_qubits", list(c.qubits)))
        elif c.kind == ops.GATE:
            script.append(("apply_controlled_gate", c.matrix.tolist(), list(c.qubits), list(c.controls)))
        else:
            raise AssertionError("unexpected command in a scheduled stream: %r" % (c,))
    script += [("run",), ("get_qubits_ids",), ("cheat_local",)]
    return script, shape

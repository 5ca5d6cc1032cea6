"""`_sched_cpp` re-implementation vs the reference scheduler: cluster choices, swap choices and the
complete GreedyScheduler command stream must be bit-exact (BASELINE.json north_star).
Live differential runs need oracle/_ref (the unmodified reference scheduler, compiled by
oracle/Makefile; it travels to the GPU box); golden schedules in tests/golden pin the same thing
where it is absent."""
import copy
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import ref

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not ref.have_ref(), reason="oracle/_ref not built")


def _mine():
    from hiqsimulator_b200 import _sched_cpp
    return _sched_cpp


def _random_gates(rng, n_qubits, n_gates, max_targets=3, max_ctrls=2, diag_prob=0.0):
    gate, ctrl, diag = [], [], []
    for _ in range(n_gates):
        k = int(rng.integers(1, max_targets + 1))
        c = int(rng.integers(0, max_ctrls + 1))
        qs = [int(x) for x in rng.choice(n_qubits, size=min(k + c, n_qubits), replace=False)]
        gate.append(qs[:k])
        ctrl.append(qs[k:])
        diag.append(bool(rng.random() < diag_prob))
    return gate, ctrl, diag


@needs_ref
@pytest.mark.parametrize("seed", range(40))
def test_cluster_scheduler_fuzz(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(4, 24))
    g = int(rng.integers(0, 4))
    ids = [int(x) for x in rng.permutation(n + 3)[:n]]  # non-contiguous ids
    n_glob = min(g, n - 2)
    locals_, globals_ = ids[: n - n_glob], ids[n - n_glob:]
    if seed % 5 == 0:
        globals_ = globals_ + [-1]  # an unallocated global slot flows into the scheduler
    gate, ctrl, diag = _random_gates(rng, n, int(rng.integers(1, 120)), diag_prob=0.0 if seed % 3 else 0.3)
    gate = [[ids[q] for q in t] for t in gate]
    ctrl = [[ids[q] for q in t] for t in ctrl]
    size = int(rng.integers(2, 6))
    R = ref.load_ref_sched()
    exp = R.ClusterScheduler(gate, ctrl, diag, locals_, globals_, size).ScheduleCluster()
    got = _mine().ClusterScheduler(gate, ctrl, diag, locals_, globals_, size).ScheduleCluster()
    assert list(got) == list(exp)


@pytest.mark.parametrize("block", range(8))
def test_cluster_bounded_search_equals_replay(block):
    """The bounded search (chain bound + analytic first-visit order, csrc/sched.cpp) picks the very cluster
    the replay of the reference's enumeration picks, ties included; the replay itself is pinned to the
    reference above.  Layered circuits (a 1-qubit gate on every qubit, then pairs) are the tie-heavy case."""
    S = _mine()
    try:
        for seed in range(block * 60, block * 60 + 60):
            rng = np.random.default_rng(50000 + seed)
            n = int(rng.integers(3, 30))
            g = int(rng.integers(0, 4))
            ids = [int(x) for x in rng.permutation(n + 3)[:n]]
            n_glob = min(g, n - 2)
            locals_, globals_ = ids[: n - n_glob], ids[n - n_glob:]
            if seed % 5 == 0:
                globals_ = globals_ + [-1]
            if seed % 4 == 3:  # layered: singles everywhere, then a random matching
                gate, ctrl, diag = [], [], []
                for _ in range(int(rng.integers(1, 6))):
                    gate += [[q] for q in range(n)]
                    ctrl += [[] for _ in range(n)]
                    diag += [bool(rng.random() < 0.2) for _ in range(n)]
                    perm = [int(x) for x in rng.permutation(n)]
                    for a, b in zip(perm[0::2], perm[1::2]):
                        if rng.random() < 0.5:
                            gate.append([a]); ctrl.append([b]); diag.append(bool(rng.random() < 0.5))
                        else:
                            gate.append([a, b]); ctrl.append([]); diag.append(False)
                drop = int(rng.integers(0, len(gate)))
                gate, ctrl, diag = gate[drop:], ctrl[drop:], diag[drop:]
            else:
                gate, ctrl, diag = _random_gates(rng, n, int(rng.integers(1, 200)), max_targets=int(rng.integers(1, 4)),
                                                 max_ctrls=int(rng.integers(0, 3)), diag_prob=[0.0, 0.3, 0.8][seed % 3])
            gate = [[ids[q] for q in t] for t in gate]
            ctrl = [[ids[q] for q in t] for t in ctrl]
            size = int(rng.integers(1, 7))
            S.set_mode(1)
            exp = S.ClusterScheduler(gate, ctrl, diag, locals_, globals_, size).ScheduleCluster()
            S.set_mode(0)
            cs = S.ClusterScheduler(gate, ctrl, diag, locals_, globals_, size)
            got = cs.ScheduleCluster()
            assert list(got) == list(exp), (seed, n, size)
            assert cs.evaluated() <= max(1, cs.candidates())
    finally:
        S.set_mode(0)


@needs_ref
@pytest.mark.parametrize("seed", range(40))
def test_swap_scheduler_fuzz(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(4, 26))
    gate, ctrl, diag = _random_gates(rng, n, int(rng.integers(1, 150)), diag_prob=0.0 if seed % 3 else 0.3)
    num_locals = int(rng.integers(3, n + 1))
    # every gate must fit the local set
    gate = [t[:num_locals] for t in gate]
    splits = int(rng.choice([0, 1, 7, 100, 10 ** 4, 10 ** 6]))
    R = ref.load_ref_sched()
    for fuse in (True, False):
        exp = R.SwapScheduler(gate, ctrl, diag, splits, num_locals, fuse).ScheduleSwap()
        got = _mine().SwapScheduler(gate, ctrl, diag, splits, num_locals, fuse).ScheduleSwap()
        assert list(got) == list(exp)


@needs_ref
@pytest.mark.parametrize("block", range(6))
def test_swap_scheduler_layered_small_budgets(block):
    """Layered circuits (long single-continuation runs, dead paths) under tight and ample split budgets: the
    looped walk, the zero-budget tails and the dead-path exit of csrc/sched.cpp keep the reference's answer."""
    R = ref.load_ref_sched()
    for seed in range(block * 20, block * 20 + 20):
        rng = np.random.default_rng(7000 + seed)
        n = int(rng.integers(4, 22))
        gate, ctrl, diag = [], [], []
        for _ in range(int(rng.integers(1, 5))):
            gate += [[q] for q in range(n)]
            ctrl += [[] for _ in range(n)]
            diag += [bool(rng.random() < 0.2) for _ in range(n)]
            perm = [int(x) for x in rng.permutation(n)]
            for a, b in zip(perm[0::2], perm[1::2]):
                if rng.random() < 0.5:
                    gate.append([a]); ctrl.append([b]); diag.append(bool(rng.random() < 0.5))
                else:
                    gate.append([a, b]); ctrl.append([]); diag.append(False)
        num_locals = int(rng.integers(3, n + 1))
        splits = int(rng.choice([0, 1, 2, 3, 7, 20, 100, 1000, 10 ** 4, 10 ** 6]))
        for fuse in (True, False):
            exp = R.SwapScheduler(gate, ctrl, diag, splits, num_locals, fuse).ScheduleSwap()
            got = _mine().SwapScheduler(gate, ctrl, diag, splits, num_locals, fuse).ScheduleSwap()
            assert list(got) == list(exp), (seed, splits, fuse)


@pytest.mark.parametrize("kind,n,R,ml,cluster", [("random", 14, 4, 12, 4), ("random", 17, 8, 14, 4), ("random", 12, 2, 11, 3),
                                                 ("qft", 16, 4, 14, 4), ("qft", 13, 1, 13, 5), ("grover", 9, 2, 9, 4),
                                                 ("random", 15, 1, 15, 4), ("random", 16, 8, 13, 2)])
def test_planner_equals_python_loop(kind, n, R, ml, cluster):
    """`_sched_cpp.GreedyPlanner` (one C++ object emitting perm / controlled-Z role swaps / clusters / swap plans) drives
    the backend through exactly the command stream of the reference's Python loop (_greedyscheduler.py:203-242)."""
    from hiqsimulator_b200 import circuits
    if kind == "random":
        nq, cmds = circuits.random_circuit(n, 8, seed=n + R)
    elif kind == "qft":
        nq, cmds = circuits.qft_circuit(n)
    else:
        nq, cmds = circuits.grover_circuit(n, 2)
    runs = []
    for use_planner in (True, False):
        log, be = greedy_log_ex(nq, cmds, R, ml, cluster, use_planner)
        runs.append((log, be.get_qubits_ids(), [(d["kind"], list(d["slots"]), int(d["ctrl_mask"]), list(d["aux"]),
                                                 np.asarray(d["payload"]).tobytes()) for d in be._simulator.trace()]))
    assert runs[0][0] == runs[1][0]          # perm / cluster / swap sequence
    assert runs[0][1] == runs[1][1]          # final qubit maps
    assert runs[0][2] == runs[1][2]          # every descriptor the engine emitted (fused matrices bit for bit)


def _reference_python_engine(n, cmds, R, max_local, cluster, supremacy=False):
    """the UNMODIFIED reference _greedyscheduler.py (stand-ins for the ProjectQ classes it imports, compiled reference
    schedulers) on the same command list, in its own process: oracle/run_reference_greedy.py"""
    import json
    import subprocess
    import sys
    import tempfile
    job = {"n": n, "R": R, "max_local": max_local, "cluster": cluster, "supremacy": supremacy,
           "gates": [[i, [int(q) for q in c.qubits], [int(q) for q in c.controls], bool(c.is_z)] for i, c in enumerate(cmds)]}
    with tempfile.TemporaryDirectory() as tmp:
        jp, op_ = os.path.join(tmp, "job.json"), os.path.join(tmp, "out.json")
        with open(jp, "w") as f:
            json.dump(job, f)
        res = subprocess.run([sys.executable, "-m", "oracle.run_reference_greedy", jp, op_], cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-3000:]
        with open(op_) as f:
            return json.load(f)


@pytest.mark.parametrize("kind,n,R,ml,cluster", [("random", 12, 1, 12, 4), ("random", 13, 2, 12, 4), ("random", 14, 4, 12, 3), ("random", 15, 8, 12, 4),
                                                 ("qft", 14, 4, 12, 4), ("qft", 12, 1, 12, 5), ("grover", 9, 2, 9, 4), ("random", 13, 4, 11, 2),
                                                 ("supremacy", 13, 4, 11, 4), ("supremacy", 12, 1, 12, 4)])
def test_planner_equals_the_unmodified_reference_python_engine(kind, n, R, ml, cluster):
    """Row A20 against the reference's own file: hiq/projectq/cengines/_greedyscheduler.py is loaded byte for byte (with
    stand-ins for the few ProjectQ classes it imports and the compiled reference schedulers) and fed the circuit; the
    product's planner must emit the same relabelling, the same clusters in the same order, the same swaps, and leave the
    controlled-Z gates with the same target / control roles and the backend with the same slot maps."""
    from oracle import ref, run_reference_greedy
    if not run_reference_greedy.available() or not ref.have_ref():
        pytest.skip("needs /root/reference and oracle/_ref")
    from hiqsimulator_b200 import backends, cengines, circuits
    from oracle import statevec
    supremacy = kind == "supremacy"  # trailing controlled-Z gates are dropped (reference _greedyscheduler.py:151-173)
    if kind in ("random", "supremacy"):
        nq, cmds = circuits.random_circuit(n, 6, seed=n + R)
    elif kind == "qft":
        nq, cmds = circuits.qft_circuit(n)
    else:
        nq, cmds = circuits.grover_circuit(n, 2)
    want = _reference_python_engine(nq, cmds, R, ml, cluster, supremacy)
    # the same run is committed as a golden fixture (tests/make_golden_greedy.py), so that the pin also holds where
    # /root/reference is absent (test_planner_equals_the_golden_reference_engine_logs)
    golden = os.path.join(ROOT, "tests", "golden", "greedy_%s_%d_r%d_l%d_c%d.json" % (kind, n, R, ml, cluster))
    if os.path.exists(golden):
        with open(golden) as f:
            assert json.load(f) == want
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=ml, max_fused_qubits=cluster,
                               backend_class=lambda s, l, c: statevec.SimulatorMPI(s, l, c, R))
    gs = cengines.GreedyScheduler(cluster_size=cluster, supremacy_circuit=supremacy)
    eng = cengines.HiQMainEngine(be, [gs])
    eng.allocate_qureg(nq)
    mine = copy.deepcopy(cmds)
    for i, c in enumerate(mine):
        c.uid = i
    eng.receive(mine)
    eng.flush()
    got = [[k, [int(x) for x in v]] for k, v in gs.log]
    assert got == want["log"]
    assert len([e for e in got if e[0] == "cluster"]) > 3
    for i, c in enumerate(mine):
        assert [list(c.qubits), list(c.controls)] == want["gates"][str(i)], i
    assert list(be._simulator.get_qubits_ids()) == want["maps"]


GREEDY_GOLDEN = [("random", 12, 1, 12, 4), ("random", 13, 2, 12, 4), ("random", 14, 4, 12, 3), ("random", 15, 8, 12, 4), ("qft", 14, 4, 12, 4),
                 ("qft", 12, 1, 12, 5), ("grover", 9, 2, 9, 4), ("random", 13, 4, 11, 2), ("supremacy", 13, 4, 11, 4), ("supremacy", 12, 1, 12, 4)]


def greedy_case_circuit(kind, n, R):
    from hiqsimulator_b200 import circuits
    if kind in ("random", "supremacy"):
        return circuits.random_circuit(n, 6, seed=n + R)
    if kind == "qft":
        return circuits.qft_circuit(n)
    return circuits.grover_circuit(n, 2)


@pytest.mark.parametrize("kind,n,R,ml,cluster", GREEDY_GOLDEN)
def test_planner_equals_the_golden_reference_engine_logs(kind, n, R, ml, cluster):
    """the same comparison against the committed output of the unmodified reference engine (tests/golden/greedy_*.json,
    written by tests/make_golden_greedy.py): runs wherever the repository is, /root/reference or not"""
    from hiqsimulator_b200 import backends, cengines
    from oracle import statevec
    with open(os.path.join(ROOT, "tests", "golden", "greedy_%s_%d_r%d_l%d_c%d.json" % (kind, n, R, ml, cluster))) as f:
        want = json.load(f)
    nq, cmds = greedy_case_circuit(kind, n, R)
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=ml, max_fused_qubits=cluster,
                               backend_class=lambda s, l, c: statevec.SimulatorMPI(s, l, c, R))
    gs = cengines.GreedyScheduler(cluster_size=cluster, supremacy_circuit=(kind == "supremacy"))
    eng = cengines.HiQMainEngine(be, [gs])
    eng.allocate_qureg(nq)
    mine = copy.deepcopy(cmds)
    for i, c in enumerate(mine):
        c.uid = i
    eng.receive(mine)
    eng.flush()
    assert [[k, [int(x) for x in v]] for k, v in gs.log] == want["log"]
    for i, c in enumerate(mine):
        assert [list(c.qubits), list(c.controls)] == want["gates"][str(i)], i
    assert list(be._simulator.get_qubits_ids()) == want["maps"]


def test_planner_with_supremacy_preprocessing_equals_python_loop():
    """supremacy_circuit=True (trailing controlled-Z gates dropped, reference _greedyscheduler.py:151-173) is a preprocessing
    of the cached list: the planner path and the reference's loop emit the same schedule"""
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, circuits
    from oracle import greedy_loop
    nq, cmds = circuits.random_circuit(12, 6, seed=3)
    logs = []
    for cls in (cengines.GreedyScheduler, greedy_loop.ReferenceLoopScheduler):
        be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=10, max_fused_qubits=4,
                                   backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, 0, 4, M.FLAG_DRY_RUN))
        gs = cls(cluster_size=4, supremacy_circuit=True)
        eng = cengines.HiQMainEngine(be, [gs])
        eng.allocate_qureg(nq)
        c2 = copy.deepcopy(cmds)
        for i, c in enumerate(c2):
            c.uid = i
        eng.receive(c2)
        eng.flush()
        logs.append([[k, [int(x) for x in v]] for k, v in gs.log])
    assert logs[0] == logs[1] and len(logs[0]) > 5


def greedy_log_ex(n, cmds, R, max_local, cluster, use_planner):
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=max_local, max_fused_qubits=cluster,
                               backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, R - 1, R, M.FLAG_DRY_RUN))
    if use_planner:
        gs = cengines.GreedyScheduler(cluster_size=cluster)
    else:
        from oracle import greedy_loop
        gs = greedy_loop.ReferenceLoopScheduler(cluster_size=cluster)
    eng = cengines.HiQMainEngine(be, [gs])
    eng.allocate_qureg(n)
    cmds = copy.deepcopy(cmds)
    for i, c in enumerate(cmds):
        c.uid = i
    half = len(cmds) // 2
    eng.receive(cmds[:half])
    eng.flush()                 # a second scheduling round starts from the maps the first one left
    eng.receive(cmds[half:])
    eng.flush()
    return [[k, [int(x) for x in v]] for k, v in gs.log], be


def greedy_log(n, cmds, R, max_local, sched_module, cluster=4, supremacy=False):
    """Full GreedyScheduler run against a dry-run engine; returns the emitted schedule."""
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines
    M.init_world(0, R, b"", 0, M.FLAG_DRY_RUN)
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=max_local, max_fused_qubits=cluster)
    M.init_world(0, 1, b"", 0, 0)
    from oracle import greedy_loop
    gs = greedy_loop.ReferenceLoopScheduler(cluster_size=cluster, sched_module=sched_module, supremacy_circuit=supremacy)
    eng = cengines.HiQMainEngine(be, [gs])
    eng.allocate_qureg(n)
    cmds = copy.deepcopy(cmds)
    for i, c in enumerate(cmds):
        c.uid = i
    eng.receive(cmds)
    eng.flush()
    return [[k, [int(x) for x in v]] for k, v in gs.log], be


def _circuit(name):
    from hiqsimulator_b200 import circuits
    if name == "qft18":
        return circuits.qft_circuit(18) + (1, 18)
    if name == "qft20_r4":
        return circuits.qft_circuit(20) + (4, 18)
    if name == "rand20_r8":
        return circuits.random_circuit(20, 12) + (8, 17)
    if name == "rand16":
        return circuits.random_circuit(16, 20) + (1, 16)
    if name == "grover9":
        return circuits.grover_circuit(9, 3) + (1, 12)
    if name == "grover9_r2":
        return circuits.grover_circuit(9, 3) + (2, 9)
    raise KeyError(name)


CIRCUITS = ["qft18", "qft20_r4", "rand20_r8", "rand16", "grover9", "grover9_r2"]


@needs_ref
@pytest.mark.parametrize("name", CIRCUITS)
def test_greedy_schedule_matches_reference_scheduler(name):
    n, cmds, R, ml = _circuit(name)
    mine, _ = greedy_log(n, cmds, R, ml, _mine())
    theirs, _ = greedy_log(n, cmds, R, ml, ref.load_ref_sched())
    assert mine == theirs


@pytest.mark.parametrize("name", CIRCUITS)
def test_greedy_schedule_matches_golden(name):
    path = os.path.join(HERE, "golden", "sched_%s.json" % name)
    n, cmds, R, ml = _circuit(name)
    mine, _ = greedy_log(n, cmds, R, ml, _mine())
    with open(path) as f:
        assert mine == json.load(f)


def test_scheduled_circuit_is_executable_and_correct():
    """Dry-run descriptors of a scheduled multi-rank circuit, replayed with the oracle kernels,
    equal the unscheduled circuit applied gate by gate (the scheduler only reorders commuting work)."""
    import scripts
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, circuits
    from oracle import statevec
    n, cmds = circuits.random_circuit(10, 6, seed=5)
    R = 4
    traces = []
    for r in range(R):
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=8, max_fused_qubits=4)
        eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler()])
        eng.allocate_qureg(n)
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        traces.append(be._simulator.trace())
        perm = be.get_qubits_ids()
    M.init_world(0, 1, b"", 0, 0)
    state = scripts.replay_traces(traces, R)
    # plain application on a single-rank oracle, qubit q at bit q
    o = statevec.SimulatorMPI(1, n, 5, 1)
    o.allocate_qureg(list(range(n)), 0)
    for c in cmds:
        o.apply_controlled_gate(c.matrix, c.qubits, c.controls)
        o.run()
    ref_state = o.vec[0]
    # re-index: scheduled bit position p holds qubit perm[p]
    idx = np.arange(1 << n)
    src = np.zeros_like(idx)
    for p, q in enumerate(perm):
        src |= ((idx >> p) & 1) << q
    assert np.abs(state - ref_state[src]).max() <= 1e-12

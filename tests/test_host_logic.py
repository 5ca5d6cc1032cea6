"""Host logic of the engine WITHOUT a GPU: slot maps, per-rank gate slicing, fusion, swap plans.

The engine runs in dry-run mode (HIQ_FLAG_DRY_RUN: every device operation is recorded as a
descriptor instead of being launched), one engine per rank; the descriptor streams are replayed
with the oracle kernels and compared with the reference's final state from the golden fixtures.
Slot maps must be bit-exact."""
import numpy as np
import pytest

import scripts
from golden_util import golden_names, load_golden


def _dry_engines(script, R):
    from hiqsimulator_b200 import _cppsim_mpi as M
    engines = []
    for r in range(R):
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        engines.append(M.SimulatorMPI(*script[0][1:]))
    M.init_world(0, 1, b"", 0, 0)
    return engines


def _gate_prefix(script):
    """ops up to (and including) the first cheat_local: gates, runs, swaps, slot-map queries"""
    out = []
    for op in script:
        out.append(op)
        if op[0] == "cheat_local":
            break
    return out


@pytest.mark.parametrize("name", golden_names())
def test_dry_run_traces_reproduce_reference_state(name):
    R, script, exp = load_golden(name)
    prefix = _gate_prefix(script)
    engines = _dry_engines(script, R)
    for j, op in enumerate(prefix[1:], start=1):
        if op[0] == "cheat_local":
            break
        for e in engines:
            got = getattr(e, op[0])(*op[1:])
            if op[0] == "get_qubits_ids":
                assert list(got) == list(exp[j])  # slot maps are bit-exact
    state = scripts.replay_traces([e.trace() for e in engines], R)
    j = len(prefix) - 1
    id2pos, vec = exp[j]
    assert np.abs(state - vec).max() <= 1e-12
    # id -> position map of cheat_local
    for e in engines:
        d, _ = ({}, None)
        ids = e.get_qubits_ids()
        nl = len(e.get_local_qubits_ids())
        got = {q: p for p, q in enumerate(ids) if q != -1}
        assert got == id2pos and nl == len(e.get_local_qubits_ids())


@pytest.mark.parametrize("name", golden_names())
def test_launch_traces_reproduce_reference_state(name):
    """the LAUNCH trace of the same runs (what the device would be handed after diagonal folding and tile-run grouping:
    batched diagonal passes, dense launches that carry factors, tile programs) replays to the same state"""
    R, script, exp = load_golden(name)
    prefix = _gate_prefix(script)
    engines = _dry_engines(script, R)
    for op in prefix[1:]:
        if op[0] == "cheat_local":
            break
        for e in engines:
            getattr(e, op[0])(*op[1:])
    for e in engines:
        e.synchronize()  # closes the launch accounting: queued gates go out
    info = {"tile_emulator": True}
    state = scripts.replay_traces([e.launch_trace() for e in engines], R, info)
    id2pos, vec = exp[len(prefix) - 1]
    assert np.abs(state - vec).max() <= 1e-12
    for e in engines:  # every pass of the plan is carried by some launch, and folding only ever reduces their number
        st = e.stats()
        assert st["gate_launches"] <= st["dense_passes"] + st["diag_passes"] + st["scale_passes"]


def test_allocation_policy_matches_reference_example():
    """R=8, max_cluster=4, allocate_qureg(range(35))."""
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(3, 8, b"", 0, M.FLAG_DRY_RUN)
    e = M.SimulatorMPI(1, 32, 4)
    M.init_world(0, 1, b"", 0, 0)
    assert e.get_global_qubits_ids() == [-1, -1, -1]
    e.allocate_qureg(list(range(35)), 0)
    assert e.get_local_qubits_ids() == [0, 1, 2, 3] + list(range(7, 35))
    assert e.get_global_qubits_ids() == [4, 5, 6]
    with pytest.raises(RuntimeError):
        e.allocate_qubit(99)


def test_error_conventions_dry_run():
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(0, 2, b"", 0, M.FLAG_DRY_RUN)
    e = M.SimulatorMPI(1, 8, 4)
    M.init_world(0, 1, b"", 0, 0)
    e.allocate_qureg(list(range(6)), 0)       # locals 0..3,5 ; global 4
    assert e.get_global_qubits_ids() == [4]
    X = [[0, 1], [1, 0]]
    with pytest.raises(RuntimeError, match="non-diagonal"):
        e.apply_controlled_gate(X, [4], [])
    with pytest.raises(RuntimeError, match="unique"):
        e.swap_qubits([4, 0, 4, 1])
    with pytest.raises(RuntimeError, match="Can't find"):
        e.swap_qubits([0, 1])
    # emulate_math works here (the reference's throws "not supported", SimulatorMPI.hpp:217-225);
    # what it rejects is a function that is not reversible on the register
    e.emulate_math(lambda x: x, [[0]], [])
    with pytest.raises(RuntimeError, match="not reversible"):
        e.emulate_math(lambda x: [0], [[0]], [])
    # six fused qubits cannot run (reference: "Run(): cannot apply 6 qubits gate")
    e2 = _dry_engines([("ctor", 1, 8, 4)], 1)[0]
    e2.allocate_qureg(list(range(7)), 0)
    for q in range(6):
        e2.apply_controlled_gate(X, [q], [])
    with pytest.raises(RuntimeError, match="cannot apply 6"):
        e2.run()
    # data-dependent calls need the device
    with pytest.raises(RuntimeError, match="dry-run"):
        e.get_probability([True], [0])


def test_set_qubits_perm_is_relabel_only():
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(1, 4, b"", 0, M.FLAG_DRY_RUN)
    e = M.SimulatorMPI(1, 8, 2)
    M.init_world(0, 1, b"", 0, 0)
    e.allocate_qureg(list(range(6)), 0)
    ids = e.get_qubits_ids()
    assert ids == [0, 1, 4, 5, 2, 3]
    n = len(e.trace())
    ids[0], ids[4] = ids[4], ids[0]
    e.set_qubits_perm(ids)
    assert e.get_local_qubits_ids() == [2, 1, 4, 5] and e.get_global_qubits_ids() == [0, 3]
    assert len(e.trace()) == n  # no data motion


def test_fullsize_property_checks_hold_on_the_oracle():
    """the size-independent checks of tests/test_fullsize_gpu.py, run here against the numpy oracle at 12 qubits:
    the closed form (bit-reversed QFT output), the marginals and the circuit-then-inverse identity are properties
    of the reference pipeline, not of the CUDA engine"""
    import test_fullsize_gpu as F
    from oracle import statevec
    worst, marg, p_after = F.check_qft_closed_form(statevec.SimulatorMPI, 12, 256)
    assert worst <= 1e-12 and marg <= 1e-12 and abs(p_after - 1.0) <= 1e-12
    d_amp, d_p = F.check_circuit_then_inverse(statevec.SimulatorMPI, 10, 6)
    assert d_amp <= 1e-12 and d_p <= 1e-12


@pytest.mark.parametrize("L,gpos,slots,piece_bits", [(8, [0], [0], 5), (8, [0], [5], 7), (9, [1, 0], [0, 1], 4), (9, [0, 2], [6, 1], 5),
                                                      (10, [0, 1, 2], [0, 1, 2], 4), (10, [2, 0, 1], [7, 0, 3], 6)])
def test_packed_exchange_scheme(L, gpos, slots, piece_bits):
    """Index logic of Engine::exchange_packed (csrc/engine.cpp), restated with numpy: every rank packs piece i for peer k
    into staging[(i % 2) * n_peers + k], and unpacks peer k's data from the PEER's staging at the same (i % 2, k) —
    peer k of rank r is r ^ bits(x_k), so k also names r in that peer's list.  The result must be the transposition of
    global-index bit (L + gpos_j) with bit slots[j], whatever the piece size."""
    g = 3
    R = 1 << g
    q = len(gpos)
    rng = np.random.default_rng(L + sum(slots))
    full = rng.normal(size=R << L) + 1j * rng.normal(size=R << L)
    vec = [full[r << L:(r + 1) << L].copy() for r in range(R)]
    order = sorted(range(q), key=lambda j: slots[j])
    srt = sorted(slots)
    idx = np.arange(1 << L)
    free = np.zeros(1 << L, dtype=np.int64)  # free index of every local index (swapped slots removed)
    pos = 0
    for b in range(L):
        if b in slots:
            continue
        free |= ((idx >> b) & 1) << pos
        pos += 1

    def select(pat, begin, count):
        sel = (free >= begin) & (free < begin + count)
        for j, s in enumerate(srt):
            sel &= ((idx >> s) & 1) == ((pat >> j) & 1)
        return sel

    def peers_of(r):
        out = []
        for x in range(1, 1 << q):
            pr = r
            for i in range(q):
                if (x >> i) & 1:
                    pr ^= 1 << gpos[i]
            pat = sum(((pr >> gpos[order[j]]) & 1) << j for j in range(q))
            out.append((pr, pat))
        return out

    chunk = 1 << (L - q)
    piece = min(chunk, 1 << piece_bits)
    n_peers = (1 << q) - 1
    staging = [np.zeros(2 * n_peers * piece, dtype=np.complex128) for _ in range(R)]
    for i in range(chunk // piece):
        buf = (i % 2) * n_peers * piece
        for r in range(R):  # pack, then the group barrier
            for k, (pr, pat) in enumerate(peers_of(r)):
                staging[r][buf + k * piece: buf + (k + 1) * piece] = vec[r][select(pat, i * piece, piece)]
        for r in range(R):  # unpack from the peers' staging buffers
            for k, (pr, pat) in enumerate(peers_of(r)):
                assert peers_of(pr)[k][0] == r
                vec[r][select(pat, i * piece, piece)] = staging[pr][buf + k * piece: buf + (k + 1) * piece]
    got = np.concatenate(vec)
    gidx = np.arange(R << L, dtype=np.int64)
    src = gidx.copy()
    for j in range(q):
        hi, lo = L + gpos[j], slots[j]
        bh, bl = (src >> hi) & 1, (src >> lo) & 1
        src = src & ~((1 << hi) | (1 << lo)) | (bl << hi) | (bh << lo)
    assert np.array_equal(got, full[src])


@pytest.mark.parametrize("L,gpos,slots", [(8, [0], [0]), (8, [2], [5]), (9, [1, 0], [0, 1]), (9, [0, 2], [6, 1]), (9, [2, 1], [8, 3]),
                                          (10, [0, 1, 2], [0, 1, 2]), (10, [2, 0, 1], [7, 0, 3]), (10, [1, 2, 0], [9, 8, 4])])
def test_inplace_exchange_scheme(L, gpos, slots):
    """Index logic of Engine::exchange_p2p (csrc/engine.cpp) + swap_p2p_kernel (csrc/swap_kernels.cu), restated with numpy
    on 8 virtual ranks: rank r meets every rank that differs from it in a non-empty subset of the swapped global bits;
    of a pair, the lower rank handles the lower half of the free indices, the higher rank the rest; for every free index
    it handles, a rank trades its amplitude whose swapped slots spell the PEER's bits for the peer's amplitude whose
    swapped slots spell ITS bits.  Every rank runs its own kernel; together they must perform the transposition of
    global-index bit (L + gpos_j) with bit slots[j] — each pair element moved exactly once."""
    g = 3
    R = 1 << g
    q = len(gpos)
    rng = np.random.default_rng(3 * L + sum(slots))
    full = rng.normal(size=R << L) + 1j * rng.normal(size=R << L)
    vec = [full[r << L:(r + 1) << L].copy() for r in range(R)]
    order = sorted(range(q), key=lambda j: slots[j])
    srt = sorted(slots)

    def pattern_of(r):  # bit j <-> j-th lowest swapped slot
        return sum(((r >> gpos[order[j]]) & 1) << j for j in range(q))

    def spread(pat):
        return sum(((pat >> j) & 1) << srt[j] for j in range(q))

    def deposit(f):
        f = np.asarray(f, dtype=np.int64).copy()
        for pos in srt:
            f = ((f >> pos) << (pos + 1)) | (f & ((1 << pos) - 1))
        return f

    n = 1 << (L - q)
    half = (n + 1) // 2
    moved = [np.zeros(1 << L, dtype=np.int32) for _ in range(R)]
    for r in range(R):  # one kernel per rank; the pairs it touches are disjoint from every other rank's
        for x in range(1, 1 << q):
            pr = r
            for i in range(q):
                if (x >> i) & 1:
                    pr ^= 1 << gpos[i]
            begin, count = (0, half) if r < pr else (half, n - half)
            base = deposit(begin + np.arange(count))
            mine = base | spread(pattern_of(pr))
            theirs = base | spread(pattern_of(r))
            a, b = vec[r][mine].copy(), vec[pr][theirs].copy()
            vec[r][mine], vec[pr][theirs] = b, a
            np.add.at(moved[r], mine, 1)
            np.add.at(moved[pr], theirs, 1)
    got = np.concatenate(vec)
    gidx = np.arange(R << L, dtype=np.int64)
    src = gidx.copy()
    for j in range(q):
        hi, lo = L + gpos[j], slots[j]
        bh, bl = (src >> hi) & 1, (src >> lo) & 1
        src = src & ~((1 << hi) | (1 << lo)) | (bl << hi) | (bh << lo)
    assert np.array_equal(got, full[src])
    for r in range(R):  # an amplitude moves once unless its swapped slots spell its own rank's bits (it stays)
        idx = np.arange(1 << L)
        stays = np.ones(1 << L, dtype=bool)
        for j in range(q):
            stays &= ((idx >> slots[j]) & 1) == ((r >> gpos[j]) & 1)
        assert np.array_equal(moved[r] == 0, stays) and moved[r].max() == 1


@pytest.mark.parametrize("kind,n,R", [("random", 13, 2), ("random", 14, 4), ("qft", 13, 8), ("random", 12, 1)])
def test_scheduled_script_dry_run_equals_compiled_reference(kind, n, R):
    """the script tests/test_fullsize_multigpu.py diffs on the GPUs (bench pipeline -> scheduled stream), here on dry-run
    engines: descriptor traces replayed with the oracle kernels == the compiled reference run as R processes"""
    from oracle import ref
    if not ref.have_ref():
        pytest.skip("oracle/_ref is not built")
    script, shape = scripts.scheduled_script(kind, n, R)
    res = scripts.merge_rank_outputs(ref.run_script(script, R, 1))
    engines = _dry_engines(script, R)
    ids = None
    for op in script[1:-1]:
        for e in engines:
            got = getattr(e, op[0])(*op[1:])
            if op[0] == "get_qubits_ids":
                ids = list(got)
    state = scripts.replay_traces([e.trace() for e in engines], R)
    assert ids == list(res[-2])
    assert np.abs(state - res[-1][1]).max() <= 1e-12


@pytest.mark.parametrize("kind,n,R", [("random", 16, 1), ("qft", 17, 1), ("random", 15, 2), ("qft", 16, 4), ("random", 16, 8)])
def test_scheduled_launch_traces_equal_compiled_reference(kind, n, R):
    """the bench pipeline on dry-run engines with enough local qubits for tile runs to form (L >= 11): the launch trace —
    tile programs executed through the launcher's parameter image and tests/tile_emulator.py — equals the compiled
    reference run as R processes.  Covers on the CPU what test_engine_tile_runs_match_oracle covers on the GPU:
    the engine's grouping (which diagonals may ride along which gate, which gates share a tile) and the launcher's
    encoding of the runs the scheduler really produces."""
    from oracle import ref
    if not ref.have_ref():
        pytest.skip("oracle/_ref is not built")
    script, shape = scripts.scheduled_script(kind, n, R)
    res = scripts.merge_rank_outputs(ref.run_script(script, R, 1))
    engines = _dry_engines(script, R)
    for op in script[1:-1]:
        for e in engines:
            getattr(e, op[0])(*op[1:])
    for e in engines:
        e.synchronize()
    info = {"tile_emulator": True}
    state = scripts.replay_traces([e.launch_trace() for e in engines], R, info)
    assert np.abs(state - res[-1][1]).max() <= 1e-12
    st = engines[0].stats()
    assert st["tile_launches"] >= 1 and info["tile_images_emulated"] >= R  # tile runs did form, on every rank
    assert info["launch_forms"]["tile"] == sum(e.stats()["tile_launches"] for e in engines)


@pytest.mark.parametrize("seed", range(16))
def test_launch_trace_equals_plan_trace_on_random_scripts(seed):
    """random gate streams with diagonal gates, controls, swaps (tests/scripts.random_script): the launches the engine forms
    (batched diagonals, dense gates carrying factors, tile runs — all three forms occur) compute what the plan's
    one-pass-per-fused-gate sequence computes; the plan traces themselves are pinned to the reference elsewhere in this file"""
    R = [1, 2, 4, 1][seed % 4]
    nq = 13 + seed % 4 + (R.bit_length() - 1)
    script = scripts.random_script(nq, R, 300 + seed, ngates=150, queries=False)
    engines = _dry_engines(script, R)
    for op in script[1:]:
        if op[0] == "cheat_local":
            continue
        for e in engines:
            getattr(e, op[0])(*op[1:])
    for e in engines:
        e.synchronize()
    info = {"tile_emulator": True}
    launched = scripts.replay_traces([e.launch_trace() for e in engines], R, info)
    planned = scripts.replay_traces([e.trace() for e in engines], R)
    assert np.abs(launched - planned).max() <= 1e-12
    assert info.get("tile_images_emulated", 0) >= 1


@pytest.mark.parametrize("R", [1, 4])
def test_bench_parity_checks_pinned_on_the_oracle(R):
    """bench.py's `parity` object (QFT closed form via get_amplitude, marginals, post-measurement state, random circuit
    followed by its inverse) evaluated on the numpy oracle: the checks themselves hold at 1e-12 on a correct engine"""
    import bench
    from hiqsimulator_b200 import backends
    from oracle import statevec
    n = 12
    L = n - (R.bit_length() - 1)
    res = bench.parity_checks(n, L, lambda: backends.SimulatorMPI(gate_fusion=True, rnd_seed=5, num_local_qubits=L, max_fused_qubits=4,
                                                                  backend_class=lambda s, ml, mc: statevec.SimulatorMPI(s, ml, mc, R)),
                              samples=256)
    assert res["ok"], res
    assert res["qft_closed_form"]["max_abs_err"] <= 1e-12 and res["random_then_inverse"]["abs_amp0_minus_1"] <= 1e-12


def test_shor_period_read_out_is_exact_beyond_53_bits():
    """circuits.period_from_bits on ideal phase-estimation outcomes of the bench's 32-qubit Shor instance
    (N = 32771 * 32779, a = 7, 62 measured bits): the exact rational recovers the order for every k coprime to it;
    the float sum of the reference example (examples/shor_mpi.py:95-105) cannot hold 62 bits and misses some."""
    import math
    import random
    from fractions import Fraction

    from hiqsimulator_b200 import circuits
    p, q, a = 32771, 32779, 7
    N = p * q
    n = int(math.ceil(math.log(N, 2)))
    assert n == 31
    lam = (p - 1) * (q - 1) // math.gcd(p - 1, q - 1)
    r = lam
    for f in range(2, 40000):  # order of a = lambda(N) stripped of every prime factor that is not needed
        while r % f == 0 and pow(a, r // f, N) == 1:
            r //= f
    assert pow(a, r, N) == 1 and r > 1 << 20
    rnd = random.Random(5)
    float_misses = 0
    for _ in range(200):
        k = rnd.randrange(1, r)
        if math.gcd(k, r) != 1:
            continue
        Y = (k * (1 << (2 * n)) + r // 2) // r  # the most likely outcome: the 2n-bit fraction closest to k / r
        bits = [(Y >> j) & 1 for j in range(2 * n)]
        assert circuits.period_from_bits(bits, N) == r
        y = sum(bits[2 * n - 1 - i] * 1.0 / (1 << (i + 1)) for i in range(2 * n))
        float_misses += Fraction(y).limit_denominator(N - 1).denominator != r
    assert float_misses > 0


def test_too_wide_cluster_is_refused_before_the_product_is_formed():
    """Run() on a cluster of more than 5 qubits raises the reference's error (SimulatorMPI.cpp:516-523) WITHOUT first
    multiplying the cluster out — 14 queued one-qubit gates would otherwise cost a 2^14 x 2^14 matrix (4 GiB) per attempt,
    and the cluster stays queued after the error as in the reference, so every later call would pay it again"""
    import time
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(0, 1, b"", 0, M.FLAG_DRY_RUN)
    try:
        e = M.SimulatorMPI(1, 16, 16)
        e.allocate_qureg(list(range(16)), 0)
        h = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
        for q in range(14):
            e.apply_controlled_matrix(h, [q], [])
        t0 = time.perf_counter()
        for _ in range(3):
            with pytest.raises(RuntimeError, match="cannot apply 14 qubits gate"):
                e.run()
        assert time.perf_counter() - t0 < 1.0
    finally:
        M.init_world(0, 1, b"", 0, 0)


def _all_ranks(R, seed, max_local, max_cluster):
    from hiqsimulator_b200 import _cppsim_mpi as M
    es = []
    for r in range(R):
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        es.append(M.SimulatorMPI(seed, max_local, max_cluster))
    M.init_world(0, 1, b"", 0, 0)
    return es


def _outcomes(es, name, *args):
    out = []
    for e in es:
        try:
            getattr(e, name)(*args)
            out.append(None)
        except RuntimeError as ex:
            out.append(str(ex))
    return out


def test_invalid_arguments_are_refused_on_every_rank_alike():
    """A call that raises on some ranks only leaves the others waiting in the next collective.  The reference meets an
    unallocated qubit in Run(), on the ranks whose global-control bits kept the gate (SimulatorMPI.cpp:466-468, after the
    filter of :736-739); this engine refuses at the call, before any rank-dependent decision (found by
    tools/fuzz_invalid_arguments.py --ranks)."""
    es = _all_ranks(4, 1, 4, 3)
    for e in es:
        e.allocate_qureg(list(range(6)), 0)
    loc, glo = es[0].get_local_qubits_ids(), [q for q in es[0].get_global_qubits_ids() if q >= 0]
    assert len(loc) == 4 and len(glo) == 2
    D = np.diag(np.exp(1j * np.arange(4)))
    H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)

    def refused_everywhere(name, *args, match):
        out = _outcomes(es, name, *args)
        assert all(o is not None and match in o for o in out), out

    # unallocated target behind a global control: only the ranks with that control bit set would ever look at the gate
    refused_everywhere("apply_controlled_matrix", H, [42], [glo[0]], match="Can't find 42")
    refused_everywhere("apply_controlled_matrix", H, [loc[0]], [glo[0], 77], match="Can't find 77")
    refused_everywhere("apply_controlled_matrix", H, [-1], [], match="Can't find -1")  # -1 marks an empty global position
    # a qubit twice, or target and control at once
    refused_everywhere("apply_controlled_matrix", D, [loc[0], loc[0]], [], match="must be distinct")
    refused_everywhere("apply_controlled_matrix", H, [loc[1]], [loc[1]], match="must be distinct")
    # wider than the cluster (3) with a global target: the reference queues it as it is and fails in Run() on some ranks
    D4 = np.diag(np.exp(1j * np.arange(16)))
    refused_everywhere("apply_controlled_matrix", D4, [loc[0], loc[1], loc[2], glo[1]], [glo[0]], match="wider than the cluster")
    # allocation of a live or negative id, relabelling that is not a permutation
    refused_everywhere("allocate_qubit", loc[2], match="already allocated")
    refused_everywhere("allocate_qubit", glo[0], match="already allocated")
    refused_everywhere("allocate_qubit", -1, match="non-negative")
    refused_everywhere("set_qubits_perm", [loc[0]] * 4 + glo, match="not a permutation")
    refused_everywhere("set_qubits_perm", loc + [glo[0], 99], match="not a permutation")
    # nothing was queued or changed by the refused calls: the engines still take a valid relabelling and a valid gate
    for e in es:
        assert e.get_local_qubits_ids() == loc
        e.set_qubits_perm(loc[::-1] + glo)
        assert e.get_local_qubits_ids() == loc[::-1]
        e.apply_controlled_matrix(H, [loc[0]], [glo[0]])
        e.run()
    runs = [sum(1 for d in e.trace() if d["kind"] == DESC_DENSE) for e in es]
    assert runs == [0, 1, 0, 1]  # global position 0 <-> rank bit 0


DESC_DENSE, DESC_SWAP = 1, 4  # HIQ_DESC_DENSE / HIQ_DESC_SWAP (include/hiq_b200.h)


def test_gates_waiting_in_the_fusion_go_out_before_a_swap():
    """The reference's wrapper runs before every swap (_simulator_mpi.py:505-507); a direct caller of the class that swaps
    with a gate still queued must get that gate applied to the layout it was given for, on every rank, not an
    ArrayFindSure error in the next Run() on the ranks that kept it."""
    from hiqsimulator_b200 import _cppsim_mpi as M
    es = _all_ranks(2, 1, 4, 3)
    for e in es:
        e.allocate_qureg(list(range(5)), 0)
    loc, glo = es[0].get_local_qubits_ids(), [q for q in es[0].get_global_qubits_ids() if q >= 0]
    H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    for e in es:
        e.clear_trace()
        e.apply_controlled_matrix(H, [loc[-1]], [glo[0]])   # kept by rank 1 only; no run()
        e.swap_qubits([glo[0], loc[-1]])                    # loc[-1] becomes global
        e.run()
    kinds = [[d["kind"] for d in e.trace() if d["kind"] in (DESC_DENSE, DESC_SWAP)] for e in es]
    assert kinds[0] == [DESC_SWAP]
    assert kinds[1] == [DESC_DENSE, DESC_SWAP]
    assert es[1].trace()[0]["slots"] == [len(loc) - 1]

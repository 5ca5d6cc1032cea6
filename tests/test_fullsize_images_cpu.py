"""The tile programs of the BENCH circuits at their real size (33 qubits, 128 GiB slab), checked on the CPU.

A dry-run engine runs the scheduled QFT-33 / random-33 of bench.py and records its launches; every tile program goes
through the launcher's parameter image (hiqk_tile_program_image at L = 33: 64-bit offsets, slots up to 32, select bits and
diagonal factors far outside the tile) and tests/tile_emulator.py on a few sampled tiles — the slab is never materialised,
amplitudes come from a hash of their index.  The expectation is computed independently on the reduced problem of one tile:
the tile's 2^T amplitudes as a T-qubit state, every diagonal factor looked up from the full index, every gate restricted
to its in-tile targets with the outside (select) bits fixed to the tile's values — plain oracle kernels, no launcher logic."""
import copy

import numpy as np
import pytest

import scripts
import tile_emulator
from oracle import statevec


def _source(idx):
    """amplitude of global index idx: a fixed pseudo-random function (splitmix64-style hash)"""
    z = np.asarray(idx, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    re = (z & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24) - 0.5
    im = ((z >> np.uint64(24)) & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24) - 0.5
    return re + 1j * im


def _launch_trace(kind, n, ranks=1, rank=0):
    import bench
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, ops
    cmds = bench.build_circuit(kind, n)
    g = ranks.bit_length() - 1
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=n - g, max_fused_qubits=4,
                               backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, rank, ranks, M.FLAG_DRY_RUN))
    eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
    eng.receive([ops.AllocateQureg(list(range(n)), 0)])
    eng.receive(copy.deepcopy(cmds))
    eng.flush()
    be._simulator.synchronize()
    return be._simulator.launch_trace(), be._simulator.stats()


def _expected_tile(L, steps, S, t):
    """the tile whose number is t (its bits deposited on the slots outside S, ascending), processed by the oracle"""
    T = len(S)
    outside = [s for s in range(L) if s not in S]
    tbase = 0
    for j, s in enumerate(outside):
        tbase |= ((t >> j) & 1) << s
    i = np.arange(1 << T, dtype=np.uint64)
    glob = np.full(1 << T, tbase, dtype=np.uint64)
    for b, s in enumerate(S):
        glob |= ((i >> np.uint64(b)) & np.uint64(1)) << np.uint64(s)
    x = _source(glob)
    for slots, m, ops in steps:
        for sl, table in ops:
            sel = np.zeros(1 << T, dtype=np.int64)
            for l, s in enumerate(sl):
                sel |= ((glob >> np.uint64(s)) & np.uint64(1)).astype(np.int64) << l
            x = x * np.asarray(table)[sel]
        k = len(slots)
        inside = [l for l in range(k) if slots[l] in S]
        fixed = sum((((tbase >> slots[l]) & 1) << l) for l in range(k) if slots[l] not in S)
        out_mask = sum(1 << l for l in range(k) if slots[l] not in S)
        full = np.arange(1 << k)
        rows_in = full[(full & out_mask) == fixed]
        # an outside bit must be a select bit: no entry couples its two values
        assert np.all(m[np.ix_(full[(full & out_mask) != fixed], rows_in)] == 0), "a mixing bit of the run lies outside the tile"
        red = np.zeros((1 << len(inside), 1 << len(inside)), dtype=np.complex128)
        for b in range(1 << len(inside)):
            for c in range(1 << len(inside)):
                bf = fixed | sum(((b >> j) & 1) << inside[j] for j in range(len(inside)))
                cf = fixed | sum(((c >> j) & 1) << inside[j] for j in range(len(inside)))
                red[b, c] = m[bf, cf]
        if inside:
            statevec.apply_dense(x, [S.index(slots[l]) for l in inside], red, 0)
        else:
            x = x * red[0, 0]
    return glob, x


@pytest.mark.parametrize("kind,n,ranks,rank", [("qft", 33, 1, 0), ("random", 33, 1, 0), ("random", 34, 2, 1), ("random", 35, 4, 2), ("random", 35, 8, 0),
                                               ("random", 35, 8, 5), ("qft", 35, 8, 7)])
def test_bench_tile_programs_at_full_size(kind, n, ranks, rank):
    """the workloads of bench.py at N = 1, 2, 4, 8 (and the QFT of its parity object), one rank's launches each"""
    from hiqsimulator_b200 import kernels as K
    L = n - (ranks.bit_length() - 1)
    trace, st = _launch_trace(kind, n, ranks, rank)
    runs = [d for d in trace if d["kind"] == scripts.KIND["launch"] and d["form"] == scripts.LAUNCH_TILE]
    assert len(runs) == st["tile_launches"] and len(runs) >= 3
    rng = np.random.default_rng(n + rank)
    for d in runs:
        steps = scripts.decode_launch(d)
        raw = K.tile_program_image(L, steps)
        im = tile_emulator.Image(raw)
        T = im.T
        S = [int(x) for x in im.p("tslot", np.uint8, 16)[:T]]
        n_tiles = 1 << (L - T)
        tiles = [0, n_tiles - 1] + [int(x) for x in rng.integers(0, n_tiles, size=3)]
        idx, val = tile_emulator.run_image(raw, None, tiles=tiles, source=_source)
        for row, t in enumerate(tiles):
            glob, want = _expected_tile(L, steps, S, t)
            order = np.argsort(idx[row])
            assert np.array_equal(idx[row][order], np.sort(glob)), "the launcher's tile does not cover the amplitudes of tile %d" % t
            got = val[row][order]
            exp = want[np.argsort(glob)]
            assert np.abs(got - exp).max() <= 1e-12


def _expected_amplitudes(idx, slots, m, ctrl_mask):
    """out[i] = sum_c m[b(i)][c] * in[i with the target bits spelling c] where the control mask is satisfied, in[i] elsewhere"""
    idx = np.asarray(idx, dtype=np.uint64)
    k = len(slots)
    tmask = np.uint64(sum(1 << s for s in slots))
    b = np.zeros(idx.shape, dtype=np.int64)
    for l, s in enumerate(slots):
        b |= ((idx >> np.uint64(s)) & np.uint64(1)).astype(np.int64) << l
    base = idx & ~tmask
    out = np.zeros(idx.shape, dtype=np.complex128)
    for c in range(1 << k):
        off = np.uint64(sum(((c >> l) & 1) << slots[l] for l in range(k)))
        out += m[b, c] * _source(base | off)
    # the tile kernel also stores the amplitudes of its tile that fail a low control bit: unchanged
    return np.where((idx & np.uint64(ctrl_mask)) == np.uint64(ctrl_mask), out, _source(idx))


@pytest.mark.parametrize("kind,n,ranks,rank", [("random", 33, 1, 0), ("random", 35, 4, 1), ("random", 35, 8, 3), ("qft", 33, 1, 0)])
def test_bench_single_gate_launches_at_full_size(kind, n, ranks, rank):
    """every single-gate dense launch of the bench circuits (the kernels a random circuit spends its time in: DIRECT with the
    three-multiplication product, the tensor-core kernel for slot-0 targets) through hiqk_dense_image at the real slab
    size and the kernel emulator on sampled work items, against the defining sum evaluated per stored amplitude"""
    import dense_emulator
    from hiqsimulator_b200 import kernels as K
    L = n - (ranks.bit_length() - 1)
    trace, st = _launch_trace(kind, n, ranks, rank)
    gates = [d for d in trace if d["kind"] == scripts.KIND["dense"]]
    if kind == "random":
        assert len(gates) >= 80
    rng = np.random.default_rng(n + rank)
    variants = set()
    for d in gates:
        slots = [int(s) for s in d["slots"]]
        k = len(slots)
        m = np.asarray(d["payload"]).reshape(1 << k, 1 << k)
        cm = int(d["ctrl_mask"])
        raw = K.dense_image(L, slots, m, cm)
        stats = {}
        nc = bin(cm).count("1")
        items = 1 << (L - k - nc)
        if raw[4] == dense_emulator.DMMA:
            items >>= 3
        elif raw[4] == dense_emulator.TILED:
            items = None
        sample = [0, 1] if items is None else [0, items - 1] + [int(x) for x in rng.integers(0, items, size=30)]
        idx, val = dense_emulator.run_dense_image(raw, None, stats, sample=sample, source=_source)
        variants.add(stats["variant"])
        assert np.unique(idx).size == idx.size
        assert np.abs(val - _expected_amplitudes(idx, slots, m, cm)).max() <= 1e-12
    if kind == "random":
        assert {dense_emulator.DIRECT, dense_emulator.DMMA} <= variants

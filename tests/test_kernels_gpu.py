"""Parity of the device-level launchers (C ABI, hiqk_*) against the numpy oracle.

Every case runs one kernel pass on cuda:0 through libhiq_b200.so and compares the whole slab
with oracle/statevec.py (itself pinned against the compiled reference).  Tolerance 1e-12
absolute on amplitudes (BASELINE.json north_star); reductions 1e-12 absolute.
"""
import itertools

import numpy as np
import pytest

from oracle import statevec

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _torch():
    import torch
    return torch


def rand_state(L, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << L) + 1j * rng.normal(size=1 << L)
    return v / np.linalg.norm(v)


def rand_matrix(k, seed):
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def dense_cases():
    from hiqsimulator_b200 import kernels as K
    cases = []
    L = 14
    slot_sets = {
        1: [[0], [1], [2], [5], [13]],
        2: [[0, 1], [1, 0], [3, 9], [13, 2], [12, 13]],
        3: [[0, 1, 2], [2, 0, 7], [4, 5, 6], [13, 6, 1], [11, 12, 13]],
        4: [[0, 1, 2, 3], [3, 1, 2, 0], [5, 2, 9, 12], [10, 11, 12, 13], [13, 0, 6, 3], [4, 8, 6, 10]],
        5: [[0, 1, 2, 3, 4], [9, 10, 11, 12, 13], [1, 5, 3, 12, 8], [2, 4, 6, 8, 10]],
    }
    for k, sets in slot_sets.items():
        for slots in sets:
            for variant in (K.AUTO, K.DIRECT, K.TILED, K.DMMA):
                for ctrl in (0, "one", "three"):
                    cases.append((L, k, tuple(slots), variant, ctrl))
    return cases


def _ctrl_mask(L, slots, kind, seed):
    if kind == 0:
        return 0
    free = [s for s in range(L) if s not in slots]
    rng = np.random.default_rng(seed)
    n = 1 if kind == "one" else 3
    pick = rng.choice(free, size=n, replace=False)
    m = 0
    for s in pick:
        m |= 1 << int(s)
    return m


@pytest.mark.parametrize("L,k,slots,variant,ctrl", dense_cases())
def test_apply_dense(L, k, slots, variant, ctrl):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    seed = hash((L, k, slots, str(ctrl))) & 0xFFFF
    ref = rand_state(L, seed)
    m = rand_matrix(k, seed + 1)
    cm = _ctrl_mask(L, slots, ctrl, seed)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_dense(dev, list(slots), m, cm, variant)
    statevec.apply_dense(ref, list(slots), m, cm)
    got = dev.cpu().numpy()
    assert np.abs(got - ref).max() <= TOL


@pytest.mark.parametrize("L", [3, 5, 6, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_apply_dense_small_slabs(L, k):
    if k > L:
        pytest.skip("k > L")
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    rng = np.random.default_rng(L * 10 + k)
    for trial in range(4):
        slots = [int(s) for s in rng.choice(L, size=k, replace=False)]
        for variant in (K.AUTO, K.DMMA):
            ref = rand_state(L, trial)
            m = rand_matrix(k, trial + 7)
            dev = torch.from_numpy(ref.copy()).cuda()
            K.apply_dense(dev, slots, m, 0, variant)
            statevec.apply_dense(ref, slots, m, 0)
            assert np.abs(dev.cpu().numpy() - ref).max() <= TOL, (slots, variant)


def test_apply_dense_many_controls():
    """Huge-gate path: k=1 with a wide control mask (reference: SimulatorMPI.cpp:752-756)."""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 16
    ref = rand_state(L, 3)
    m = rand_matrix(1, 4)
    cm = ((1 << L) - 1) & ~(1 << 7)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_dense(dev, [7], m, cm)
    statevec.apply_dense(ref, [7], m, cm)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("k,slots,ctrl", [(1, (0,), 0), (1, (9,), "one"), (2, (3, 0), 0), (3, (13, 1, 6), "three"),
                                          (4, (2, 3, 4, 5), 0), (5, (0, 13, 5, 7, 2), "one")])
def test_apply_diag(k, slots, ctrl):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    ref = rand_state(L, 11 + k)
    rng = np.random.default_rng(k)
    d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k))
    cm = _ctrl_mask(L, slots, ctrl, 5)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_diag(dev, list(slots), d, cm)
    statevec.apply_diag(ref, list(slots), d, cm)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


def _rand_diag_ops(L, n_ops, seed, avoid=()):
    rng = np.random.default_rng(seed)
    ops = []
    for j in range(n_ops):
        k = int(rng.integers(0, 6))
        slots = [int(x) for x in rng.choice(np.arange(L), size=k, replace=False)]
        d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k)) * rng.uniform(0.5, 1.5)
        ops.append((slots, d))
    return ops


@pytest.mark.parametrize("L,n_ops,seed", [(14, 1, 0), (14, 5, 1), (15, 16, 2), (9, 7, 3), (5, 3, 4), (13, 12, 5)])
def test_apply_diag_batch(L, n_ops, seed):
    """one pass == the reference's n_ops successive kernel_core_diag passes"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    ref = rand_state(L, 40 + seed)
    ops = _rand_diag_ops(L, n_ops, seed)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_diag_batch(dev, ops)
    for slots, d in ops:
        if slots:
            statevec.apply_diag(ref, slots, d, 0)
        else:
            ref *= d[0]
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("k,slots", [(1, (2,)), (1, (13,)), (2, (3, 9)), (2, (12, 2)), (3, (4, 5, 6)), (3, (13, 6, 2)),
                                     (4, (5, 2, 9, 12)), (4, (10, 11, 12, 13)), (4, (4, 8, 6, 10)),
                                     (1, (1,)), (3, (1, 6, 9)), (3, (9, 1, 2)), (4, (1, 5, 9, 12)), (4, (4, 3, 2, 1))])
@pytest.mark.parametrize("n_pre", [1, 4, 16])
def test_apply_dense_prediag(k, slots, n_pre):
    """one pass == n_pre diagonal passes followed by the dense pass (overlapping and disjoint slots)"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    assert K.dense_prediag_supported(L, slots)
    ref = rand_state(L, 70 + k)
    m = rand_matrix(k, 3 * k + n_pre)
    ops = _rand_diag_ops(L, n_pre, 100 * k + n_pre)
    # make sure one op lies entirely inside the targets and one entirely outside
    ops[0] = (list(slots[:max(1, k - 1)]), np.exp(1j * np.linspace(0.1, 2.0, 1 << max(1, k - 1))))
    if n_pre > 1:
        outside = [s for s in range(L) if s not in slots][:3]
        ops[1] = (outside, np.exp(1j * np.linspace(0.3, 3.0, 8)))
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_dense_prediag(dev, list(slots), m, ops)
    for sl, d in ops:
        if sl:
            statevec.apply_diag(ref, sl, d, 0)
        else:
            ref *= d[0]
    statevec.apply_dense(ref, list(slots), m, 0)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


def multiplexed_matrix(k, select_bits, seed):
    """unitary that is block diagonal in `select_bits`: one Haar block on the other bits per select value"""
    mix = [l for l in range(k) if l not in select_bits]
    m = np.zeros((1 << k, 1 << k), dtype=np.complex128)
    for v in range(1 << len(select_bits)):
        u = rand_matrix(len(mix), seed + 31 * v) if mix else np.array([[np.exp(1j * (seed + v))]])
        hi = sum(((v >> i) & 1) << select_bits[i] for i in range(len(select_bits)))
        for b in range(1 << len(mix)):
            for c in range(1 << len(mix)):
                fb = hi | sum(((b >> i) & 1) << mix[i] for i in range(len(mix)))
                fc = hi | sum(((c >> i) & 1) << mix[i] for i in range(len(mix)))
                m[fb, fc] = u[b, c]
    return m


BLOCK_CASES = [(2, (3, 9), (0,)), (2, (12, 2), (1,)), (3, (4, 5, 6), (1,)), (3, (13, 6, 2), (0, 2)), (3, (2, 9, 5), (2,)),
               (4, (5, 2, 9, 12), (0,)), (4, (5, 2, 9, 12), (1, 3)), (4, (10, 11, 12, 13), (0, 1, 2)), (4, (4, 8, 6, 10), (2,)),
               (4, (13, 3, 7, 2), (0, 3)), (4, (6, 7, 8, 9), (0, 1, 2, 3))]


@pytest.mark.parametrize("k,slots,select", BLOCK_CASES)
@pytest.mark.parametrize("ctrl", [0, "one"])
def test_apply_dense_block_structure(k, slots, select, ctrl):
    """fused gates that are block diagonal in some index bits take the reduced product: same result as the
    full product (DIRECT_FULL) and as the oracle, wherever the select bits sit in the matrix index"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    m = multiplexed_matrix(k, list(select), 7 * k + len(select))
    ks, order = K.dense_block_shape(m)
    assert ks == max(1, k - len(select)) and sorted(order) == list(range(k))
    ref = rand_state(L, 200 + k)
    cm = _ctrl_mask(L, slots, ctrl, 9)
    exp = ref.copy()
    statevec.apply_dense(exp, list(slots), m, cm)
    for variant in (K.AUTO, K.DIRECT, K.DIRECT_FULL):
        dev = torch.from_numpy(ref.copy()).cuda()
        K.apply_dense(dev, list(slots), m, cm, variant)
        assert np.abs(dev.cpu().numpy() - exp).max() <= TOL, variant


@pytest.mark.parametrize("k,slots,select", [c for c in BLOCK_CASES if c[0] >= 2])
@pytest.mark.parametrize("n_pre", [1, 6])
def test_apply_dense_prediag_block_structure(k, slots, select, n_pre):
    """folded diagonals + block-structured matrix: the class-E tables follow the permuted target order"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 15
    m = multiplexed_matrix(k, list(select), 11 * k + len(select))
    ref = rand_state(L, 300 + k)
    ops = _rand_diag_ops(L, n_pre, 500 * k + n_pre)
    ops[0] = (list(slots[:max(1, k - 1)]), np.exp(1j * np.linspace(0.1, 2.0, 1 << max(1, k - 1))))
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_dense_prediag(dev, list(slots), m, ops)
    for sl, d in ops:
        if sl:
            statevec.apply_diag(ref, sl, d, 0)
        else:
            ref *= d[0]
    statevec.apply_dense(ref, list(slots), m, 0)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


def _tile_reference(ref, steps):
    for slots, m, pre in steps:
        for sl, d in pre:
            if sl:
                statevec.apply_diag(ref, sl, d, 0)
            else:
                ref *= d[0]
        statevec.apply_dense(ref, list(slots), m, 0)


TILE_RUNS = [
    # (L, [(k, slots, select bits of the matrix or None = dense, number of diagonal ops)], expected tile bits)
    (15, [(4, (11, 12, 13, 14), None, 0)], 11),
    (15, [(4, (11, 12, 13, 14), (2, 3), 3), (4, (9, 10, 13, 14), (0, 1), 5), (4, (9, 10, 11, 12), (2, 3), 12)], 11),   # 6 high slots
    (16, [(4, (12, 13, 14, 15), (0, 1), 2), (4, (10, 11, 12, 13), (2, 3), 4), (4, (8, 9, 10, 11), (2, 3), 6)], 11),     # QFT-like chain
    (20, [(4, (16, 17, 18, 19), (2, 3), 5), (4, (14, 15, 16, 17), (2, 3), 7), (4, (12, 13, 14, 15), (2, 3), 6), (4, (10, 11, 12, 13), (2, 3), 4)], 12),
    (19, [(4, (3, 9, 14, 18), (0, 2, 3), 3), (3, (17, 5, 11), (0, 1), 2), (2, (18, 16), (0,), 2), (4, (2, 8, 13, 17), (1, 2, 3), 9)], 11),  # select bits outside the tile
    (16, [(4, (7, 8, 9, 10), (2, 3), 2), (4, (5, 6, 7, 8), (2, 3), 3), (4, (3, 4, 5, 6), (2, 3), 4), (4, (1, 2, 3, 4), (2, 3), 5)], 11),
    (14, [(3, (0, 1, 2), (1, 2), 6), (4, (1, 2, 3, 4), None, 2)], 11),                                                  # slot-0 targets
    (15, [(4, (0, 5, 9, 13), None, 1), (4, (2, 6, 9, 12), None, 2)], 11),                                               # two full products
    (17, [(2, (3, 16), None, 1), (1, (14,), None, 0), (3, (15, 0, 7), (1,), 4), (4, (12, 13, 14, 15), (0,), 7)], 11),
    (16, [(4, (0, 1, 14, 15), None, 12), (4, (2, 3, 12, 13), None, 9)], 11),
    (18, [(4, (10, 11, 16, 17), None, 16), (4, (12, 13, 14, 15), (1, 2), 16)], 11),
]


def monomial_matrix(k, seed):
    """a permutation of the basis states with random phases (what fused X / Y / Z / phase gates make)"""
    rng = np.random.default_rng(seed)
    d = 1 << k
    m = np.zeros((d, d), dtype=np.complex128)
    m[rng.permutation(d), np.arange(d)] = np.exp(1j * rng.uniform(0, 2 * np.pi, size=d))
    return m


TILE_RUNS += [
    (16, [(4, (9, 11, 12, 14), "mono", 3), (4, (12, 13, 14, 15), (2, 3), 4), (3, (10, 4, 13), "mono", 6), (4, (1, 3, 4, 6), "mono", 0)], 12),
    (17, [(4, (13, 14, 15, 16), "mono", 0), (4, (5, 9, 12, 16), None, 5), (2, (0, 16), "mono", 2)], 11),
]


def tile_case(case):
    """(L, steps, expected tile bits) of TILE_RUNS[case]; tests/test_tile_program_cpu.py runs the same cases on the CPU"""
    L, run, tile_bits = TILE_RUNS[case]
    steps = []
    for i, (k, slots, select, n_pre) in enumerate(run):
        if select is None:
            m = rand_matrix(k, 17 * case + i)
        elif select == "mono":
            m = monomial_matrix(k, 19 * case + i)
        else:
            m = multiplexed_matrix(k, list(select), 13 * case + i)
        ops = _rand_diag_ops(L, n_pre, 1000 * case + i)
        ops = [(sl[:4], d[:1 << len(sl[:4])]) for sl, d in ops]  # tables of at most 16 entries (cluster size 4)
        if n_pre:  # one op entirely inside the targets, one overlapping them partly
            ops[0] = (list(slots[:max(1, k - 1)]), np.exp(1j * np.linspace(0.1, 2.0, 1 << max(1, k - 1))))
        if n_pre > 1:
            other = [s for s in range(L) if s not in slots]
            ops[1] = ([slots[0], other[0], other[-1]], np.exp(1j * np.linspace(0.3, 3.0, 8)))
        steps.append((slots, m, ops))
    return L, steps, tile_bits


@pytest.mark.parametrize("case", range(len(TILE_RUNS)))
def test_tile_program_matches_oracle(case):
    """a run of dense gates, each preceded by diagonal factors, in ONE tile-resident pass == the reference's
    one-sweep-per-fused-gate sequence (numpy oracle), every target / select / class-E configuration"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L, steps, tile_bits = tile_case(case)
    assert K.tile_program_fits(L, steps) == tile_bits
    ref = rand_state(L, 800 + case)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_tile_program(dev, steps)
    _tile_reference(ref, steps)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


def test_tile_program_multi_iteration(tiny_grid):
    """more tiles than CTAs: every CTA walks several tiles (per-tile selectors, tile reuse)"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 18
    steps = []
    for i, slots in enumerate([(13, 14, 15, 16), (11, 12, 13, 14), (9, 10, 11, 12)]):
        ops = _rand_diag_ops(L, 5, 40 + i)
        ops = [(sl[:4], d[:1 << len(sl[:4])]) for sl, d in ops]
        ops[0] = ([slots[1], 17, 2], np.exp(1j * np.linspace(0.2, 2.5, 8)))
        steps.append((slots, multiplexed_matrix(4, [2, 3], 50 + i), ops))
    ref = rand_state(L, 77)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_tile_program(dev, steps)
    K.apply_tile_program(dev, steps[:2])
    _tile_reference(ref, steps)
    _tile_reference(ref, steps[:2])
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


def test_tile_program_rejects_what_does_not_fit():
    from hiqsimulator_b200 import kernels as K
    m = rand_matrix(4, 1)
    wide = [((20, 21, 22, 23), m, []), ((10, 11, 12, 13), m, []), ((15, 16, 17, 18), m, [])]  # 12 high mixing slots
    assert K.tile_program_fits(26, wide) == 0
    assert K.tile_program_fits(26, wide[:2]) == 12
    assert K.tile_program_fits(10, [((0, 1, 2, 3), m, [])]) == 0  # slab smaller than a tile


def test_apply_dense_block_structure_multi_iteration(tiny_grid):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 17
    for k, slots, select in [(4, (5, 2, 9, 12), (1, 3)), (3, (13, 6, 2), (0,)), (4, (12, 13, 14, 15), (0, 1, 2))]:
        m = multiplexed_matrix(k, list(select), 5 * k)
        ref = rand_state(L, 400 + k)
        dev = torch.from_numpy(ref.copy()).cuda()
        K.apply_dense(dev, list(slots), m, 0, K.AUTO)
        ops = _rand_diag_ops(L, 5, 77 * k)
        K.apply_dense_prediag(dev, list(slots), m, ops)
        statevec.apply_dense(ref, list(slots), m, 0)
        for sl, d in ops:
            if sl:
                statevec.apply_diag(ref, sl, d, 0)
            else:
                ref *= d[0]
        statevec.apply_dense(ref, list(slots), m, 0)
        assert np.abs(dev.cpu().numpy() - ref).max() <= TOL, (k, slots, select)


@pytest.fixture
def tiny_grid():
    """3 CTAs only: every persistent kernel iterates many times per CTA (the path 2^30+ slabs take)"""
    from hiqsimulator_b200 import kernels as K
    K.debug_set_max_grid(3)
    yield
    K.debug_set_max_grid(0)


@pytest.mark.parametrize("k,slots,ctrl", [(3, (4, 5, 6), 0), (3, (13, 6, 2), "one"), (4, (5, 2, 9, 12), 0), (4, (10, 11, 12, 13), "three"),
                                          (4, (4, 8, 6, 10), "one"), (2, (3, 9), 0), (5, (9, 10, 11, 12, 13), 0)])
def test_apply_dense_multi_iteration(tiny_grid, k, slots, ctrl):
    """grid-stride / cp.async-staged paths: each thread processes many tuples"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 16
    ref = rand_state(L, 90 + k)
    m = rand_matrix(k, 17 * k)
    cm = _ctrl_mask(L, slots, ctrl, 3)
    for variant in (K.AUTO, K.DIRECT, K.TILED):
        dev = torch.from_numpy(ref.copy()).cuda()
        K.apply_dense(dev, list(slots), m, cm, variant)
        exp = ref.copy()
        statevec.apply_dense(exp, list(slots), m, cm)
        assert np.abs(dev.cpu().numpy() - exp).max() <= TOL, variant


@pytest.mark.parametrize("k,slots", [(2, (3, 9)), (3, (13, 6, 2)), (4, (5, 2, 9, 12)), (4, (12, 13, 14, 15)), (4, (4, 8, 6, 10))])
@pytest.mark.parametrize("n_pre", [1, 5, 16])
def test_apply_dense_prediag_multi_iteration(tiny_grid, k, slots, n_pre):
    """folded diagonals with many chunks per CTA (per-chunk hoisting + prefetch across chunk boundaries)"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 17
    ref = rand_state(L, 170 + k)
    m = rand_matrix(k, 5 * k + n_pre)
    ops = _rand_diag_ops(L, n_pre, 300 * k + n_pre)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.apply_dense_prediag(dev, list(slots), m, ops)
    K.apply_diag_batch(dev, ops[:max(1, n_pre // 2)])
    ref2 = ref
    for sl, d in ops:
        if sl:
            statevec.apply_diag(ref2, sl, d, 0)
        else:
            ref2 *= d[0]
    statevec.apply_dense(ref2, list(slots), m, 0)
    for sl, d in ops[:max(1, n_pre // 2)]:
        if sl:
            statevec.apply_diag(ref2, sl, d, 0)
        else:
            ref2 *= d[0]
    assert np.abs(dev.cpu().numpy() - ref2).max() <= TOL


def test_streaming_kernels_multi_iteration(tiny_grid):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 15
    ref = rand_state(L, 5)
    dev = torch.from_numpy(ref.copy()).cuda()
    d = np.exp(1j * np.linspace(0, 3, 8))
    K.apply_diag(dev, [1, 7, 12], d, 1 << 4)
    statevec.apply_diag(ref, [1, 7, 12], d, 1 << 4)
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL
    assert abs(K.prob_masked(dev, 0b101, 0b001) - float((np.abs(ref[(np.arange(1 << L) & 0b101) == 1]) ** 2).sum())) <= TOL


def test_dense_prediag_rejects_low_slots():
    from hiqsimulator_b200 import kernels as K
    assert not K.dense_prediag_supported(14, (0, 5, 6, 7))
    assert not K.dense_prediag_supported(14, (1, 2, 3, 4, 5))


def test_scale_fill_collapse():
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 13
    ref = rand_state(L, 21)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.scale(dev, 0.3 - 0.4j)
    assert np.abs(dev.cpu().numpy() - ref * (0.3 - 0.4j)).max() <= TOL
    dev = torch.from_numpy(ref.copy()).cuda()
    mask, val = 0b1010010, 0b1000010
    K.collapse(dev, mask, val, 1.7)
    i = np.arange(1 << L)
    exp = np.where((i & mask) == val, ref * 1.7, 0)
    assert np.abs(dev.cpu().numpy() - exp).max() <= TOL
    K.fill(dev, 16, 100, 0.25 + 0.5j)
    got = dev.cpu().numpy()
    assert np.all(got[16:116] == 0.25 + 0.5j) and np.abs(got[:16] - exp[:16]).max() == 0


@pytest.mark.parametrize("L", [1, 4, 9, 14, 20])
def test_reductions(L):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    ref = rand_state(L, 31 + L)
    dev = torch.from_numpy(ref.copy()).cuda()
    p = np.abs(ref) ** 2
    assert abs(K.prob_masked(dev) - p.sum()) <= TOL
    i = np.arange(1 << L)
    for mask, val in [(1, 1), (1, 0), ((1 << L) - 1, 5 % (1 << L)), (0b101 & ((1 << L) - 1), 0b100 & ((1 << L) - 1))]:
        assert abs(K.prob_masked(dev, mask, val) - p[(i & mask) == val].sum()) <= TOL
    for slot in {0, L // 2, L - 1}:
        got = K.bit_norms(dev, slot)
        bit = (i >> slot) & 1
        assert abs(got[0] - p[bit == 0].sum()) <= TOL and abs(got[1] - p[bit == 1].sum()) <= TOL
    e = (p[p > 0] * np.log2(p[p > 0])).sum()
    assert abs(K.entropy_sum(dev) - e) <= 1e-10
    for nb in {1, 2, min(1 << L, 1 << 15)}:
        if nb > (1 << L):
            continue
        got = K.block_norms(dev, nb)
        assert np.abs(got - p.reshape(nb, -1).sum(axis=1)).max() <= TOL


@pytest.mark.parametrize("slot,keep", [(0, 0), (0, 1), (5, 1), (11, 0), (12, 1)])
def test_compact_bit(slot, keep):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 13
    ref = rand_state(L, 41)
    dev = torch.from_numpy(ref.copy()).cuda()
    scratch = torch.empty(300, dtype=torch.complex128, device="cuda")  # forces many chunks
    K.compact_bit(dev, slot, keep, scratch)
    exp = ref.reshape(-1, 2, 1 << slot)[:, keep, :].reshape(-1)
    assert np.abs(dev.cpu().numpy()[: 1 << (L - 1)] - exp).max() == 0


@pytest.mark.parametrize("slots", [(0,), (12,), (3, 7), (0, 1, 2), (11, 4, 9)])
def test_swap_pack_unpack(slots):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 13
    q = len(slots)
    ref = rand_state(L, 51)
    dev = torch.from_numpy(ref.copy()).cuda()
    n = 1 << (L - q)
    srt = sorted(slots)
    i = np.arange(1 << L)
    for pat in range(1 << q):
        sel = np.ones(1 << L, dtype=bool)
        for b, s in enumerate(srt):
            sel &= ((i >> s) & 1) == ((pat >> b) & 1)
        buf = torch.zeros(n, dtype=torch.complex128, device="cuda")
        K.swap_pack(dev, list(slots), pat, 0, n, buf)
        assert np.abs(buf.cpu().numpy() - ref[sel]).max() == 0
        # unpack pattern `pat` data into the positions of the complementary pattern
        other = pat ^ ((1 << q) - 1)
        K.swap_unpack(dev, list(slots), other, 0, n, buf)
        sel2 = np.ones(1 << L, dtype=bool)
        for b, s in enumerate(srt):
            sel2 &= ((i >> s) & 1) == ((other >> b) & 1)
        got = dev.cpu().numpy()
        assert np.abs(got[sel2] - ref[sel]).max() == 0
        dev = torch.from_numpy(ref.copy()).cuda()


@pytest.mark.parametrize("slots", [(0,), (2, 0), (1, 9), (0, 1, 2), (11, 4, 1)])
@pytest.mark.parametrize("piece", [None, 300])
def test_swap_move_all_peers(slots, piece):
    """hiqk_swap_move: every peer's piece gathered / scattered by one launch == numpy selection (bit-exact), whole range
    and an odd sub-range"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    q = len(slots)
    ref = rand_state(L, 52)
    dev = torch.from_numpy(ref.copy()).cuda()
    n = 1 << (L - q)
    begin, count = (0, n) if piece is None else (37, piece)
    srt = sorted(slots)
    i = np.arange(1 << L)
    pats = list(range(1, 1 << q))  # this GPU has pattern 0; peer k has pattern k + 1

    def select(pat):
        sel = np.ones(1 << L, dtype=bool)
        for b, s in enumerate(srt):
            sel &= ((i >> s) & 1) == ((pat >> b) & 1)
        return sel
    bufs = [torch.zeros(count, dtype=torch.complex128, device="cuda") for _ in pats]
    K.swap_move(dev, list(slots), pats, begin, count, bufs, True)
    for pat, buf in zip(pats, bufs):
        assert np.array_equal(buf.cpu().numpy(), ref[select(pat)][begin:begin + count])
    # scatter different data back into the same positions
    new = [torch.from_numpy(rand_state(L, 60 + pat)[:count].copy()).cuda() for pat in pats]
    K.swap_move(dev, list(slots), pats, begin, count, new, False)
    got = dev.cpu().numpy()
    exp = ref.copy()
    for pat, buf in zip(pats, new):
        pos = np.nonzero(select(pat))[0][begin:begin + count]
        exp[pos] = buf.cpu().numpy()
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("slots", [(13,), (0,), (5, 12), (1, 0), (11, 3, 7), (0, 1, 2), (12, 13, 11)])
@pytest.mark.parametrize("grid", [0, 2])
def test_swap_p2p_in_place(slots, grid):
    """the peer-mapped in-place exchange == transposing index bit (L + gpos_j) with bit slots[j]
    (virtual ranks = separate buffers on one GPU; on a multi-GPU box the peers are NVLink-mapped slabs)"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L, q = 14, len(slots)
    R = 1 << q
    rng = np.random.default_rng(sum(slots) + q)
    full = rng.normal(size=R << L) + 1j * rng.normal(size=R << L)
    ranks = [torch.from_numpy(full[r << L:(r + 1) << L].copy()).cuda() for r in range(R)]
    order = sorted(range(q), key=lambda j: slots[j])  # pair indices by ascending slot

    def pat(r):  # bit j' of the pattern <-> j'-th lowest swapped slot <-> gpos of that pair
        return sum(((r >> order[j]) & 1) << j for j in range(q))
    n = 1 << (L - q)
    half = (n + 1) // 2
    K.debug_set_max_grid(grid)
    try:
        for r in range(R):
            peers = [p for p in range(R) if p != r]
            begins = [0 if r < p else half for p in peers]
            counts = [half if r < p else n - half for p in peers]
            K.swap_p2p(ranks[r], [ranks[p] for p in peers], list(slots), [pat(p) for p in peers], pat(r), begins, counts)
        torch.cuda.synchronize()
    finally:
        K.debug_set_max_grid(0)
    got = np.concatenate([t.cpu().numpy() for t in ranks])
    idx = np.arange(R << L, dtype=np.int64)
    src = idx.copy()
    for j in range(q):
        hi, lo = L + j, slots[j]
        bh, bl = (src >> hi) & 1, (src >> lo) & 1
        src = src & ~((1 << hi) | (1 << lo)) | (bl << hi) | (bh << lo)
    assert np.array_equal(got, full[src])


# ---------------------------------------------------------------------------------------------
# Operator-level passes (hiqk_pauli_expect / hiqk_pauli_apply / hiqk_permute_gather) against the
# kernel-level statements in oracle/statevec.py
# ---------------------------------------------------------------------------------------------
def _terms(L, n, seed):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(0, 1 << L)), complex(rng.normal(), rng.normal())) for _ in range(n)]


PAULI_XMASKS = [0, 1, 2, 1 << 13, (1 << 13) | 1, 0b1010001, (1 << 12) | (1 << 5) | 6, (1 << 14) - 1]


@pytest.mark.parametrize("xmask", PAULI_XMASKS)
@pytest.mark.parametrize("n_terms", [1, 3, 64])
@pytest.mark.parametrize("grid", [0, 3])
def test_pauli_expect_local(xmask, n_terms, grid):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    ref = rand_state(L, 7 + xmask % 97)
    terms = _terms(L, n_terms, xmask + n_terms)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.debug_set_max_grid(grid)
    try:
        got = K.pauli_expect(dev, xmask, terms)
    finally:
        K.debug_set_max_grid(0)
    exp = statevec.pauli_expect(ref, xmask, terms)
    assert abs(got - exp) <= TOL
    assert np.array_equal(dev.cpu().numpy(), ref)  # read-only


@pytest.mark.parametrize("xmask", [0, 5, (1 << 13) | 2])
@pytest.mark.parametrize("begin,count", [(0, 1 << 14), (0, 1 << 12), (3 << 12, 1 << 12), (5 << 10, 3 << 10), (100, 0)])
def test_pauli_expect_ranges_and_staged_source(xmask, begin, count):
    """a partner GPU's amplitudes staged in a separate buffer (src) and sub-ranges of the local slab"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    ref = rand_state(L, 21)
    other = rand_state(L, 22)
    terms = _terms(L, 4, 9)
    dev = torch.from_numpy(ref.copy()).cuda()
    src = torch.from_numpy(other[begin:begin + count].copy()).cuda() if count else torch.zeros(1, dtype=torch.complex128, device="cuda")
    got = K.pauli_expect(dev, xmask, terms, src=src, begin=begin, count=count)
    exp = statevec.pauli_expect(ref, xmask, terms, src=other[begin:begin + count], begin=begin, count=count)
    assert abs(got - exp) <= TOL
    got = K.pauli_expect(dev, xmask, terms, begin=begin, count=count)
    exp = statevec.pauli_expect(ref, xmask, terms, begin=begin, count=count)
    assert abs(got - exp) <= TOL


@pytest.mark.parametrize("xmask", PAULI_XMASKS)
@pytest.mark.parametrize("n_terms", [1, 5, 64])
@pytest.mark.parametrize("grid", [0, 3])
def test_pauli_apply_in_place(xmask, n_terms, grid):
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    ref = rand_state(L, 31 + xmask % 89)
    terms = _terms(L, n_terms, 2 * xmask + n_terms)
    dev = torch.from_numpy(ref.copy()).cuda()
    K.debug_set_max_grid(grid)
    try:
        K.pauli_apply(dev, xmask, terms)
    finally:
        K.debug_set_max_grid(0)
    statevec.pauli_apply(ref, xmask, terms)
    # up to 64 coefficients of magnitude ~1 are summed per amplitude
    assert np.abs(dev.cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("xmask", [0, 3, (1 << 13) | (1 << 6), (1 << 14) - 1])
def test_pauli_apply_accumulator(xmask):
    """overwrite pass, accumulate pass from the slab, accumulate passes from a staged partner slab in pieces"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 14
    n = 1 << L
    ref = rand_state(L, 41)
    other = rand_state(L, 42)
    t1, t2, t3 = _terms(L, 2, 1), _terms(L, 3, 2), _terms(L, 1, 3)
    dev = torch.from_numpy(ref.copy()).cuda()
    acc = torch.full((n,), float("nan"), dtype=torch.complex128, device="cuda")
    exp = np.full(n, np.nan + 0j)
    K.pauli_apply(dev, xmask, t1, acc=acc, accumulate=False)
    statevec.pauli_apply(ref, xmask, t1, acc=exp, accumulate=False)
    K.pauli_apply(dev, xmask ^ 4, t2, acc=acc, accumulate=True)
    statevec.pauli_apply(ref, xmask ^ 4, t2, acc=exp, accumulate=True)
    piece = n // 4
    for b in range(0, n, piece):
        src = torch.from_numpy(other[b:b + piece].copy()).cuda()
        K.pauli_apply(dev, xmask, t3, acc=acc, accumulate=True, src=src, begin=b, count=piece)
        statevec.pauli_apply(ref, xmask, t3, acc=exp, accumulate=True, src=other[b:b + piece], begin=b, count=piece)
    assert np.abs(acc.cpu().numpy() - exp).max() <= TOL
    assert np.array_equal(dev.cpu().numpy(), ref)


def _perm_case(name, L, g):
    """(kind, pos, ctrl_mask, a, N, forward table or None)"""
    top = L + g
    if name == "add_contig":
        return 1, list(range(2, 9)), 1 << 11, 37, 0, None
    if name == "add_neg_scattered":
        return 1, [9, 0, 4, top - 1, 2], (1 << 1) | (1 << 7), (1 << 64) - 3, 0, None
    if name == "add_mod":
        return 2, [5, 6, 7, 8, 9, 10], 1, 17, 53, None
    if name == "mul_mod":
        return 3, [3, 1, top - 1, 8, 10, 0, 6], 1 << 4, 29, 101, None
    if name == "mul_mod_contig_global":
        return 3, list(range(L - 3, top)), 0, 7, (1 << (g + 3)) - 3, None
    if name == "whole_index_add":
        return 1, list(range(top)), 0, 12345, 0, None
    if name == "table":
        rng = np.random.default_rng(5)
        return 0, [2, top - 1, 7, 0, 5], 1 << 3, 0, 0, [int(x) for x in rng.permutation(32)]
    raise KeyError(name)


@pytest.mark.parametrize("name", ["add_contig", "add_neg_scattered", "add_mod", "mul_mod", "mul_mod_contig_global",
                                  "whole_index_add", "table"])
@pytest.mark.parametrize("g", [0, 1, 2])
def test_permute_gather(name, g):
    """register permutations over R = 2^g slabs (virtual ranks = separate buffers on one GPU)"""
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    L = 12
    R = 1 << g
    kind, pos, cmask, a, N, fwd = _perm_case(name, L, g)
    rng = np.random.default_rng(g + len(pos))
    full = rng.normal(size=R << L) + 1j * rng.normal(size=R << L)
    slabs_np = [full[r << L:(r + 1) << L] for r in range(R)]
    slabs = [torch.from_numpy(s.copy()).cuda() for s in slabs_np]
    inv = dev_table = None
    if fwd is not None:
        inv = [0] * len(fwd)
        for v, w in enumerate(fwd):
            inv[w] = v
        dev_table = torch.tensor(inv, dtype=torch.int32, device="cuda")
    a_oracle = a if a < (1 << 63) else a - (1 << 64)  # the C ABI takes the two's complement
    outs = []
    for r in range(R):
        dst = torch.empty(1 << L, dtype=torch.complex128, device="cuda")
        K.permute_gather(dst, slabs, r, kind, pos, cmask, a, N, dev_table)
        exp = statevec.permute_gather(slabs_np, r, L, kind, pos, cmask, a_oracle, N, inv)
        got = dst.cpu().numpy()
        assert np.array_equal(got, exp), (name, g, r)
        outs.append(got)
    # a permutation: every amplitude appears exactly once in the result
    assert np.array_equal(np.sort_complex(np.concatenate(outs)), np.sort_complex(full))


def test_permute_gather_rejects_bad_arguments():
    torch = _torch()
    from hiqsimulator_b200 import kernels as K
    from hiqsimulator_b200._lib import HiqError
    s = torch.zeros(1 << 8, dtype=torch.complex128, device="cuda")
    d = torch.zeros_like(s)
    with pytest.raises(HiqError, match="not invertible"):
        K.permute_gather(d, [s], 0, K.PERM_MUL_MOD, [0, 1, 2, 3], 0, 6, 15)
    with pytest.raises(HiqError, match="alias"):
        K.permute_gather(s, [s], 0, K.PERM_ADD, [0, 1], 0, 1, 0)
    with pytest.raises(HiqError, match="control"):
        K.permute_gather(d, [s], 0, K.PERM_ADD, [0, 1], 2, 1, 0)
    with pytest.raises(HiqError, match="missing the slab"):
        K.permute_gather(d, [s, None], 0, K.PERM_ADD, [0, 8], 0, 1, 0)



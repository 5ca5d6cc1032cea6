"""Host logic of hiqk_apply_dense WITHOUT a GPU: variant resolution, block-structure permutation, free-index deposit and
control masks, the tile kernel's swizzle and offsets, the tensor-core kernel's real embedding and fragment ownership.

`hiqk_dense_image` hands back the resolved variant and its kernel parameters; tests/dense_emulator.py executes them the way
the kernels do.  Cases: those of the GPU suite (every target-slot class x variant x control mask, small slabs, the
huge-gate control mask, block-structured matrices) and random ones."""
import numpy as np
import pytest

import dense_emulator
from oracle import statevec
from test_kernels_gpu import BLOCK_CASES, _ctrl_mask, dense_cases, multiplexed_matrix, rand_matrix, rand_state

TOL = 1e-12


def _check(L, slots, m, cm, variant, seed, stats=None):
    from hiqsimulator_b200 import kernels as K
    ref = rand_state(L, seed)
    got = ref.copy()
    dense_emulator.run_dense_image(K.dense_image(L, list(slots), m, cm, variant), got, stats)
    statevec.apply_dense(ref, list(slots), m, cm)
    return float(np.abs(got - ref).max())


@pytest.mark.parametrize("L,k,slots,variant,ctrl", dense_cases())
def test_dense_image_of_the_gpu_cases(L, k, slots, variant, ctrl):
    from hiqsimulator_b200 import kernels as K
    seed = (L * 131 + k * 17 + sum((i + 1) * s for i, s in enumerate(slots))) & 0xFFFF
    cm = _ctrl_mask(L, slots, ctrl, seed)
    stats = {}
    assert _check(L, slots, rand_matrix(k, seed + 1), cm, variant, seed, stats) <= TOL
    if variant == K.AUTO:
        assert stats["variant"] == K.lib().hiqk_dense_pick_variant(L, k, K._ints(slots))
    elif variant == K.TILED:
        assert stats["variant"] == dense_emulator.TILED
    elif variant == K.DMMA and k >= 2:
        assert stats["variant"] == dense_emulator.DMMA
    if stats["variant"] == dense_emulator.DIRECT and k == 4:
        assert stats["m3"] == 1  # full 16 x 16 products take the three-multiplication form


@pytest.mark.parametrize("L", [3, 5, 6, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_dense_image_small_slabs(L, k):
    if k > L:
        pytest.skip("k > L")
    from hiqsimulator_b200 import kernels as K
    rng = np.random.default_rng(L * 10 + k)
    for trial in range(4):
        slots = [int(s) for s in rng.choice(L, size=k, replace=False)]
        for variant in (K.AUTO, K.DMMA):
            assert _check(L, slots, rand_matrix(k, trial + 7), 0, variant, trial) <= TOL, (slots, variant)


def test_dense_image_many_controls():
    L = 16
    cm = ((1 << L) - 1) & ~(1 << 7)
    assert _check(L, [7], rand_matrix(1, 4), cm, 0, 3) <= TOL


@pytest.mark.parametrize("k,slots,select", BLOCK_CASES)
@pytest.mark.parametrize("ctrl", [0, "one"])
def test_dense_image_block_structure(k, slots, select, ctrl):
    from hiqsimulator_b200 import kernels as K
    L = 14
    m = multiplexed_matrix(k, list(select), 7 * k + len(select))
    cm = _ctrl_mask(L, slots, ctrl, 9)
    for variant in (K.AUTO, K.DIRECT, K.DIRECT_FULL):
        stats = {}
        assert _check(L, slots, m, cm, variant, 200 + k, stats) <= TOL, variant
        if variant == K.DIRECT:
            assert stats["ks"] == max(1, k - len(select))
        if variant == K.DIRECT_FULL:
            assert stats["ks"] == k


@pytest.mark.parametrize("seed", range(40))
def test_random_dense_images(seed):
    from hiqsimulator_b200 import kernels as K
    rng = np.random.default_rng(8000 + seed)
    seen = set()
    for rep in range(10):
        L = int(rng.integers(1, 17))
        k = int(rng.integers(1, min(5, L) + 1))
        slots = [int(x) for x in rng.choice(L, size=k, replace=False)]
        free = [s for s in range(L) if s not in slots]
        nc = int(rng.integers(0, min(len(free), 4) + 1)) if rng.random() < 0.5 else 0
        cm = 0
        for s in rng.choice(free, size=nc, replace=False) if nc else []:
            cm |= 1 << int(s)
        variant = int(rng.choice([K.AUTO, K.AUTO, K.DIRECT, K.TILED, K.DMMA, K.DIRECT_FULL]))
        if variant == K.TILED and (L < 10 or min(max(k + 7, 10), L) < k + 3):
            variant = K.AUTO
        t = rng.random()
        if t < 0.6 or k == 1:
            m = rand_matrix(k, int(rng.integers(1 << 30)))
        else:
            n_sel = int(rng.integers(1, k))
            m = multiplexed_matrix(k, sorted(int(x) for x in rng.choice(k, size=n_sel, replace=False)), int(rng.integers(1 << 30)))
        stats = {}
        err = _check(L, slots, m, cm, variant, 10 * seed + rep, stats)
        assert err <= TOL, (seed, rep, L, slots, cm, variant, stats)
        seen.add(stats["variant"])
    assert seen  # which kernels the random cases reached is reported by the next test


def test_random_cases_reach_every_kernel():
    from hiqsimulator_b200 import kernels as K
    reached = set()
    for L, k, slots in [(14, 4, [1, 5, 9, 12]), (14, 4, [0, 5, 9, 12]), (14, 2, [0, 9]), (14, 5, [2, 4, 6, 8, 10]), (6, 4, [0, 1, 2, 3])]:
        stats = {}
        assert _check(L, slots, rand_matrix(k, 1), 0, K.AUTO, 1, stats) <= TOL
        reached.add(stats["variant"])
    assert reached == {dense_emulator.DIRECT, dense_emulator.TILED, dense_emulator.DMMA}

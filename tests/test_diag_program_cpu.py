"""Host logic of the launchers that carry a diagonal program (hiqk_apply_diag_batch, hiqk_apply_dense_prediag) WITHOUT a GPU.

The launchers hand back the kernel parameters they would launch with (hiqk_diag_batch_image, hiqk_dense_prediag_image);
tests/diag_emulator.py executes them the way the kernels do.  Checked against the numpy oracle on the cases of the GPU suite
and on random programs: every class of factor (per CTA and chunk, per thread and chunk, per element, on the gate's
targets), slabs smaller than one chunk, block-structured gates, the three-multiplication form."""
import numpy as np
import pytest

import diag_emulator
from oracle import statevec
from test_kernels_gpu import BLOCK_CASES, _rand_diag_ops, multiplexed_matrix, rand_matrix, rand_state

TOL = 1e-12


def _apply_ops(ref, ops):
    for slots, d in ops:
        if slots:
            statevec.apply_diag(ref, list(slots), d, 0)
        else:
            ref *= d[0]


def _check_batch(L, ops, seed):
    from hiqsimulator_b200 import kernels as K
    ref = rand_state(L, seed)
    got = ref.copy()
    diag_emulator.run_diag_batch_image(K.diag_batch_image(L, ops), got)
    _apply_ops(ref, ops)
    return float(np.abs(got - ref).max())


@pytest.mark.parametrize("L,n_ops,seed", [(14, 1, 0), (14, 5, 1), (15, 16, 2), (9, 7, 3), (5, 3, 4), (13, 12, 5)])
def test_diag_batch_image_of_the_gpu_cases(L, n_ops, seed):
    assert _check_batch(L, _rand_diag_ops(L, n_ops, seed), 40 + seed) <= TOL


@pytest.mark.parametrize("seed", range(30))
def test_random_diag_batch_images(seed):
    rng = np.random.default_rng(5000 + seed)
    for rep in range(8):
        L = int(rng.integers(1, 17))
        n_ops = int(rng.integers(1, 17))
        ops = []
        for j in range(n_ops):
            k = int(rng.integers(0, min(5, L) + 1))
            # cluster the ops on few positions now and then, so that no untouched position is left for the u part
            pool = np.arange(L) if rng.random() < 0.6 else np.arange(max(0, L - 6), L)
            k = min(k, len(pool))
            slots = [int(x) for x in rng.choice(pool, size=k, replace=False)]
            ops.append((slots, np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k)) * rng.uniform(0.5, 1.5)))
        assert _check_batch(L, ops, 10 * seed + rep) <= TOL, (seed, rep, L, [o[0] for o in ops])


def _check_prediag(L, slots, m, ops, seed, stats=None):
    from hiqsimulator_b200 import kernels as K
    ref = rand_state(L, seed)
    got = ref.copy()
    diag_emulator.run_dense_prediag_image(K.dense_prediag_image(L, list(slots), m, ops), got, stats)
    _apply_ops(ref, ops)
    statevec.apply_dense(ref, list(slots), m, 0)
    return float(np.abs(got - ref).max())


@pytest.mark.parametrize("k,slots", [(1, (2,)), (1, (13,)), (2, (3, 9)), (2, (12, 2)), (3, (4, 5, 6)), (3, (13, 6, 2)),
                                     (4, (5, 2, 9, 12)), (4, (10, 11, 12, 13)), (4, (4, 8, 6, 10)),
                                     (1, (1,)), (3, (1, 6, 9)), (3, (9, 1, 2)), (4, (1, 5, 9, 12)), (4, (4, 3, 2, 1))])
@pytest.mark.parametrize("n_pre", [1, 4, 16])
def test_dense_prediag_image_of_the_gpu_cases(k, slots, n_pre):
    L = 14
    m = rand_matrix(k, 3 * k + n_pre)
    ops = _rand_diag_ops(L, n_pre, 100 * k + n_pre)
    ops[0] = (list(slots[:max(1, k - 1)]), np.exp(1j * np.linspace(0.1, 2.0, 1 << max(1, k - 1))))
    if n_pre > 1:
        outside = [s for s in range(L) if s not in slots][:3]
        ops[1] = (outside, np.exp(1j * np.linspace(0.3, 3.0, 8)))
    stats = {}
    assert _check_prediag(L, slots, m, ops, 70 + k, stats) <= TOL
    assert stats["m3"] == (1 if k == 4 else 0)


@pytest.mark.parametrize("k,slots,select", [c for c in BLOCK_CASES if c[0] >= 2])
@pytest.mark.parametrize("n_pre", [1, 6])
def test_dense_prediag_image_block_structure(k, slots, select, n_pre):
    L = 15
    m = multiplexed_matrix(k, list(select), 11 * k + len(select))
    ops = _rand_diag_ops(L, n_pre, 500 * k + n_pre)
    ops[0] = (list(slots[:max(1, k - 1)]), np.exp(1j * np.linspace(0.1, 2.0, 1 << max(1, k - 1))))
    stats = {}
    assert _check_prediag(L, slots, m, ops, 300 + k, stats) <= TOL
    assert stats["ks"] == max(1, k - len(select))  # the reduced product is what the kernel would run


@pytest.mark.parametrize("seed", range(30))
def test_random_dense_prediag_images(seed):
    from hiqsimulator_b200 import kernels as K
    rng = np.random.default_rng(6000 + seed)
    done = 0
    classes = np.zeros(4, dtype=np.int64)
    while done < 8:
        L = int(rng.integers(3, 17))
        k = int(rng.integers(1, min(4, L) + 1))
        slots = [int(x) for x in rng.choice(np.arange(L), size=k, replace=False)]
        if not K.dense_prediag_supported(L, slots):
            continue
        t = rng.random()
        if t < 0.5 or k == 1:
            m = rand_matrix(k, int(rng.integers(1 << 30)))
        else:
            n_sel = int(rng.integers(1, k))
            m = multiplexed_matrix(k, sorted(int(x) for x in rng.choice(np.arange(k), size=n_sel, replace=False)), int(rng.integers(1 << 30)))
        n_pre = int(rng.integers(1, 17))
        ops = []
        for j in range(n_pre):
            where = rng.random()
            if where < 0.35:
                pool = list(slots)
            elif where < 0.7:
                pool = list(range(L))
            else:
                pool = sorted(set(list(slots) + [int(x) for x in rng.choice(np.arange(L), size=min(3, L), replace=False)]))
            ko = min(int(rng.integers(0, 6)), len(pool))
            sl = [int(x) for x in rng.choice(pool, size=ko, replace=False)]
            ops.append((sl, np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << ko)) * rng.uniform(0.5, 1.5)))
        stats = {}
        err = _check_prediag(L, slots, m, ops, 10 * seed + done, stats)
        assert err <= TOL, (seed, done, L, slots, [o[0] for o in ops], stats)
        classes += np.array(stats["classes"])
        done += 1
    assert classes[3] > 0  # class-E factors occurred


def test_images_reject_what_the_launchers_reject():
    from hiqsimulator_b200 import kernels as K
    from hiqsimulator_b200._lib import HiqError
    d = np.exp(1j * np.arange(2))
    with pytest.raises(HiqError, match="distinct"):
        K.diag_batch_image(10, [([3, 3], np.ones(4))])
    with pytest.raises(HiqError, match="DIRECT"):
        K.dense_prediag_image(13, [0, 1, 2, 3], rand_matrix(4, 1), [([5], d)])  # slot-0 target of a 4-qubit gate: tensor-core kernel
    with pytest.raises(HiqError, match="1..16"):
        K.diag_batch_image(10, [([j % 10], d) for j in range(17)])

"""Host logic of the tile-program launcher WITHOUT a GPU.

`hiqk_tile_program_image` hands back the kernel parameters a launch would use; tests/tile_emulator.py executes them the
way `tile_program_kernel` does (index logic restated, arithmetic in numpy).  Checked against the numpy oracle: the hand-made
runs of the GPU suite (tests/test_kernels_gpu.py::TILE_RUNS) and a few hundred random programs — every mix of block /
full / monomial gates, select bits inside and outside the tile, padding, and diagonal ops of every class."""
import numpy as np
import pytest

import tile_emulator
from test_kernels_gpu import TILE_RUNS, _tile_reference, monomial_matrix, multiplexed_matrix, rand_matrix, rand_state, tile_case

TOL = 1e-12


def _run(L, steps, seed, stats=None):
    from hiqsimulator_b200 import kernels as K
    raw = K.tile_program_image(L, steps)
    ref = rand_state(L, seed)
    got = ref.copy()
    tile_emulator.run_image(raw, got, stats)
    _tile_reference(ref, steps)
    return float(np.abs(got - ref).max())


@pytest.mark.parametrize("case", range(len(TILE_RUNS)))
def test_tile_image_of_the_gpu_cases_matches_oracle(case):
    from hiqsimulator_b200 import kernels as K
    L, steps, tile_bits = tile_case(case)
    assert K.tile_program_fits(L, steps) == tile_bits
    stats = {}
    assert _run(L, steps, 800 + case, stats) <= TOL
    assert stats["tile_bits"] == tile_bits


def _random_program(rng, L):
    """a run the launcher accepts: 1..4 gates with 1..4 targets each, random structure, 0..16 diagonal ops per gate"""
    from hiqsimulator_b200 import kernels as K
    for _ in range(200):
        n_steps = int(rng.integers(1, 5))
        # targets drawn from a window so that runs often fit, sometimes with far-away (select / outside) bits
        window = sorted(int(x) for x in rng.choice(np.arange(L), size=min(L, int(rng.integers(4, 10))), replace=False))
        steps = []
        for s in range(n_steps):
            k = int(rng.integers(1, 5))
            pool = window if rng.random() < 0.7 else list(range(L))
            slots = tuple(int(x) for x in rng.choice(pool, size=k, replace=False))
            kind = rng.choice(["dense", "block", "mono", "block1"])
            seed = int(rng.integers(1 << 30))
            if kind == "dense" or k == 1:
                m = rand_matrix(k, seed)
            elif kind == "mono":
                m = monomial_matrix(k, seed)
            else:
                n_sel = int(rng.integers(1, k)) if kind == "block" else k - 1
                select = sorted(int(x) for x in rng.choice(np.arange(k), size=n_sel, replace=False))
                m = multiplexed_matrix(k, select, seed)
            n_pre = int(rng.choice([0, 1, 3, 7, 16]))
            ops = []
            for j in range(n_pre):
                ko = int(rng.integers(0, 5))
                where = rng.random()
                if where < 0.3 and k:      # on the gate's targets (class E)
                    base = list(slots)
                elif where < 0.6:          # anywhere
                    base = list(range(L))
                else:                      # mixed: some targets, some others
                    base = list(slots) + [int(x) for x in rng.choice(np.arange(L), size=3, replace=False)]
                base = sorted(set(base))
                ko = min(ko, len(base))
                sl = [int(x) for x in rng.choice(base, size=ko, replace=False)]
                d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << ko)) * rng.uniform(0.5, 1.5)
                ops.append((sl, d))
            steps.append((slots, m, ops))
        if K.tile_program_fits(L, steps):
            return steps
    raise AssertionError("no acceptable program found")


@pytest.mark.parametrize("seed", range(40))
def test_random_tile_images_match_oracle(seed):
    rng = np.random.default_rng(7000 + seed)
    for rep in range(6):
        L = int(rng.integers(11, 17))
        steps = _random_program(rng, L)
        err = _run(L, steps, 100 * seed + rep)
        assert err <= TOL, (seed, rep, L, [(s[0], len(s[2])) for s in steps])


def test_qft_like_runs_gather_without_bank_conflicts():
    """the launcher's swizzle: for the chains the engine forms on a QFT (two mixing + two select bits per cluster, neighbouring
    clusters) every quarter-warp gather phase hits eight distinct 16-byte bank groups"""
    L = 16
    for slots_run in ([(12, 13, 14, 15), (10, 11, 12, 13), (8, 9, 10, 11)], [(7, 8, 9, 10), (5, 6, 7, 8), (3, 4, 5, 6), (1, 2, 3, 4)]):
        steps = [(sl, multiplexed_matrix(4, [2, 3], 5 + i), []) for i, sl in enumerate(slots_run)]
        stats = {}
        assert _run(L, steps, 3, stats) <= TOL
        assert stats["gather_phases_with_bank_conflicts"] == 0


def test_tile_image_rejects_what_the_launcher_rejects():
    from hiqsimulator_b200 import kernels as K
    from hiqsimulator_b200._lib import HiqError
    m = rand_matrix(4, 1)
    wide = [((20, 21, 22, 23), m, []), ((10, 11, 12, 13), m, []), ((15, 16, 17, 18), m, [])]
    with pytest.raises(HiqError, match="do not fit"):
        K.tile_program_image(26, wide)
    with pytest.raises(HiqError, match="distinct"):
        K.tile_program_image(14, [((3, 3, 5, 6), m, [])])

"""Target for `ncu --set full` (round 2): the kernels a bench step is made of, over a 2^L slab (L = 29 by default: 8 GiB, so
ncu's save / restore between replay passes stays cheap).
    ncu --set full --clock-control none --import-source on -k regex:'tile_program|dense_direct|dense_dmma' -o out python tools/ncu_target_r02.py
Order of launches (per rep): 3M full product (staged), block-loop prediag (mix 2, QFT-like), tensor-core k=4 (slot-0 target),
tile program of 3 QFT-like gates, tile program of 2 full products, single-gate tile program with a slot-0 target."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiqsimulator_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=29)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
L = args.L
state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
rng = np.random.default_rng(0)


def haar(d):
    z = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, _ = np.linalg.qr(z)
    return q


def blocks(ks):
    m = np.zeros((16, 16), dtype=np.complex128)
    for v in range(16 >> ks):
        m[v << ks:(v + 1) << ks, v << ks:(v + 1) << ks] = haar(1 << ks)
    return m


def diag(k):
    return np.exp(1j * rng.uniform(0, 6.28, size=1 << k))


u4 = haar(16)
m2 = blocks(2)
tg = [L - 8, L - 7, L - 6, L - 5]
hi = [s for s in range(12, L) if s not in tg]
ops = [([tg[2], tg[3], 3, 15], diag(4))] + [([int(x) for x in rng.choice(hi, size=4, replace=False)], diag(4)) for _ in range(7)]
chain = [([L - 4, L - 3, L - 2, L - 1], m2, ops), ([L - 6, L - 5, L - 4, L - 3], m2, ops), ([L - 8, L - 7, L - 6, L - 5], m2, ops)]
pair = [([5, 9, 17, 25], u4, []), ([7, 12, 17, 22], haar(16), [])]
low = [([0, 9, 17, 25], u4, [])]
for _ in range(args.reps):
    K.apply_dense(state, [3, 9, 17, 25], u4, 0, K.DIRECT)
    K.apply_dense_prediag(state, tg, m2, ops)
    K.apply_dense(state, [0, 9, 17, 25], u4, 0, K.AUTO)
    K.apply_tile_program(state, chain)
    K.apply_tile_program(state, pair)
    K.apply_tile_program(state, low)
torch.cuda.synchronize()
print("ok", K.prob_masked(state))

"""Target for `ncu --set full`: a few launches of the dominant gate kernels over a 2^L slab
(L=30 by default: 16 GiB, so ncu's save/restore between replay passes stays cheap).
    ncu --set full --clock-control none --import-source on -k regex:'diag_kernel|dense_direct' \
        --launch-skip 2 -c 2 -o gpurun_out/r01_full python tools/ncu_target.py
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiqsimulator_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=30)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
L = args.L
state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
rng = np.random.default_rng(0)
z = rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16))
u4, _ = np.linalg.qr(z)
d4 = np.exp(1j * rng.uniform(0, 6.28, size=16))
for _ in range(args.reps):
    K.apply_diag(state, [3, 9, 17, 25], d4, 0)
    K.apply_dense(state, [3, 9, 17, 25], u4, 0, K.DIRECT)
    K.apply_dense(state, [0, 9, 17, 25], u4, 0, K.TILED)
torch.cuda.synchronize()
print("ok", K.prob_masked(state))
# folded-diagonal launches (QFT-like: one op overlapping two targets + 11 chunk-constant ops)
tg = [L - 8, L - 7, L - 6, L - 5]
hi = [s for s in range(12, L) if s not in tg]
ops = [([tg[0], tg[1], 3, 15], d4)] + [([int(x) for x in rng.choice(hi, size=4, replace=False)], d4) for _ in range(11)]
for _ in range(args.reps):
    K.apply_dense_prediag(state, tg, u4, ops)
    K.apply_diag_batch(state, ops[1:])
torch.cuda.synchronize()
print("ok2", K.prob_masked(state))
# block-structured matrix (two mixing bits, as 15 of the 21 dense passes of QFT-33): reduced product, plain and folded
m2 = np.zeros((16, 16), dtype=np.complex128)
for v in range(4):
    zz = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    qq, _ = np.linalg.qr(zz)
    m2[4 * v:4 * v + 4, 4 * v:4 * v + 4] = qq
for _ in range(args.reps):
    K.apply_dense(state, tg, m2, 0, K.DIRECT)
    K.apply_dense_prediag(state, tg, m2, ops)
    K.apply_dense(state, [0, 9, 17, 25], u4, 0, K.AUTO)  # slot-0 target: tensor-core kernel
torch.cuda.synchronize()
print("ok3", K.prob_masked(state))

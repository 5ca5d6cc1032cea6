"""Micro-benchmark of the folded-diagonal launches at one L: plain dense vs dense+diagonals (by op class)
and the batched diagonal kernel.  CUDA-event timing, JSON lines to gpurun_out/prediag_<tag>_L<L>.jsonl."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiqsimulator_b200 import kernels as K  # noqa: E402
from tools.sweep_kernels import time_launch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=30)
    ap.add_argument("--tag", default="r01")
    args = ap.parse_args()
    L = args.L
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/prediag_%s_L%d.jsonl" % (args.tag, L), "w")
    state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
    rng = np.random.default_rng(0)
    z = rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16))
    u4, _ = np.linalg.qr(z)
    nbytes = 32.0 * (1 << L)

    def diag(k):
        return np.exp(1j * rng.uniform(0, 6.28, size=1 << k))

    def rec(name, fn, **kw):
        med, best = time_launch(fn, reps=5, warm=2)
        r = dict(name=name, L=L, ms=med, best_ms=best, gbs=nbytes / med / 1e6, **kw)
        print(json.dumps(r), flush=True)
        out.write(json.dumps(r) + "\n")

    tg = [L - 8, L - 7, L - 6, L - 5]          # dense targets (QFT-like: mid/high slots)
    hi = [s for s in range(12, L) if s not in tg]
    rec("dense_plain", lambda: K.apply_dense(state, tg, u4, 0, K.DIRECT))
    for n in (1, 4, 12, 16):
        ops = [([int(x) for x in rng.choice(hi, size=4, replace=False)], diag(4)) for _ in range(n)]
        rec("dense+S0x%d_hi" % n, lambda: K.apply_dense_prediag(state, tg, u4, ops), n_ops=n)
    for n in (1, 4, 12):
        ops = [([int(x) for x in rng.choice(np.arange(0, L - 9), size=4, replace=False)], diag(4)) for _ in range(n)]
        rec("dense+S*x%d_any" % n, lambda: K.apply_dense_prediag(state, tg, u4, ops), n_ops=n)
    ops = [([tg[0], tg[1], 3, 15], diag(4))]
    rec("dense+E1(ov2)", lambda: K.apply_dense_prediag(state, tg, u4, ops), n_ops=1)
    ops = [([tg[0], tg[1], 3, 15], diag(4)), ([tg[2], tg[3], 5, 16], diag(4))]
    rec("dense+E2(ov2)", lambda: K.apply_dense_prediag(state, tg, u4, ops), n_ops=2)
    ops = [([tg[0], tg[1], 3, 15], diag(4))] + [([int(x) for x in rng.choice(hi, size=4, replace=False)], diag(4)) for _ in range(11)]
    rec("dense+E1+S0x11 (qft-like)", lambda: K.apply_dense_prediag(state, tg, u4, ops), n_ops=12)
    # block-structured matrices (select bits = the high matrix bits here): reduced product vs the full one
    def blocks(ks):
        m = np.zeros((16, 16), dtype=np.complex128)
        for v in range(16 >> ks):
            zz = rng.normal(size=(1 << ks, 1 << ks)) + 1j * rng.normal(size=(1 << ks, 1 << ks))
            q, _ = np.linalg.qr(zz)
            m[v << ks:(v + 1) << ks, v << ks:(v + 1) << ks] = q
        return m
    for ks in (1, 2, 3):
        mb = blocks(ks)
        rec("dense_mix%d" % ks, lambda: K.apply_dense(state, tg, mb, 0, K.DIRECT), mixing_bits=ks)
        rec("dense_mix%d_full_product" % ks, lambda: K.apply_dense(state, tg, mb, 0, K.DIRECT_FULL), mixing_bits=ks)
        rec("dense_mix%d+E1+S0x11 (qft-like)" % ks, lambda: K.apply_dense_prediag(state, tg, mb, ops), n_ops=12, mixing_bits=ks)
    for n in (1, 4, 10, 16):
        ops = [([int(x) for x in rng.choice(np.arange(0, L), size=4, replace=False)], diag(4)) for _ in range(n)]
        rec("diag_batch_x%d" % n, lambda: K.apply_diag_batch(state, ops), n_ops=n)
    ops = [([int(rng.integers(0, 3))] + [int(x) for x in rng.choice(np.arange(3, L), size=3, replace=False)], diag(4)) for _ in range(10)]
    rec("diag_batch_x10_lowbit", lambda: K.apply_diag_batch(state, ops), n_ops=10)
    rec("diag_single", lambda: K.apply_diag(state, [3, 9, 17, 25], diag(4), 0))
    out.close()


if __name__ == "__main__":
    main()

"""Micro-benchmark of the tile-resident gate programs at one L against the one-launch-per-gate kernels (CUDA events, median
of 5 after 2 warm-ups, JSON lines to gpurun_out/tile_<tag>_L<L>.jsonl): QFT-like chains of 1..4 block-structured gates with
their diagonal factors, pairs / triples of full products, single gates with a slot-0 target."""
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
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    L = args.L
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/tile_%s_L%d.jsonl" % (args.tag, L), "w")
    state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
    rng = np.random.default_rng(0)
    nbytes = 32.0 * (1 << L)

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

    def rec(name, fn, gates, **kw):
        med, best = time_launch(fn, reps=5, warm=2)
        r = dict(name=name, L=L, ms=med, best_ms=best, gates=gates, ms_per_gate=med / gates, gbs_per_pass=nbytes / med / 1e6,
                 effective_gbs=gates * nbytes / med / 1e6, **kw)
        print(json.dumps(r), flush=True)
        out.write(json.dumps(r) + "\n")

    def qft_ops(tg, n, avoid=()):
        """one op on two of the gate's targets (class E), the others away from every target of the run (per-tuple scalars)"""
        free = [s for s in range(L) if s not in tg and s not in avoid]
        return [([tg[2], tg[3], free[0], free[-1]], diag(4))] + [([int(x) for x in rng.choice(free, size=4, replace=False)], diag(4)) for _ in range(n - 1)]

    m2 = blocks(2)
    chain = []
    run_slots = list(range(L - 10, L))
    for i in range(4):
        tg = [L - 4 - 2 * i, L - 3 - 2 * i, L - 2 - 2 * i, L - 1 - 2 * i]
        chain.append((tg, m2, qft_ops(tg, 8, run_slots)))
    for n in (1, 2, 3):
        rec("tile_qft_chain_%d" % n, lambda: K.apply_tile_program(state, chain[:n]), n, ops_per_gate=8)
    rec("one_launch_per_gate_qft_chain_3", lambda: [K.apply_dense_prediag(state, tg, m, ops) for tg, m, ops in chain[:3]], 3, ops_per_gate=8)
    low = list(range(0, 11))
    low_chain = [([7, 8, 9, 10], m2, qft_ops([7, 8, 9, 10], 4, low)), ([5, 6, 7, 8], m2, qft_ops([5, 6, 7, 8], 3, low)),
                 ([3, 4, 5, 6], m2, qft_ops([3, 4, 5, 6], 2, low)), ([1, 2, 3, 4], m2, qft_ops([1, 2, 3, 4], 2, low))]
    rec("tile_qft_low_chain_4", lambda: K.apply_tile_program(state, low_chain), 4)
    noop_chain = [(tg, m, []) for tg, m, _ in chain]
    for n in (1, 2, 3):
        rec("tile_mix2_chain_%d_no_diagonals" % n, lambda: K.apply_tile_program(state, noop_chain[:n]), n)
    u = [haar(16) for _ in range(3)]
    full = [([5, 9, 17, 25], u[0], []), ([7, 12, 17, 22], u[1], []), ([6, 9, 12, 25], u[2], [])]
    rec("direct_full_product", lambda: K.apply_dense(state, full[0][0], u[0], 0, K.DIRECT), 1)
    for n in (1, 2, 3):
        rec("tile_full_product_x%d" % n, lambda: K.apply_tile_program(state, full[:n]), n)
    # how the number of far-apart regions a tile gathers from (2^high slots) and the run length (2^low slots) act on a
    # one-gate pass: targets (hi..hi+3) share the tile with `extra` further high slots brought in by a no-op partner gate
    ident = np.eye(2, dtype=np.complex128)
    for extra in (0, 2, 4):
        steps = [([L - 4, L - 3, L - 2, L - 1], m2, [])]
        if extra:
            steps.append(([L - 5 - i for i in range(min(extra, 4))], np.kron(np.eye(1 << (min(extra, 4) - 1)), ident) if extra > 1 else ident, []))
        rec("tile_mix2_high_slots_%d" % (4 + extra), lambda: K.apply_tile_program(state, steps), 1, tile_bits=K.tile_program_fits(L, steps))
    low = ([0, 9, 17, 25], u[0], [])
    rec("dmma_slot0", lambda: K.apply_dense(state, low[0], u[0], 0, K.AUTO), 1)
    rec("tile_slot0_single", lambda: K.apply_tile_program(state, [low]), 1)
    rec("tile_slot0_pair", lambda: K.apply_tile_program(state, [low, ([1, 9, 13, 22], u[1], [])]), 2)
    out.close()


if __name__ == "__main__":
    main()

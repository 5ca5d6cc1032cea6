"""Sustained (power-capped) throughput of the dense kernel variants at one L: every case is launched
back to back for ~1 s and the second half is timed with CUDA events, so the numbers are the ones a long
circuit sees (bench.py runs under sw_power_cap), not the burst figures of sweep_kernels.py.
JSON lines to gpurun_out/sustained_<tag>_L<L>.jsonl."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiqsimulator_b200 import kernels as K  # noqa: E402


def sustained(fn, warm, reps):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=30)
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--reps", type=int, default=80)
    ap.add_argument("--targets", default=None, help="only the DIRECT kernels (full and block forms) on these target slots, e.g. 29,30,31,32")
    args = ap.parse_args()
    L = args.L
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/sustained_%s_L%d.jsonl" % (args.tag, L), "w")
    state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
    rng = np.random.default_rng(0)
    nbytes = 32.0 * (1 << L)
    names = {K.DIRECT: "direct", K.TILED: "tiled", K.DMMA: "dmma", K.DIRECT_FULL: "direct_full"}

    def rec(**kw):
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")

    if args.targets:
        for tg in args.targets.split(";"):
            slots = [int(x) for x in tg.split(",")]
            z = rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16))
            q, _ = np.linalg.qr(z)
            m2 = np.zeros((16, 16), dtype=np.complex128)
            for v in range(4):
                m2[4 * v:4 * v + 4, 4 * v:4 * v + 4] = q[:4, :4]
            ms = sustained(lambda: K.apply_dense(state, slots, q, 0, K.DIRECT), args.reps, args.reps)
            rec(kind="dense", L=L, k=4, slots=slots, variant="direct", ms=ms, gbs=nbytes / ms / 1e6)
            ms = sustained(lambda: K.apply_dense(state, slots, m2, 0, K.DIRECT), args.reps, args.reps)
            rec(kind="dense_blocks", L=L, k=4, slots=slots, mixing_bits=2, variant="direct", ms=ms, gbs=nbytes / ms / 1e6)
        d = np.exp(1j * rng.uniform(0, 6.28, size=16))
        ms = sustained(lambda: K.apply_diag(state, [3, 9, 17, 25], d, 0), args.reps, args.reps)
        rec(kind="diag", L=L, k=4, ms=ms, gbs=nbytes / ms / 1e6)
        out.close()
        return
    for k in (4, 3):
        z = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _ = np.linalg.qr(z)
        for slots in ([0, 9, 17, 25][:k], [1, 9, 17, 25][:k], [0, 1, 17, 25][:k], [5, 9, 17, 25][:k]):
            for variant in (K.DIRECT, K.TILED, K.DMMA):
                if variant == K.TILED and min(slots) >= 5 and k == 3:
                    continue
                ms = sustained(lambda: K.apply_dense(state, slots, q, 0, variant), args.reps, args.reps)
                rec(kind="dense", L=L, k=k, slots=slots, variant=names[variant], ms=ms, gbs=nbytes / ms / 1e6)
    # block-structured matrices on the DIRECT path (mixing bits first)
    for ks in (1, 2, 3):
        m = np.zeros((16, 16), dtype=np.complex128)
        for v in range(16 >> ks):
            zz = rng.normal(size=(1 << ks, 1 << ks)) + 1j * rng.normal(size=(1 << ks, 1 << ks))
            qq, _ = np.linalg.qr(zz)
            m[v << ks:(v + 1) << ks, v << ks:(v + 1) << ks] = qq
        for variant in (K.DIRECT, K.DIRECT_FULL):
            ms = sustained(lambda: K.apply_dense(state, [5, 9, 17, 25], m, 0, variant), args.reps, args.reps)
            rec(kind="dense_blocks", L=L, k=4, mixing_bits=ks, variant=names[variant], ms=ms, gbs=nbytes / ms / 1e6)
    d = np.exp(1j * rng.uniform(0, 6.28, size=16))
    ms = sustained(lambda: K.apply_diag(state, [3, 9, 17, 25], d, 0), args.reps, args.reps)
    rec(kind="diag", L=L, k=4, ms=ms, gbs=nbytes / ms / 1e6)
    out.close()


if __name__ == "__main__":
    main()

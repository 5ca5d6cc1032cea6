"""Micro-sweep of the gate kernels on one B200 (mirrors the reference's opyt.cpp:194-212 sweep):
k x lowest-target-slot class x variant at a given L; effective GB/s = 32 B * 2^L / t.
Writes JSON lines to gpurun_out/sweep_<tag>.jsonl.  Not a bench.py replacement."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiqsimulator_b200 import kernels as K  # noqa: E402


def time_launch(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=30)
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--ks", default="1,2,3,4,5")
    args = ap.parse_args()
    L = args.L
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/sweep_%s_L%d.jsonl" % (args.tag, L), "w")
    peaks = {}
    for name, what in (("copy_gbs", 0), ("dfma_tflops", 1), ("dmma_tflops", 2)):
        peaks[name] = K.microbench(what, 5)
    print(json.dumps({"peaks": peaks}), flush=True)
    out.write(json.dumps({"peaks": peaks}) + "\n")
    state = torch.full((1 << L,), 2.0 ** (-L / 2), dtype=torch.complex128, device="cuda")
    rng = np.random.default_rng(0)
    nbytes = 32.0 * (1 << L)
    for k in [int(x) for x in args.ks.split(",")]:
        z = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _ = np.linalg.qr(z)
        for low in (0, 1, 2, 3, 5, 8, 13, L - k):
            slots = list(range(low, low + k))
            for vname, variant in (("direct", K.DIRECT), ("tiled", K.TILED), ("dmma", K.DMMA)):
                if variant == K.DMMA and k < 2:
                    continue
                med, best = time_launch(lambda: K.apply_dense(state, slots, q, 0, variant))
                rec = {"kind": "dense", "L": L, "k": k, "low_slot": low, "variant": vname, "ms": med, "best_ms": best,
                       "eff_gbs": nbytes / med / 1e6, "tflops": 8.0 * (1 << k) * (1 << L) / med / 1e9}
                print(json.dumps(rec), flush=True)
                out.write(json.dumps(rec) + "\n")
        # scattered targets
        slots = sorted(int(s) for s in rng.choice(np.arange(2, L), size=k, replace=False))
        for vname, variant in (("direct", K.DIRECT), ("dmma", K.DMMA)):
            if variant == K.DMMA and k < 2:
                continue
            med, best = time_launch(lambda: K.apply_dense(state, slots, q, 0, variant))
            rec = {"kind": "dense", "L": L, "k": k, "slots": slots, "variant": vname, "ms": med, "best_ms": best,
                   "eff_gbs": nbytes / med / 1e6}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        d = np.exp(1j * rng.uniform(0, 6.28, size=1 << k))
        med, best = time_launch(lambda: K.apply_diag(state, list(range(3, 3 + k)), d, 0))
        rec = {"kind": "diag", "L": L, "k": k, "ms": med, "eff_gbs": nbytes / med / 1e6}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
    med, best = time_launch(lambda: K.prob_masked(state))
    rec = {"kind": "prob", "L": L, "ms": med, "eff_gbs": nbytes / 2 / med / 1e6}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    med, best = time_launch(lambda: K.scale(state, 1.0))
    rec = {"kind": "scale", "L": L, "ms": med, "eff_gbs": nbytes / med / 1e6}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()

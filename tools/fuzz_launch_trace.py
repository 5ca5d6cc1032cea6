"""Fuzz of the engine's launch grouping and of the tile-program launcher WITHOUT a GPU: random gate streams (dense / monomial /
controlled-block / diagonal gates, controls, swaps; 1-4 virtual ranks, 11-14 local qubits, cluster size 3-5) run on dry-run
engines; the LAUNCH trace (tile programs executed through hiqk_tile_program_image + tests/tile_emulator.py) must equal the
PLAN trace (one pass per fused gate, the reference's sequence).    python tools/fuzz_launch_trace.py <first seed> <last seed>"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import scripts
from hiqsimulator_b200 import _cppsim_mpi as M
from hiqsimulator_b200.gates import haar_unitary

def engines_for(R, nq, max_local, max_cluster, seed):
    es = []
    for r in range(R):
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        es.append(M.SimulatorMPI(seed, max_local, max_cluster))
    M.init_world(0, 1, b"", 0, 0)
    return es

def mono(k, rng):
    d = 1 << k
    m = np.zeros((d, d), dtype=complex)
    m[rng.permutation(d), np.arange(d)] = np.exp(1j * rng.uniform(0, 6.28, d))
    return m

def one(seed):
    rng = np.random.default_rng(seed)
    R = int(rng.choice([1, 1, 2, 4]))
    g = R.bit_length() - 1
    L = int(rng.integers(11, 15))
    nq = L + g
    mc = int(rng.choice([3, 4, 4, 4, 5]))
    es = engines_for(R, nq, L, mc, seed)
    def call(name, *a):
        out = None
        for e in es:
            out = getattr(e, name)(*a)
        return out
    call("allocate_qureg", list(range(nq)), 0)
    pend = set()
    ng = int(rng.integers(40, 160))
    for _ in range(ng):
        loc = list(es[0].get_local_qubits_ids()); glo = [q for q in es[0].get_global_qubits_ids() if q >= 0]
        kind = rng.random()
        if kind < 0.55:
            k = int(rng.integers(1, min(4, mc) + 1))
            if rng.random() < 0.7:
                # neighbouring slots
                s0 = int(rng.integers(0, len(loc) - k + 1)); ids = [loc[s0 + i] for i in range(k)]
                rng.shuffle(ids); ids = [int(x) for x in ids]
            else:
                ids = [int(x) for x in rng.choice(loc, size=k, replace=False)]
            t = rng.random()
            if t < 0.5: m = haar_unitary(1 << k, rng)
            elif t < 0.75: m = mono(k, rng)
            else:
                # controlled-U folded: block structure
                m = np.eye(1 << k, dtype=complex)
                if k > 1:
                    u = haar_unitary(1 << (k - 1), rng); m[(1 << (k - 1)):, (1 << (k - 1)):] = u
                else: m = haar_unitary(2, rng)
            ctrls = []
            if rng.random() < 0.15:
                rest = [q for q in loc + glo if q not in ids]
                ctrls = [int(x) for x in rng.choice(rest, size=min(len(rest), int(rng.integers(1, 3))), replace=False)]
            need = set(ids) | set(ctrls)
            if len(pend | need) > mc: call("run"); pend = set()
            pend |= need
            call("apply_controlled_gate", m.tolist(), ids, ctrls)
            nlc = sum(1 for c in ctrls if c in loc)
            if len(ids) + nlc > mc: call("run"); pend = set()
        elif kind < 0.85:
            allq = loc + glo
            k = int(rng.integers(1, min(3, mc) + 1))
            ids = [int(x) for x in rng.choice(allq, size=k, replace=False)]
            d = np.exp(1j * rng.uniform(0, 6.28, 1 << k))
            need = set(ids)
            if len(pend | need) > mc: call("run"); pend = set()
            pend |= need
            call("apply_controlled_gate", np.diag(d).tolist(), ids, [])
        elif kind < 0.95:
            call("run"); pend = set()
        elif glo:
            q = int(rng.integers(1, len(glo) + 1))
            gs = [int(x) for x in rng.choice(glo, size=q, replace=False)]
            ls = [int(x) for x in rng.choice(loc, size=q, replace=False)]
            pairs = []
            for a, b in zip(gs, ls): pairs += [a, b]
            call("run"); pend = set()
            call("swap_qubits", pairs)
    call("run")
    for e in es: e.synchronize()
    info = {"tile_emulator": True}
    a = scripts.replay_traces([e.launch_trace() for e in es], R, info)
    b = scripts.replay_traces([e.trace() for e in es], R)
    err = float(np.abs(a - b).max())
    return err, info.get("launch_forms"), R, L, mc

if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    worst = 0; forms = {}
    t0 = time.time()
    for s in range(lo, hi):
        try:
            err, f, R, L, mc = one(s)
        except AssertionError as e:
            print("ASSERT", s, e); continue
        for k, v in (f or {}).items(): forms[k] = forms.get(k, 0) + v
        if err > 1e-12: print("MISMATCH", s, err, R, L, mc)
        worst = max(worst, err)
    print("seeds", lo, hi, "worst", worst, forms, round(time.time() - t0, 1), "s")

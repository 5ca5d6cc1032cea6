"""Crash / hang fuzz of the reference-facing class WITHOUT a GPU: random call sequences with invalid arguments (unallocated,
duplicated, negative and huge qubit ids, matrices of the wrong size, bit strings of the wrong length, empty lists, operators
on qubits outside the register) on dry-run engines of 1, 2 and 4 virtual ranks.  Every call must either succeed or raise;
a crash shows as a signal, a hang as the caller's timeout.    python tools/fuzz_invalid_arguments.py <first seed> <last seed>
With --ranks: every call goes to the engines of ALL ranks of a 2 / 4 / 8-rank world and must be accepted by all of them or
refused by all of them (a call that raises on some ranks only would leave the others waiting in the next collective on
hardware); prints DIVERGE lines otherwise."""
import sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiqsimulator_b200 import _cppsim_mpi as M
faulthandler.enable()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
ALL_RANKS = '--ranks' in sys.argv
n_div = 0
def rid(rng, nq):
    r = rng.random()
    if r < 0.7: return int(rng.integers(0, nq))
    if r < 0.85: return int(rng.integers(-3, nq + 4))
    return int(rng.choice([2**31, -2**31, 2**62, 10**6]))
def rlist(rng, nq, maxn=6):
    n = int(rng.integers(0, maxn))
    return [rid(rng, nq) for _ in range(n)]
def rmat(rng, k=None):
    if k is None: k = int(rng.integers(0, 6))
    d = 1 << k
    t = rng.random()
    if t < 0.5:
        m = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    elif t < 0.8:
        m = np.diag(np.exp(1j * rng.uniform(0, 6, d)))
    else:
        dd = int(rng.integers(1, 40)); m = rng.normal(size=(dd, dd)) + 0j
    return m
n_exc = n_ok = 0
class AllRanks:
    def __init__(s, es, seed): s.es = es; s.seed = seed
    def __getattr__(s, name):
        def g(*a):
            global n_div
            outs = []
            for e in s.es:
                try:
                    getattr(e, name)(*a); outs.append(('ok', None))
                except Exception as ex:
                    outs.append((type(ex).__name__, str(ex)))
            if len(set(o[0] for o in outs)) > 1:
                n_div += 1
                print('DIVERGE seed', s.seed, name, [x.shape if hasattr(x, 'shape') else x for x in a], outs, flush=True)
            if outs[0][0] != 'ok':
                raise RuntimeError(outs[0][1])
        return g

for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    R = int(rng.choice([2, 4, 8] if ALL_RANKS else [1, 2, 4]))
    r = int(rng.integers(0, R))
    L = int(rng.integers(3, 9)); mc = int(rng.integers(1, 6))
    if ALL_RANKS:
        es = []
        for rr in range(R):
            M.init_world(rr, R, b"", 0, M.FLAG_DRY_RUN)
            es.append(M.SimulatorMPI(seed, L, mc))
        e = AllRanks(es, seed)
    else:
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        e = M.SimulatorMPI(seed, L, mc)
    nq = L + R.bit_length() - 1
    for step in range(int(rng.integers(5, 60))):
        c = int(rng.integers(0, 22))
        print(seed, step, c, flush=True, file=sys.stderr) if os.environ.get("FZ_VERBOSE") else None
        try:
            if c == 0: e.allocate_qureg(list(range(nq)) if rng.random() < 0.6 else rlist(rng, nq, 12), 0)
            elif c == 1: e.allocate_qubit(rid(rng, nq + 2))
            elif c == 2: e.deallocate_qubit(rid(rng, nq))
            elif c in (3, 4, 5):
                ids = rlist(rng, nq, 6)
                m = rmat(rng, len(ids) if rng.random() < 0.8 else None)
                e.apply_controlled_matrix(np.ascontiguousarray(m, dtype=complex), ids, rlist(rng, nq, 4))
            elif c == 6: e.run()
            elif c == 7:
                p = rlist(rng, nq, 8)
                if len(p) == 0: continue   # documented: the reference does not terminate on an empty swap list
                e.swap_qubits(p)
            elif c == 8: e.measure_qubits(rlist(rng, nq, 6))
            elif c == 9:
                ids = rlist(rng, nq, 6); e.get_probability([bool(rng.integers(0, 2)) for _ in range(len(ids) + int(rng.integers(-1, 2)))], ids)
            elif c == 10:
                ids = rlist(rng, nq, nq + 2); e.get_amplitude([bool(rng.integers(0, 2)) for _ in range(len(ids))], ids)
            elif c == 11:
                ids = rlist(rng, nq, 6); e.collapse_wavefunction(ids, [bool(rng.integers(0, 2)) for _ in range(len(ids) + int(rng.integers(-1, 2)))])
            elif c == 12: e.set_qubits_perm(rlist(rng, nq, nq + 3))
            elif c == 13: e.get_qubits_ids(); e.get_local_qubits_ids(); e.get_global_qubits_ids()
            elif c == 14:
                terms = [([(rid(rng, nq), str(rng.choice(list("XYZQ")))) for _ in range(int(rng.integers(0, 4)))], complex(rng.normal())) for _ in range(int(rng.integers(0, 4)))]
                e.get_expectation_value(terms, rlist(rng, nq, nq + 1))
            elif c == 15:
                terms = [([(rid(rng, nq), str(rng.choice(list("XYZ")))) for _ in range(int(rng.integers(0, 4)))], complex(rng.normal())) for _ in range(int(rng.integers(0, 4)))]
                e.apply_qubit_operator(terms, rlist(rng, nq, nq + 1))
            elif c == 16:
                terms = [([(rid(rng, nq), str(rng.choice(list("XYZ")))) for _ in range(int(rng.integers(0, 4)))], complex(rng.normal())) for _ in range(int(rng.integers(0, 4)))]
                e.emulate_time_evolution(terms, float(rng.normal()), rlist(rng, nq, nq + 1), rlist(rng, nq, 3))
            elif c == 17:
                n = int(rng.integers(0, 5)); e.set_wavefunction(np.ones(1 << n, dtype=complex) / np.sqrt(1 << n), rlist(rng, nq, n + 2))
            elif c == 18:
                e.emulate_math_add_constant(int(rng.integers(-5, 50)), rlist(rng, nq, 5), rlist(rng, nq, 3))
            elif c == 19:
                e.emulate_math_multiply_by_constant_modN(int(rng.integers(0, 20)), int(rng.integers(0, 20)), rlist(rng, nq, 5), rlist(rng, nq, 3))
            elif c == 20:
                e.emulate_math_add_constant_modN(int(rng.integers(0, 20)), int(rng.integers(0, 20)), rlist(rng, nq, 5), rlist(rng, nq, 3))
            elif c == 21:
                e.trace(); e.launch_trace(); e.stats(); e.synchronize()
            n_ok += 1
        except (RuntimeError, ValueError, TypeError, OverflowError, IndexError) as ex:
            n_exc += 1
    del e
M.init_world(0, 1, b"", 0, 0)
print("seeds", lo, hi, "calls ok", n_ok, "refused", n_exc, *(("rank-divergent", n_div) if ALL_RANKS else ()))

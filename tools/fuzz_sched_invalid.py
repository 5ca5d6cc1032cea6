"""Crash / hang fuzz of the scheduler module (_sched_cpp: SwapScheduler, ClusterScheduler, GreedyPlanner) with degenerate and
invalid inputs: empty gate lists, ids outside the register, duplicated ids, list lengths that do not match, cluster sizes and
split counts of -1 / 0 / 40, registers with no local qubit.  Every call must return or raise.
    python tools/fuzz_sched_invalid.py <first seed> <last seed>"""
import sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiqsimulator_b200 import _sched_cpp as S
faulthandler.enable()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
ok = refused = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    nl = int(rng.integers(0, 8)); ng = int(rng.integers(0, 4))
    locals_ = [int(x) for x in rng.permutation(nl + ng)[:nl]]
    globals_ = [q for q in range(nl + ng) if q not in locals_]
    if rng.random() < 0.2: globals_ = globals_ + [int(rng.integers(-2, nl + ng + 3))]
    if rng.random() < 0.2 and locals_: locals_ = locals_ + [locals_[0]]
    n = int(rng.integers(0, 30))
    def ids(maxn):
        k = int(rng.integers(0, maxn + 1))
        r = rng.random()
        pool = nl + ng + (3 if r < 0.2 else 0)
        if pool == 0: return []
        if r < 0.8 and k <= pool: return [int(x) for x in rng.choice(pool, size=k, replace=False)]
        return [int(rng.integers(-1, pool + 1)) for _ in range(k)]
    gate = [ids(4) for _ in range(n)]
    ctrl = [ids(3) for _ in range(n if rng.random() < 0.9 else max(0, n - 1))]
    isz = [bool(rng.integers(0, 2)) for _ in range(n if rng.random() < 0.9 else n + 1)]
    cs = int(rng.choice([-1, 0, 1, 2, 3, 4, 5, 6, 40]))
    ns = int(rng.choice([-1, 0, 1, 2, 3, 4]))
    which = int(rng.integers(0, 3))
    print(seed, which, nl, ng, n, cs, ns, file=sys.stderr, flush=True) if os.environ.get("FZ_VERBOSE") else None
    try:
        if which == 0:
            s = S.SwapScheduler(gate, ctrl, isz, cs, ns, bool(rng.integers(0, 2)))
            s.ScheduleSwap()
        elif which == 1:
            s = S.ClusterScheduler(gate, ctrl, isz, locals_, globals_, cs)
            for _ in range(n + 2):
                r = s.ScheduleCluster()
                if not r: break
        else:
            p = S.GreedyPlanner(gate, ctrl, isz, locals_, globals_, cs, ns, bool(rng.integers(0, 2)))
            for _ in range(20 * n + 50):
                k, d = p.next()
                if k == 0: break
        ok += 1
    except (RuntimeError, ValueError, TypeError, IndexError, OverflowError, MemoryError) as ex:
        refused += 1
print("seeds", lo, hi, "ok", ok, "refused", refused)

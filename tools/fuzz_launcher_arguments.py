"""Crash fuzz of the launchers' host side WITHOUT a GPU: the parameter-image functions (hiqk_dense_image, hiqk_diag_batch_image,
hiqk_dense_prediag_image, hiqk_tile_program_image) and the host predicates (hiqk_tile_program_fits, hiqk_dense_pick_variant,
hiqk_dense_prediag_supported, hiqk_dense_is_monomial) run the same argument checks and fill functions as the launches; fed
slab sizes of -1..100, target counts of -1..31, slots that are negative / beyond the slab / repeated, op counts of -1..1000,
null matrices and op arrays.  Every call must return a status.    python tools/fuzz_launcher_arguments.py <first seed> <last seed>"""
import sys, ctypes, faulthandler, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiqsimulator_b200 import _lib
faulthandler.enable()
lib = ctypes.CDLL(_lib.LIB_PATH)
class DiagOp(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int), ("slots", ctypes.c_int * 5), ("lut", ctypes.c_double * 64)]
class TileStep(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int), ("slots", ctypes.c_int * 5), ("matrix", ctypes.c_void_p), ("pre", ctypes.c_void_p), ("n_pre", ctypes.c_int)]
for f in ("hiqk_dense_image_bytes", "hiqk_diag_batch_image_bytes", "hiqk_dense_prediag_image_bytes", "hiqk_tile_program_image_bytes"):
    getattr(lib, f).restype = ctypes.c_size_t
lo, hi = int(sys.argv[1]), int(sys.argv[2])
nb = [lib.hiqk_dense_image_bytes(), lib.hiqk_diag_batch_image_bytes(), lib.hiqk_dense_prediag_image_bytes(), lib.hiqk_tile_program_image_bytes()]
bufs = [ctypes.create_string_buffer(n) for n in nb]
ok = bad = 0
def rslot(rng, L):
    r = rng.random()
    if r < 0.8: return int(rng.integers(0, max(1, L)))
    return int(rng.choice([-1, L, L + 1, 63, 64, 1000, -2**31]))
def rop(rng, L):
    o = DiagOp()
    o.k = int(rng.choice([0, 1, 2, 3, 4, 5, 5, 6, -1, 100])) if rng.random() < 0.3 else int(rng.integers(0, 6))
    for l in range(5): o.slots[l] = rslot(rng, L)
    for i in range(64): o.lut[i] = float(rng.normal())
    return o
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    L = int(rng.choice([-1, 0, 1, 2, 3, 5, 8, 10, 12, 14, 20, 30, 33, 40, 41, 64, 100]))
    k = int(rng.choice([-1, 0, 1, 2, 3, 4, 5, 6, 31]))
    slots = (ctypes.c_int * 8)(*[rslot(rng, L) for _ in range(8)])
    m = (ctypes.c_double * 2048)(*rng.normal(size=2048))
    n_ops = int(rng.choice([-1, 0, 1, 2, 5, 16, 17, 1000]))
    ops = (DiagOp * 1000)(*[rop(rng, L) for _ in range(20)])  # the arrays are as long as the counts claim: only the geometry lies
    which = int(rng.integers(0, 6))
    if os.environ.get("FZ_VERBOSE"): print(seed, which, L, k, list(slots), n_ops, file=sys.stderr, flush=True)
    if which == 0:
        rc = lib.hiqk_dense_image(L, k, slots, m, ctypes.c_uint64(int(rng.integers(0, 2**62)) if rng.random() < 0.5 else 0), int(rng.integers(-1, 7)), bufs[0], ctypes.c_size_t(nb[0]))
    elif which == 1:
        rc = lib.hiqk_diag_batch_image(L, ops, n_ops, bufs[1], ctypes.c_size_t(nb[1]))
    elif which == 2:
        rc = lib.hiqk_dense_prediag_image(L, k, slots, m, ops, n_ops, bufs[2], ctypes.c_size_t(nb[2]))
    elif which in (3, 4):
        ns = int(rng.choice([-1, 0, 1, 2, 3, 4, 5, 100]))
        steps = (TileStep * 6)()
        for i in range(6):
            steps[i].k = int(rng.choice([-1, 0, 1, 2, 3, 4, 5, 6]))
            for l in range(5): steps[i].slots[l] = rslot(rng, L)
            steps[i].matrix = ctypes.cast(m, ctypes.c_void_p) if rng.random() < 0.9 else None
            steps[i].n_pre = int(rng.choice([-1, 0, 0, 1, 3, 16, 17]))
            steps[i].pre = ctypes.cast(ops, ctypes.c_void_p) if rng.random() < 0.9 else None
        if which == 3:
            rc = lib.hiqk_tile_program_image(L, ns, steps, bufs[3], ctypes.c_size_t(nb[3]))
        else:
            rc = 0 if lib.hiqk_tile_program_fits(L, ns, steps) else 1
    else:
        lib.hiqk_dense_pick_variant(L, k, slots); lib.hiqk_dense_prediag_supported(L, k, slots); lib.hiqk_dense_is_monomial(k, m)
        rc = 0
    if rc == 0: ok += 1
    else: bad += 1
print("seeds", lo, hi, "ok", ok, "refused", bad)

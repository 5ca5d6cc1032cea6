"""ctypes view of the C ABI in ``include/hiq_b200.h`` (device-level ``hiqk_*`` launchers).

The library is built in-tree by ``hiqsimulator_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
Failing to find it is an error — there is deliberately no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhiq_b200.so")


class LibraryMissing(RuntimeError):
    pass


class HiqError(RuntimeError):
    pass


_lib = None

_u64 = C.c_uint64
_vp = C.c_void_p
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)

class DiagOp(C.Structure):
    """hiqk_diag_op of include/hiq_b200.h"""
    _fields_ = [("k", C.c_int), ("slots", C.c_int * 5), ("lut", C.c_double * 64)]


class TileStep(C.Structure):
    """hiqk_tile_step of include/hiq_b200.h"""
    _fields_ = [("k", C.c_int), ("slots", C.c_int * 5), ("matrix", C.POINTER(C.c_double)), ("pre", C.POINTER(DiagOp)), ("n_pre", C.c_int)]


class PauliTerm(C.Structure):
    """hiqk_pauli_term of include/hiq_b200.h"""
    _fields_ = [("zmask", C.c_uint64), ("re", C.c_double), ("im", C.c_double)]


class Perm(C.Structure):
    """hiqk_perm of include/hiq_b200.h"""
    _fields_ = [("kind", C.c_int), ("n_bits", C.c_int), ("pos", C.c_int * 40), ("ctrl_mask", C.c_uint64),
                ("a", C.c_uint64), ("N", C.c_uint64), ("table", C.c_void_p)]


_SIGNATURES = {
    "hiq_last_error": (C.c_char_p, []),
    "hiq_version": (C.c_char_p, []),
    "hiq_device_count": (C.c_int, []),
    "hiqk_apply_dense": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _dp, _u64, C.c_int, _vp]),
    "hiqk_dense_pick_variant": (C.c_int, [C.c_int, C.c_int, _ip]),
    "hiqk_apply_diag": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _dp, _u64, _vp]),
    "hiqk_apply_diag_batch": (C.c_int, [_vp, C.c_int, C.POINTER(DiagOp), C.c_int, _vp]),
    "hiqk_apply_dense_prediag": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _dp, C.POINTER(DiagOp), C.c_int, _vp]),
    "hiqk_dense_prediag_supported": (C.c_int, [C.c_int, C.c_int, _ip]),
    "hiqk_scale": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _vp]),
    "hiqk_workspace_bytes": (C.c_size_t, []),
    "hiqk_prob_masked": (C.c_int, [_vp, C.c_int, _u64, _u64, _vp, _vp, _vp]),
    "hiqk_block_norms": (C.c_int, [_vp, C.c_int, _u64, _vp, _vp]),
    "hiqk_bit_norms": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "hiqk_entropy": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "hiqk_collapse": (C.c_int, [_vp, C.c_int, _u64, _u64, C.c_double, _vp]),
    "hiqk_fill": (C.c_int, [_vp, _u64, _u64, C.c_double, C.c_double, _vp]),
    "hiqk_compact_bit": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _u64, _vp]),
    "hiqk_swap_pack": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _u64, _u64, _u64, _vp, _vp]),
    "hiqk_swap_unpack": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _u64, _u64, _u64, _vp, _vp]),
    "hiqk_swap_p2p": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _ip, C.POINTER(_u64), _u64, C.POINTER(_u64),
                               C.POINTER(_u64), _vp]),
    "hiqk_dense_is_monomial": (C.c_int, [C.c_int, _dp]),
    "hiqk_tile_program_fits": (C.c_int, [C.c_int, C.c_int, C.POINTER(TileStep)]),
    "hiqk_apply_tile_program": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(TileStep), _vp]),
    "hiqk_dense_image_bytes": (C.c_size_t, []),
    "hiqk_dense_image": (C.c_int, [C.c_int, C.c_int, _ip, _dp, _u64, C.c_int, _vp, C.c_size_t]),
    "hiqk_diag_batch_image_bytes": (C.c_size_t, []),
    "hiqk_diag_batch_image": (C.c_int, [C.c_int, C.POINTER(DiagOp), C.c_int, _vp, C.c_size_t]),
    "hiqk_dense_prediag_image_bytes": (C.c_size_t, []),
    "hiqk_dense_prediag_image": (C.c_int, [C.c_int, C.c_int, _ip, _dp, C.POINTER(DiagOp), C.c_int, _vp, C.c_size_t]),
    "hiqk_tile_program_image_bytes": (C.c_size_t, []),
    "hiqk_tile_program_image": (C.c_int, [C.c_int, C.c_int, C.POINTER(TileStep), _vp, C.c_size_t]),
    "hiqk_axpy_masked": (C.c_int, [_vp, _vp, C.c_int, _u64, _u64, C.c_double, C.c_double, _vp]),
    "hiqk_swap_move": (C.c_int, [_vp, C.c_int, C.c_int, _ip, C.c_int, C.POINTER(_u64), _u64, _u64, C.POINTER(_vp), C.c_int, _vp]),
    "hiqk_pauli_expect": (C.c_int, [_vp, C.c_int, _u64, C.POINTER(PauliTerm), C.c_int, _vp, _u64, _u64, _vp, _vp, _vp]),
    "hiqk_pauli_apply": (C.c_int, [_vp, C.c_int, _u64, C.POINTER(PauliTerm), C.c_int, _vp, C.c_int, _vp, _u64, _u64, _vp]),
    "hiqk_permute_gather": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.POINTER(Perm), _vp]),
    "hiq_modinv": (C.c_int, [_u64, _u64, C.POINTER(_u64)]),
    "hiqk_dense_block_shape": (C.c_int, [C.c_int, _dp, _ip]),
    "hiqk_dense_direct_mixing_bits": (C.c_int, [C.c_int, _dp]),
    "hiqk_microbench": (C.c_int, [C.c_int, C.c_int, _dp]),
    "hiqk_launch_count": (_u64, []),
    "hiqk_debug_set_max_grid": (C.c_int, [C.c_int]),
}


def load_library(path: str | None = None):
    """Load ``libhiq_b200.so`` and declare every signature of the C ABI."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LibraryMissing(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def lib():
    return load_library()


def check(rc: int):
    if rc != 0:
        raise HiqError(lib().hiq_last_error().decode())


def exported_symbols():
    return sorted(_SIGNATURES)

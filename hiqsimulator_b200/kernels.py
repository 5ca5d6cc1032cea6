"""Python-side launchers for the device-level C ABI (``hiqk_*``) operating on torch CUDA tensors.

torch is used here only as the owner of device memory and streams; every computation is one of
this repository's own CUDA kernels reached through ``libhiq_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import DiagOp, PauliTerm, Perm, TileStep, check, lib

AUTO, DIRECT, TILED, DMMA, DIRECT_FULL = 0, 1, 2, 3, 4


def _slab(t):
    import torch
    assert t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()
    n = t.numel()
    L = int(round(math.log2(n)))
    assert 1 << L == n
    return C.c_void_p(t.data_ptr()), L


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ints(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _cplx(m):
    a = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def apply_dense(state, slots, matrix, ctrl_mask=0, variant=AUTO):
    p, L = _slab(state)
    keep, mp = _cplx(matrix)
    check(lib().hiqk_apply_dense(p, L, len(slots), _ints(slots), mp, ctrl_mask, variant, _stream()))


def dense_block_shape(matrix):
    """(number of mixing index bits, bit order with the mixing bits first) of a 2^k x 2^k matrix — host only."""
    keep, mp = _cplx(matrix)
    k = int(keep.shape[0]).bit_length() - 1
    order = (C.c_int * k)()
    ks = lib().hiqk_dense_block_shape(k, mp, order)
    if ks < 0:
        check(ks)
    return ks, [int(x) for x in order]


def apply_diag(state, slots, diag, ctrl_mask=0):
    p, L = _slab(state)
    keep, dp = _cplx(diag)
    check(lib().hiqk_apply_diag(p, L, len(slots), _ints(slots), dp, ctrl_mask, _stream()))


def _diag_ops(ops):
    """ops: iterable of (slots, diag) -> ctypes array of hiqk_diag_op"""
    arr = (DiagOp * len(ops))()
    for o, (slots, diag) in zip(arr, ops):
        d = np.ascontiguousarray(np.asarray(diag, dtype=np.complex128)).view(np.float64)
        o.k = len(slots)
        assert d.size == 2 << o.k
        for l, sl in enumerate(slots):
            o.slots[l] = int(sl)
        for i, v in enumerate(d):
            o.lut[i] = float(v)
    return arr


def apply_diag_batch(state, ops):
    """One pass: psi[i] *= prod_j diag_j[bits of i at slots_j]; ops = [(slots, diag), ...]"""
    p, L = _slab(state)
    arr = _diag_ops(ops)
    check(lib().hiqk_apply_diag_batch(p, L, arr, len(ops), _stream()))


def apply_dense_prediag(state, slots, matrix, pre):
    """One pass: psi <- M * prod_j D_j * psi (DIRECT kernel; pre = [(slots, diag), ...])"""
    p, L = _slab(state)
    keep, mp = _cplx(matrix)
    arr = _diag_ops(pre)
    check(lib().hiqk_apply_dense_prediag(p, L, len(slots), _ints(slots), mp, arr, len(pre), _stream()))


def _tile_steps(steps):
    """steps: [(slots, matrix, [(slots, diag), ...]), ...] -> (ctypes array of hiqk_tile_step, objects to keep alive)"""
    arr = (TileStep * len(steps))()
    keep = []
    for st, (slots, matrix, pre) in zip(arr, steps):
        m, mp = _cplx(matrix)
        ops = _diag_ops(pre) if pre else None
        keep += [m, ops]
        st.k = len(slots)
        for l, sl in enumerate(slots):
            st.slots[l] = int(sl)
        st.matrix = mp
        st.pre = ops if ops is not None else C.POINTER(DiagOp)()
        st.n_pre = len(pre) if pre else 0
    return arr, keep


def tile_program_fits(L, steps) -> int:
    """tile size in bits (11 / 12) when the run of gates can share one pass, 0 when it cannot — host only"""
    arr, keep = _tile_steps(steps)
    return int(lib().hiqk_tile_program_fits(L, len(steps), arr))


def apply_tile_program(state, steps):
    """One pass: for every step in order, psi <- M_s * prod_j D_sj * psi (tile-resident gate program)"""
    p, L = _slab(state)
    arr, keep = _tile_steps(steps)
    check(lib().hiqk_apply_tile_program(p, L, len(steps), arr, _stream()))


def dense_image(L, slots, matrix, ctrl_mask=0, variant=AUTO) -> bytes:
    """the variant apply_dense resolves to and the kernel parameters it would launch with — host only;
    tests/dense_emulator.py interprets them"""
    keep, mp = _cplx(matrix)
    n = lib().hiqk_dense_image_bytes()
    buf = C.create_string_buffer(n)
    check(lib().hiqk_dense_image(L, len(slots), _ints(slots), mp, ctrl_mask, variant, buf, n))
    return buf.raw


def diag_batch_image(L, ops) -> bytes:
    """the kernel parameters apply_diag_batch would launch with — host only; tests/diag_emulator.py interprets them"""
    arr = _diag_ops(ops)
    n = lib().hiqk_diag_batch_image_bytes()
    buf = C.create_string_buffer(n)
    check(lib().hiqk_diag_batch_image(L, arr, len(ops), buf, n))
    return buf.raw


def dense_prediag_image(L, slots, matrix, pre) -> bytes:
    """the kernel parameters apply_dense_prediag would launch with — host only; tests/diag_emulator.py interprets them"""
    keep, mp = _cplx(matrix)
    arr = _diag_ops(pre)
    n = lib().hiqk_dense_prediag_image_bytes()
    buf = C.create_string_buffer(n)
    check(lib().hiqk_dense_prediag_image(L, len(slots), _ints(slots), mp, arr, len(pre), buf, n))
    return buf.raw


def tile_program_image(L, steps) -> bytes:
    """the kernel-parameter image apply_tile_program would launch with (header | parameters | table pool) — host only,
    nothing is computed; tests/tile_emulator.py interprets it"""
    arr, keep = _tile_steps(steps)
    n = lib().hiqk_tile_program_image_bytes()
    buf = C.create_string_buffer(n)
    check(lib().hiqk_tile_program_image(L, len(steps), arr, buf, n))
    return buf.raw


def dense_prediag_supported(L, slots) -> bool:
    return bool(lib().hiqk_dense_prediag_supported(L, len(slots), _ints(slots)))


def scale(state, factor):
    p, L = _slab(state)
    f = complex(factor)
    check(lib().hiqk_scale(p, L, f.real, f.imag, _stream()))


_ws = {}


def _workspace(device):
    import torch
    key = str(device)
    if key not in _ws:
        _ws[key] = torch.empty(lib().hiqk_workspace_bytes() // 8, dtype=torch.float64, device=device)
    return _ws[key]


def prob_masked(state, mask=0, val=0) -> float:
    import torch
    p, L = _slab(state)
    out = torch.empty(1, dtype=torch.float64, device=state.device)
    check(lib().hiqk_prob_masked(p, L, mask, val, C.c_void_p(out.data_ptr()),
                                 C.c_void_p(_workspace(state.device).data_ptr()), _stream()))
    return float(out.item())


def entropy_sum(state) -> float:
    import torch
    p, L = _slab(state)
    out = torch.empty(1, dtype=torch.float64, device=state.device)
    check(lib().hiqk_entropy(p, L, C.c_void_p(out.data_ptr()),
                             C.c_void_p(_workspace(state.device).data_ptr()), _stream()))
    return float(out.item())


def bit_norms(state, slot):
    import torch
    p, L = _slab(state)
    out = torch.empty(2, dtype=torch.float64, device=state.device)
    check(lib().hiqk_bit_norms(p, L, slot, C.c_void_p(out.data_ptr()),
                               C.c_void_p(_workspace(state.device).data_ptr()), _stream()))
    return out.cpu().numpy()


def block_norms(state, n_blocks):
    import torch
    p, L = _slab(state)
    out = torch.empty(n_blocks, dtype=torch.float64, device=state.device)
    check(lib().hiqk_block_norms(p, L, n_blocks, C.c_void_p(out.data_ptr()), _stream()))
    return out.cpu().numpy()


def collapse(state, mask, val, scale_factor):
    p, L = _slab(state)
    check(lib().hiqk_collapse(p, L, mask, val, float(scale_factor), _stream()))


def fill(state, begin, count, value):
    v = complex(value)
    check(lib().hiqk_fill(C.c_void_p(state.data_ptr()), begin, count, v.real, v.imag, _stream()))


def compact_bit(state, slot, keep, scratch):
    p, L = _slab(state)
    check(lib().hiqk_compact_bit(p, L, slot, int(keep), C.c_void_p(scratch.data_ptr()), scratch.numel(), _stream()))


def swap_pack(state, slots, pat, begin, count, buf):
    p, L = _slab(state)
    check(lib().hiqk_swap_pack(p, L, len(slots), _ints(slots), pat, begin, count, C.c_void_p(buf.data_ptr()), _stream()))


def swap_unpack(state, slots, pat, begin, count, buf):
    p, L = _slab(state)
    check(lib().hiqk_swap_unpack(p, L, len(slots), _ints(slots), pat, begin, count, C.c_void_p(buf.data_ptr()), _stream()))


def swap_move(state, slots, peer_pats, begin, count, bufs, pack):
    """all peers' pieces in one launch: gather into bufs[k] (pack) or scatter from them (unpack)"""
    p, L = _slab(state)
    n = len(bufs)
    ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
    pats = (C.c_uint64 * n)(*[int(x) for x in peer_pats])
    check(lib().hiqk_swap_move(p, L, len(slots), _ints(slots), n, pats, begin, count, ptrs, 1 if pack else 0, _stream()))


def swap_p2p(local, peers, slots, peer_pats, my_pat, begins, counts):
    """in-place exchange of `local` with the peer slabs (torch tensors; on one GPU they simply are other buffers)"""
    p, L = _slab(local)
    n = len(peers)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in peers])
    pats = (C.c_uint64 * n)(*[int(x) for x in peer_pats])
    b = (C.c_uint64 * n)(*[int(x) for x in begins])
    c = (C.c_uint64 * n)(*[int(x) for x in counts])
    check(lib().hiqk_swap_p2p(p, ptrs, n, L, len(slots), _ints(slots), pats, int(my_pat), b, c, _stream()))


def _pauli_terms(terms):
    """terms: [(zmask, coefficient), ...] -> ctypes array of hiqk_pauli_term"""
    arr = (PauliTerm * len(terms))()
    for o, (z, c) in zip(arr, terms):
        c = complex(c)
        o.zmask, o.re, o.im = int(z), c.real, c.imag
    return arr


def pauli_expect(state, xmask, terms, src=None, begin=0, count=None) -> complex:
    """sum_i conj(state[i ^ xmask]) F(i) S[i] over [begin, begin+count); S = src (a staged partner slab) or state"""
    import torch
    p, L = _slab(state)
    count = state.numel() - begin if count is None else count
    out = torch.empty(2, dtype=torch.float64, device=state.device)
    check(lib().hiqk_pauli_expect(p, L, int(xmask), _pauli_terms(terms), len(terms),
                                  C.c_void_p(src.data_ptr()) if src is not None else None, begin, count,
                                  C.c_void_p(out.data_ptr()), C.c_void_p(_workspace(state.device).data_ptr()), _stream()))
    re, im = out.cpu().tolist()
    return complex(re, im)


def pauli_apply(state, xmask, terms, acc=None, accumulate=False, src=None, begin=0, count=None):
    """acc None: state <- P state in place; else acc[i ^ xmask] (+)= F(i) S[i] over [begin, begin+count)"""
    p, L = _slab(state)
    count = state.numel() - begin if count is None else count
    check(lib().hiqk_pauli_apply(p, L, int(xmask), _pauli_terms(terms), len(terms),
                                 C.c_void_p(acc.data_ptr()) if acc is not None else None, int(bool(accumulate)),
                                 C.c_void_p(src.data_ptr()) if src is not None else None, begin, count, _stream()))


PERM_TABLE, PERM_ADD, PERM_ADD_MOD, PERM_MUL_MOD = 0, 1, 2, 3


def permute_gather(dst, slabs, rank, kind, pos, ctrl_mask=0, a=0, N=0, table=None):
    """dst <- gather of rank `rank` through the inverse register map; slabs = one tensor per rank (or None);
    a, N are the FORWARD constants, table (torch int32/uint32 tensor on the device) the INVERSE map"""
    p, L = _slab(dst)
    n = len(slabs)
    ptrs = (C.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in slabs])
    d = Perm()
    d.kind, d.n_bits, d.ctrl_mask, d.a, d.N = int(kind), len(pos), int(ctrl_mask), int(a), int(N)
    for b, q in enumerate(pos):
        d.pos[b] = int(q)
    d.table = table.data_ptr() if table is not None else None
    check(lib().hiqk_permute_gather(p, ptrs, n, int(rank), L, C.byref(d), _stream()))


def debug_set_max_grid(max_ctas: int) -> None:
    """cap the grid of the persistent kernels (0 = natural): lets small slabs take the multi-iteration paths"""
    check(lib().hiqk_debug_set_max_grid(int(max_ctas)))


def microbench(what: int, iters: int = 5) -> float:
    out = C.c_double(0.0)
    check(lib().hiqk_microbench(what, iters, C.byref(out)))
    return out.value

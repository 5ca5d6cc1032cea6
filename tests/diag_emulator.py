"""numpy interpreters of the parameter images of the two launchers that carry a diagonal program (test infrastructure).

`hiqk_diag_batch_image` / `hiqk_dense_prediag_image` (include/hiq_b200.h) write the kernel parameters a launch of
`diag_batch_kernel` (csrc/stream_kernels.cu) / `dense_direct_pre_kernel`, `dense_direct_pre_blocks_kernel`
(csrc/apply_dense.cu) would receive.  The functions below execute such an image the way the kernels do: the index is split
into (chunk | per-thread tuples | thread), every diagonal factor is looked up with the OR of its partial selectors
(DiagProg, csrc/hiq_device.cuh), class-E factors go through the joint-pattern tables — index logic restated statement by
statement, arithmetic in complex128 numpy, vectorised over (chunk, thread).  A launcher that splits the index wrongly,
misclassifies a factor or builds a wrong table makes the result differ from the oracle."""
from __future__ import annotations

import struct

import numpy as np


def _insert_zero_bits(f, positions):
    f = np.asarray(f, dtype=np.uint64).copy()
    for pos in positions:
        pos = np.uint64(pos)
        low = f & ((np.uint64(1) << pos) - np.uint64(1))
        f = ((f >> pos) << (pos + np.uint64(1))) | low
    return f


class _Prog:
    """DiagProg inside an image"""

    def __init__(self, raw, at, off, max_ops, n_usel, n_lut):
        def i32(name):
            return struct.unpack_from("<i", raw, at + off[name])[0]
        self.n, self.n_s0, self.n_s1, self.n_e, self.n_s0a = (i32(k) for k in ("n", "n_s0", "n_s1", "n_e", "n_s0a"))
        self.slots = np.frombuffer(raw, np.uint8, max_ops * 8, at + off["slots"]).reshape(max_ops, 8).astype(np.uint64)
        self.usel = np.frombuffer(raw, np.uint8, max_ops * n_usel, at + off["usel"]).reshape(max_ops, n_usel).astype(np.int64)
        self.lut = np.frombuffer(raw, np.complex128, max_ops * n_lut, at + off["lut"]).reshape(max_ops, n_lut)
        assert 0 <= self.n_s0a <= self.n_s0 and self.n_s0 + self.n_s1 + self.n_e == self.n <= max_ops

    def select(self, j, idx):
        """diag_select: selector bit l = index bit slots[j][l], l < 5 (unused slots read bit 63 = 0)"""
        idx = np.asarray(idx, dtype=np.uint64)
        sel = np.zeros(idx.shape, dtype=np.int64)
        for l in range(5):
            sel |= ((idx >> self.slots[j][l]) & np.uint64(1)).astype(np.int64) << l
        return sel

    def chunk_state(self, chunk_idx):
        """diag_prog_chunk: per-chunk partial selectors and the product of the CTA-uniform factors"""
        selh = [self.select(j, chunk_idx) for j in range(self.n)]
        s_hi = np.ones(np.asarray(chunk_idx).shape, dtype=np.complex128)
        for j in range(self.n_s0a):
            s_hi = s_hi * self.lut[j][selh[j]]
        return selh, s_hi

    def s0(self, selh, s_hi, selt, threads):
        """diag_prog_s0: [chunks, threads]"""
        s = np.repeat(s_hi[:, None], threads, axis=1)
        for j in range(self.n_s0a, self.n_s0):
            s = s * self.lut[j][selh[j][:, None] | selt[j][None, :]]
        return s

    def s1(self, selh, selt, u, s):
        for j in range(self.n_s0, self.n_s0 + self.n_s1):
            s = s * self.lut[j][selh[j][:, None] | selt[j][None, :] | self.usel[j][u]]
        return s


_PROG_FIELDS = ["n", "n_s0", "n_s1", "n_e", "n_s0a", "slots", "usel", "lut"]


def run_diag_batch_image(raw: bytes, psi: np.ndarray) -> None:
    """psi *= the diagonal factors, the way diag_batch_kernel applies them"""
    head = struct.unpack_from("<32I", raw, 0)
    assert head[0] == 0x42445148, "not a diag-batch image"
    threads, _size, max_ops, n_usel, n_lut, _sz_ins = head[1:7]
    off = dict(zip(["n", "n_chunks", "n_u", "ins", "uoff", "prog"], head[7:13]))
    poff = dict(zip(_PROG_FIELDS, head[13:21]))
    base = 32 * 4
    n = struct.unpack_from("<Q", raw, base + off["n"])[0]
    n_chunks = struct.unpack_from("<Q", raw, base + off["n_chunks"])[0]
    n_u = struct.unpack_from("<i", raw, base + off["n_u"])[0]
    ins_n = struct.unpack_from("<i", raw, base + off["ins"])[0]
    ins_pos = np.frombuffer(raw, np.uint8, 64, base + off["ins"] + 4)[:ins_n]
    uoff = np.frombuffer(raw, np.uint64, n_usel, base + off["uoff"])
    prog = _Prog(raw, base + off["prog"], poff, max_ops, n_usel, n_lut)
    assert n == psi.shape[0] and ins_n == n_u and prog.n_e == 0
    assert threads == 256
    tid = np.arange(threads, dtype=np.uint64)
    selt = [prog.select(j, tid) for j in range(prog.n)]
    chunk = np.arange(n_chunks, dtype=np.uint64)
    cidx = _insert_zero_bits(chunk << np.uint64(8), ins_pos)
    selh, s_hi = prog.chunk_state(cidx)
    s0 = prog.s0(selh, s_hi, selt, threads)
    base_idx = cidx[:, None] | tid[None, :]
    touched = np.zeros(n, dtype=np.int32)
    for u in range(1 << n_u):
        idx = base_idx | uoff[u]
        ok = idx < np.uint64(n)
        f = prog.s1(selh, selt, u, s0) if prog.n_s1 else s0
        psi[idx[ok]] = psi[idx[ok]] * f[ok]
        np.add.at(touched, idx[ok].astype(np.int64), 1)
    assert touched.min() == 1 and touched.max() == 1, "the index split does not visit every amplitude exactly once"


_PRE_FIELDS = ["n_free", "ins", "off", "m", "msum", "n_t", "fast", "toff", "e_npat", "e_pat", "e_cmap", "prog"]


def run_dense_prediag_image(raw: bytes, psi: np.ndarray, stats: dict | None = None) -> None:
    """psi <- M * prod_j D_j * psi, the way dense_direct_pre_kernel / dense_direct_pre_blocks_kernel do it"""
    head = struct.unpack_from("<64I", raw, 0)
    assert head[0] == 0x50445148, "not a dense-prediag image"
    K, threads, ks, m3, _size, max_ops, n_usel, n_lut, n_tmax, _sz_ins = head[1:11]
    off = dict(zip(_PRE_FIELDS, head[11:11 + len(_PRE_FIELDS)]))
    poff = dict(zip(_PROG_FIELDS, head[11 + len(_PRE_FIELDS):11 + len(_PRE_FIELDS) + len(_PROG_FIELDS)]))
    base = 64 * 4
    D = 1 << K
    n_free = struct.unpack_from("<Q", raw, base + off["n_free"])[0]
    ins_n = struct.unpack_from("<i", raw, base + off["ins"])[0]
    ins_pos = np.frombuffer(raw, np.uint8, 64, base + off["ins"] + 4)[:ins_n]
    toff_c = np.frombuffer(raw, np.uint64, D, base + off["off"])
    m = np.frombuffer(raw, np.complex128, D * D, base + off["m"]).reshape(D, D)
    n_t = struct.unpack_from("<i", raw, base + off["n_t"])[0]
    fast = struct.unpack_from("<i", raw, base + off["fast"])[0]
    toff = np.frombuffer(raw, np.uint64, n_tmax, base + off["toff"])
    e_npat = struct.unpack_from("<i", raw, base + off["e_npat"])[0]
    e_pat = np.frombuffer(raw, np.uint8, max_ops * D, base + off["e_pat"]).reshape(max_ops, D).astype(np.int64)
    e_cmap = np.frombuffer(raw, np.uint8, D, base + off["e_cmap"]).astype(np.int64)
    prog = _Prog(raw, base + off["prog"], poff, max_ops, n_usel, n_lut)
    assert ins_n == K + n_t and (n_free << (K + n_t)) == psi.shape[0]
    assert prog.n_s1 == 0 or not fast
    if m3:
        msum = np.frombuffer(raw, np.float64, D * D, base + off["msum"]).reshape(D, D)
        assert K == 4 and ks == 4 and np.array_equal(msum, m.real + m.imag)
    je = prog.n_s0 + prog.n_s1
    tid = np.arange(threads, dtype=np.uint64)
    valid = tid < np.uint64(n_free)
    selt = [prog.select(j, _insert_zero_bits(tid, ins_pos)) for j in range(prog.n)]
    n_chunks = (n_free + threads - 1) // threads
    chunk = np.arange(n_chunks, dtype=np.uint64)
    selh, s_hi = prog.chunk_state(_insert_zero_bits(chunk * np.uint64(threads), ins_pos))
    s0 = prog.s0(selh, s_hi, selt, threads)
    bidx = _insert_zero_bits(chunk[:, None] * np.uint64(threads) + tid[None, :], ins_pos)

    def patterns(s, t):
        out = []
        for e in range(e_npat):
            f = s
            for j in range(je, prog.n):
                f = f * prog.lut[j][selh[j][:, None] | selt[j][None, :] | prog.usel[j][t] | e_pat[j][e]]
            out.append(f)
        return out

    sT = patterns(s0, 0) if (fast and prog.n_e) else None
    touched = np.zeros(psi.shape[0], dtype=np.int32)
    DS = 1 << ks
    for t in range(1 << n_t):
        basei = (bidx | toff[t])[:, valid]                     # [chunks, valid threads]
        s = s0
        if not fast:
            s = prog.s1(selh, selt, t, s0)
            if prog.n_e:
                sT = patterns(s, t)
        x = []
        for c in range(D):
            idx = (basei + toff_c[c]).astype(np.int64)
            np.add.at(touched, idx.ravel(), 1)
            fac = s if prog.n_e == 0 else sT[e_cmap[c]]
            x.append(psi[idx] * fac[:, valid])
        for b in range(D):
            lo = b & ~(DS - 1)
            acc = np.zeros_like(x[0])
            for c in range(lo, lo + DS):                       # apply_rows<K, KS>: row b meets the columns of its block only
                acc = acc + m[b, c] * x[c]
            psi[(basei + toff_c[b]).astype(np.int64)] = acc
    assert touched.min() == 1 and touched.max() == 1, "the index split does not visit every amplitude exactly once"
    if stats is not None:
        stats.update({"K": K, "ks": ks, "m3": m3, "n_t": n_t, "fast": fast, "classes": (prog.n_s0a, prog.n_s0 - prog.n_s0a, prog.n_s1, prog.n_e),
                      "e_npat": e_npat})

"""numpy interpreter of a tile-program parameter image (test infrastructure).

`hiqk_tile_program_image` (include/hiq_b200.h) writes the kernel parameters a launch of `tile_program_kernel`
(hiqsimulator_b200/csrc/tile_program.cu) would receive.  This module executes that image the way the kernel does — tile by
tile, one 16-element tuple per thread and gate, through the launcher's swizzle, block selection, selector bytes and
table pool — so that the launcher's host logic (tile choice, planned bit order, diagonal-op classes, monomial form) is
checked against the oracle where there is no GPU.  It restates the kernel's INDEX logic, statement by statement;
the arithmetic is plain complex128 numpy.  Vectorised over (tile, thread)."""
from __future__ import annotations

import struct

import numpy as np

_P_FIELDS = ["n_tiles", "n_steps", "lo", "outer", "swz_mask", "ioff", "pi", "tslot", "step", "n_lut", "lut_pad", "m", "msum"]
_S_FIELDS = ["ks", "n_ops", "n_e", "sel_off", "n_t", "n_em", "in_mask", "n_in", "n_out", "out_slot", "mono_row", "tpos", "ploff",
             "lut_off", "lpos", "outer", "esel"]


class Image:
    def __init__(self, raw: bytes):
        head = struct.unpack_from("<64I", raw, 0)
        assert head[0] == 0x50545148, "not a tile-program image"
        self.T, self.sizeof_params, self.sizeof_step, self.max_steps, self.max_ops, self.lut_entries, self.sizeof_insert = head[1:8]
        self.p_off = dict(zip(_P_FIELDS, head[8:8 + len(_P_FIELDS)]))
        self.s_off = dict(zip(_S_FIELDS, head[8 + len(_P_FIELDS):8 + len(_P_FIELDS) + len(_S_FIELDS)]))
        self.base = 64 * 4
        self.raw = raw
        lut_at = self.base + self.sizeof_params
        self.lut = np.frombuffer(raw, dtype=np.complex128, count=self.lut_entries, offset=lut_at)

    def _arr(self, off, dtype, count):
        return np.frombuffer(self.raw, dtype=dtype, count=count, offset=self.base + off)

    def p(self, name, dtype, count=1):
        a = self._arr(self.p_off[name], dtype, count)
        return a if count > 1 else a[0]

    def s(self, step, name, dtype, count=1):
        a = self._arr(self.p_off["step"] + step * self.sizeof_step + self.s_off[name], dtype, count)
        return a if count > 1 else a[0]

    def matrix(self, step):
        """p.m[step] as a 16 x 16 complex array (leading dimension 16)"""
        return self._arr(self.p_off["m"] + step * 256 * 16, np.complex128, 256).reshape(16, 16)

    def msum(self, step):
        return self._arr(self.p_off["msum"] + step * 256 * 8, np.float64, 256).reshape(16, 16)


def _parity(x):
    x = x.copy()
    for sh in (16, 8, 4, 2, 1):
        x ^= x >> sh
    return x & 1


def _phys(j, masks):
    j = np.asarray(j, dtype=np.uint32)
    return j ^ (_parity(j & masks[0]) | (_parity(j & masks[1]) << 1) | (_parity(j & masks[2]) << 2))


def _insert_zero_bits(f, positions):
    f = np.asarray(f, dtype=np.uint64).copy()
    for pos in positions:
        pos = np.uint64(pos)
        low = f & ((np.uint64(1) << pos) - np.uint64(1))
        f = ((f >> pos) << (pos + np.uint64(1))) | low
    return f


def run_image(raw: bytes, psi: np.ndarray | None, stats: dict | None = None, tiles=None, source=None):
    """apply the program to `psi` (complex128, 2^L amplitudes) in place, the way tile_program_kernel does.
    Slabs too large to hold (the bench's 2^33): pass psi=None, tiles = the tile numbers to process and source(indices) ->
    amplitudes; returns (indices [tiles, 2^T], values [tiles, 2^T]) of what the kernel would store."""
    im = Image(raw)
    T = im.T
    THREADS = 1 << (T - 4)
    n_tiles = int(im.p("n_tiles", np.uint64))
    n_steps = int(im.p("n_steps", np.int32))
    lo = int(im.p("lo", np.int32))
    assert psi is None or n_tiles << T == psi.shape[0]
    if source is None:
        source = lambda g: psi[g]  # noqa: E731
    outer_n = int(im._arr(im.p_off["outer"], np.int32, 1)[0])
    outer_pos = im._arr(im.p_off["outer"] + 4, np.uint8, 64)[:outer_n]
    swz = [np.uint32(x) for x in im.p("swz_mask", np.uint32, 3)]
    ioff = im.p("ioff", np.uint64, 16)
    pi = im.p("pi", np.uint16, 16).astype(np.uint32)
    tslot = im.p("tslot", np.uint8, 16)
    lut = im.lut
    n_lut = int(im.p("n_lut", np.int32))
    lut_pad = int(im.p("lut_pad", np.int32))
    assert n_lut <= lut_pad <= im.lut_entries

    tid = np.arange(THREADS, dtype=np.uint32)
    pt = _phys(tid, swz)
    goff_t = np.zeros(THREADS, dtype=np.uint64)
    for b in range(T - 4):
        goff_t |= ((tid >> b) & 1).astype(np.uint64) << np.uint64(tslot[b])
    # the load / store phases must cover every tile position exactly once
    cover = np.sort(np.concatenate([pt ^ pi[i] for i in range(16)]))
    assert np.array_equal(cover, np.arange(1 << T)), "load/store slices do not tile the buffer"

    t_idx = np.arange(n_tiles, dtype=np.uint64) if tiles is None else np.asarray(tiles, dtype=np.uint64)
    assert int(t_idx.max()) < n_tiles
    n_tiles = t_idx.shape[0]  # tiles processed from here on
    tbase = _insert_zero_bits(t_idx, outer_pos) << np.uint64(lo)          # [tiles]
    tile = np.zeros((n_tiles, 1 << T), dtype=np.complex128)
    for i in range(16):
        g = tbase[:, None] + goff_t[None, :] + ioff[i]
        tile[:, pt ^ pi[i]] = source(g)

    conflicts = 0
    for s in range(n_steps):
        ks = int(im.s(s, "ks", np.int32))
        n_ops = int(im.s(s, "n_ops", np.int32))
        n_e = int(im.s(s, "n_e", np.int32))
        n_t = int(im.s(s, "n_t", np.int32))
        n_em = int(im.s(s, "n_em", np.int32))
        in_mask = int(im.s(s, "in_mask", np.int32))
        n_in = int(im.s(s, "n_in", np.int32))
        n_out = int(im.s(s, "n_out", np.int32))
        out_slot = im.s(s, "out_slot", np.uint8, 4)
        mono_row = im.s(s, "mono_row", np.uint8, 16)
        tpos = im.s(s, "tpos", np.uint8, 4)
        ploff = im.s(s, "ploff", np.uint16, 16).astype(np.uint32)
        lut_off = im.s(s, "lut_off", np.uint16, im.max_ops).astype(np.int64)
        lpos = im.s(s, "lpos", np.uint8, im.max_ops * 5).reshape(im.max_ops, 5)
        outer = im.s(s, "outer", np.uint8, im.max_ops * 5).reshape(im.max_ops, 5)
        esel = im.s(s, "esel", np.uint8, im.max_ops * 16).reshape(im.max_ops, 16).astype(np.int64)
        m = im.matrix(s)
        assert 0 <= n_t <= n_ops - n_e and 0 <= n_em <= n_e <= n_ops <= im.max_ops
        assert list(tpos) == sorted(set(int(x) for x in tpos)) and tpos[-1] < T, "tuple bits must be four distinct tile positions"

        # per-thread context (once per launch in the kernel)
        base = tid.copy()
        for i in range(4):
            low = base & ((np.uint32(1) << np.uint32(tpos[i])) - np.uint32(1))
            base = ((base >> np.uint32(tpos[i])) << np.uint32(tpos[i] + 1)) | low
        pb = _phys(base, swz)
        my_sel = np.zeros((n_ops, THREADS), dtype=np.int64)
        for j in range(n_ops):
            for l in range(5):
                if lpos[j][l] != 0xFF:
                    my_sel[j] |= ((base >> np.uint32(lpos[j][l])) & 1).astype(np.int64) << l
        # per-tile selectors
        selh = np.zeros((n_ops, n_tiles), dtype=np.int64)
        for j in range(n_ops):
            for l in range(5):
                selh[j] |= ((tbase >> np.uint64(outer[j][l])) & np.uint64(1)).astype(np.int64) << l
        stile = np.ones(n_tiles, dtype=np.complex128)
        for j in range(n_t):
            stile = stile * lut[lut_off[j] + selh[j]]

        # the 16 tile positions of every thread's tuple must be distinct across threads (no two threads write one cell)
        cells = np.stack([pb ^ ploff[c] for c in range(16)])            # [16, threads]
        assert np.unique(cells).size == 16 * THREADS, "tuples of different threads overlap"
        # bank groups of a quarter warp (8 lanes, 16-byte cells): count the conflicting gathers (evidence, not an assertion)
        for c in range(16):
            banks = (cells[c].reshape(-1, 8) & 7)
            conflicts += int(sum(len(set(row)) != 8 for row in banks))

        x = np.stack([tile[:, cells[c]] for c in range(16)])            # [16, tiles, threads]

        def sel_of(j):
            return selh[j][:, None] | my_sel[j][None, :]                # [tiles, threads]

        if n_ops:
            n_s = n_ops - n_e
            sc = np.repeat(stile[:, None], THREADS, axis=1)
            for j in range(n_t, n_s):
                sc = sc * lut[lut_off[j] + sel_of(j)]
            if ks in (1, 2):
                DS, NB = 4, 4
                n_es = n_e - n_em
                f = [sc.copy() for _ in range(NB)]
                for j in range(n_s, n_s + n_es):
                    sel0 = lut_off[j] + sel_of(j)
                    for blk in range(NB):
                        f[blk] = f[blk] * lut[sel0 + esel[j][blk * DS]]
                for j in range(n_s + n_es, n_ops):
                    sel0 = lut_off[j] + sel_of(j)
                    for c in range(16):
                        x[c] = x[c] * lut[sel0 + esel[j][c]]
                for c in range(16):
                    x[c] = x[c] * f[c >> 2]
            else:
                for j in range(n_s, n_ops):
                    sel0 = sel_of(j)
                    for c in range(16):
                        x[c] = x[c] * lut[lut_off[j] + (sel0 | esel[j][c])]
                for c in range(16):
                    x[c] = x[c] * sc

        out = np.zeros_like(x)
        if ks == 0:
            for c in range(16):
                out[mono_row[c]] = m.reshape(-1)[c] * x[c]
            assert sorted(int(r) for r in mono_row) == list(range(16)), "monomial rows must be a permutation"
        elif ks in (1, 2):
            DS, NB = 4, 4
            osel = np.zeros(n_tiles, dtype=np.int64)
            for i in range(3):
                if i < n_out:
                    osel |= ((tbase >> np.uint64(out_slot[i])) & np.uint64(1)).astype(np.int64) << i
            obase = osel << n_in
            flat = m.reshape(-1)
            for blk in range(NB):
                q = (blk & in_mask) | obase                              # [tiles]
                assert int(q.max()) * DS * 17 + (DS - 1) * 16 + DS - 1 < 256, "block index outside the matrix"
                for r in range(DS):
                    acc = np.zeros_like(x[0])
                    for j in range(DS):
                        coef = flat[q * (DS * 17) + r * 16 + j]          # [tiles]
                        acc = acc + coef[:, None] * x[blk * DS + j]
                    out[blk * DS + r] = acc
        else:
            # full product; the kernel's three-multiplication form needs msum = re + im of the same matrix
            assert np.array_equal(im.msum(s), m.real + m.imag)
            out = np.einsum("bc,ctn->btn", m, x)
        for c in range(16):
            tile[:, cells[c]] = out[c]

    out_idx = np.zeros((n_tiles, 1 << T), dtype=np.uint64)
    out_val = np.zeros((n_tiles, 1 << T), dtype=np.complex128)
    for i in range(16):
        g = tbase[:, None] + goff_t[None, :] + ioff[i]
        if psi is not None:
            psi[g] = tile[:, pt ^ pi[i]]
        out_idx[:, i * THREADS:(i + 1) * THREADS] = g
        out_val[:, i * THREADS:(i + 1) * THREADS] = tile[:, pt ^ pi[i]]
    if stats is not None:
        stats["tile_bits"] = T
        stats["gather_phases_with_bank_conflicts"] = conflicts
        stats["n_tiles"] = n_tiles
    return out_idx, out_val

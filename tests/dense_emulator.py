"""numpy interpreter of the parameter image of hiqk_apply_dense (test infrastructure).

`hiqk_dense_image` (include/hiq_b200.h) resolves the kernel variant the launcher would take for a gate (DIRECT — with the
block-structure permutation and the three-multiplication flag —, TILED, DMMA) and writes that kernel's parameters.  The
functions below execute them the way `dense_direct_kernel` / `dense_direct_staged_kernel`, `dense_tiled_kernel` and
`dense_dmma_kernel` (csrc/apply_dense.cu) do: free-index deposit and control mask, tile load through the bank swizzle,
tuple gathers, the real embedding and the `mma.sync.m8n8k4.f64` fragment ownership of the tensor-core kernel (which lane
loads and stores what) — index logic restated statement by statement, arithmetic in numpy."""
from __future__ import annotations

import struct

import numpy as np

DIRECT, TILED, DMMA = 1, 2, 3


def _insert_zero_bits(f, positions):
    f = np.asarray(f, dtype=np.uint64).copy()
    for pos in positions:
        pos = np.uint64(pos)
        low = f & ((np.uint64(1) << pos) - np.uint64(1))
        f = ((f >> pos) << (pos + np.uint64(1))) | low
    return f


def _ins(raw, at):
    n = struct.unpack_from("<i", raw, at)[0]
    return np.frombuffer(raw, np.uint8, 64, at + 4)[:n]


def _u64(raw, at):
    return struct.unpack_from("<Q", raw, at)[0]


def _i32(raw, at):
    return struct.unpack_from("<i", raw, at)[0]


class _Virtual:
    """a slab too large to hold: reads come from source(indices), writes are collected"""

    def __init__(self, source):
        self.source = source
        self.idx, self.val = [], []

    def __getitem__(self, i):
        return self.source(np.asarray(i))

    def __setitem__(self, i, v):
        self.idx.append(np.asarray(i).ravel().copy())
        self.val.append(np.asarray(v).ravel().copy())


def run_dense_image(raw: bytes, psi, stats: dict | None = None, sample=None, source=None):
    """psi: complex128 array updated in place — or None with `sample` (work items to process: free indices for DIRECT,
    groups of 8 for DMMA, tiles for TILED) and `source(indices)`: returns (indices, values) of what would be stored"""
    virtual = psi is None
    if virtual:
        psi = _Virtual(source)
    out = _run(raw, psi, stats, sample, virtual)
    if virtual:
        return np.concatenate(psi.idx), np.concatenate(psi.val)
    return out


def _run(raw, psi, stats, sample, virtual):
    head = struct.unpack_from("<32I", raw, 0)
    assert head[0] == 0x4e445148, "not a dense image"
    variant, K, ks, m3 = head[1:5]
    base = 32 * 4
    D = 1 << K
    if stats is not None:
        stats.update({"variant": variant, "K": K, "ks": ks, "m3": m3})
    touched = None if virtual else np.zeros(psi.shape[0], dtype=np.int32)
    if variant == DIRECT:
        off = dict(zip(["n_free", "ctrl_mask", "ins", "off", "m", "msum"], head[7:13]))
        n_free = _u64(raw, base + off["n_free"])
        ctrl_mask = np.uint64(_u64(raw, base + off["ctrl_mask"]))
        ins = _ins(raw, base + off["ins"])
        eoff = np.frombuffer(raw, np.uint64, D, base + off["off"])
        m = np.frombuffer(raw, np.complex128, D * D, base + off["m"]).reshape(D, D)
        if m3:
            msum = np.frombuffer(raw, np.float64, D * D, base + off["msum"]).reshape(D, D)
            assert K == 4 and ks == 4 and np.array_equal(msum, m.real + m.imag)
        f = np.arange(n_free, dtype=np.uint64) if sample is None else np.asarray(sample, dtype=np.uint64)
        assert int(f.max()) < n_free
        b = (_insert_zero_bits(f, ins) | ctrl_mask).astype(np.int64)
        x = [psi[b + int(eoff[c])] for c in range(D)]
        if not virtual:
            for c in range(D):
                np.add.at(touched, b + int(eoff[c]), 1)
        DS = 1 << ks
        for r in range(D):
            lo = r & ~(DS - 1)
            acc = np.zeros_like(x[0])
            for c in range(lo, lo + DS):
                acc = acc + m[r, c] * x[c]
            psi[b + int(eoff[r])] = acc
        if not virtual:
            assert touched.max() == 1, "two tuples overlap"
            # untouched amplitudes are exactly those whose control bits are not all set
            idx = np.arange(psi.shape[0], dtype=np.uint64)
            assert np.array_equal(touched == 1, (idx & ctrl_mask) == ctrl_mask)
        return
    if variant == TILED:
        names = ["n_tiles", "hi_ctrl_mask", "lo_ctrl_mask", "lo", "tile_bits", "nswz", "swz_src", "swz_dst", "outer", "inner", "hoff", "loff", "m"]
        off = dict(zip(names, head[7:7 + len(names)]))
        n_tiles = _u64(raw, base + off["n_tiles"])
        hi_ctrl = np.uint64(_u64(raw, base + off["hi_ctrl_mask"]))
        lo_ctrl = np.uint32(struct.unpack_from("<I", raw, base + off["lo_ctrl_mask"])[0])
        lo = _i32(raw, base + off["lo"])
        tb = _i32(raw, base + off["tile_bits"])
        nswz = _i32(raw, base + off["nswz"])
        swz_src = struct.unpack_from("<3I", raw, base + off["swz_src"])
        swz_dst = struct.unpack_from("<3I", raw, base + off["swz_dst"])
        outer = _ins(raw, base + off["outer"])
        inner = _ins(raw, base + off["inner"])
        hoff = np.frombuffer(raw, np.uint64, 32, base + off["hoff"])
        loff = np.frombuffer(raw, np.uint32, D, base + off["loff"])
        m = np.frombuffer(raw, np.complex128, D * D, base + off["m"]).reshape(D, D)

        def swz(j):
            j = np.asarray(j, dtype=np.uint32)
            x = np.zeros_like(j)
            for i in range(nswz):
                x |= ((j >> np.uint32(swz_src[i])) & np.uint32(1)) << np.uint32(swz_dst[i])
            return j ^ x

        tile_amps = 1 << tb
        j = np.arange(tile_amps, dtype=np.uint32)
        pj = swz(j).astype(np.int64)
        assert np.array_equal(np.sort(pj), np.arange(tile_amps)), "the swizzle is not a permutation of the tile"
        lo_mask = np.uint32((1 << lo) - 1)
        t = np.arange(n_tiles, dtype=np.uint64) if sample is None else np.asarray(sample, dtype=np.uint64)
        assert int(t.max()) < n_tiles
        n_tiles = t.shape[0]
        tbase = (_insert_zero_bits(t, outer) << np.uint64(lo)) | hi_ctrl
        g = (tbase[:, None] + (j & lo_mask).astype(np.uint64)[None, :] + hoff[(j >> np.uint32(lo)).astype(np.int64)][None, :]).astype(np.int64)
        if not virtual:
            np.add.at(touched, g.ravel(), 1)
            assert touched.max() == 1, "two tiles overlap"
        tile = np.zeros((n_tiles, tile_amps), dtype=np.complex128)
        tile[:, pj] = psi[g]
        u = np.arange(tile_amps >> K, dtype=np.uint64)
        lb = _insert_zero_bits(u, inner).astype(np.uint32)
        keep = (lb & lo_ctrl) == lo_ctrl
        lb = lb[keep]
        pb = swz(lb)
        cells = [(pb ^ loff[c]).astype(np.int64) for c in range(D)]
        x = [tile[:, cells[c]] for c in range(D)]
        for r in range(D):
            acc = np.zeros_like(x[0])
            for c in range(D):
                acc = acc + m[r, c] * x[c]
            tile[:, cells[r]] = acc
        psi[g] = tile[:, pj]
        if stats is not None:
            stats["tile_bits"] = tb
        return
    assert variant == DMMA
    off = dict(zip(["n_groups", "ctrl_mask", "ins", "off", "m"], head[7:12]))
    n_groups = _u64(raw, base + off["n_groups"])
    ctrl_mask = np.uint64(_u64(raw, base + off["ctrl_mask"]))
    ins = _ins(raw, base + off["ins"])
    eoff = np.frombuffer(raw, np.uint64, D, base + off["off"])
    m = np.frombuffer(raw, np.complex128, D * D, base + off["m"]).reshape(D, D)
    KT, NT = 2 * D // 4, 2 * D // 8
    # B fragments as the kernel builds them: lane holds Mreal[8 nt + lane / 4][4 kt + lane % 4]
    bs = np.zeros((KT, NT, 32))
    for kt in range(KT):
        for nt in range(NT):
            for lane in range(32):
                r, q = 8 * nt + (lane >> 2), 4 * kt + (lane & 3)
                e = m[r >> 1, q >> 1]
                bs[kt, nt, lane] = e.real if (r & 1) == (q & 1) else (e.imag if (r & 1) else -e.imag)
    grp = np.arange(n_groups, dtype=np.uint64) if sample is None else np.asarray(sample, dtype=np.uint64)
    assert int(grp.max()) < n_groups
    n_groups = grp.shape[0]
    lane = np.arange(32)
    t, j = lane >> 2, lane & 3
    tb_idx = (_insert_zero_bits(grp[:, None] * np.uint64(8) + t[None, :].astype(np.uint64), ins) | ctrl_mask).astype(np.int64)  # [groups, lanes]
    # A fragment: lane (t, j) loads component (j & 1) of element 2 kt + (j >> 1) of tuple t (an 8-byte load)
    a = np.zeros((KT, n_groups, 32))
    for kt in range(KT):
        elem = 2 * kt + (j >> 1)
        amp = psi[tb_idx + eoff[elem].astype(np.int64)[None, :]]
        a[kt] = np.where((j & 1)[None, :] == 0, amp.real, amp.imag)
    # mma.m8n8k4 (row.col): D[row][col] += sum_k A[row][k] B[k][col]; A: lane -> (row = lane / 4, k = lane % 4);
    # B: lane -> (k = lane % 4, col = lane / 4); D: lane -> (row = lane / 4, cols 2 (lane % 4), + 1)
    d = np.zeros((NT, n_groups, 8, 8))
    for kt in range(KT):
        A = a[kt].reshape(n_groups, 8, 4)                    # [group, row, k]
        for nt in range(NT):
            B = bs[kt, nt].reshape(8, 4).T                   # [k, col]
            d[nt] += A @ B
    for nt in range(NT):
        # lane (t, j) stores (d0, d1) = D[t][2 j], D[t][2 j + 1] as the amplitude at element 4 nt + j of tuple t
        val = d[nt][:, t, 2 * j] + 1j * d[nt][:, t, 2 * j + 1]          # [groups, lanes]
        dst = tb_idx + eoff[4 * nt + j].astype(np.int64)[None, :]
        if not virtual:
            np.add.at(touched, dst.ravel(), 1)
        psi[dst] = val
    if not virtual:
        assert touched.max() == 1, "two lanes store the same amplitude"
        idx = np.arange(psi.shape[0], dtype=np.uint64)
        assert np.array_equal(touched == 1, (idx & ctrl_mask) == ctrl_mask)

"""Synthetic circuits of the benchmark configurations, as Command lists.

ProjectQ is not available, so the harness builds the gate matrices itself with numpy
(`default_rng` seeds are fixed and stated).  Every generator returns (n_qubits, [Command...])
without allocation / measurement commands.
"""
from __future__ import annotations

import math

import numpy as np

from . import gates as G
from .ops import Gate


def random_circuit(n: int, depth: int = 20, seed: int = 20261017):
    """C2/C4: layer d = Haar SU(2) on every qubit, then CZ / CNOT / Haar SU(4) (cycled) on a random
    perfect matching of the qubits."""
    cmds = []
    kinds = 0
    for d in range(depth):
        rng = np.random.default_rng(seed + d)
        for q in range(n):
            cmds.append(Gate(G.haar_unitary(2, rng), [q], name="U2"))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            a, b = int(perm[i]), int(perm[i + 1])
            kind = kinds % 3
            kinds += 1
            if kind == 0:
                cmds.append(Gate(G.Z, [b], [a], name="CZ", is_z=True))
            elif kind == 1:
                cmds.append(Gate(G.X, [b], [a], name="CNOT"))
            else:
                cmds.append(Gate(G.haar_unitary(4, rng), [a, b], name="U4"))
    return n, cmds


def qft_circuit(n: int, x: int | None = None):
    """C3: basis state |x> (X gates), then H + controlled-R(pi/2^j) ladder, no final swaps.
    ProjectQ's QFT decomposition: for i = n-1..0: H(i); for j < i: CR(pi / 2^(i-j)) target i control j."""
    if x is None:
        x = 0x5A5A5A5A5A5A5A5A & ((1 << n) - 1)
    cmds = []
    for q in range(n):
        if (x >> q) & 1:
            cmds.append(Gate(G.X, [q], name="X"))
    for i in range(n - 1, -1, -1):
        cmds.append(Gate(G.H, [i], name="H"))
        for j in range(i - 1, -1, -1):
            cmds.append(Gate(G.R(math.pi / (1 << (i - j))), [j], [i], name="CR"))
    return n, cmds


def qft_expected_amplitude(n: int, x: int, y_bits: int) -> complex:
    """Amplitude of basis state y after qft_circuit(n, x) (the swap-less QFT leaves the output
    bit-reversed): amp(y) = 2^(-n/2) exp(2 pi i x rev(y) / 2^n)."""
    rev = int(format(y_bits, "0%db" % n)[::-1], 2)
    phase = (x * rev) % (1 << n)
    return 2.0 ** (-n / 2) * np.exp(2j * math.pi * phase / (1 << n))


def grover_circuit(n_search: int, iterations: int, marked: int | None = None):
    """C1 (structure of examples/grover_mpi.py:23-94): n_search data qubits + 1 oracle qubit."""
    n = n_search
    oracle = n
    if marked is None:
        marked = ((1 << n) - 1) & ~0b10  # all ones except bit 1 (the reference oracle flips bit 1)
    cmds = []
    for q in range(n):
        cmds.append(Gate(G.H, [q], name="H"))
    cmds.append(Gate(G.X, [oracle], name="X"))
    cmds.append(Gate(G.H, [oracle], name="H"))
    zero_bits = [q for q in range(n) if not (marked >> q) & 1]
    for _ in range(iterations):
        for q in zero_bits:
            cmds.append(Gate(G.X, [q], name="X"))
        cmds.append(Gate(G.X, [oracle], list(range(n)), name="CnX"))
        for q in zero_bits:
            cmds.append(Gate(G.X, [q], name="X"))
        for q in range(n):
            cmds.append(Gate(G.H, [q], name="H"))
        for q in range(n):
            cmds.append(Gate(G.X, [q], name="X"))
        cmds.append(Gate(G.Z, [n - 1], list(range(n - 1)), name="CnZ", is_z=True))
        for q in range(n):
            cmds.append(Gate(G.X, [q], name="X"))
        for q in range(n):
            cmds.append(Gate(G.H, [q], name="H"))
        for q in range(n):
            cmds.append(Gate(G.Ph(math.pi / n), [q], name="Ph"))
    return n + 1, cmds


def inverse(cmds):
    """Dagger of a command list (for circuit * circuit^-1 identity checks)."""
    out = []
    for c in reversed(cmds):
        out.append(Gate(c.matrix.conj().T, list(c.qubits), list(c.controls), name=c.name + "^", is_z=c.is_z))
    return out


def run_shor(eng, N: int, a: int, n: int | None = None, verbose: bool = False):
    """C5: the quantum subroutine of Shor's algorithm as the reference example runs it
    (reference: examples/shor_mpi.py:45-105): an n-qubit register |1>, one phase-estimation qubit, 2n rounds of
    H / controlled MultiplyByConstantModN(a^(2^(2n-1-k)) mod N) / conditional R / H / Measure / conditional X.
    The reference decomposes the multiplication through projectq.libs.math (2n+3 qubits with ancillas); here it is
    emulated by the engine (ops.MultiplyByConstantModN -> emulate_math), so the circuit needs n + 1 qubits.
    `eng` is a HiQMainEngine.  Returns (candidate period r, the 2n measured bits)."""
    from . import ops
    if n is None:
        n = int(math.ceil(math.log(N, 2)))
    x = eng.allocate_qureg(n)
    eng.receive([Gate(G.X, [x[0]], name="X")])
    measurements = [0] * (2 * n)
    ctrl = eng.allocate_qubit()
    for k in range(2 * n):
        current_a = pow(a, 1 << (2 * n - 1 - k), N)
        cmds = [Gate(G.H, [ctrl], name="H"), ops.MultiplyByConstantModN(current_a, N, x, [ctrl])]
        for i in range(k):
            if measurements[i]:
                cmds.append(Gate(G.R(-math.pi / (1 << (k - i))), [ctrl], name="R"))
        cmds.append(Gate(G.H, [ctrl], name="H"))
        cmds.append(ops.Measure([ctrl]))
        eng.receive(cmds)
        eng.flush()
        measurements[k] = int(eng.measurements[ctrl])
        if measurements[k]:
            eng.receive([Gate(G.X, [ctrl], name="X")])
        if verbose:
            print(measurements[k], end="", flush=True)
    eng.receive([ops.Measure(list(x))])
    eng.flush()
    return period_from_bits(measurements, N), measurements


def period_from_bits(measurements, N: int) -> int:
    """Period candidate from the measured bits of the semi-classical phase estimation (reference:
    examples/shor_mpi.py:95-105): round k measured bit 2n-1-k of the phase, so measurements[2n-1-i] is the bit of weight
    2^-(i+1); the candidate is the denominator of the closest fraction with denominator < N.  The reference example sums
    the bits as floats; beyond 2n = 53 bits that drops the low bits the continued fraction needs (a 32-qubit run measures
    62), so the phase is kept as an exact rational here."""
    from fractions import Fraction
    m = len(measurements)
    y = Fraction(sum(int(measurements[m - 1 - i]) << (m - 1 - i) for i in range(m)), 1 << m)
    return y.limit_denominator(N - 1).denominator


def inverse_circuit(cmds):
    """The circuit that undoes `cmds`: reversed order, conjugate-transposed matrices (same targets / controls)."""
    out = []
    for c in reversed(cmds):
        m = np.asarray(c.matrix)
        out.append(Gate(m.conj().T.copy(), list(c.qubits), list(c.controls), name=getattr(c, "name", "") + "^-1",
                        is_z=bool(getattr(c, "is_z", False))))
    return out

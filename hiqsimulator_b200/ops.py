"""Minimal, ProjectQ-free command objects for the host pipeline.

They carry exactly what the reference backend reads from a ProjectQ ``Command``
(reference: hiq/projectq/backends/_sim/_simulator_mpi.py:416-494): a gate matrix, target ids,
control ids — or one of the meta operations Allocate / AllocateQureg / Deallocate / Measure /
Flush / MetaSwap (reference: hiq/projectq/ops/_gates.py:20-73).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GATE, ALLOCATE, ALLOCATE_QUREG, DEALLOCATE, MEASURE, FLUSH, METASWAP, MATH, TIME_EVOLUTION = range(9)


@dataclass
class Command:
    kind: int
    qubits: list = field(default_factory=list)     # target ids (GATE/MEASURE/ALLOCATE*/DEALLOCATE), pairs for METASWAP
    controls: list = field(default_factory=list)
    matrix: np.ndarray | None = None
    name: str = ""
    init: complex = 0                               # ALLOCATE_QUREG initial amplitude
    is_z: bool = False                              # ZGate (target/control roles may be exchanged)
    quregs: list = field(default_factory=list)      # MATH: one id list per register
    math: tuple = ()                                # MATH: ("add", a) | ("add_mod", a, N) | ("mul_mod", a, N) | ("fn", callable); TIME_EVOLUTION: (time, terms)

    @property
    def fast_forwarding(self) -> bool:
        """ProjectQ FastForwardingGate family: Measure, Flush, Deallocate, MetaSwap.  Math gates are emulated on
        the whole state (reference call site: _simulator_mpi.py:459-468), so everything pending runs first."""
        return self.kind in (MEASURE, FLUSH, DEALLOCATE, METASWAP, MATH, TIME_EVOLUTION)


def Gate(matrix, qubits, controls=(), name="", is_z=False):
    return Command(GATE, list(qubits), list(controls), np.asarray(matrix, dtype=np.complex128), name, 0, is_z)


def Allocate(qid):
    return Command(ALLOCATE, [qid], name="Allocate")


def AllocateQureg(ids, init=0):
    return Command(ALLOCATE_QUREG, list(ids), name="AllocateQureg", init=init)


def Deallocate(qid):
    return Command(DEALLOCATE, [qid], name="Deallocate")


def Measure(ids):
    return Command(MEASURE, list(ids), name="Measure")


def Flush():
    return Command(FLUSH, name="Flush")


def MetaSwap(pairs):
    return Command(METASWAP, list(pairs), name="MetaSwap")


# ProjectQ's math gates (projectq.libs.math: AddConstant, AddConstantModN, MultiplyByConstantModN and the generic
# BasicMathGate), emulated by the engine instead of being decomposed into adders (reference: the wrapper's
# BasicMathGate branch, _simulator_mpi.py:459-468, which the reference engine answers with "not supported")
def AddConstant(a, qureg, controls=()):
    return Command(MATH, [q for q in qureg], list(controls), name="AddConstant", quregs=[list(qureg)], math=("add", int(a)))


def AddConstantModN(a, N, qureg, controls=()):
    return Command(MATH, [q for q in qureg], list(controls), name="AddConstantModN", quregs=[list(qureg)],
                   math=("add_mod", int(a), int(N)))


def MultiplyByConstantModN(a, N, qureg, controls=()):
    return Command(MATH, [q for q in qureg], list(controls), name="MultiplyByConstantModN", quregs=[list(qureg)],
                   math=("mul_mod", int(a), int(N)))


def TimeEvolution(time, hamiltonian, qureg, controls=()):
    """ProjectQ's TimeEvolution(time, hamiltonian) | qureg: exp(-i time H), H = list of (term, coefficient) with term =
    sequence of (index into qureg, 'X'|'Y'|'Z') (a QubitOperator's .terms.items()).  Emulated on the whole state
    (reference call site: _simulator_mpi.py:469-475)."""
    terms = hamiltonian.terms.items() if hasattr(hamiltonian, "terms") else hamiltonian
    return Command(TIME_EVOLUTION, list(qureg), list(controls), name="TimeEvolution",
                   math=(float(time), [(list(t), c) for t, c in terms]))


def BasicMath(fn, quregs, controls=()):
    """fn maps the list of register values to the list of new values (BasicMathGate.get_math_function)"""
    return Command(MATH, [q for qr in quregs for q in qr], list(controls), name="BasicMathGate",
                   quregs=[list(qr) for qr in quregs], math=("fn", fn))

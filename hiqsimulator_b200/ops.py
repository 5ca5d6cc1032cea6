"""Minimal, ProjectQ-free command objects for the host pipeline.

They carry exactly what the reference backend reads from a ProjectQ ``Command``
(reference: hiq/projectq/backends/_sim/_simulator_mpi.py:416-494): a gate matrix, target ids,
control ids — or one of the meta operations Allocate / AllocateQureg / Deallocate / Measure /
Flush / MetaSwap (reference: hiq/projectq/ops/_gates.py:20-73).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GATE, ALLOCATE, ALLOCATE_QUREG, DEALLOCATE, MEASURE, FLUSH, METASWAP = range(7)


@dataclass
class Command:
    kind: int
    qubits: list = field(default_factory=list)     # target ids (GATE/MEASURE/ALLOCATE*/DEALLOCATE), pairs for METASWAP
    controls: list = field(default_factory=list)
    matrix: np.ndarray | None = None
    name: str = ""
    init: complex = 0                               # ALLOCATE_QUREG initial amplitude
    is_z: bool = False                              # ZGate (target/control roles may be exchanged)

    @property
    def fast_forwarding(self) -> bool:
        """ProjectQ FastForwardingGate family: Measure, Flush, Deallocate, MetaSwap."""
        return self.kind in (MEASURE, FLUSH, DEALLOCATE, METASWAP)


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

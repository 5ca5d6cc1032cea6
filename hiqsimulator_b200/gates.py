"""Gate matrices (numpy complex128) used by the ProjectQ-free harness, benches and tests.

ProjectQ is not available in this image; the reference obtains the same matrices from
``cmd.gate.matrix`` (reference: hiq/projectq/backends/_sim/_simulator_mpi.py:477).  Conventions
follow ProjectQ: Rx/Ry/Rz(theta) = exp(-i theta/2 sigma), R(phi) = diag(1, e^{i phi}), Ph(phi) = e^{i phi} I.
"""
from __future__ import annotations

import cmath
import math

import numpy as np

I2 = np.eye(2, dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2.0)
S = np.array([[1, 0], [0, 1j]], dtype=np.complex128)
T = np.array([[1, 0], [0, cmath.exp(0.25j * math.pi)]], dtype=np.complex128)
SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def Rx(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def Ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def Rz(t):
    return np.array([[cmath.exp(-0.5j * t), 0], [0, cmath.exp(0.5j * t)]], dtype=np.complex128)


def R(phi):
    return np.array([[1, 0], [0, cmath.exp(1j * phi)]], dtype=np.complex128)


def Ph(phi):
    return cmath.exp(1j * phi) * I2


def haar_unitary(dim: int, rng: np.random.Generator) -> np.ndarray:
    """Haar-random unitary: QR of a complex Ginibre matrix with the phase fix."""
    z = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diag(r)
    return q * (d / np.abs(d))

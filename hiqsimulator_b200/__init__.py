"""hiqsimulator_b200 — B200-native (sm_100a) state-vector engine behind HiQsimulator's
``_cppsim_mpi.SimulatorMPI`` operator API.

Only the hot path of the reference lives here: the CUDA kernels and the
C-ABI library (``csrc/`` -> ``libhiq_b200.so``, declared in ``include/hiq_b200.h``), the
pybind11 modules mirroring the reference's ``_cppsim_mpi`` / ``_sched_cpp`` and the thin
Python host mirror of the reference backend.  There is no CPU fallback: importing the
compute entry points without the built library raises.
"""
from ._lib import lib, load_library, LibraryMissing  # noqa: F401

__version__ = "0.1.0"

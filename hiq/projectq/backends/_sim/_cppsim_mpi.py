"""Import-path shim: ``hiq.projectq.backends._sim._cppsim_mpi`` -> the B200 engine.

The reference's Python backend does ``from ._cppsim_mpi import SimulatorMPI as SimulatorBackend``
(reference: hiq/projectq/backends/_sim/_simulator_mpi.py:39; the extension is built to that path by setup.py:87-94).
This module makes the same import resolve to ``hiqsimulator_b200._cppsim_mpi`` — same class name, constructor
``SimulatorMPI(seed, max_local, max_cluster_size)`` and methods.  The directories above it carry no ``__init__.py`` on
purpose (PEP 420 namespace portions): dropped next to the reference's own ``hiq`` tree, or copied into it, the file
replaces the reference's compiled module and nothing else.  World set-up replaces ``mpirun``: call
``hiqsimulator_b200.world.init_world()`` once per process (one process per GPU, launched by torchrun)."""
from hiqsimulator_b200._cppsim_mpi import *  # noqa: F401,F403
from hiqsimulator_b200._cppsim_mpi import SimulatorMPI, init_world, unique_id  # noqa: F401

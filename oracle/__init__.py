"""CPU oracle for the HiQsimulator `SimulatorMPI` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product (``hiqsimulator_b200``) never imports it and has
no CPU fallback.

Contents
--------
``statevec.py``   numpy restatement of the reference engine (virtual ranks).
``sched.py``      pure-Python restatement of the reference `_sched_cpp` schedulers
                  and of the ProjectQ-free GreedyScheduler driver loop.
``ref.py``        loader / multi-process launcher for ``oracle/_ref`` — the
                  UNMODIFIED reference sources compiled against ``shim/``.
``Makefile``      recipe that builds ``oracle/_ref`` from ``/root/reference``.

Parity pinning: the restatements are checked in ``tests/`` against (a) the
known answers of the reference's own test-suite
(``hiq/projectq/backends/_sim/_simulator_mpi_test.py``), (b) golden vectors in
``tests/golden`` that were produced by ``oracle/_ref`` (generator:
``tests/golden/make_golden.py``) and (c) ``oracle/_ref`` itself whenever it is
present.
"""

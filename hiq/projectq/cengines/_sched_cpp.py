"""Import-path shim: ``hiq.projectq.cengines._sched_cpp`` -> the re-implemented host scheduler.

The reference does ``from ._sched_cpp import SwapScheduler, ClusterScheduler`` (reference:
hiq/projectq/cengines/__init__.py:15, _greedyscheduler.py:26; built to that path by setup.py).  Same two classes,
constructor signatures and ``ScheduleSwap()`` / ``ScheduleCluster()`` results (bit-exact, tests/test_scheduler.py), plus
``GreedyPlanner``.  No ``__init__.py`` above it on purpose (PEP 420 namespace portion)."""
from hiqsimulator_b200._sched_cpp import *  # noqa: F401,F403
from hiqsimulator_b200._sched_cpp import ClusterScheduler, GreedyPlanner, SwapScheduler  # noqa: F401

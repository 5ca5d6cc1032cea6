"""The reference's GreedyScheduler stage / cluster loop, restated without ProjectQ (TEST INFRASTRUCTURE).

Behavioural spec: reference hiq/projectq/cengines/_greedyscheduler.py:95-242 — the Python loop that calls
``ClusterScheduler`` / ``SwapScheduler`` once per cluster / stage.  The product drives the same decisions through one C++
object (``_sched_cpp.GreedyPlanner``, hiqsimulator_b200/cengines.py); this module keeps the reference's loop for
  * the cross-check ``tests/test_scheduler.py::test_planner_equals_python_loop`` (logs, maps, every emitted descriptor), and
  * driving the UNMODIFIED reference scheduler (oracle/_ref/_sched_cpp, which has no planner) in bench.py's reference arm.
Nothing under hiqsimulator_b200/ imports it."""
from __future__ import annotations

import time

from hiqsimulator_b200 import ops
from hiqsimulator_b200.cengines import GreedyScheduler as _Base


class ReferenceLoopScheduler(_Base):
    """same wiring, caching and ``receive`` as the product class; ``_force_scheduling`` is the reference's Python loop
    over ``sched_module.ClusterScheduler`` / ``SwapScheduler``"""

    def __init__(self, supremacy_circuit=False, num_splits=10 ** 6, cluster_size=4, sched_module=None):
        super().__init__(supremacy_circuit, num_splits, cluster_size, sched_module)

    # -- reference: _greedyscheduler.py:95-112
    def _prepare_ctrlz(self):
        local_qubits = self.backend.get_local_qubits_ids()
        global_qubits = self.backend.get_global_qubits_ids()
        for cmd in self._cmd_list:
            if cmd.is_z:
                assert len(cmd.qubits) == 1
                if cmd.qubits[0] in global_qubits:
                    for i, c in enumerate(cmd.controls):
                        if c in local_qubits:
                            cmd.controls[i], cmd.qubits[0] = cmd.qubits[0], cmd.controls[i]
                            break

    # -- reference: _greedyscheduler.py:119-137
    def _call_cluster_scheduler(self):
        self._prepare_ctrlz()
        local_qubits = self.backend.get_local_qubits_ids()
        global_qubits = self.backend.get_global_qubits_ids()
        while True:
            gate, gate_ctrl, gate_diag = self._get_commands()
            t0 = time.perf_counter()
            cs = self._sched.ClusterScheduler(gate, gate_ctrl, gate_diag, local_qubits, global_qubits, self.CLUSTER_SIZE)
            avail = cs.ScheduleCluster()
            self.cluster_seconds += time.perf_counter() - t0
            if len(avail) == 0:
                return
            self.n_clusters += 1
            self.log.append(("cluster", [self._cmd_list[i].uid for i in avail]))
            for i in avail:
                self.send([self._cmd_list[i]])
            self.send([ops.Flush()])
            for i in reversed(sorted(avail)):
                del self._cmd_list[i]

    # -- reference: _greedyscheduler.py:175-193
    def _call_swap_scheduler(self):
        local_qubits = self.backend.get_local_qubits_ids()
        gate, gate_ctrl, gate_diag = self._get_commands()
        t0 = time.perf_counter()
        new_locals = self._sched.SwapScheduler(gate, gate_ctrl, gate_diag, self.NUM_SPLITS, len(local_qubits), True).ScheduleSwap()
        if len(new_locals) == 0:
            new_locals = self._sched.SwapScheduler(gate, gate_ctrl, gate_diag, self.NUM_SPLITS, len(local_qubits), False).ScheduleSwap()
        self.swap_seconds += time.perf_counter() - t0
        g_to_l = sorted(set(new_locals) - set(local_qubits))
        l_to_g = []
        if len(g_to_l) > 0:
            lst = sorted(set(local_qubits) - set(new_locals))
            assert len(lst) >= len(g_to_l)
            l_to_g = lst[:len(g_to_l)]
        return g_to_l, l_to_g

    # -- reference: _greedyscheduler.py:203-242
    def _force_scheduling(self):
        if len(self._cmd_list) == 0:
            return
        self._check_commands()
        if self._supremacy_circuit:
            self._remove_ending_cz()
        if not self._was_scheduling:
            self._was_scheduling = True
            ids_list = list(self.backend.get_qubits_ids())
            g_to_l, l_to_g = self._call_swap_scheduler()
            for i in range(len(l_to_g)):
                p1 = ids_list.index(g_to_l[i])
                p2 = ids_list.index(l_to_g[i])
                ids_list[p1], ids_list[p2] = ids_list[p2], ids_list[p1]
            self.backend.set_qubits_perm(ids_list)
            self.log.append(("perm", list(ids_list)))
        self._call_cluster_scheduler()
        while len(self._cmd_list) > 0:
            g_to_l, l_to_g = self._call_swap_scheduler()
            assert len(g_to_l) > 0
            pairs = []
            for i in range(len(g_to_l)):
                pairs += [g_to_l[i], l_to_g[i]]
            self.n_swaps += 1
            self.log.append(("swap", list(pairs)))
            self.send([ops.MetaSwap(pairs)])
            self._call_cluster_scheduler()
        assert len(self._cmd_list) == 0

"""GreedyScheduler: the engine that caches gates and hands them to the backend stage by stage, cluster by cluster.

Behavioural spec: reference hiq/projectq/cengines/_greedyscheduler.py:95-265.  The stage / cluster loop
itself (which cluster next, which qubits swap) runs inside `_sched_cpp.GreedyPlanner` (csrc/sched.cpp), one step per
`next()`, on a host thread that runs ahead of the device; the reference's Python loop, restated, lives in
oracle/greedy_loop.py as a cross-check and as the driver of the unmodified reference scheduler.
Gates are cached; Allocate / AllocateQureg / fast-forwarding commands force scheduling:
  first time only, the SwapScheduler picks the initial local set and the backend is *relabelled*
  (set_qubits_perm, no data motion); then ClusterScheduler is asked repeatedly which cached gates
  form the next cluster (sent followed by one Flush); when nothing is schedulable a MetaSwap with
  pairs [g->l id, l->g id, ...] starts the next stage.
Every gate is reported to the schedulers as non-diagonal (reference: _greedyscheduler.py:28).
"""
from __future__ import annotations

import time

from . import ops


class GreedyScheduler:
    def __init__(self, supremacy_circuit=False, num_splits=10 ** 6, cluster_size=4, sched_module=None, prefetch=True):
        if sched_module is None:
            from . import _sched_cpp as sched_module
        self._sched = sched_module
        self._cmd_list = []
        self._was_scheduling = False
        self._supremacy_circuit = supremacy_circuit
        self._prefetch = prefetch        # planner steps are computed ahead of the device on a host thread
        self.NUM_SPLITS = num_splits
        self.CLUSTER_SIZE = cluster_size
        self._deallocations_cache = []
        self.backend = None
        self.next_engine = None
        # instrumentation (host seconds spent inside the C++ schedulers, emitted schedule shape)
        self.cluster_seconds = 0.0
        self.swap_seconds = 0.0
        self.n_clusters = 0
        self.n_swaps = 0
        self.log = []  # ("cluster", [ids of gates...]) / ("swap", pairs) / ("perm", ids)

    # -- wiring --------------------------------------------------------------------------------
    def send(self, cmds):
        self.next_engine.receive(cmds)

    def _get_commands(self):
        return ([list(c.qubits) for c in self._cmd_list], [list(c.controls) for c in self._cmd_list],
                [False] * len(self._cmd_list))

    # -- reference: _greedyscheduler.py:151-173
    def _remove_ending_cz(self):
        i = len(self._cmd_list) - 1
        used = set()
        while i >= 0:
            cmd = self._cmd_list[i]
            allq = list(cmd.controls) + list(cmd.qubits)
            bad = True
            if cmd.is_z:
                if any(q in used for q in allq):
                    bad = False
            else:
                bad = False
            if bad:
                self._cmd_list.pop(i)
            else:
                used.update(allq)
            i -= 1

    # -- reference: _greedyscheduler.py:195-201
    def _check_commands(self):
        locals_size = len(self.backend.get_local_qubits_ids())
        for gate in self._get_commands()[0]:
            if locals_size < len(gate):
                raise Exception("Can't apply {}-qubits gate (only {} local qubits)".format(len(gate), locals_size))
            if len(gate) > 5:
                raise Exception("Can't apply {}-qubits gate (no more that 5 qubits allowed)".format(len(gate)))

    # -- the same loop driven by the C++ planner (`_sched_cpp.GreedyPlanner`, csrc/sched.cpp): one next() per step,
    #    so every cluster reaches the device while the following one is being searched
    def _force_scheduling_planned(self):
        cmds = self._cmd_list
        gate, gate_ctrl, _ = self._get_commands()
        planner = self._sched.GreedyPlanner(gate, gate_ctrl, [bool(c.is_z) for c in cmds], self.backend.get_local_qubits_ids(),
                                            self.backend.get_global_qubits_ids(), self.CLUSTER_SIZE, self.NUM_SPLITS,
                                            not self._was_scheduling)
        self._was_scheduling = True
        if self._prefetch:
            # the planner runs ahead on a host thread (next() releases the GIL): while the device works through the
            # clusters of a stage, the search for the next stage's local set is already under way, so neither the
            # swap search nor the cluster search sits between two launches.  Steps are consumed in emission order.
            import queue
            import threading
            steps = queue.Queue(maxsize=256)

            def produce():
                try:
                    while True:
                        step = planner.next()
                        steps.put(step)
                        if step[0] == 0:
                            return
                except BaseException as e:  # surfaces in the consumer
                    steps.put((-1, e))
            worker = threading.Thread(target=produce, name="hiq-planner", daemon=True)
            worker.start()
            next_step = steps.get
        else:
            worker = None
            next_step = planner.next
        while True:
            kind, data = next_step()
            if kind == -1:
                raise data
            if kind == 0:
                break
            if kind == 1:
                self.backend.set_qubits_perm(list(data))
                self.log.append(("perm", list(data)))
            elif kind == 2:
                cmd = cmds[data[0]]
                cmd.controls[data[1]], cmd.qubits[0] = cmd.qubits[0], cmd.controls[data[1]]
            elif kind == 3:
                self.n_clusters += 1
                self.log.append(("cluster", [cmds[i].uid for i in data]))
                for i in data:
                    self.send([cmds[i]])
                self.send([ops.Flush()])
            elif kind == 4:
                self.n_swaps += 1
                self.log.append(("swap", list(data)))
                self.send([ops.MetaSwap(list(data))])
        if worker is not None:
            worker.join()
        self.cluster_seconds += planner.cluster_seconds()
        self.swap_seconds += planner.swap_seconds()
        self._cmd_list = []

    # -- reference: _greedyscheduler.py:203-242 (the stage / cluster loop itself runs inside the planner)
    def _force_scheduling(self):
        if len(self._cmd_list) == 0:
            return
        self._check_commands()
        if self._supremacy_circuit:
            self._remove_ending_cz()
            if len(self._cmd_list) == 0:
                return
        self._force_scheduling_planned()

    def _send_deallocations(self):
        for c in sorted(self._deallocations_cache, key=lambda cmd: cmd.qubits[0], reverse=True):
            self.send([c])
        del self._deallocations_cache[:]

    # -- reference: _greedyscheduler.py:250-265
    def receive(self, command_list):
        for cmd in command_list:
            if not hasattr(cmd, "uid"):
                cmd.uid = GreedyScheduler._next_uid
                GreedyScheduler._next_uid += 1
            if cmd.kind == ops.DEALLOCATE:
                self._deallocations_cache.append(cmd)
            elif cmd.kind in (ops.ALLOCATE, ops.ALLOCATE_QUREG) or cmd.fast_forwarding:
                self._force_scheduling()
                self._send_deallocations()
                self.send([cmd])
            else:
                if len(self._deallocations_cache) > 0:
                    self._force_scheduling()
                    self._send_deallocations()
                self._cmd_list.append(cmd)

    _next_uid = 0


class HiQMainEngine:
    """Thin stand-in for the reference's HiQMainEngine(backend, [GreedyScheduler()]) wiring
    (reference: hiq/projectq/cengines/_hiq_main_engine.py:22-81): scheduler -> backend."""

    def __init__(self, backend, engine_list=None):
        self.backend = backend
        self.scheduler = (engine_list or [GreedyScheduler()])[-1]
        self.scheduler.backend = backend
        self.scheduler.next_engine = backend
        backend.main_engine = self
        self._next_id = 0
        self.measurements = {}

    def allocate_qubit(self):
        q = self._next_id
        self._next_id += 1
        self.receive([ops.Allocate(q)])
        return q

    def allocate_qureg(self, n, init=0.0):
        ids = list(range(self._next_id, self._next_id + n))
        self._next_id += n
        self.receive([ops.AllocateQureg(ids, init)])
        return ids

    def receive(self, cmds):
        self.scheduler.receive(cmds)

    def flush(self):
        self.receive([ops.Flush()])

    def set_measurement_result(self, qid, value):
        self.measurements[qid] = bool(value)

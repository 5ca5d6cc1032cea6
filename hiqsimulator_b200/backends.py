"""SimulatorMPI: host-side mirror of the reference's ProjectQ backend engine, without ProjectQ.

Same constructor keywords, query methods and command handling as
reference hiq/projectq/backends/_sim/_simulator_mpi.py:47-513, driving the B200 engine through
the `_cppsim_mpi` pybind module (and therefore through the C ABI).  `cheat()` gathers the rank
slabs with torch.distributed instead of mpi4py.
"""
from __future__ import annotations

import random

import numpy as np

from . import ops


class SimulatorMPI:
    def __init__(self, gate_fusion=False, rnd_seed=None, num_local_qubits=33, max_fused_qubits=4, backend_class=None):
        if rnd_seed is None:
            rnd_seed = random.randint(0, 4294967295)
        if backend_class is None:
            # the B200 engine; imported here so that a caller who brings its own engine class (the compiled reference in
            # bench.py's reference arm, the numpy oracle in the tests) never loads this repository's native modules
            from . import _cppsim_mpi as _M
            backend_class = _M.SimulatorMPI
        cls = backend_class
        self._simulator = cls(rnd_seed, num_local_qubits, max_fused_qubits)
        self._gate_fusion = gate_fusion
        self.main_engine = None
        self.h2d_bytes = 0  # gate-matrix bytes handed to the engine (bench accounting)

    # -- queries (reference: _simulator_mpi.py:148-346) ------------------------------------------
    def get_probability(self, bit_string, qureg):
        bit_string = [bool(int(b)) for b in bit_string]
        return self._simulator.get_probability(bit_string, list(qureg))

    def get_amplitude(self, bit_string, qureg):
        bit_string = [bool(int(b)) for b in bit_string]
        return self._simulator.get_amplitude(bit_string, list(qureg))

    def collapse_wavefunction(self, qureg, values):
        return self._simulator.collapse_wavefunction(list(qureg), [bool(int(v)) for v in values])

    def cheat_local(self):
        return self._simulator.cheat_local()

    def cheat(self):
        """(id2pos, full state vector) — concatenation of the rank slabs on every rank (reference: :348-380 does an
        MPI Allgather of cheat_local; here the engine gathers over NCCL)."""
        if hasattr(self._simulator, "cheat"):
            id2pos, vec = self._simulator.cheat()
            return id2pos, np.asarray(vec)
        return self.cheat_local()  # single-rank stand-ins (the compiled reference, the numpy oracle)

    # reference: _simulator_mpi.py:148-223 — qubit_operator is a list of (term, coefficient) with
    # term = sequence of (index into qureg, 'X'|'Y'|'Z') (a projectq QubitOperator's .terms.items())
    @staticmethod
    def _terms(qubit_operator, num_qubits):
        items = qubit_operator.terms.items() if hasattr(qubit_operator, "terms") else qubit_operator
        operator = [(list(term), coeff) for (term, coeff) in items]
        for term, _ in operator:
            if len(term) and max(t[0] for t in term) >= num_qubits:
                raise Exception("qubit_operator acts on more qubits than contained in the qureg.")
        return operator

    def get_expectation_value(self, qubit_operator, qureg):
        return self._simulator.get_expectation_value(self._terms(qubit_operator, len(qureg)), list(qureg))

    def apply_qubit_operator(self, qubit_operator, qureg):
        return self._simulator.apply_qubit_operator(self._terms(qubit_operator, len(qureg)), list(qureg))

    def set_wavefunction(self, wavefunction, qureg):
        """reference: _simulator_mpi.py:279-305"""
        self._simulator.set_wavefunction(np.asarray(wavefunction, dtype=np.complex128), list(qureg))

    def get_qubits_ids(self):
        return self._simulator.get_qubits_ids()

    def get_local_qubits_ids(self):
        return self._simulator.get_local_qubits_ids()

    def get_global_qubits_ids(self):
        return self._simulator.get_global_qubits_ids()

    def set_qubits_perm(self, ids):
        self._simulator.set_qubits_perm(list(ids))

    # -- command handling (reference: _simulator_mpi.py:416-513) -----------------------------------
    def _handle(self, cmd):
        if cmd.kind == ops.FLUSH:
            pass
        elif cmd.kind == ops.METASWAP:
            self._simulator.swap_qubits(list(cmd.qubits))
        elif cmd.kind == ops.MEASURE:
            assert len(cmd.controls) == 0
            out = self._simulator.measure_qubits(list(cmd.qubits))
            for q, b in zip(cmd.qubits, out):
                if self.main_engine is not None:
                    self.main_engine.set_measurement_result(q, b)
            cmd.result = list(out)
        elif cmd.kind == ops.ALLOCATE:
            self._simulator.allocate_qubit(cmd.qubits[0])
        elif cmd.kind == ops.ALLOCATE_QUREG:
            self._simulator.allocate_qureg(list(cmd.qubits), cmd.init)
        elif cmd.kind == ops.DEALLOCATE:
            self._simulator.deallocate_qubit(cmd.qubits[0])
        elif cmd.kind == ops.MATH:  # reference: the BasicMathGate branch, _simulator_mpi.py:459-468
            ctrls = list(cmd.controls)
            if cmd.math[0] == "fn":
                self._simulator.emulate_math(cmd.math[1], [list(qr) for qr in cmd.quregs], ctrls)
            elif cmd.math[0] == "add":
                self._simulator.emulate_math_add_constant(cmd.math[1], list(cmd.quregs[0]), ctrls)
            elif cmd.math[0] == "add_mod":
                self._simulator.emulate_math_add_constant_modN(cmd.math[1], cmd.math[2], list(cmd.quregs[0]), ctrls)
            elif cmd.math[0] == "mul_mod":
                self._simulator.emulate_math_multiply_by_constant_modN(cmd.math[1], cmd.math[2], list(cmd.quregs[0]), ctrls)
            else:
                raise Exception("unknown math gate %r" % (cmd.math,))
        elif cmd.kind == ops.TIME_EVOLUTION:  # reference: the TimeEvolution branch, _simulator_mpi.py:469-475
            t, op = cmd.math
            self._simulator.emulate_time_evolution(self._terms(op, len(cmd.qubits)), t, list(cmd.qubits), list(cmd.controls))
        elif cmd.kind == ops.GATE and len(cmd.matrix) <= 2 ** 5:
            if not 2 ** len(cmd.qubits) == len(cmd.matrix):
                raise Exception("Simulator: Error applying {} gate: {}-qubit gate applied to {} qubits.".format(
                    cmd.name, int(np.log2(len(cmd.matrix))), len(cmd.qubits)))
            self.h2d_bytes += cmd.matrix.nbytes
            if hasattr(self._simulator, "apply_controlled_matrix"):
                self._simulator.apply_controlled_matrix(cmd.matrix, list(cmd.qubits), list(cmd.controls))
            else:  # the reference binding only takes nested lists (reference: _simulator_mpi.py:485-488)
                self._simulator.apply_controlled_gate(cmd.matrix.tolist(), list(cmd.qubits), list(cmd.controls))
            if not self._gate_fusion:
                self._simulator.run()
        else:
            raise Exception("This simulator only supports controlled k-qubit gates with k < 6!")

    def receive(self, command_list):
        for cmd in command_list:
            if cmd.kind == ops.FLUSH or cmd.fast_forwarding:
                self._simulator.run()  # flush gate --> run all saved gates
            self._handle(cmd)

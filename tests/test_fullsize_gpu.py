"""Parity at the sizes BASELINE.json names (30 and 33 qubits on one B200), where no host can hold the state for a
full diff: size-independent properties of the same scheduled pipeline —
  * QFT of a basis state has a closed form: sampled amplitudes, single-qubit marginals, measurement;
  * a circuit followed by its inverse returns |0...0>.
The property checks themselves are pinned on the CPU: tests/test_host_logic.py runs the same functions against the numpy
oracle at 12 qubits.  Amplitude tolerance 1e-12 absolute (BASELINE.json north_star)."""
import copy
import gc

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make_pipeline(backend_class, n, cluster=4, seed=1):
    from hiqsimulator_b200 import backends, cengines
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=seed, num_local_qubits=n, max_fused_qubits=cluster,
                               backend_class=backend_class)
    return be, cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=cluster)])


def qft_input(n):
    return 0x5A5A5A5A5A5A5A5A & ((1 << n) - 1)  # circuits.qft_circuit's default input state


def check_qft_closed_form(backend_class, n, samples):
    """worst |amplitude - closed form| over sampled basis states, worst |marginal - 1/2|, P(measured outcome)"""
    from hiqsimulator_b200 import circuits, ops
    be, eng = make_pipeline(backend_class, n)
    try:
        _, cmds = circuits.qft_circuit(n)
        eng.allocate_qureg(n)
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        ids = list(range(n))
        rng = np.random.default_rng(1000 + n)
        worst = 0.0
        for _ in range(samples):
            y = int(rng.integers(0, 1 << n))
            got = be.get_amplitude([(y >> q) & 1 for q in range(n)], ids)
            worst = max(worst, abs(got - circuits.qft_expected_amplitude(n, qft_input(n), y)))
        marg = max(abs(be.get_probability([0], [q]) - 0.5) for q in (0, n // 2, n - 1))
        eng.receive([ops.Measure(ids)])
        bits = [int(eng.measurements[q]) for q in ids]
        p_after = be.get_probability(bits, ids)
        return worst, marg, p_after
    finally:
        be.main_engine = None  # break the engine <-> backend cycle so that the slab is released now
        del eng, be
        gc.collect()


def check_circuit_then_inverse(backend_class, n, depth):
    """|<0...0|psi> - 1| and |P(0...0) - 1| after a random circuit followed by its inverse"""
    from hiqsimulator_b200 import circuits
    be, eng = make_pipeline(backend_class, n)
    try:
        _, cmds = circuits.random_circuit(n, depth=depth)
        eng.allocate_qureg(n)
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        eng.receive(circuits.inverse_circuit(cmds))
        eng.flush()
        ids = list(range(n))
        amp = be.get_amplitude([0] * n, ids)
        p = be.get_probability([0] * n, ids)
        return abs(amp - 1.0), abs(p - 1.0)
    finally:
        be.main_engine = None
        del eng, be
        gc.collect()


def _need_gib(n):
    import torch
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(0, 1, b"", 0, 0)
    need = 16.0 * (1 << n) / 2 ** 30 + 2.0
    free = torch.cuda.mem_get_info()[0] / 2 ** 30
    if free < need:
        pytest.skip("needs %.0f GiB of free device memory, %.0f available" % (need, free))


@pytest.mark.parametrize("n", [30, 33])
def test_qft_closed_form_at_full_size(n):
    _need_gib(n)
    worst, marg, p_after = check_qft_closed_form(None, n, 2048)
    assert worst <= TOL and marg <= TOL and abs(p_after - 1.0) <= TOL, (worst, marg, p_after)


@pytest.mark.parametrize("n", [30, 33])
def test_random_circuit_then_inverse_at_full_size(n):
    _need_gib(n)
    # BASELINE.json configs[1]: 30 qubits, depth 20; north_star: "a 33-qubit random circuit on 1 B200"
    d_amp, d_p = check_circuit_then_inverse(None, n, 20)
    assert d_amp <= TOL and d_p <= TOL, (d_amp, d_p)

"""The benchmark pipelines of BASELINE.json configs[0] (Grover, examples/grover_mpi.py) and configs[4] (Shor by emulation,
examples/shor_mpi.py) through the whole host layer (HiQMainEngine -> GreedyScheduler -> backend) WITHOUT a GPU:
the workload generators and the Python mirror of the reference's backend / scheduler are pinned on the compiled reference
engine (Grover: the reference's own CPU path, C1) and on the numpy oracle (Shor: the reference engine cannot emulate
math gates, SimulatorMPI.hpp:217-225).  tests/test_engine_gpu.py runs the same pipelines on the B200 against the oracle."""
import copy
import math

import numpy as np
import pytest

from oracle import statevec


def _pipeline(backend_class, seed, L, cluster):
    from hiqsimulator_b200 import backends, cengines
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=seed, num_local_qubits=L, max_fused_qubits=cluster, backend_class=backend_class)
    return be, cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=cluster)])


@pytest.mark.parametrize("n_search,iters,cluster", [(9, 6, 4), (11, 3, 4), (10, 4, 3)])
def test_grover_pipeline_on_the_compiled_reference_engine(n_search, iters, cluster):
    """C1: the same command stream, scheduled by this repository's planner, on the UNMODIFIED reference engine and on the
    numpy oracle: slot maps and measured bits identical, amplitudes within 1e-12, and the textbook success probability
    sin^2((2k + 1) asin(2^(-n/2))) of the marked state."""
    from hiqsimulator_b200 import circuits, ops
    from oracle import ref
    if not ref.have_ref():
        pytest.skip("oracle/_ref is not built")
    n, cmds = circuits.grover_circuit(n_search, iters)
    out = []
    for cls in (ref.load_ref_sim().SimulatorMPI, statevec.SimulatorMPI):
        be, eng = _pipeline(cls, 1, n, cluster)
        eng.allocate_qureg(n)
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        id2pos, vec = be.cheat()
        vec = np.asarray(vec, dtype=np.complex128).copy()
        eng.receive([ops.Measure(list(range(n)))])
        out.append((dict(id2pos), [bool(eng.measurements[q]) for q in range(n)], vec))
    assert out[0][:2] == out[1][:2]
    assert np.abs(out[0][2] - out[1][2]).max() <= 1e-12
    marked = ((1 << n_search) - 1) & ~0b10
    pos = out[0][0]
    idx = np.arange(1 << n)
    data = np.zeros_like(idx)
    for q in range(n_search):
        data |= ((idx >> pos[q]) & 1) << q
    p = float((np.abs(out[0][2][data == marked]) ** 2).sum())
    assert abs(p - math.sin((2 * iters + 1) * math.asin(2.0 ** (-n_search / 2))) ** 2) <= 1e-9
    # the measured data bits spell the marked element when it dominates
    if p > 0.9:
        assert sum(int(out[0][1][q]) << q for q in range(n_search)) == marked


@pytest.mark.parametrize("N,a,order", [(15, 7, 4), (15, 2, 4), (21, 2, 6), (35, 4, 6)])
def test_shor_pipeline_finds_divisors_of_the_order(N, a, order):
    """C5 at test size on the numpy oracle: semi-classical phase estimation with emulated modular multiplication returns
    the denominator of k / r in lowest terms — a divisor of the order of a modulo N for every seed, the order itself for some"""
    from hiqsimulator_b200 import circuits
    assert pow(a, order, N) == 1 and all(pow(a, d, N) != 1 for d in range(1, order))
    n = int(math.ceil(math.log(N, 2)))
    found = set()
    for seed in range(8):
        be, eng = _pipeline(statevec.SimulatorMPI, seed, n + 1, 3)
        r, bits = circuits.run_shor(eng, N, a, n)
        assert len(bits) == 2 * n and order % r == 0, (seed, r)
        found.add(r)
        # the data register ends in a basis state that is a power of a (the measured x = a^j mod N)
        x = sum(int(eng.measurements[q]) << q for q in range(n))
        assert x in {pow(a, j, N) for j in range(order)}
    assert order in found

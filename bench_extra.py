"""Secondary bench line: BASELINE.json configs[4] — Shor's algorithm with the controlled modular multiplication emulated
by the engine (reference structure: examples/shor_mpi.py:45-105), n data qubits + 1 phase-estimation qubit, 2n rounds of
H / controlled MultiplyByConstantModN / conditional R / H / Measure / conditional X.  Called from bench.py
(`--circuit shor [--qubits Q]`, default 32 qubits = a 31-bit modulus; launch under torchrun for N > 1).

Reported: seconds per circuit (metric, lower is better), the per-round split — permutation pass (emulate_math), the
measurement (the reference's three-pass algorithm, SimulatorMPI.cpp:897-1008) and the gates — and the measured period.
Every step checks its own result: the continued-fraction step returns the order of a modulo N or one of its divisors, so
the candidate r must divide lambda(N) = lcm(p-1, q-1) in every run; a^r = 1 (mod N) holds in the runs that hit the order
itself (the line reports both counts)."""
from __future__ import annotations

import gc
import json
import math
import os
import time

# register width -> (p, q, a): N = p * q is a semiprime of exactly that many bits, a is coprime to N
MODULI = {31: (32771, 32779, 7), 29: (16411, 16417, 7), 27: (8209, 8219, 7), 25: (4099, 4111, 7), 23: (2053, 2063, 7), 21: (1031, 1033, 7), 19: (521, 523, 7), 17: (257, 263, 7), 15: (131, 137, 7), 13: (67, 71, 7), 11: (37, 41, 7), 9: (17, 19, 7), 5: (3, 7, 2)}


def main(args):
    import torch
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, circuits, world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this engine has no CPU path)")
    rank, size = world.init_world(M.FLAG_TIMING)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    total = args.qubits or 32
    n = total - 1
    if n not in MODULI:
        raise SystemExit("shor: pick --qubits from %s" % sorted(k + 1 for k in MODULI))
    p, q, a = MODULI[n]
    N = p * q
    lam = (p - 1) * (q - 1) // math.gcd(p - 1, q - 1)  # the order of a divides lambda(N)
    assert N.bit_length() == n and math.gcd(a, N) == 1
    g = size.bit_length() - 1
    L = total - g
    steps = max(1, args.steps)
    per_step, found, consistent = [], 0, 0
    split = {"measure_s": 0.0, "permutation_and_gates_s": 0.0}
    rounds = 2 * n
    last = None
    for it in range(args.warmup + steps):
        world.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1234 + it, num_local_qubits=L, max_fused_qubits=4)
        eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
        r, bits = circuits.run_shor(eng, N, a, n)
        be._simulator.synchronize()
        world.barrier()
        dt = time.perf_counter() - t0
        st = be._simulator.stats()
        ok = pow(a, r, N) == 1
        divides = lam % r == 0  # continued fractions return the order or one of its divisors (numpy oracle: 12 of 12 runs)
        if it >= args.warmup:
            per_step.append(dt)
            found += int(ok)
            consistent += int(divides)
            split["measure_s"] += st["measures_s"]
            split["permutation_and_gates_s"] += dt - st["measures_s"]
        last = {"period_candidate": r, "a_pow_r_is_1": ok, "divides_lambda_N": divides, "measure_calls": rounds + 1}
        be.main_engine = None
        del eng, be
        gc.collect()
    t = torch.tensor([sum(per_step)], dtype=torch.float64, device="cuda")
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item()) / len(per_step)
    if rank == 0:
        line = {"metric": "shor_circuit_seconds", "value": sec, "unit": "s", "n_gpus": size, "steps": steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "shor-%d" % total, "qubits": total, "local_qubits": L, "modulus": N, "base": a, "rounds": rounds,
                           "math": "controlled MultiplyByConstantModN emulated as one permutation pass (reference example decomposes it)"},
                "per_round_ms": {"total": 1e3 * sec / rounds, "measure": 1e3 * split["measure_s"] / len(per_step) / (rounds + 1),
                                 "permutation_and_gates": 1e3 * split["permutation_and_gates_s"] / len(per_step) / rounds},
                "period_found_in": "%d of %d runs" % (found, len(per_step)),
                "candidate_divides_the_order_bound_in": "%d of %d runs" % (consistent, len(per_step)), "last_run": last}
        print(json.dumps(line))

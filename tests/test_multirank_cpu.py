"""World-size-2 run of the N>1 plumbing on CPU (gloo): torchrun launches two ranks, each builds a
dry-run engine for its rank, the descriptor traces are gathered on rank 0 and replayed against
the reference's golden state."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,R", [("r2_q10_gates", 2), ("r2_q9", 2), ("r4_q12_gates", 4)])
def test_dry_run_world(name, R):
    from torchrun_util import run_torchrun
    res = run_torchrun(R, os.path.join(HERE, "mp_worker.py"), [name, "dry"], env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=300)
    assert res.returncode == 0 and "MP_WORKER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("name,R", [("ops:8:21", 2), ("ops:9:22", 4)])
def test_dry_run_world_operator_calls(name, R):
    """get_expectation_value / apply_qubit_operator / emulate_math / set_wavefunction host logic, one dry-run engine
    per gloo rank, traces replayed on rank 0 against the numpy oracle"""
    from torchrun_util import run_torchrun
    res = run_torchrun(R, os.path.join(HERE, "mp_worker.py"), [name, "dry"], env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=300)
    assert res.returncode == 0 and "MP_WORKER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("name,R", [("swapskew:10:5", 4)])
def test_dry_run_world_swaps_with_changing_peers(name, R):
    """the script of the multi-GPU skew test (exchanges on alternating global bits, half of the ranks held back): its
    descriptor traces, replayed on rank 0, give the oracle's state — the script and its expectation are pinned on the CPU"""
    from torchrun_util import run_torchrun
    res = run_torchrun(R, os.path.join(HERE, "mp_worker.py"), [name, "dry"], env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=300)
    assert res.returncode == 0 and "MP_WORKER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]

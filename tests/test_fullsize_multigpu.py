"""Parity of the multi-GPU engine at size (BASELINE.json "35q@8"):
  * full-size properties on R GPUs (QFT-(32 + log2 R) closed form over 2048 amplitudes fetched across ranks, marginals,
    post-measurement state; random circuit followed by its inverse) — exercises multi-chunk peer-mapped slabs, the
    in-place NVLink swap at L = 32 and the reductions;
  * DIRECT diff against the compiled, unmodified reference (oracle/_ref, one OS process per rank over the shared-memory
    Boost.MPI stand-in) on the bench's own scheduled random circuit at 26-28 qubits, R = 1, 2, 4, 8: every rank's slab
    within 1e-12 of the reference rank's, slot maps equal.
Multi-GPU cases skip when the box has fewer GPUs."""
import json
import os
import pickle
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import scripts

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _free_gib():
    import torch
    return torch.cuda.mem_get_info()[0] / 2 ** 30


def _ram_gib():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            return int(line.split()[1]) / 2 ** 20
    return 16.0


def _torchrun(R, *worker_args, timeout=1500):
    from torchrun_util import run_torchrun
    res = run_torchrun(R, os.path.join(HERE, "mp_fullsize_worker.py"), list(worker_args), timeout=timeout)
    assert res.returncode == 0 and "FULLSIZE_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("FULLSIZE_OK")][-1]
    return json.loads(line[len("FULLSIZE_OK "):])


@pytest.mark.parametrize("R", [2, 4, 8])
def test_fullsize_properties_multi_gpu(R):
    if _gpu_count() < R:
        pytest.skip("needs %d GPUs" % R)
    n = 32 + R.bit_length() - 1  # L = 32: 64 GiB per GPU, many VMM chunks per slab, swaps of 1 / 2 / 3 qubits
    if _free_gib() < 70:
        pytest.skip("needs 70 GiB of free device memory per GPU")
    rep = _torchrun(R, "props", str(n))
    assert rep["ok"], rep
    assert rep["qft_closed_form"]["swaps"] + rep["random_then_inverse"]["swaps"] >= 1, rep  # the swap path was on the way


@pytest.mark.parametrize("R", [1, 2, 4, 8])
def test_direct_diff_against_compiled_reference(R):
    """the bench's scheduled random circuit, on R reference ranks (host CPU) and on R GPUs: slab by slab within 1e-12"""
    from oracle import ref
    if _gpu_count() < R:
        pytest.skip("needs %d GPUs" % R)
    if not ref.have_ref():
        pytest.skip("oracle/_ref is not built")
    g = R.bit_length() - 1
    # 2^25 (2^26 on one rank) amplitudes per reference rank: its cheat_local() hands back a Python list (~50 B per amplitude)
    n = {1: 26, 2: 26, 4: 27, 8: 28}[R]
    while n > 20 and R * (1 << (n - g)) * 80 / 2 ** 30 > 0.5 * _ram_gib():
        n -= 1
    script, shape = scripts.scheduled_script("random", n, R)
    if R > 1:
        assert shape["swaps"] >= 1
    threads = max(1, (os.cpu_count() or 1) // R)
    res = ref.run_script(script, R, threads, timeout=1500)
    with tempfile.TemporaryDirectory(prefix="hiq_fullsize_") as d:
        ids_pos = [j for j, op in enumerate(script) if op[0] == "get_qubits_ids"][-1]
        for r in range(R):
            errs = [o for o in res[r] if isinstance(o, tuple) and len(o) == 2 and o[0] == "error"]
            assert not errs, errs[:2]
            np.save(os.path.join(d, "ref%d.npy" % r), np.asarray(res[r][-1][1], dtype=np.complex128))
        with open(os.path.join(d, "ref_ids.json"), "w") as f:
            json.dump({"ids": [int(x) for x in res[0][ids_pos]], "id2pos": {int(k): int(v) for k, v in res[0][-1][0].items()}}, f)
        with open(os.path.join(d, "script.pkl"), "wb") as f:
            pickle.dump(script, f)
        del res
        rep = _torchrun(R, "diff", d)
    assert rep["max_abs_err"] <= 1e-12, rep

"""Loader and multi-process launcher for ``oracle/_ref`` (TEST INFRASTRUCTURE).

``oracle/_ref`` holds the UNMODIFIED reference engine and scheduler compiled by
``oracle/Makefile`` against the stand-in headers in ``oracle/shim``.  This module

* imports those extension modules under private names (so they never clash with
  the product's ``_cppsim_mpi`` / ``_sched_cpp``);
* runs a *script* (a list of method calls) on R reference ranks, one OS process
  per rank, wired together by the shared-memory communicator in
  ``shim/boost/mpi.hpp`` — that is how multi-rank slot maps, swap data movement,
  reductions and measurement outcomes of the reference are obtained without MPI.

Script format: ``[("ctor", seed, max_local, max_cluster), (method, *args), ...]``.
Pseudo methods: ``("cheat_local",)`` returns ``(id2pos, np.ndarray)``; ``("timed", method, *args)`` returns the wall
seconds the call took on that rank (bench.py's CPU swap baseline).
Every op yields its return value, or ``("error", message)`` if it raised.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import pickle
import subprocess
import sys
import sysconfig
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_EXT = sysconfig.get_config_var("EXT_SUFFIX")


def ref_path(name: str) -> str:
    return os.path.join(_REF_DIR, name + _EXT)


def have_ref() -> bool:
    return os.path.exists(ref_path("_cppsim_mpi")) and os.path.exists(ref_path("_sched_cpp"))


def _load(name: str):
    full = "hiq_oracle_ref." + name
    if full in sys.modules:
        return sys.modules[full]
    loader = importlib.machinery.ExtensionFileLoader(full, ref_path(name))
    spec = importlib.util.spec_from_file_location(full, ref_path(name), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[full] = mod
    return mod


def load_ref_sim():
    """The reference's own pybind module `_cppsim_mpi` (class SimulatorMPI)."""
    return _load("_cppsim_mpi")


def load_ref_sched():
    """The reference's own pybind module `_sched_cpp`."""
    return _load("_sched_cpp")


def _to_py(v):
    if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], dict):
        return dict(v[0]), np.asarray(v[1], dtype=np.complex128)
    return v


def execute_script(script):
    """Run `script` on a reference SimulatorMPI in THIS process (rank from env)."""
    mod = load_ref_sim()
    sim = None
    out = []
    for op in script:
        name, args = op[0], op[1:]
        try:
            if name == "ctor":
                sim = mod.SimulatorMPI(*args)
                out.append(None)
            elif name == "timed":  # ("timed", method, *args) -> wall seconds of that call on this rank
                t0 = time.perf_counter()
                getattr(sim, args[0])(*args[1:])
                out.append(time.perf_counter() - t0)
            else:
                out.append(_to_py(getattr(sim, name)(*args)))
        except RuntimeError as e:  # the reference throws std::runtime_error
            out.append(("error", str(e)))
    del sim
    return out


def run_script(script, world_size: int = 1, omp_threads: int | None = None, timeout: float = 600.0):
    """Run `script` on `world_size` reference ranks; returns results[rank][op]."""
    if world_size == 1 and omp_threads is None:
        return [execute_script(script)]
    with tempfile.TemporaryDirectory(prefix="hiqref_") as tmp:
        spath = os.path.join(tmp, "script.pkl")
        with open(spath, "wb") as f:
            pickle.dump(script, f)
        shm = None
        if world_size > 1:
            shm = "/dev/shm/hiqref_%d_%s" % (os.getpid(), os.path.basename(tmp))
            with open(shm, "wb") as f:
                f.truncate(4096 + (1 << 20) * world_size)
        procs = []
        try:
            for r in range(world_size):
                env = dict(os.environ)
                env["HIQ_REF_SIZE"] = str(world_size)
                env["HIQ_REF_RANK"] = str(r)
                if shm:
                    env["HIQ_REF_SHM"] = shm
                env["OMP_NUM_THREADS"] = str(omp_threads if omp_threads else 1)
                if world_size > 1:
                    env.pop("OMP_PROC_BIND", None)  # every rank would bind its team to the same cores
                env["PYTHONPATH"] = os.path.dirname(_HERE) + os.pathsep + env.get("PYTHONPATH", "")
                procs.append(subprocess.Popen(
                    [sys.executable, "-m", "oracle.ref", spath, os.path.join(tmp, "out%d.pkl" % r)],
                    env=env, cwd=os.path.dirname(_HERE)))
            for p in procs:
                rc = p.wait(timeout=timeout)
                if rc != 0:
                    raise RuntimeError("reference rank exited with code %d" % rc)
            res = []
            for r in range(world_size):
                with open(os.path.join(tmp, "out%d.pkl" % r), "rb") as f:
                    res.append(pickle.load(f))
            return res
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            if shm and os.path.exists(shm):
                os.unlink(shm)


def run_module_on_ranks(module: str, job_path: str, world_size: int, omp_threads: int | None = None, timeout: float = 600.0):
    """Start `python -m <module> <job_path> <out_r.pkl>` once per reference rank (same process wiring as run_script: one OS
    process per rank, ranks attached to one shared-memory arena through the environment) and return the unpickled outputs
    by rank.  Used by oracle/run_reference_pipeline.py, which runs the reference's Python layers on top of the engine."""
    with tempfile.TemporaryDirectory(prefix="hiqref_") as tmp:
        shm = None
        if world_size > 1:
            shm = "/dev/shm/hiqref_%d_%s" % (os.getpid(), os.path.basename(tmp))
            with open(shm, "wb") as f:
                f.truncate(4096 + (1 << 20) * world_size)
        procs = []
        try:
            for r in range(world_size):
                env = dict(os.environ)
                env["HIQ_REF_SIZE"] = str(world_size)
                env["HIQ_REF_RANK"] = str(r)
                if shm:
                    env["HIQ_REF_SHM"] = shm
                env["OMP_NUM_THREADS"] = str(omp_threads if omp_threads else 1)
                if world_size > 1:
                    env.pop("OMP_PROC_BIND", None)
                env["PYTHONPATH"] = os.path.dirname(_HERE) + os.pathsep + env.get("PYTHONPATH", "")
                procs.append(subprocess.Popen([sys.executable, "-m", module, job_path, os.path.join(tmp, "out%d.pkl" % r)],
                                              env=env, cwd=os.path.dirname(_HERE)))
            for p in procs:
                rc = p.wait(timeout=timeout)
                if rc != 0:
                    raise RuntimeError("reference rank exited with code %d" % rc)
            res = []
            for r in range(world_size):
                with open(os.path.join(tmp, "out%d.pkl" % r), "rb") as f:
                    res.append(pickle.load(f))
            return res
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            if shm and os.path.exists(shm):
                os.unlink(shm)


if __name__ == "__main__":
    with open(sys.argv[1], "rb") as f:
        _script = pickle.load(f)
    _res = execute_script(_script)
    with open(sys.argv[2], "wb") as f:
        pickle.dump(_res, f)

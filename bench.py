#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 state-vector engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU SimulatorMPI (oracle/_ref)
    python bench.py --circuit shor|grover ...                 # secondary lines (BASELINE.json configs[4] / configs[0])

Metric: gate-apply effective HBM GB/s = 32 B x 2^L x (fused-gate passes of the plan) / time, summed over all ranks
(BASELINE.json).  One *step* = one complete execution of the scheduled circuit over the resident state
vector.  ONE workload family for every N (BASELINE.json configs[1]/[3] generator, depth 20, cluster size 4):
  N=1  33-qubit random circuit (L=33, 128 GiB slab)   north_star: "a 33-qubit random circuit on 1 B200"
  N=2  34-qubit random circuit (L=33 per GPU)
  N=4  35-qubit random circuit (L=33 per GPU)
  N=8  35-qubit random circuit (L=32 per GPU)          "35q@8"
The N=1 line additionally carries `qft33` (BASELINE.json configs[2], 33-qubit QFT) with its effective AND physical GB/s.

`value`    : the pre-scheduled command stream replayed on the resident state (host fusion and kernel launches inside the
             timed region, scheduling outside), CUDA events on the engine stream, max over ranks.
`e2e`      : the same circuit through the reference-facing API end to end, every step: SimulatorMPI(...) ->
             allocate_qureg -> GreedyScheduler -> gates/flush/swaps -> Measure(all), host matrices in and measured bits out
             (host wall clock around a device synchronize + barrier, max over ranks); `e2e_breakdown` says where it goes.
`roofline` : dominant kernel, algorithmic bytes (32 B x 2^L per pass) / its mean launch time from CUDA events recorded
             around every pass inside the timed region, vs the measured HBM peak in MEASURED_PEAKS.json.
`parity`   : computed OUTSIDE the timed regions on fresh engines of the same process group: QFT closed form on 2048
             sampled amplitudes fetched with get_amplitude (crosses ranks, includes the swaps), single-qubit marginals,
             post-measurement probability, and random circuit followed by its inverse; all must be <= 1e-12.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference engine and scheduler (oracle/_ref) on the host cores,
             same metric, on a bounded sample of the same workload (same generator at fewer qubits); the sample and
             the OpenMP thread count actually in effect are stated in the line.  That arm loads none of this
             repository's native modules.
"""
from __future__ import annotations

import argparse
import copy
import ctypes
import gc
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gate_apply_effective_hbm_gbs"
UNIT = "GB/s"
TOL = 1e-12


# ---------------------------------------------------------------------------------------------
def workload_for(n_gpus: int, qubits: int | None, circuit: str | None):
    """(name, n_qubits, L, circuit kind) — one family for every N: the random circuit of BASELINE.json configs[1]/[3]"""
    table = {1: 33, 2: 34, 4: 35, 8: 35}
    n = table.get(n_gpus, 32 + n_gpus.bit_length() - 1)
    kind = circuit or "random"
    if kind == "grover":
        n = 20   # BASELINE.json configs[0]
    if kind == "shor":
        n = 32   # BASELINE.json configs[4]
    if qubits:
        n = qubits
    g = n_gpus.bit_length() - 1
    return "%s-%d" % (kind, n), n, n - g, kind


def build_circuit(kind: str, n: int):
    from hiqsimulator_b200 import circuits
    if kind == "qft":
        return circuits.qft_circuit(n)[1]
    if kind == "random":
        return circuits.random_circuit(n, depth=20)[1]
    if kind == "grover":
        return circuits.grover_circuit(n - 1, 8)[1]
    raise SystemExit("unknown circuit %r" % kind)


class RecordingBackend:
    """Backend proxy that records the post-scheduler command stream while the wrapped backend keeps the slot maps the
    schedulers query."""

    def __init__(self, inner):
        self.inner = inner
        self.stream = []
        self.main_engine = None

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def set_qubits_perm(self, ids):
        self.stream.append(("perm", list(ids)))
        self.inner.set_qubits_perm(ids)

    def receive(self, cmds):
        self.stream.extend(cmds)
        self.inner.receive(cmds)


def schedule_circuit(n, L, cmds, backend_class, sched_module=None, cluster=4):
    """Run the GreedyScheduler once against `backend_class` (a dry-run engine of this repository, or the reference
    engine itself in the reference arm) -> recorded command stream, schedule shape."""
    from hiqsimulator_b200 import backends, cengines, ops
    inner = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=L, max_fused_qubits=cluster, backend_class=backend_class)
    rec = RecordingBackend(inner)
    if sched_module is None:
        gs = cengines.GreedyScheduler(cluster_size=cluster)
    else:  # the reference arm: the reference's own Python loop around the unmodified reference scheduler
        from oracle import greedy_loop
        gs = greedy_loop.ReferenceLoopScheduler(cluster_size=cluster, sched_module=sched_module)
    eng = cengines.HiQMainEngine(rec, [gs])
    t0 = time.perf_counter()
    eng.receive([ops.AllocateQureg(list(range(n)), 0)])
    eng.receive(copy.deepcopy(cmds))
    eng.flush()
    host_s = time.perf_counter() - t0
    stream = [c for c in rec.stream if not (hasattr(c, "kind") and c.kind == ops.ALLOCATE_QUREG)]
    shape = {"gates": len(cmds), "passes": gs.n_clusters, "swaps": gs.n_swaps,
             "swap_qubits": [len(v) // 2 for k, v in gs.log if k == "swap"],
             "host_schedule_s": round(host_s, 3), "cluster_sched_s": round(gs.cluster_seconds, 3),
             "swap_sched_s": round(gs.swap_seconds, 3)}
    return stream, shape, inner


def replay(backend, stream):
    for c in stream:
        if isinstance(c, tuple):
            backend.set_qubits_perm(c[1])
        else:
            backend.receive([c])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.path = tempfile.mktemp(prefix="hiq_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "power_w_max": max(power)}
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref): rank 0 only, no native module of this repository in the process
# ---------------------------------------------------------------------------------------------
def host_threads():
    # counted once, before any OpenMP runtime binds the main thread to a core (OMP_PROC_BIND), and handed to re-executed /
    # child interpreters through the environment
    if "HIQ_BENCH_HOST_CPUS" not in os.environ:
        try:
            os.environ["HIQ_BENCH_HOST_CPUS"] = str(len(os.sched_getaffinity(0)))
        except AttributeError:
            os.environ["HIQ_BENCH_HOST_CPUS"] = str(os.cpu_count() or 1)
    return int(os.environ["HIQ_BENCH_HOST_CPUS"])


def fix_openmp_environment():
    """The reference engine is an OpenMP program (reference: _simulator_mpi.py:52-58 sets OMP_NUM_THREADS / OMP_PROC_BIND).
    torchrun exports OMP_NUM_THREADS=1 to its children, and libgomp reads the variable once, when it is loaded — so the
    environment is set to the intended team size and the interpreter re-executed before anything OpenMP is imported."""
    want = os.environ.get("HIQ_BENCH_CPU_THREADS") or str(host_threads())
    if os.environ.get("OMP_NUM_THREADS") == want and os.environ.get("HIQ_BENCH_OMP_FIXED") == "1":
        return
    os.environ["OMP_NUM_THREADS"] = want
    os.environ["OMP_PROC_BIND"] = "spread"
    os.environ["HIQ_BENCH_OMP_FIXED"] = "1"
    sys.stdout.flush()
    os.execv(sys.executable, [sys.executable] + sys.argv)


def omp_threads_in_effect():
    """omp_get_max_threads() of the libgomp the reference module is linked against (already loaded -> same instance)"""
    try:
        return int(ctypes.CDLL("libgomp.so.1").omp_get_max_threads())
    except OSError:
        return None


def available_ram_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 32 << 30


def flush_positions(stream):
    from hiqsimulator_b200 import ops
    return [i for i, c in enumerate(stream) if not isinstance(c, tuple) and c.kind == ops.FLUSH]


def run_reference_steps(kind, steps, warmup, budget_s, n_cpu=None, log=None):
    """Times the unmodified reference SimulatorMPI (oracle/_ref) replaying a stream planned by the unmodified reference
    scheduler (oracle/_ref/_sched_cpp, driven by the reference's Python loop restated in cengines.py)."""
    from hiqsimulator_b200 import backends
    from oracle import ref
    if not ref.have_ref():
        raise RuntimeError("oracle/_ref is not built")
    refsim = ref.load_ref_sim()
    refsched = ref.load_ref_sched()
    iters = steps + warmup
    # ---- probe: a few fused passes of the same generator at 26 qubits (1 GiB state, far beyond the last-level cache)
    if n_cpu is None:
        pn = 26
        pstream, pshape, pbe = schedule_circuit(pn, pn, build_circuit(kind, pn)[:12 * pn], refsim.SimulatorMPI, refsched)
        fl = flush_positions(pstream)
        t0 = time.perf_counter()
        replay(pbe, pstream)
        rate = 32.0 * (1 << pn) * max(1, len(fl)) / (time.perf_counter() - t0)  # bytes / s
        del pbe
        gc.collect()
        ram = available_ram_bytes()
        n_cpu = 24
        for n in (30, 29, 28, 27, 26, 25):
            passes = 4.2 * n if kind == "qft" else 3.3 * n
            if 16 * (1 << n) * 1.3 > ram / 2:
                continue
            # the scheduler of the reference costs ~0.1-0.3 s per cluster at these sizes (outside the timed region)
            if n <= 28 or iters * passes * 32.0 * (1 << n) / rate + 0.2 * passes <= budget_s:
                n_cpu = n
                break
        if log is not None:
            log["probe"] = "%d passes of %s-%d: %.1f GB/s" % (len(fl), kind, pn, rate / 1e9)
    cmds = build_circuit(kind, n_cpu)
    t_s = time.perf_counter()
    stream, shape, be = schedule_circuit(n_cpu, n_cpu, cmds, refsim.SimulatorMPI, refsched)
    shape["reference_scheduler_s"] = round(time.perf_counter() - t_s, 2)
    fl = flush_positions(stream)
    # bounded sample: the first P fused passes of the plan when the whole circuit x (steps + warmup) exceeds the budget
    t0 = time.perf_counter()
    replay(be, stream[:fl[min(3, len(fl) - 1)] + 1])
    per_pass = (time.perf_counter() - t0) / min(4, len(fl))
    p_fit = int(budget_s / max(1e-9, iters * per_pass))
    n_pass = max(min(8, len(fl)), min(len(fl), p_fit))
    part = stream[:fl[n_pass - 1] + 1]
    times = []
    for it in range(iters):
        t0 = time.perf_counter()
        replay(be, part)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    nbytes = 32.0 * (1 << n_cpu) * n_pass
    total = sum(times)
    return {"value": nbytes * len(times) / total / 1e9, "ms_per_step": 1e3 * total / len(times), "shape": shape, "qubits": n_cpu,
            "passes_per_step": n_pass, "passes_in_plan": len(fl)}


def reference_swap_baseline(world, L=24):
    """The reference's own swap (pack -> all_to_all -> unpack, SwapperMT) on `world` ranks of the multi-process oracle;
    GB/s per rank per direction by the reference's formula (reference: SimulatorMPI.cpp:1069-1080, there in Gbit/s)."""
    from oracle import ref
    g = world.bit_length() - 1
    n = L + g
    threads = max(1, host_threads() // world)
    import numpy as np
    h = (np.array([[1, 1], [1, -1]]) / np.sqrt(2.0)).tolist()
    script = [("ctor", 1, L, 4), ("allocate_qureg", list(range(n)), 0)]
    for q in range(min(n, 8)):
        script.append(("apply_controlled_gate", h, [q], []))
        if q % 4 == 3:
            script.append(("run",))
    script += [("run",), ("get_local_qubits_ids",), ("get_global_qubits_ids",)]
    probe = ref.run_script(script[:2] + script[-2:], world, threads)
    loc, glo = list(probe[0][-2]), [q for q in probe[0][-1] if q >= 0]
    pairs = []
    for i, gq in enumerate(glo):
        pairs += [gq, loc[len(loc) - 1 - i]]
    back = []
    for i in range(0, len(pairs), 2):
        back += [pairs[i + 1], pairs[i]]
    script += [("timed", "swap_qubits", pairs), ("timed", "swap_qubits", back), ("timed", "swap_qubits", pairs)]
    res = ref.run_script(script, world, threads)
    secs = [max(r[-k] for r in res) for k in (3, 2, 1)]
    t = min(secs)
    gbs = 16.0 * (1 << L) * (1 - 2.0 ** -g) / t / 1e9
    return {"value": gbs, "unit": "GB/s per rank per direction", "ranks": world, "local_qubits": L, "swapped_qubits": g,
            "seconds": t, "threads_per_rank": threads, "kind": "reference",
            "formula": "16 B x 2^L x (1 - 2^-q) / t (reference: SimulatorMPI.cpp:1069-1080)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fix_openmp_environment()
    name, n, L, kind = workload_for(args.gpus, args.qubits, args.circuit)
    info = {}
    try:
        if kind == "shor":
            raise RuntimeError("the reference engine cannot emulate math gates (SimulatorMPI.hpp:217-225 throws)")
        n_fixed = args.cpu_qubits or (n if kind == "grover" else None)
        r = run_reference_steps(kind, args.steps, args.warmup, args.cpu_budget, n_fixed, info)
        threads = omp_threads_in_effect()
    except Exception as e:  # the oracle build is missing: say so, exit 0
        print(json.dumps({"impl": "reference", "unavailable": "%s: %s" % (type(e).__name__, e)}))
        return
    sample = ("%s-%d (same generator as the %d-qubit workload), %d of its %d fused passes per step, planned by the unmodified "
              "reference scheduler (%.1f s, outside the timed region), state resident in host RAM, %s OpenMP threads in effect"
              % (kind, r["qubits"], n, r["passes_per_step"], r["passes_in_plan"], r["shape"]["reference_scheduler_s"], threads))
    cpu = {"value": r["value"], "unit": UNIT, "cores": threads or int(os.environ["OMP_NUM_THREADS"]), "kind": "reference", "sample": sample,
           "host_cpus": host_threads(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), **info}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "qubits": n, "local_qubits": L, "cluster_size": 4, "sample_qubits": r["qubits"],
                   "same_config": r["qubits"] == n,
                   "caveat": None if r["qubits"] == n else
                   "CPU sample is %d qubits (host RAM / time bound), the GPU arm runs %d: GB/s is size-normalised, seconds are not" % (r["qubits"], n)},
        "cpu_baseline": cpu,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.gpus > 1 and not args.no_swap_baseline:
        try:
            line["swap_reference_cpu"] = reference_swap_baseline(args.gpus)
        except Exception as e:
            line["swap_reference_cpu"] = {"value": None, "error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line))


def cpu_leg_subprocess(args, kind, n):
    """cpu_baseline of the GPU arm: the reference arm in a child process (its own OpenMP environment, no repo module)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus), "--steps", "2", "--warmup", "1",
           "--cpu-budget", "25", "--circuit", kind, "--qubits", str(n)]
    if args.cpu_qubits:
        cmd += ["--cpu-qubits", str(args.cpu_qubits)]
    env = dict(os.environ)
    env["RANK"] = "0"
    env.pop("HIQ_BENCH_OMP_FIXED", None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        last = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(last)
        if "unavailable" in d:
            raise RuntimeError(d["unavailable"])
        out = d["cpu_baseline"]
        if "swap_reference_cpu" in d:
            out["swap_reference_cpu"] = d["swap_reference_cpu"]
        return out
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": "unavailable: %s: %s" % (type(e).__name__, e)}


# ---------------------------------------------------------------------------------------------
# parity checks (outside every timed region; all ranks take part, the numbers are identical on every rank)
# ---------------------------------------------------------------------------------------------
def release(be, eng=None):
    """Breaks the engine <-> backend <-> scheduler cycles and drops the simulator so that its slab is unmapped NOW (the next
    engine maps its own 128 GiB); callers rebind their own names to None as well."""
    be.main_engine = None
    if eng is not None:
        eng.scheduler.backend = eng.scheduler.next_engine = None
        eng.backend = None
    be._simulator = None
    gc.collect()


def parity_checks(n, L, fresh_backend, samples=2048, skip_random=False):
    import numpy as np
    from hiqsimulator_b200 import cengines, circuits, ops
    out = {"tolerance": TOL, "qubits": n, "local_qubits": L}
    # (1) QFT closed form through the whole pipeline (scheduler, fusion, folded diagonals, swaps, get_amplitude broadcast)
    t0 = time.perf_counter()
    be = fresh_backend()
    eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
    _, cmds = circuits.qft_circuit(n)
    x = 0x5A5A5A5A5A5A5A5A & ((1 << n) - 1)
    eng.receive([ops.AllocateQureg(list(range(n)), 0)])
    eng.receive(copy.deepcopy(cmds))
    eng.flush()
    ids = list(range(n))
    rng = np.random.default_rng(1000 + n)
    worst = 0.0
    for _ in range(samples):
        y = int(rng.integers(0, 1 << n))
        got = be.get_amplitude([(y >> q) & 1 for q in range(n)], ids)
        worst = max(worst, abs(got - circuits.qft_expected_amplitude(n, x, y)))
    marg = max(abs(be.get_probability([0], [q]) - 0.5) for q in (0, n // 2, n - 1))
    eng.receive([ops.Measure(ids)])
    bits = [int(eng.measurements[q]) for q in ids]
    p_after = be.get_probability(bits, ids)
    st = getattr(be._simulator, "stats", lambda: {"total_swaps": -1})()
    out["qft_closed_form"] = {"circuit": "qft-%d" % n, "sampled_amplitudes": samples, "max_abs_err": worst, "max_marginal_err": marg,
                              "post_measurement_prob_err": abs(p_after - 1.0), "swaps": int(st["total_swaps"]),
                              "seconds": round(time.perf_counter() - t0, 2)}
    release(be, eng)
    be = eng = None
    ok = worst <= TOL and marg <= TOL and abs(p_after - 1.0) <= TOL
    # (2) random circuit followed by its inverse
    if not skip_random:
        t0 = time.perf_counter()
        be = fresh_backend()
        eng = cengines.HiQMainEngine(be, [cengines.GreedyScheduler(cluster_size=4)])
        _, cmds = circuits.random_circuit(n, depth=20)
        eng.receive([ops.AllocateQureg(list(range(n)), 0)])
        eng.receive(copy.deepcopy(cmds))
        eng.flush()
        eng.receive(circuits.inverse_circuit(cmds))
        eng.flush()
        amp = be.get_amplitude([0] * n, ids)
        p0 = be.get_probability([0] * n, ids)
        st = getattr(be._simulator, "stats", lambda: {"total_swaps": -1})()
        out["random_then_inverse"] = {"circuit": "random-%d depth 20 followed by its inverse" % n, "abs_amp0_minus_1": abs(amp - 1.0),
                                      "abs_p0_minus_1": abs(p0 - 1.0), "swaps": int(st["total_swaps"]),
                                      "seconds": round(time.perf_counter() - t0, 2)}
        release(be, eng)
        be = eng = None
        ok = ok and abs(amp - 1.0) <= TOL and abs(p0 - 1.0) <= TOL
    out["ok"] = bool(ok)
    return out


# ---------------------------------------------------------------------------------------------
def time_replay(be, stream, steps, warmup, world, torch, M, local, sampler=None):
    """W untimed + K timed replays of `stream` on the resident state; returns device ms (this rank), launches, stats delta, timings"""
    sim = be._simulator
    ext = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local))
    for _ in range(warmup):
        replay(be, stream)  # each replay starts by re-applying the initial relabelling (no data motion)
    sim.synchronize()
    sim.collect_timings()
    world.barrier()
    if sampler:
        sampler.start()
    launches0 = M.launch_count()
    stats0 = sim.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record(ext)
    for _ in range(steps):
        replay(be, stream)
    e1.record(ext)
    e1.synchronize()
    sim.synchronize()
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    launches = M.launch_count() - launches0
    stats1 = sim.stats()
    timings = sim.collect_timings()
    world.barrier()
    delta = {k: stats1[k] - stats0[k] for k in stats1}
    return ms_total, launches, delta, timings, clocks, t_host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=None, help="override the number of qubits of the workload")
    ap.add_argument("--circuit", default=None, choices=[None, "qft", "random", "shor", "grover"])
    ap.add_argument("--cpu-qubits", type=int, default=None, help="size of the CPU-baseline sample")
    ap.add_argument("--cpu-budget", type=float, default=120.0, help="seconds the timed CPU steps may take in total")
    ap.add_argument("--e2e-steps", type=int, default=None, help="end-to-end steps (default min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-swap-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-qft-line", action="store_true")
    ap.add_argument("--no-batch", action="store_true",
                    help="one launch per fused gate of the plan (HIQ_FLAG_NO_BATCH): no folding of diagonal passes, no tile groups")
    args = ap.parse_args()

    if args.impl == "reference":
        reference_arm(args)
        return
    if args.circuit == "shor":
        import bench_extra
        bench_extra.main(args)
        return

    import torch
    import torch.distributed as dist
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, ops, world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this engine has no CPU path)")

    rank, size = world.init_world(M.FLAG_TIMING | (M.FLAG_NO_BATCH if args.no_batch else 0))
    assert size == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, size)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    name, n, L, kind = workload_for(size, args.qubits, args.circuit)
    cmds = build_circuit(kind, n)

    def dry(s, ml, mc):
        return M.SimulatorMPI(s, ml, mc, rank, size, M.FLAG_DRY_RUN)

    def fresh_backend(local_qubits=L):
        return backends.SimulatorMPI(gate_fusion=True, rnd_seed=12345, num_local_qubits=local_qubits, max_fused_qubits=4)

    def reduce_max_sum(ms, sums):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        w = torch.tensor(sums, dtype=torch.float64, device="cuda")
        if size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        return float(t.item()), [float(x) for x in w.tolist()]

    # ---- schedule once (host), outside the timed region of `value`
    stream, shape, _ = schedule_circuit(n, L, cmds, dry)
    passes = shape["passes"]

    # ---- value: replay on the resident state
    be = fresh_backend()
    be._simulator.allocate_qureg(list(range(n)), 0)
    be._simulator.synchronize()
    ms_total, launches, ds, timings, clocks, t_host = time_replay(be, stream, args.steps, args.warmup, world, torch, M, local, ClockSampler(local))
    plan_passes = ds["dense_passes"] + ds["diag_passes"] + ds["scale_passes"]
    ms_total, (total_bytes, total_launches, total_swap_bytes, phys_bytes) = reduce_max_sum(
        ms_total, [32.0 * (1 << L) * plan_passes, float(launches), ds["swap_bytes_sent"], 32.0 * (1 << L) * ds["gate_launches"]])
    value = total_bytes / (ms_total * 1e-3) / 1e9
    ms_per_step = ms_total / args.steps

    # ---- roofline of the dominant kernel (this rank's launches; rank 0 reports)
    # timings: (kind, k, variant, ms, n_ref) per launch; n_ref = passes of the reference's plan the launch carried
    roofline, breakdown, groups = roofline_of(timings, L, args)
    swap_gbs = None
    # the exchange itself; the time a rank waits for its peers to reach the swap (rank skew) is reported apart
    swap_ms = sum(t[3] for t in timings if t[0] == 4 and t[2] == 0)
    swap_wait_ms = sum(t[3] for t in timings if t[0] == 4 and t[2] == 1)
    if swap_ms > 0:
        swap_gbs = ds["swap_bytes_sent"] / (swap_ms * 1e-3) / 1e9
    release(be)
    be = None
    torch.cuda.empty_cache()

    # ---- N = 1: the 33-qubit QFT of BASELINE.json configs[2] as a second measurement
    qft33 = None
    if size == 1 and kind == "random" and n == 33 and not args.no_qft_line:
        qcmds = build_circuit("qft", n)
        qstream, qshape, _ = schedule_circuit(n, L, qcmds, dry)
        be = fresh_backend()
        be._simulator.allocate_qureg(list(range(n)), 0)
        be._simulator.synchronize()
        qsteps = max(1, min(args.steps, 5))
        qms, ql, qd, qt, _, _ = time_replay(be, qstream, qsteps, 2, world, torch, M, local)
        qpasses = qd["dense_passes"] + qd["diag_passes"] + qd["scale_passes"]
        qroof, qbreak, _ = roofline_of(qt, L, args)
        qft33 = {"workload": "qft-33", "ms_per_step": qms / qsteps, "steps": qsteps, "warmup": 2,
                 "effective_gbs": 32.0 * (1 << L) * qpasses / (qms * 1e-3) / 1e9,
                 "physical_gbs": 32.0 * (1 << L) * qd["gate_launches"] / (qms * 1e-3) / 1e9,
                 "fused_passes": qshape["passes"], "plan_passes_per_step": int(qpasses // qsteps),
                 "hbm_passes_per_step": int(qd["gate_launches"] // qsteps),
                 "tile_launches_per_step": int(qd.get("tile_launches", 0) // qsteps),
                 "note": "effective = plan passes x 32 B x 2^L / t (diagonal fused gates and consecutive clusters share HBM passes); "
                         "physical = launches x 32 B x 2^L / t",
                 "roofline": qroof, "kernel_breakdown": qbreak}
        release(be)
        be = None
        torch.cuda.empty_cache()

    # ---- e2e: the full reference-facing pipeline, every step from host inputs to measured bits
    e2e, e2e_breakdown = None, None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
        per_step, h2d, d2h = [], 0.0, 0.0
        parts = {k: 0.0 for k in ("construct_s", "allocate_s", "circuit_enqueue_s", "measure_and_sync_s", "device_busy_s", "host_cluster_search_s",
                                  "host_swap_search_s", "peer_map_s", "slab_map_s", "swap_call_s")}
        e2e_passes = 0
        for it in range(1 + e2e_steps):
            step_cmds = copy.deepcopy(cmds)
            world.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            be2 = fresh_backend()
            t1 = time.perf_counter()
            gs = cengines.GreedyScheduler(cluster_size=4)
            eng = cengines.HiQMainEngine(be2, [gs])
            eng.receive([ops.AllocateQureg(list(range(n)), 0)])
            t2 = time.perf_counter()
            eng.receive(step_cmds)
            eng.flush()
            t3 = time.perf_counter()
            eng.receive([ops.Measure(list(range(n)))])
            be2._simulator.synchronize()
            world.barrier()
            t4 = time.perf_counter()
            st = be2._simulator.stats()
            dev_ms = sum(t[3] for t in be2._simulator.collect_timings())
            if it >= 1:
                per_step.append(t4 - t0)
                h2d += be2.h2d_bytes
                d2h += st["d2h_bytes"] + n  # + the measured bits returned to the caller
                for k, v in (("construct_s", t1 - t0), ("allocate_s", t2 - t1), ("circuit_enqueue_s", t3 - t2), ("measure_and_sync_s", t4 - t3),
                             ("device_busy_s", dev_ms * 1e-3), ("host_cluster_search_s", gs.cluster_seconds),
                             ("host_swap_search_s", gs.swap_seconds), ("peer_map_s", st["peer_map_s"]), ("slab_map_s", st["slab_grow_s"]),
                             ("swap_call_s", st["swaps_s"])):
                    parts[k] += v
            e2e_passes = st["dense_passes"] + st["diag_passes"] + st["scale_passes"]
            release(be2, eng)
            be2 = eng = gs = None
        tt, (ww,) = reduce_max_sum(sum(per_step), [32.0 * (1 << L) * e2e_passes * len(per_step)])
        e2e = {"value": ww / tt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / len(per_step)), "d2h_bytes_per_step": int(d2h / len(per_step)),
               "seconds_per_step": tt / len(per_step), "steps": len(per_step),
               "includes": "engine construction, allocate_qureg, host scheduling + fusion, all passes/swaps, Measure(all)"}
        e2e_breakdown = {k: round(v / len(per_step), 4) for k, v in parts.items()}
        e2e_breakdown["note"] = ("rank 0, seconds per step; wall = construct + allocate + circuit_enqueue + measure_and_sync; device_busy = sum of "
                                 "the per-launch CUDA-event times; the host searches run on a planner thread ahead of the device; "
                                 "peer_map / slab_map / swap_call are inside the wall parts")

    # ---- parity on the same process group, outside the timed regions
    parity = None
    if not args.no_parity:
        parity = parity_checks(n, L, fresh_backend)

    # ---- CPU baseline (rank 0 at every N): the unmodified reference on a bounded sample, in a child process
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_baseline = cpu_leg_subprocess(args, kind, n)
    world.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": name, "qubits": n, "local_qubits": L, "slab_gib_per_gpu": 16.0 * (1 << L) / 2 ** 30,
                       "cluster_size": 4, "gates": shape["gates"], "fused_passes": passes,
                       "hbm_passes_per_step": int(ds["gate_launches"]) // args.steps,
                       "batching": "off (one launch per fused gate)" if args.no_batch else
                                   "diagonal fused gates ride along the next dense launch; consecutive dense gates whose targets fit one "
                                   "shared-memory tile share a launch",
                       "swaps": shape["swaps"],
                       "swap_qubits": shape["swap_qubits"],
                       "l2_policy": "inputs larger than L2 (slab >> 126 MB), no flush" if 16.0 * (1 << L) > 4 * 126e6 else
                                    "slab of %.0f MiB is L2-resident: a parity-size configuration, not a bandwidth measurement" % (16.0 * (1 << L) / 2 ** 20),
                       "timing": "CUDA events on the engine stream, max over ranks",
                       "series": "one family for every N: random-33 / 34 / 35 / 35 at N = 1 / 2 / 4 / 8 (BASELINE.json configs[1]/[3] generator); "
                                 "the N=1 line carries the 33-qubit QFT of configs[2] under `qft33`"},
            "physical_hbm_gbs": phys_bytes / (ms_total * 1e-3) / 1e9,
            "clocks": clocks, "e2e": e2e, "e2e_breakdown": e2e_breakdown, "gpu_launches": int(total_launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity, "qft33": qft33,
            "circuit_seconds": {"device_only": ms_per_step * 1e-3, "host_schedule_s": shape["host_schedule_s"],
                                "end_to_end": e2e["seconds_per_step"] if e2e else None},
            "swap_nvlink_gbs_per_gpu": swap_gbs,
            "swap_wait_for_peers_ms_per_step": swap_wait_ms / args.steps,
            "swap_transport": {"peer_mapped_in_place": int(ds["swaps_p2p"]), "staged_nccl": int(ds["swaps_staged"]),
                               "packed_push": int(ds.get("swaps_packed", 0)), "nvlink_peak_gbs_per_dir": 900.0,
                               "frac_of_nvlink": (swap_gbs / 900.0) if swap_gbs else None},
            "kernel_breakdown": breakdown,
            "host_enqueue_seconds_per_step": t_host / args.steps,
        }
        print(json.dumps(line))


def roofline_of(timings, L, args):
    groups = {}
    for kind_id, k, variant, ms, n_ref in timings:
        folded = kind_id == 1 and n_ref > 1
        # dense DIRECT launches of block-structured matrices carry their mixing bits in variant bits 8..
        groups.setdefault((kind_id, k if kind_id != 2 else 0, variant & 0xff, folded, variant >> 8), []).append((ms, n_ref))
    names = {1: "dense", 2: "diag_batch", 3: "scale", 4: "swap", 12: "tile"}
    vnames = {0: "", 1: "direct", 2: "tiled", 3: "dmma"}

    def gname(g):
        if g[0] == 4:
            return "swap_q%d%s" % (g[1], "_wait_for_peers" if g[2] == 1 else "")
        if g[0] == 2:
            return "diag_batch" if not args.no_batch else "diag"
        if g[0] == 12:
            return "tile_program_%dgates" % g[1]
        return "%s_k%d_%s%s%s" % (names.get(g[0], "?"), g[1], vnames.get(g[2], ""), ("_mix%d" % g[4]) if g[4] else "",
                                  "+prediag" if g[3] else "")
    peak, peak_src = measured_peaks()
    gate_groups = {g: v for g, v in groups.items() if g[0] in (1, 2, 3, 12)}
    roofline = None
    breakdown = []
    if gate_groups:
        tot = lambda v: sum(ms for ms, _ in v)  # noqa: E731
        dom = max(gate_groups, key=lambda g: tot(gate_groups[g]))
        mean_ms = tot(gate_groups[dom]) / len(gate_groups[dom])
        mean_ref = sum(r for _, r in gate_groups[dom]) / len(gate_groups[dom])
        achieved = 32.0 * (1 << L) / (mean_ms * 1e-3) / 1e9  # bytes one launch moves: ONE pass over the slab
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": gname(dom), "launches": len(gate_groups[dom]), "mean_launch_ms": mean_ms, "peak_source": peak_src,
                    "share_of_step": tot(gate_groups[dom]) / max(1e-9, sum(tot(v) for v in groups.values())),
                    "reference_passes_per_launch": mean_ref, "effective_gbs_per_launch": achieved * mean_ref,
                    "note": "achieved counts the bytes a launch really moves (32 B x 2^L); a launch that also carries folded "
                            "passes of the plan does their work in the same pass (effective = achieved x passes per launch)"}
        # measured DRAM traffic of the same kernel (ncu --set full capture, profiles/ncu_traffic.json)
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                cap = json.load(f)["kernels"].get(re.sub(r"_mix\d", "", gname(dom)))  # same kernel, fewer flops: same traffic
            if cap:
                per_amp = cap["dram_gbytes"] * 1e9 / (1 << cap["L"])
                roofline["traffic"] = per_amp * (1 << L)
                roofline["traffic_source"] = "%s: %.3f GB at L=%d (%.2f B/amplitude), scaled to L=%d" % (
                    cap["source"], cap["dram_gbytes"], cap["L"], per_amp, L)
        except (OSError, KeyError, ValueError):
            pass
        for g, v in sorted(groups.items(), key=lambda kv: -tot(kv[1])):
            per = 32.0 * (1 << L) if g[0] != 4 else 16.0 * (1 << L) * (1 - 2.0 ** -g[1])
            m = tot(v) / len(v)
            breakdown.append({"kernel": gname(g), "launches": len(v), "reference_passes": sum(r for _, r in v),
                              "total_ms": round(tot(v), 3), "mean_ms": round(m, 4), "gbs": round(per / (m * 1e-3) / 1e9, 1)})
    return roofline, breakdown, groups


if __name__ == "__main__":
    main()

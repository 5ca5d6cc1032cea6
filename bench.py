#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 state-vector engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU SimulatorMPI

Metric: gate-apply effective HBM GB/s = 32 B x 2^L x (fused-gate passes) / time, summed over all
ranks (BASELINE.json).  One *step* = one complete execution of the scheduled
circuit over the resident state vector:
  N=1  33-qubit QFT            (L=33, 137 GB slab)      BASELINE.json configs[2], "33q@1"
  N=2  34-qubit random circuit (L=33 per GPU)           configs[3]
  N=4  35-qubit random circuit (L=33 per GPU)           configs[3]
  N=8  35-qubit random circuit (L=32 per GPU)           configs[3], "35q@8"
`value`    : the pre-scheduled command stream replayed on the resident state (host fusion and
             kernel launches inside the timed region, scheduling outside), CUDA events on the
             engine stream, max over ranks.
`e2e`      : the same circuit through the reference-facing API end to end, every step:
             SimulatorMPI(...) -> allocate_qureg -> GreedyScheduler -> gates/flush/swaps ->
             Measure(all) with host matrices in and measured bits out (host wall clock around a
             device synchronize + barrier, max over ranks).
`roofline` : dominant kernel, algorithmic bytes (32 B x 2^L per pass) / its mean launch time from
             CUDA events recorded around every pass inside the timed region, vs the measured HBM
             peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the unmodified reference engine (oracle/_ref) on the host
             cores, same metric, on a bounded sample (a smaller QFT) of the same workload.
"""
from __future__ import annotations

import argparse
import copy
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


# ---------------------------------------------------------------------------------------------
def workload_for(n_gpus: int, qubits: int | None, circuit: str | None):
    """(name, n_qubits, L, circuit kind)"""
    table = {1: ("qft", 33), 2: ("random", 34), 4: ("random", 35), 8: ("random", 35)}
    kind, n = table.get(n_gpus, ("random", 32 + n_gpus.bit_length() - 1))
    if circuit:
        kind = circuit
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
    raise SystemExit("unknown circuit %r" % kind)


class RecordingBackend:
    """Backend proxy that records the post-scheduler command stream while a dry-run engine keeps
    the slot maps the schedulers query."""

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


def schedule_circuit(n, L, cmds, rank, world, sched_module=None, cluster=4):
    """Run the GreedyScheduler once against a dry-run engine -> command stream, schedule shape."""
    from hiqsimulator_b200 import _cppsim_mpi as M
    from hiqsimulator_b200 import backends, cengines, ops
    inner = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=L, max_fused_qubits=cluster,
                                  backend_class=lambda s, ml, mc: M.SimulatorMPI(s, ml, mc, rank, world, M.FLAG_DRY_RUN))
    rec = RecordingBackend(inner)
    gs = cengines.GreedyScheduler(cluster_size=cluster, sched_module=sched_module)
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
    return stream, shape


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
# CPU reference arm (oracle/_ref): rank 0 only
# ---------------------------------------------------------------------------------------------
def pick_cpu_qubits(kind, total_steps, budget_s=150.0):
    """Largest sample whose (steps x estimated time) fits the budget: ~15 GB/s effective on 16 cores."""
    for n in (28, 27, 26, 25, 24, 22, 20):
        passes = 4.2 * n if kind == "qft" else 3.3 * n
        est = passes * 32.0 * (1 << n) / 12e9
        if est * total_steps <= budget_s:
            return n
    return 20


def run_reference_steps(kind, n_cpu, steps, warmup):
    """Times the unmodified reference SimulatorMPI (oracle/_ref) on the pre-scheduled stream."""
    from hiqsimulator_b200 import backends
    from oracle import ref
    if not ref.have_ref():
        raise RuntimeError("oracle/_ref is not built")
    refsim = ref.load_ref_sim()
    cmds = build_circuit(kind, n_cpu)
    stream, shape = schedule_circuit(n_cpu, n_cpu, cmds, 0, 1)
    be = backends.SimulatorMPI(gate_fusion=True, rnd_seed=1, num_local_qubits=n_cpu, max_fused_qubits=4,
                               backend_class=refsim.SimulatorMPI)
    be._simulator.allocate_qureg(list(range(n_cpu)), 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        replay(be, stream)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    nbytes = 32.0 * (1 << n_cpu) * shape["passes"]
    total = sum(times)
    return {"value": nbytes * len(times) / total / 1e9, "ms_per_step": 1e3 * total / len(times), "shape": shape, "qubits": n_cpu}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, n, L, kind = workload_for(args.gpus, args.qubits, args.circuit)
    cores = os.cpu_count() or 1
    try:
        n_cpu = args.cpu_qubits or pick_cpu_qubits(kind, args.steps + args.warmup)
        r = run_reference_steps(kind, n_cpu, args.steps, args.warmup)
    except Exception as e:  # the oracle build is missing: say so, exit 0
        print(json.dumps({"impl": "reference", "unavailable": "%s: %s" % (type(e).__name__, e)}))
        return
    sample = "%s-%d full circuit (%d fused passes), state resident in host RAM, %d OpenMP threads" % (
        kind, r["qubits"], r["shape"]["passes"], cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "qubits": n, "local_qubits": L, "cluster_size": 4, "sample_qubits": r["qubits"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=None, help="override the number of qubits of the workload")
    ap.add_argument("--circuit", default=None, choices=[None, "qft", "random"])
    ap.add_argument("--cpu-qubits", type=int, default=None, help="size of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-batch", action="store_true",
                    help="one launch per fused gate of the plan (HIQ_FLAG_NO_BATCH): no folding of diagonal passes")
    args = ap.parse_args()

    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
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

    # ---- schedule once (host), outside the timed region of `value`
    stream, shape = schedule_circuit(n, L, cmds, rank, size)
    passes = shape["passes"]

    def fresh_backend():
        return backends.SimulatorMPI(gate_fusion=True, rnd_seed=12345, num_local_qubits=L, max_fused_qubits=4)

    # ---- value: replay on the resident state
    be = fresh_backend()
    sim = be._simulator
    sim.allocate_qureg(list(range(n)), 0)
    sim.synchronize()
    ext = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local))
    for _ in range(args.warmup):
        replay(be, stream)  # each replay starts by re-applying the initial relabelling (no data motion)
    sim.synchronize()
    sim.collect_timings()
    world.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = M.launch_count()
    stats0 = sim.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        replay(be, stream)
    e1.record(ext)
    e1.synchronize()
    sim.synchronize()
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = M.launch_count() - launches0
    stats1 = sim.stats()
    timings = sim.collect_timings()
    world.barrier()

    # max over ranks of the device time; sums of work over ranks
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    work = torch.tensor([32.0 * (1 << L) * (stats1["dense_passes"] + stats1["diag_passes"] + stats1["scale_passes"]
                                             - stats0["dense_passes"] - stats0["diag_passes"] - stats0["scale_passes"]),
                         float(launches), stats1["swap_bytes_sent"] - stats0["swap_bytes_sent"]], dtype=torch.float64, device="cuda")
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    total_bytes, total_launches, total_swap_bytes = (float(x) for x in work.tolist())
    value = total_bytes / (ms_total * 1e-3) / 1e9
    ms_per_step = ms_total / args.steps

    # ---- roofline of the dominant kernel (this rank's launches; rank 0 reports)
    # timings: (kind, k, variant, ms, n_ref) per launch; n_ref = passes of the reference's plan the launch carried
    groups = {}
    for kind_id, k, variant, ms, n_ref in timings:
        folded = kind_id == 1 and n_ref > 1
        # dense DIRECT launches of block-structured matrices carry their mixing bits in variant bits 8..
        groups.setdefault((kind_id, k if kind_id != 2 else 0, variant & 0xff, folded, variant >> 8), []).append((ms, n_ref))
    names = {1: "dense", 2: "diag_batch", 3: "scale", 4: "swap"}
    vnames = {0: "", 1: "direct", 2: "tiled", 3: "dmma"}

    def gname(g):
        if g[0] == 4:
            return "swap_q%d%s" % (g[1], "_wait_for_peers" if g[2] == 1 else "")
        if g[0] == 2:
            return "diag_batch" if not args.no_batch else "diag"
        return "%s_k%d_%s%s%s" % (names.get(g[0], "?"), g[1], vnames.get(g[2], ""), ("_mix%d" % g[4]) if g[4] else "",
                                  "+prediag" if g[3] else "")
    peak, peak_src = measured_peaks()
    gate_groups = {g: v for g, v in groups.items() if g[0] in (1, 2, 3)}
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
                            "diagonal passes of the plan does their work in the same pass (effective = achieved x passes per launch)"}
        # measured DRAM traffic of the same kernel (ncu --set full capture at L = 30, profiles/ncu_traffic.json)
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
    swap_gbs = None
    # the exchange itself; the time a rank waits for its peers to reach the swap (rank skew) is reported apart
    swap_ms = sum(t[3] for t in timings if t[0] == 4 and t[2] == 0)
    swap_wait_ms = sum(t[3] for t in timings if t[0] == 4 and t[2] == 1)
    if swap_ms > 0:
        swap_gbs = (stats1["swap_bytes_sent"] - stats0["swap_bytes_sent"]) / (swap_ms * 1e-3) / 1e9
    del be, sim, ext
    gc.collect()
    torch.cuda.empty_cache()

    # ---- e2e: the full reference-facing pipeline, every step from host inputs to measured bits
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.steps
        per_step, h2d, d2h = [], 0.0, 0.0
        for it in range(1 + e2e_steps):
            step_cmds = copy.deepcopy(cmds)
            world.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            be2 = fresh_backend()
            eng = cengines.HiQMainEngine(be2, [cengines.GreedyScheduler(cluster_size=4)])
            eng.receive([ops.AllocateQureg(list(range(n)), 0)])
            eng.receive(step_cmds)
            eng.receive([ops.Measure(list(range(n)))])
            be2._simulator.synchronize()
            world.barrier()
            dt = time.perf_counter() - t0
            st = be2._simulator.stats()
            if it >= 1:
                per_step.append(dt)
                h2d += be2.h2d_bytes
                d2h += st["d2h_bytes"] + n  # + the measured bits returned to the caller
            e2e_passes = st["dense_passes"] + st["diag_passes"] + st["scale_passes"]
            be2.main_engine = None  # break the engine<->backend cycle: the 128 GiB slab must go before the next step
            del eng, be2
            gc.collect()
        tt = torch.tensor([sum(per_step)], dtype=torch.float64, device="cuda")
        ww = torch.tensor([32.0 * (1 << L) * e2e_passes * len(per_step)], dtype=torch.float64, device="cuda")
        if size > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ww.item()) / float(tt.item()) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / len(per_step)), "d2h_bytes_per_step": int(d2h / len(per_step)),
               "seconds_per_step": float(tt.item()) / len(per_step), "steps": len(per_step),
               "includes": "engine construction, allocate_qureg, host scheduling + fusion, all passes/swaps, Measure(all)"}

    # ---- CPU baseline (rank 0, N=1): the unmodified reference on a bounded sample
    cpu_baseline = None
    if rank == 0 and size == 1 and not args.no_cpu_baseline:
        try:
            n_cpu = args.cpu_qubits or 26
            r = run_reference_steps(kind, n_cpu, 1, 1)
            cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                            "sample": "%s-%d full circuit (%d fused passes) on the unmodified reference engine, %s OpenMP threads"
                                      % (kind, n_cpu, r["shape"]["passes"], os.environ.get("OMP_NUM_THREADS"))}
        except Exception as e:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                            "sample": "unavailable: %s" % e}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": name, "qubits": n, "local_qubits": L, "slab_gib_per_gpu": 16.0 * (1 << L) / 2 ** 30,
                       "cluster_size": 4, "gates": shape["gates"], "fused_passes": passes,
                       "hbm_passes_per_step": int(stats1["gate_launches"] - stats0["gate_launches"]) // args.steps,
                       "batching": "off (one launch per fused gate)" if args.no_batch else
                                   "diagonal fused gates folded into the next dense launch / batched per pass",
                       "swaps": shape["swaps"],
                       "swap_qubits": shape["swap_qubits"], "l2_policy": "inputs larger than L2 (slab >> 126 MB), no flush",
                       "timing": "CUDA events on the engine stream, max over ranks",
                       "series": "BASELINE.json configs: N=1 is QFT-33 (33q@1; its diagonal fused gates are folded into "
                                 "neighbouring launches, so effective GB/s exceeds the physical rate), N>=2 are random-34/35/35 "
                                 "(one HBM pass per fused gate); the like-for-like single-GPU figure of the random series is "
                                 "`bench.py --circuit random --qubits 33` (or 30: profiles/r01o_bench_n1_random30.json)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "circuit_seconds": {"device_only": ms_per_step * 1e-3, "host_schedule_s": shape["host_schedule_s"],
                                "end_to_end": e2e["seconds_per_step"] if e2e else None},
            "swap_nvlink_gbs_per_gpu": swap_gbs,
            "swap_wait_for_peers_ms_per_step": swap_wait_ms / args.steps,
            "swap_transport": {"peer_mapped_in_place": int(stats1["swaps_p2p"] - stats0["swaps_p2p"]),
                               "staged_nccl": int(stats1["swaps_staged"] - stats0["swaps_staged"]),
                               "packed_peer_read": int(stats1.get("swaps_packed", 0) - stats0.get("swaps_packed", 0)),
                               "nvlink_peak_gbs_per_dir": 900.0,
                               "frac_of_nvlink": (swap_gbs / 900.0) if swap_gbs else None},
            "kernel_breakdown": breakdown,
            "host_enqueue_seconds_per_step": t_host / args.steps,
        }
        print(json.dumps(line))


if __name__ == "__main__":
    main()

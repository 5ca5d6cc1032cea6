"""bench.py's reference arm (`--impl reference`) on the CPU: the JSON-line contract, the OpenMP team it really runs with when
the launcher exports OMP_NUM_THREADS=1 (torchrun does for nproc > 1 — the defect of round 1's SCALE ratios), the ranks other
than 0 leaving without work, and that the arm's process maps none of this repository's native modules (the timed engine
and the scheduler that planned the stream are the unmodified reference under oracle/_ref)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
SMALL = ["--impl", "reference", "--circuit", "random", "--qubits", "16", "--cpu-qubits", "14", "--steps", "1", "--warmup", "1",
         "--cpu-budget", "2"]

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")), reason="oracle/_ref is not built")


def run_bench(extra, env_over, prog=None):
    env = {k: v for k, v in os.environ.items() if not k.startswith("HIQ_BENCH_") and k not in ("RANK", "OMP_NUM_THREADS", "OMP_PROC_BIND")}
    env.update(env_over)
    cmd = [sys.executable] + (["-c", prog, BENCH] if prog else [BENCH]) + SMALL + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def last_line(out):
    return json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])


@needs_ref
def test_reference_arm_line_and_thread_count_under_torchrun_environment():
    # torchrun's environment for nproc > 1: OMP_NUM_THREADS=1 exported to every rank
    d = last_line(run_bench(["--gpus", "2"], {"OMP_NUM_THREADS": "1", "RANK": "0", "HIQ_BENCH_CPU_THREADS": "2"}))
    assert d["impl"] == "reference" and "unavailable" not in d
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["metric"] == "gate_apply_effective_hbm_gbs" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 2 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # the team the reference engine really ran with: the intended one, not the launcher's 1, and it is what the line states
    cpu = d["cpu_baseline"]
    assert cpu["kind"] == "reference" and cpu["cores"] == 2 and cpu["omp_num_threads_env"] == "2"
    assert "2 OpenMP threads in effect" in cpu["sample"] and cpu["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a bounded sample of the same generator, said so
    assert d["config"]["workload"] == "random-16" and d["config"]["sample_qubits"] == 14 and d["config"]["same_config"] is False
    assert d["config"]["caveat"]
    # N > 1: the reference's own swap (pack -> all_to_all -> unpack) on N ranks of the multi-process reference build
    sw = d["swap_reference_cpu"]
    assert sw["ranks"] == 2 and sw["swapped_qubits"] == 1 and sw["value"] > 0


def test_reference_arm_other_ranks_exit_without_work():
    out = run_bench(["--gpus", "2"], {"RANK": "1", "OMP_NUM_THREADS": "1"})
    assert out.strip() == ""


@needs_ref
def test_reference_arm_maps_no_native_module_of_this_repository():
    # environment already in its final form -> bench.py does not re-execute itself and the wrapper survives to list the mappings
    prog = ("import runpy, sys; sys.argv = sys.argv[1:]; runpy.run_path(sys.argv[0], run_name='__main__')\n"
            "libs = sorted({ln.split()[-1] for ln in open('/proc/self/maps') if ln.rstrip().endswith('.so')})\n"
            "print('MAPPED ' + ' '.join(libs))")
    out = run_bench(["--gpus", "1", "--no-swap-baseline"], {"OMP_NUM_THREADS": "2", "HIQ_BENCH_CPU_THREADS": "2", "HIQ_BENCH_OMP_FIXED": "1",
                                                            "OMP_PROC_BIND": "spread"}, prog)
    d = last_line(out)
    assert d["impl"] == "reference" and d["cpu_baseline"]["cores"] == 2
    mapped = [ln for ln in out.splitlines() if ln.startswith("MAPPED ")][-1].split()[1:]
    ours = [p for p in mapped if p.startswith(ROOT)]
    assert ours, "the reference modules under oracle/_ref should be mapped"
    assert all(os.sep + os.path.join("oracle", "_ref") + os.sep in p for p in ours), ours
    assert not any("libhiq_b200" in p or "hiqsimulator_b200" in p for p in mapped), mapped

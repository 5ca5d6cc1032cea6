"""The invalid-input fuzzers of tools/ on a few seeds each, in their own processes (a crash shows as a signal, a hang as the
timeout): the reference-facing class on dry-run engines — every call accepted or refused, on all ranks alike —, the
scheduler module, and the launchers' host side.  The long campaigns are logged under profiles/r02q_fuzz_*.log."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_tool(name, *args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", name), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, (res.returncode, res.stdout[-800:], res.stderr[-1500:])
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("seeds ")][-1]
    return res.stdout, line


def test_api_with_invalid_arguments_never_crashes_or_hangs():
    out, line = run_tool("fuzz_invalid_arguments.py", "0", "60")
    m = re.search(r"calls ok (\d+) refused (\d+)", line)
    assert int(m.group(1)) > 100 and int(m.group(2)) > 100, line   # both sides of every check are exercised


def test_every_rank_accepts_or_refuses_a_call_alike():
    out, line = run_tool("fuzz_invalid_arguments.py", "0", "60", "--ranks")
    assert "DIVERGE" not in out, out[-1500:]
    assert line.endswith("rank-divergent 0"), line


def test_scheduler_module_with_degenerate_inputs_never_crashes_or_hangs():
    out, line = run_tool("fuzz_sched_invalid.py", "0", "400")
    m = re.search(r"ok (\d+) refused (\d+)", line)
    assert int(m.group(1)) > 50 and int(m.group(2)) > 50, line


def test_launchers_host_side_refuses_invalid_geometry_with_a_status():
    out, line = run_tool("fuzz_launcher_arguments.py", "0", "1500")
    m = re.search(r"ok (\d+) refused (\d+)", line)
    assert int(m.group(1)) > 100 and int(m.group(2)) > 500, line

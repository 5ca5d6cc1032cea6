"""Launching the rank bodies of the multi-process tests under torch.distributed.run (one process per rank / GPU).

A free rendez-vous port is picked by binding port 0; another process can still grab it before torchrun binds it
(EADDRINUSE, seen once on the 8-GPU box), so the launch is retried with a fresh port."""
import os
import socket
import subprocess
import sys


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_torchrun(R, script, args, env=None, timeout=600, attempts=3):
    """-> CompletedProcess of `torchrun --nproc-per-node R script args...` (R == 1: plain python)"""
    env = dict(os.environ if env is None else env)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    res = None
    for _ in range(attempts):
        if R == 1:
            cmd = [sys.executable, script, *args]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(R), "--master-addr", "127.0.0.1",
                   "--master-port", str(free_port()), script, *args]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if res.returncode == 0 or "EADDRINUSE" not in res.stderr:
            break
    return res

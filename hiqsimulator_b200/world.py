"""Process-group plumbing: one process per GPU, launched by torchrun.

Replaces `mpirun -np R` + mpi4py in the reference (reference:
hiq/projectq/backends/_sim/_simulator_mpi.py:41-45).  torch.distributed is used ONLY to hand rank
0's NCCL unique id to the other ranks and for barriers / timing reductions in the harness; the
engine's own collectives are NCCL calls issued from C++ (csrc/comm.cpp, csrc/engine.cpp).
"""
from __future__ import annotations

import os

from . import _cppsim_mpi as _M


def env_rank():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def init_world(flags: int = 0, backend: str | None = None):
    """Initialise the engine world from torchrun's environment. Returns (rank, world_size)."""
    rank, world, local = env_rank()
    dry = bool(flags & _M.FLAG_DRY_RUN)
    if world == 1:
        _M.init_world(0, 1, b"", local if not dry else 0, flags)
        return 0, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "gloo" if dry or not torch.cuda.is_available() else "nccl"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    if dry:
        uid = b""
    else:
        torch.cuda.set_device(local)
        box = [_M.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    _M.init_world(rank, world, uid, local, flags)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def gather_objects(obj, dst: int = 0):
    """Gather a picklable object from every rank on `dst` (None elsewhere)."""
    rank, world, _ = env_rank()
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world if rank == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _ranks_of(item):
    """world size of a multi-GPU case, read from its id (golden name r<R>_..., trailing -<R>] or [<R>])"""
    import re
    name = item.name
    for pat in (r"\[r(\d)_", r"-(\d)\]$", r"\[(\d)\]$"):
        m = re.search(pat, name)
        if m:
            return int(m.group(1))
    return None


def pytest_collection_modifyitems(config, items):
    # HIQ_TEST_R=4,8 keeps, among the multi-GPU cases, only those of the listed world sizes (GPU time is charged per GPU)
    only = os.environ.get("HIQ_TEST_R")
    if only:
        keep = {int(x) for x in only.split(",") if x}
        selected, dropped = [], []
        for item in items:
            r = _ranks_of(item) if ("multi_gpu" in item.name or "direct_diff" in item.name) else None
            (dropped if (r is not None and r not in keep) else selected).append(item)
        if dropped:
            config.hook.pytest_deselected(items=dropped)
            items[:] = selected
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

import glob
import json
import os

import numpy as np

import scripts

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    # engine fixtures = <name>.json + <name>.npz (sched_*.json are scheduler fixtures)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    script = scripts.script_from_json(json.dumps(meta["script"]))
    outputs = []
    for j, m in enumerate(meta["outputs"]):
        if m is None:
            outputs.append(None)
        elif "id2pos" in m:
            outputs.append(({int(k): v for k, v in m["id2pos"].items()}, arrays["vec%d" % j]))
        elif "error" in m:
            outputs.append(("error", m["error"]))
        elif "list" in m:
            outputs.append(m["list"])
        else:
            outputs.append(m["float"])
    return meta["R"], script, outputs

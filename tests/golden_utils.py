"""Helpers to load tests/golden/*.npz fixtures (made by oracle/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np
import torch

from genrl_b200.config import genrl_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(case):
    z = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    meta = json.load(open(os.path.join(GOLDEN, f"{case}.json")))
    g = {"init": {}, "data": {}, "tape": [], "ref0": {}, "ref1": {}, "grads": {}}
    for k in z.files:
        head, rest = k.split("/", 1)
        arr = torch.from_numpy(z[k])
        if head == "tape":
            g["tape"].append((rest, arr))
        elif head == "grads":
            i, name = rest.split("/", 1)
            g["grads"].setdefault(int(i), {})[name] = arr
        else:
            g[head][rest] = arr
    g["tape"] = [(name.split("_", 1)[1], v) for name, v in sorted(g["tape"], key=lambda kv: kv[0])]
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"],
                       imag_horizon=meta["H"], agent=meta["kind"], **meta["overrides"])
    return g, meta, cfg

"""Access to the committed reference outputs in tests/golden (made by oracle/make_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-len(".in.npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.in.npz")))


class GoldenCase:
    def __init__(self, name):
        self.name = name
        ins = np.load(os.path.join(GOLDEN_DIR, f"{name}.in.npz"))
        self.raw = torch.from_numpy(ins["raw"])
        self.state = {k[len("state."):]: torch.from_numpy(ins[k]) for k in ins.files if k.startswith("state.")}
        self.extra = {k[len("extra."):]: torch.from_numpy(ins[k]) for k in ins.files if k.startswith("extra.")}
        flags = ins["flags"]
        self.track_stages = bool(flags[0])
        self.additive = self.extra.get("additive") if flags[1] else None
        self.bn = {0: None, 1: "train", 2: "eval"}[int(flags[2])]
        self.f32 = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.f32.npz")))
        self.f64 = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.f64.npz")))

    def bn_dict(self, dtype=torch.float32):
        if self.bn is None:
            return None
        return dict(training=self.bn == "train", running_mean=self.extra["running_mean"].to(dtype).clone(),
                    running_var=self.extra["running_var"].to(dtype).clone(), momentum=0.1, eps=1e-5)


def maxabs(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) if a.size else 0.0

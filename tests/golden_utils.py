"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["v040_seed0_b64", "v040_seed0_downsampled", "v040_perturbed_edge", "v040_two_sources", "small_hp"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.hp = json.loads(str(z["hp"]))
        self.sd, self.inputs, self.out, self.loss, self.grad = {}, {}, {}, {}, {}
        groups = {"sd": self.sd, "in": self.inputs, "out": self.out, "loss": self.loss, "grad": self.grad}
        for key in z.files:
            if "/" not in key:
                continue
            group, rest = key.split("/", 1)
            groups[group][rest] = z[key]
        self.sd = {k: torch.from_numpy(np.array(v)) for k, v in self.sd.items()}

    def raw(self):
        d = {k: v for k, v in self.inputs.items() if k != "decoded_reads"}
        d.setdefault("read_indices", None)
        return d


def load(name):
    return Golden(name)

"""Loading of tests/golden (outputs of the real reference, see oracle/make_golden.py) + input regeneration."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from rtm3d_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as _f:
    MANIFEST = json.load(_f)
CASES = {c["name"]: c for c in MANIFEST["cases"]}
_npz = None


def arrays(name: str) -> dict:
    global _npz
    if _npz is None:
        _npz = np.load(os.path.join(GOLDEN_DIR, "decode_golden.npz"))
    pre = name + "/"
    return {k[len(pre):]: _npz[k] for k in _npz.files if k.startswith(pre)}


def inputs(name: str):
    """Regenerate the case's inputs from its seed and check them against the recorded sha256."""
    c = CASES[name]
    logits, kpt = synth.head_outputs(c["B"], c["C"], c["H"], c["W"], c["seed"], c["kind"], kpt_channels=c["kpt"])
    h = hashlib.sha256()
    for t in logits + [kpt]:
        if t is not None:
            h.update(t.contiguous().numpy().tobytes())
    if h.hexdigest() != c["inputs_sha256"]:
        raise RuntimeError(f"golden case {name}: regenerated inputs do not match the recorded checksum "
                           f"(torch RNG drift?) -- regenerate with python -m oracle.make_golden")
    return logits, kpt


def image_rows(arr: dict, b: int) -> dict:
    n = int(arr["counts"][b])
    return {k: v[b, :n] for k, v in arr.items() if k != "counts"}

"""Generate tests/golden/boxfit_golden.npz by EXECUTING THE REAL REFERENCE's optim_decode_bbox3d (utils/model_utils.py:264-312)
on CPU.  TEST INFRASTRUCTURE ONLY.  Run in the build container:   python -m oracle.make_boxfit_golden

Inputs: KITTI-like boxes (class-wise reference dimensions +- 15 %, depth 6..45 m, yaw uniform) projected with the reference's
own calc_proj_corners (:147-152) through a KITTI camera matrix, plus pixel noise of three strengths (0.02 / 0.08 px: accepted
fits; 0.6 px: rejected by res.fun < 0.1) and a few garbage objects.  The reference's `minimize` is wrapped (not modified) so
that the raw solution vector and res.fun of every object are recorded next to what optim_decode_bbox3d returns."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

DIM_REF = [[1.53, 1.63, 3.88], [1.76, 0.66, 0.84], [1.74, 0.60, 1.76]]        # (h, w, l) per class: rtm3d_*_kitti.yaml dim_ref
REF_LOC = [0, -0.5, 20]                                                         # detect.py:74
K_CAM = np.array([721.5377, 0, 609.5593, 0, 721.5377, 172.854, 0, 0, 1], dtype=np.float64)


def make_inputs(seed=7, n=48):
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import model_utils
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, 3, size=n)
    uv = np.zeros((n, 8, 2))
    truth = np.zeros((n, 7))
    for i in range(n):
        h, w, l = np.array(DIM_REF[cls[i]]) * rng.uniform(0.85, 1.15, size=3)
        z = rng.uniform(6, 45)
        x = rng.uniform(-0.35, 0.35) * z
        y = rng.uniform(0.8, 1.9)
        ry = rng.uniform(-np.pi, np.pi)
        proj = model_utils.calc_proj_corners([h, w, l], [x, y, z], ry, K_CAM.reshape(3, 3))[:8]
        noise = (0.02, 0.08, 0.6)[i % 3]
        uv[i] = proj + rng.normal(0, noise, size=(8, 2))
        truth[i] = [ry, l, h, w, x, y, z]
    uv[-2:] = rng.uniform(0, 1000, size=(2, 8, 2))                              # garbage objects
    return cls.astype(np.int64), uv.astype(np.float32), truth


def main():
    assert ref_import.available(), "the reference checkout is not mounted"
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import model_utils
    cls, uv, truth = make_inputs()
    records = []
    real_minimize = model_utils.minimize

    def recording_minimize(*a, **k):
        res = real_minimize(*a, **k)
        records.append((np.array(res.x, dtype=np.float64), float(res.fun), int(res.nit)))
        return res

    model_utils.minimize = recording_minimize
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = model_utils.optim_decode_bbox3d(cls, uv, K_CAM.astype(np.float32), DIM_REF, REF_LOC)
    finally:
        model_utils.minimize = real_minimize
    assert len(records) == len(cls)
    x = np.stack([r[0] for r in records]); fun = np.array([r[1] for r in records]); nit = np.array([r[2] for r in records])
    path = os.path.join(ROOT, "tests", "golden", "boxfit_golden.npz")
    np.savez_compressed(path, cls=cls, uv=uv, truth=truth, K=K_CAM.astype(np.float32), dim_ref=np.array(DIM_REF), ref_loc=np.array(REF_LOC, dtype=np.float64),
                        x=x, fun=fun, nit=nit, out_class=np.asarray(out.get_field('class'), dtype=np.int64), out_Ry=np.asarray(out.get_field('Ry')),
                        out_dimension=np.asarray(out.get_field('dimension')), out_location=np.asarray(out.get_field('location')))
    print("wrote", path, "objects", len(cls), "accepted", int((fun < 0.1).sum()), "median nit", int(np.median(nit)))


if __name__ == "__main__":
    main()

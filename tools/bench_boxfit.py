"""Throughput of the batched 3D-box fit (rtm3d_fit_box3d) next to the reference's CPU implementation
(utils/model_utils.py:264-312 through the oracle port, scipy L-BFGS-B, one object at a time).
   python tools/bench_boxfit.py [--images 256] [--topk 100] [--cpu-objects 24]    -> one JSON line"""
import argparse, json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--topk", type=int, default=100)
ap.add_argument("--cpu-objects", type=int, default=24)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
from oracle import boxfit_ref as bf
from oracle.make_boxfit_golden import DIM_REF, REF_LOC, K_CAM
from rtm3d_b200 import fit_packed

# synthetic detections: KITTI-like boxes projected through the camera + 0.05 px of noise (the fit's typical input)
rng = np.random.default_rng(5)
B, K = args.images, args.topk
n = B * K
cls = rng.integers(0, 3, size=n)
dims = np.array(DIM_REF)[cls] * rng.uniform(0.85, 1.15, size=(n, 3))
z = rng.uniform(6, 45, size=n); xw = rng.uniform(-0.35, 0.35, size=n) * z; yw = rng.uniform(0.8, 1.9, size=n); ry = rng.uniform(-np.pi, np.pi, size=n)
Kc = K_CAM.reshape(3, 3)
x8 = np.stack([np.sin(ry), np.cos(ry), dims[:, 2], dims[:, 0], dims[:, 1], xw, yw, z], axis=1)
uv = np.stack([bf.reproject(x8[i], Kc) for i in range(n)]) + rng.normal(0, 0.05, size=(n, 8, 2))
dev = torch.device("cuda:0")
verts = torch.as_tensor(uv.astype(np.float32), device=dev).reshape(B, K, 8, 2)
clst = torch.as_tensor(cls.astype(np.int64), device=dev).reshape(B, K)
cam = torch.as_tensor(K_CAM.astype(np.float32), device=dev)
for _ in range(3):
    fit = fit_packed(verts, clst, None, cam, DIM_REF, REF_LOC)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    fit = fit_packed(verts, clst, None, cam, DIM_REF, REF_LOC)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
acc = float(fit.accept.float().mean())
# CPU: the oracle port of the reference's fit on a bounded sample
m = args.cpu_objects
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    out = bf.optim_decode_bbox3d(cls[:m], uv[:m], K_CAM, np.array(DIM_REF), REF_LOC)
    cpu_s = time.perf_counter() - t0
print(json.dumps({"metric": "fitted 3D boxes/sec", "value": round(n / (ms * 1e-3), 1), "unit": "objects/s", "ms_per_batch": round(ms, 4),
                  "objects_per_batch": n, "accepted_fraction": round(acc, 4), "dtype": "f64",
                  "cpu_baseline": {"value": round(m / cpu_s, 2), "unit": "objects/s", "cores": 1, "kind": "port",
                                   "sample": f"{m} objects, oracle port of optim_decode_bbox3d (scipy L-BFGS-B with the reference's options)",
                                   "accepted_fraction": round(len(out['index']) / m, 4)},
                  "config": {"workload": f"box fit behind cfg4: {B} images x {K} detections, KITTI-like boxes + 0.05 px noise"}}))

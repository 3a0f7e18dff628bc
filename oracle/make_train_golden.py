"""Generate tests/golden/train_golden.npz with the REFERENCE's own helpers (utils.data_utils._compute_gaussian_radius /
gaussian2D, models.nets.module.FocalLoss, utils.model_utils.sigmoid_hm) on CPU.  TEST INFRASTRUCTURE ONLY.
    python -m oracle.make_train_golden
_build_targets itself cannot be imported (albumentations, un-vendored KITTI devkit): its splat loop
(datasets/dataset_reader.py:255-273) is driven from oracle/train_ref.main_targets with the reference's helpers plugged in."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, train_ref  # noqa: E402


def make_objects(seed=11, N=60, B=4, C=3, H=96, W=320):
    rng = np.random.default_rng(seed)
    w = rng.uniform(3, 70, N); h = rng.uniform(3, 45, N)
    cx = rng.uniform(-5, W + 5, N); cy = rng.uniform(-3, H + 3, N)
    bbox = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1).astype(np.float32)
    cls = rng.integers(0, C, N); img = rng.integers(0, B, N)
    mask = (rng.random(N) > 0.15).astype(np.uint8); noise = (rng.random(N) > 0.8).astype(np.uint8)
    return bbox, cls.astype(np.int64), img.astype(np.int64), mask, noise, (B, C, H, W)


def main():
    assert ref_import.available()
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import data_utils, model_utils
    from models.nets.module import FocalLoss
    bbox, cls, img, mask, noise, (B, C, H, W) = make_objects()
    m_hm, m_proj, m_off, sigma, radius = train_ref.main_targets(bbox, cls, img, mask, noise, B, C, H, W,
                                                                radius_fn=data_utils._compute_gaussian_radius, gaussian_fn=data_utils.gaussian2D)
    g = torch.Generator().manual_seed(5)
    logits = torch.randn((B, C, H, W), generator=g) * 2 - 2
    target = torch.from_numpy(m_hm).float()
    lg = logits.clone().requires_grad_(True)
    loss = FocalLoss(2.0, 4.0)(model_utils.sigmoid_hm(lg.clone()), target)
    loss.backward()
    empty_loss = FocalLoss(2.0, 4.0)(model_utils.sigmoid_hm(logits.clone()), torch.zeros_like(target))
    path = os.path.join(ROOT, "tests", "golden", "train_golden.npz")
    np.savez_compressed(path, bbox=bbox, cls=cls, img_id=img, mask=mask, noise_mask=noise, shape=np.array([B, C, H, W]), m_hm=m_hm.astype(np.float32),
                        m_proj=m_proj, m_off=m_off, sigma=sigma, radius=radius, logits=logits.numpy(), loss=float(loss), grad=lg.grad.numpy(),
                        empty_loss=float(empty_loss))
    print("wrote", path, "positives", int((m_hm == 1).sum()), "loss", float(loss), "empty", float(empty_loss))


def reference_loss_class():
    """The reference's RTM3DLoss; the modules it imports but __call__ never touches (albumentations transforms, the un-vendored
    KITTI devkit) are stubbed."""
    import types
    sys.dont_write_bytecode = True
    if ref_import.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REF_ROOT)
    for name in ("preprocess", "preprocess.transforms", "datasets", "datasets.data", "datasets.data.kitti", "datasets.data.kitti.devkit_object",
                 "datasets.data.kitti.devkit_object.utils"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["preprocess.transforms"].ToAbsoluteCoords = lambda: None
    sys.modules["datasets.data.kitti.devkit_object"].utils = sys.modules["datasets.data.kitti.devkit_object.utils"]
    from models.rtm3d_loss import RTM3DLoss
    from utils.ParamList import ParamList
    return RTM3DLoss, ParamList


def reference_loss(pred, fields, weights):
    """(loss, parts, gradients w.r.t. the four logit maps) from the reference's RTM3DLoss.__call__."""
    import types
    RTM3DLoss, ParamList = reference_loss_class()
    ns = types.SimpleNamespace
    cfg = ns(MODEL=ns(FOCAL_LOSS_ALPHA=2.0, FOCAL_LOSS_BEDA=4.0),
             DATASET=ns(GAUSSIAN_SIGMA_MAX=3., GAUSSIAN_SIGMA_MIN=1., BBOX_AREA_MAX=10., BBOX_AREA_MIN=1., VERTEX_OFFSET_INFER=[1.0]),
             TRAINING=ns(W_MKF=weights[0], W_VKF=1.0, W_VFM=weights[1], W_M_OFF=weights[2], W_V_OFF=weights[3]))
    t = ParamList((1, 1))
    for k, v in fields.items():
        t.add_field(k, v.clone())
    leaves = [p.clone().requires_grad_(True) for p in pred]
    loss, parts = RTM3DLoss(cfg)([x.clone() for x in leaves], t)        # (sigmoid_hm works in place: on the clones)
    loss.backward()
    return loss.detach(), parts, [x.grad for x in leaves]


def main_loss():
    pred, fields = train_ref.make_loss_case()
    loss, parts, grads = reference_loss(pred, fields, train_ref.LOSS_WEIGHTS)
    path = os.path.join(ROOT, "tests", "golden", "loss_golden.npz")
    np.savez_compressed(path, parts=parts.numpy(), **{f"grad{i}": g.numpy() for i, g in enumerate(grads)})
    print("wrote", path, "parts", parts.tolist())


if __name__ == "__main__":
    main()
    main_loss()

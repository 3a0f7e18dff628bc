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


if __name__ == "__main__":
    main()

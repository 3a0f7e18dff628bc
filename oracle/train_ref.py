"""CPU restatement of the training-side pieces next to the decoder.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  main_targets   datasets/dataset_reader.py:215-291 (_build_targets: centres :223-228, Gaussian splat :255-273) with
                 utils/data_utils.py:7-13 (bbox_center), :97-125 (_compute_gaussian_radius, dynamic_radius), :128-141 (gaussian2D)
  focal_loss     models/nets/module.py:41-68 (FocalLoss.forward) on utils/model_utils.py:10-14 (sigmoid_hm), as called at
                 models/rtm3d_loss.py:283

Pin status: PINNED -- tests/test_train_side.py compares gaussian radius / kernel and the loss with the imported reference
functions (utils.data_utils, models.nets.module, utils.model_utils are importable; datasets.dataset_reader is not: it needs
albumentations and the un-vendored KITTI devkit, so _build_targets' loop is restated here around the reference's own helpers)
and with tests/golden/train_golden.npz (oracle/make_train_golden.py, generated with those reference helpers).
"""
from __future__ import annotations

import numpy as np
import torch


def gaussian_radius(bboxes, min_overlap=0.7):
    height, width = np.ceil(bboxes[:, 3] - bboxes[:, 1]), np.ceil(bboxes[:, 2] - bboxes[:, 0])
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return np.minimum(np.minimum(r1, r2), r3)


def gaussian2d(sigma, radius):
    off = np.arange(-radius, radius + 1, 1)
    ox, oy = np.meshgrid(off, off)
    oy, ox = oy.flatten(), ox.flatten()
    return np.exp(-1 * (ox ** 2 + oy ** 2) / (2 * (sigma ** 2))), ox.astype(np.int32), oy.astype(np.int32)


def main_targets(bbox_hm, cls, img_id, mask, noise_mask, B, C, H, W, radius_fn=None, gaussian_fn=None):
    """Returns (m_hm f64 [B,C,H,W], m_proj i64 [N,2], m_off f64 [N,2], sigma [N], radius [N]).  radius_fn / gaussian_fn: the
    reference's own helpers when available (data_utils._compute_gaussian_radius, data_utils.gaussian2D)."""
    radius_fn = radius_fn or gaussian_radius
    gaussian_fn = gaussian_fn or gaussian2d
    bbox_hm = np.asarray(bbox_hm, dtype=np.float64)
    centers = np.stack([(bbox_hm[:, 0] + bbox_hm[:, 2]) / 2, (bbox_hm[:, 1] + bbox_hm[:, 3]) / 2], axis=1)
    m_proj = centers.astype(np.int64)
    m_off = centers - m_proj
    rad = radius_fn(bbox_hm)
    sigma, radius = (2 * rad + 1) / 6, np.ceil(rad)
    m_hm = np.zeros((B, C, H, W), dtype=np.float64)
    for i in range(len(bbox_hm)):
        if not mask[i]:
            continue
        kern, xs, ys = gaussian_fn(sigma[i], radius[i])
        kern = kern.copy()
        if noise_mask[i]:
            kern[len(xs) // 2] = 0.9999
        mx, my = xs + m_proj[i, 0], ys + m_proj[i, 1]
        valid = (mx >= 0) & (mx < W) & (my >= 0) & (my < H)
        plane = m_hm[int(img_id[i]), int(cls[i])]
        plane[my[valid], mx[valid]] = np.maximum(plane[my[valid], mx[valid]], kern[valid])
    return m_hm, m_proj, m_off, sigma, radius


def focal_loss(logits: torch.Tensor, target: torch.Tensor, alpha=2.0, beta=4.0) -> torch.Tensor:
    pred = logits.sigmoid().clamp(min=1e-4, max=1 - 1e-4)
    pos, neg = target.eq(1).float(), target.lt(1).float()
    neg_w = torch.pow(1 - target, beta)
    pos_loss = (torch.log(pred) * torch.pow(1 - pred, alpha) * pos).sum()
    neg_loss = (torch.log(1 - pred) * torch.pow(pred, alpha) * neg_w * neg).sum()
    num = pos.sum()
    return -neg_loss if num == 0 else -(pos_loss + neg_loss) / num

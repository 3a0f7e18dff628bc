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


def rtm3d_loss(pred_logits, fields, weights, alpha=2.0, beta=4.0):
    """models/rtm3d_loss.py:268-340 (RTM3DLoss.__call__) restated: focal loss on the main heat-map + the three gather-L1 losses.
    ``fields``: dict of the targets' tensors; ``weights`` = (W_MKF, W_VFM, W_M_OFF, W_V_OFF).  Does not modify the logits.
    Returns (loss, [main_kf, ver_coor, main_offset, vertex_offset, loss])."""
    import torch.nn.functional as F
    m_hm_pred, ver_coor_pred, m_off_pred, v_off_pred = pred_logits
    m_projs, v_projs = fields["m_proj"].long(), fields["v_proj"].long()
    img_id = fields["img_id"].long()
    m_mask, not_noise = fields["mask"].bool(), fields["noise_mask"].bool().bitwise_not()
    mask_3d, v_mask = fields["mask_3d"].bool(), fields["v_mask"].bool()
    loss_main_kf = focal_loss(m_hm_pred, fields["m_hm"], alpha, beta)
    num_vc = ver_coor_pred.shape[1] // 2
    ofm_valid = m_mask & not_noise & mask_3d
    ofm_valid_expand = v_mask[ofm_valid].view(-1)
    vc = ver_coor_pred.permute(0, 2, 3, 1)[img_id[ofm_valid], m_projs[ofm_valid][:, 1], m_projs[ofm_valid][:, 0]].reshape(-1, 2)
    loss_ver_coor = F.l1_loss(vc[ofm_valid_expand], fields["v_coor_off"][ofm_valid].view(-1, 2)[ofm_valid_expand], reduction="mean")
    bs = img_id.view(-1, 1).repeat(1, num_vc).view(-1)
    vp = v_projs.view(-1, 2)
    ver_valid = ofm_valid.view(-1, 1).repeat(1, num_vc).view(-1) & v_mask.view(-1)
    pos_v = v_off_pred.permute(0, 2, 3, 1)[bs[ver_valid], vp[ver_valid][:, 1], vp[ver_valid][:, 0]].sigmoid()
    loss_vertex_offset = F.l1_loss(pos_v, fields["v_off"].view(-1, 2)[ver_valid], reduction="mean")
    m_valid = m_mask & not_noise
    pos_m = m_off_pred.permute(0, 2, 3, 1)[img_id[m_valid], m_projs[m_valid][:, 1], m_projs[m_valid][:, 0]].sigmoid()
    loss_main_offset = F.l1_loss(pos_m, fields["m_off"][m_valid], reduction="mean")
    w_mkf, w_vfm, w_moff, w_voff = weights
    parts = [loss_main_kf * w_mkf, loss_ver_coor * w_vfm, loss_main_offset * w_moff, loss_vertex_offset * w_voff]
    loss = parts[0] + parts[1] + parts[2] + parts[3]
    return loss, parts + [loss]


LOSS_FIELDS = ("m_hm", "m_proj", "m_off", "v_coor_off", "v_proj", "v_off", "img_id", "mask", "noise_mask", "mask_3d", "v_mask")
LOSS_WEIGHTS = (1.0, 0.5, 1.0, 1.0)          # W_MKF, W_VFM, W_M_OFF, W_V_OFF of the golden case


def make_loss_case(seed=3, B=3, C=3, H=24, W=40, N=19, V=8):
    """Seeded logits + targets of the loss golden (shared by the generator and the live pin)."""
    g = torch.Generator().manual_seed(seed)
    pred = [torch.randn(B, C, H, W, generator=g) * 2 - 2, torch.randn(B, 2 * V, H, W, generator=g) * 3, torch.randn(B, 2, H, W, generator=g),
            torch.randn(B, 2, H, W, generator=g)]
    m_hm = torch.rand(B, C, H, W, generator=g) ** 6
    m_proj = torch.stack([torch.randint(0, W, (N,), generator=g), torch.randint(0, H, (N,), generator=g)], 1)
    m_proj[1] = m_proj[0]                                             # two objects on one pixel: their gradients add up
    img_id = torch.randint(0, B, (N,), generator=g)
    img_id[1] = img_id[0]
    for i in range(N):
        m_hm[img_id[i], i % C, m_proj[i, 1], m_proj[i, 0]] = 1.0
    fields = dict(m_hm=m_hm, m_proj=m_proj, m_off=torch.rand(N, 2, generator=g), v_coor_off=torch.randn(N, V, 2, generator=g) * 5,
                  v_proj=torch.stack([torch.randint(0, W, (N, V), generator=g), torch.randint(0, H, (N, V), generator=g)], 2),
                  v_off=torch.rand(N, V, 2, generator=g), img_id=img_id, mask=(torch.rand(N, generator=g) > 0.15).float(),
                  noise_mask=(torch.rand(N, generator=g) > 0.8).float(), mask_3d=(torch.rand(N, generator=g) > 0.1).float(),
                  v_mask=torch.rand(N, V, generator=g) > 0.3)
    return pred, fields

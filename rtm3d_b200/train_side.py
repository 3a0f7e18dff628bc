"""Host-side mirror of the training-side pieces next to the decoder over librtm3d_decode.so:

* ``build_main_targets`` -- the main heat-map target of ``DatasetReader._build_targets`` (datasets/dataset_reader.py:215-291:
  Gaussian splat of every labelled object's 2D-box centre, ``data_utils.dynamic_radius`` / ``gaussian2D``), for a whole batch
  in one launch instead of a per-object numpy loop;
* ``FocalLoss`` -- ``models.nets.module.FocalLoss`` (:41-68) applied to ``sigmoid_hm(logits)`` (utils/model_utils.py:10-14) as
  models/rtm3d_loss.py:283 does, with the gradient w.r.t. the logits from a second streaming kernel (``torch.autograd.Function``).

PyTorch owns memory and streams; the arithmetic runs in the CUDA library.  No CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _native


@dataclass
class MainTargets:
    m_hm: torch.Tensor      # f32 [B,C,H,W]
    m_proj: torch.Tensor    # int32 [N,2] integer centre (x, y) on the heat-map grid
    m_off: torch.Tensor     # f32 [N,2] sub-pixel offset of the centre
    sigma: torch.Tensor     # f32 [N]
    radius: torch.Tensor    # int32 [N]


def build_main_targets(bbox_hm: torch.Tensor, cls: torch.Tensor, img_id: torch.Tensor, mask: torch.Tensor, noise_mask: torch.Tensor,
                       B: int, C: int, H: int, W: int) -> MainTargets:
    """bbox_hm f32 [N,4]: the targets' 2D boxes divided by DOWN_SAMPLE (x1,y1,x2,y2 on the heat-map grid)."""
    if not bbox_hm.is_cuda:
        raise ValueError("build_main_targets: expected CUDA tensors (rtm3d_b200 has no CPU path)")
    dev = bbox_hm.device
    N = bbox_hm.shape[0]
    bbox_hm = bbox_hm.to(torch.float32).contiguous()
    cls = cls.to(device=dev, dtype=torch.int64).contiguous()
    img_id = img_id.to(device=dev, dtype=torch.int64).contiguous()
    mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
    noise_mask = noise_mask.to(device=dev, dtype=torch.uint8).contiguous()
    out = MainTargets(m_hm=torch.empty((B, C, H, W), dtype=torch.float32, device=dev), m_proj=torch.empty((N, 2), dtype=torch.int32, device=dev),
                      m_off=torch.empty((N, 2), dtype=torch.float32, device=dev), sigma=torch.empty((N,), dtype=torch.float32, device=dev),
                      radius=torch.empty((N,), dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        rc = _native.lib().rtm3d_encode_main_targets(bbox_hm.data_ptr(), cls.data_ptr(), img_id.data_ptr(), mask.data_ptr(), noise_mask.data_ptr(),
                                                     N, B, C, H, W, out.m_hm.data_ptr(), out.m_proj.data_ptr(), out.m_off.data_ptr(),
                                                     out.sigma.data_ptr(), out.radius.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _native.check(rc, "rtm3d_encode_main_targets")
    return out


class _FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, alpha, beta):
        if not logits.is_cuda:
            raise ValueError("FocalLoss: expected CUDA tensors (rtm3d_b200 has no CPU path)")
        logits_c, target_c = logits.detach().to(torch.float32).contiguous(), target.detach().to(torch.float32).contiguous()
        if logits_c.shape != target_c.shape:
            raise ValueError("FocalLoss: logits and target must have the same shape")
        dev = logits_c.device
        acc = torch.empty(3, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _native.lib().rtm3d_focal_loss(logits_c.data_ptr(), target_c.data_ptr(), logits_c.numel(), float(alpha), float(beta),
                                                acc.data_ptr(), loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_focal_loss")
        ctx.save_for_backward(logits_c, target_c, acc)
        ctx.alpha, ctx.beta = float(alpha), float(beta)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        logits_c, target_c, acc = ctx.saved_tensors
        dev = logits_c.device
        grad = torch.empty_like(logits_c)
        up = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            rc = _native.lib().rtm3d_focal_loss_grad(logits_c.data_ptr(), target_c.data_ptr(), logits_c.numel(), ctx.alpha, ctx.beta,
                                                     acc.data_ptr(), up.data_ptr(), grad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_focal_loss_grad")
        return grad, None, None, None


class FocalLoss(torch.nn.Module):
    """``models.nets.module.FocalLoss`` fused with the clamped sigmoid in front of it: ``loss = FocalLoss(alpha, beta)(logits, m_hm)``
    equals the reference's ``FocalLoss(alpha, beta)(model_utils.sigmoid_hm(logits), m_hm)`` (models/rtm3d_loss.py:283), without
    modifying the logits in place."""

    def __init__(self, alpha=2.0, beta=4.0):
        super().__init__()
        self.alpha, self.beta = float(alpha), float(beta)

    def forward(self, logits, target):
        return _FocalLossFn.apply(logits, target, self.alpha, self.beta)

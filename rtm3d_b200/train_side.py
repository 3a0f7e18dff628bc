"""Host-side mirror of the training-side pieces next to the decoder over librtm3d_decode.so:

* ``build_main_targets`` -- the main heat-map target of ``DatasetReader._build_targets`` (datasets/dataset_reader.py:215-291:
  Gaussian splat of every labelled object's 2D-box centre, ``data_utils.dynamic_radius`` / ``gaussian2D``), for a whole batch
  in one launch instead of a per-object numpy loop;
* ``FocalLoss`` -- ``models.nets.module.FocalLoss`` (:41-68) applied to ``sigmoid_hm(logits)`` (utils/model_utils.py:10-14) as
  models/rtm3d_loss.py:283 does, with the gradient w.r.t. the logits from a second streaming kernel (``torch.autograd.Function``);
* ``gather_l1_loss`` / ``RTM3DLoss`` -- the three gather-L1 losses of ``RTM3DLoss.__call__`` (models/rtm3d_loss.py:268-340) and the
  whole loss with the reference's call signature.

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


class _GatherL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap, img, x, y, c0, valid, target, sigmoid):
        if not fmap.is_cuda:
            raise ValueError("GatherL1Loss: expected CUDA tensors (rtm3d_b200 has no CPU path)")
        dev = fmap.device
        m = fmap.detach().to(torch.float32).contiguous()
        B, C, H, W = m.shape
        i64 = lambda t: t.to(device=dev, dtype=torch.int64).contiguous()
        img, x, y = i64(img), i64(x), i64(y)
        c0 = None if c0 is None else c0.to(device=dev, dtype=torch.int32).contiguous()
        valid = valid.to(device=dev, dtype=torch.uint8).contiguous()
        target = target.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1, 2)
        n = img.numel()
        if not (x.numel() == y.numel() == valid.numel() == target.shape[0] == n and (c0 is None or c0.numel() == n)):
            raise ValueError("GatherL1Loss: entry arrays of different lengths")
        acc = torch.empty(3, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        args = (m.data_ptr(), B, C, H, W, img.data_ptr(), x.data_ptr(), y.data_ptr(), None if c0 is None else c0.data_ptr(), valid.data_ptr(),
                target.data_ptr(), n, 1 if sigmoid else 0)
        with torch.cuda.device(dev):
            rc = _native.lib().rtm3d_gather_l1_loss(*args, acc.data_ptr(), loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_gather_l1_loss")
        ctx.keep = (m, img, x, y, c0, valid, target, acc)
        ctx.args = args
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        m, acc = ctx.keep[0], ctx.keep[-1]
        dev = m.device
        grad = torch.empty_like(m)
        up = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            rc = _native.lib().rtm3d_gather_l1_loss_grad(*ctx.args, acc.data_ptr(), up.data_ptr(), grad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_gather_l1_loss_grad")
        return grad, None, None, None, None, None, None, None


def gather_l1_loss(fmap, img, x, y, valid, target, c0=None, sigmoid=False):
    """mean |act(fmap[img, c0:c0+2, y, x]) - target| over the valid entries (models/rtm3d_loss.py:302-330), differentiable
    w.r.t. ``fmap`` [B,C,H,W]; ``act`` = sigmoid or identity.  One gather kernel instead of a permuted copy of the map."""
    return _GatherL1Fn.apply(fmap, img, x, y, c0, valid, target, bool(sigmoid))


class RTM3DLoss(torch.nn.Module):
    """``models.rtm3d_loss.RTM3DLoss`` (:268-340) over the CUDA library: the focal loss on the main heat-map and the three
    gather-L1 losses, weighted with ``config.TRAINING.W_*``; same call signature and return value
    ``(loss, tensor([main_kf, ver_coor, main_offset, vertex_offset, loss]))``.  ``targets`` needs ``get_field`` for m_hm, m_proj,
    m_off, v_coor_off, v_proj, v_off, img_id, mask, noise_mask, mask_3d, v_mask (the reference's ParamList)."""

    def __init__(self, config):
        super().__init__()
        self._config = config
        self._main_kf_loss = FocalLoss(config.MODEL.FOCAL_LOSS_ALPHA, config.MODEL.FOCAL_LOSS_BEDA)

    def forward(self, pred_logits, targets):
        m_hm_pred, ver_coor_pred, m_off_pred, v_off_pred = pred_logits
        dev = m_hm_pred.device
        f = lambda name: targets.get_field(name).to(dev)
        m_projs, v_projs = f("m_proj").long(), f("v_proj").long()
        img_id = f("img_id").long()
        m_mask, not_noise, mask_3d, v_mask = f("mask").bool(), f("noise_mask").bool().bitwise_not(), f("mask_3d").bool(), f("v_mask").bool()
        N, V = v_projs.shape[0], v_projs.shape[1]
        if ver_coor_pred.shape[1] // 2 != V:
            raise ValueError("offset_fr_main channels do not match the targets' vertex count")
        loss_main_kf = self._main_kf_loss(m_hm_pred, f("m_hm"))                                           # :283
        ofm_valid = m_mask & not_noise & mask_3d                                                          # :297
        ver_valid = (ofm_valid.view(-1, 1) & v_mask).reshape(-1)                                          # :298, :313
        bs = img_id.view(-1, 1).expand(N, V).reshape(-1)
        mx, my = m_projs[:, 0].view(-1, 1).expand(N, V).reshape(-1), m_projs[:, 1].view(-1, 1).expand(N, V).reshape(-1)
        c0 = (2 * torch.arange(V, device=dev, dtype=torch.int32)).view(1, -1).expand(N, V).reshape(-1)
        loss_ver_coor = gather_l1_loss(ver_coor_pred, bs, mx, my, ver_valid, f("v_coor_off").reshape(-1, 2), c0=c0)          # :300-308
        vp = v_projs.view(-1, 2)
        loss_vertex_offset = gather_l1_loss(v_off_pred, bs, vp[:, 0], vp[:, 1], ver_valid, f("v_off").reshape(-1, 2), sigmoid=True)   # :311-320
        m_valid = m_mask & not_noise                                                                      # :323
        loss_main_offset = gather_l1_loss(m_off_pred, img_id, m_projs[:, 0], m_projs[:, 1], m_valid, f("m_off"), sigmoid=True)      # :324-328
        T = self._config.TRAINING
        loss_main_kf = loss_main_kf * T.W_MKF
        loss_ver_coor = loss_ver_coor * T.W_VFM
        loss_main_offset = loss_main_offset * T.W_M_OFF
        loss_vertex_offset = loss_vertex_offset * T.W_V_OFF
        loss = loss_main_kf + loss_ver_coor + loss_main_offset + loss_vertex_offset
        return loss, torch.stack([loss_main_kf, loss_ver_coor, loss_main_offset, loss_vertex_offset, loss]).detach()

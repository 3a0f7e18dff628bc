"""Host-side mirror of the reference's 3D-box fit (utils/model_utils.py:264-312, called at detect.py:71-74) over
``rtm3d_fit_box3d``: every detection of a batch is fitted by one GPU thread (Levenberg-Marquardt in double precision on the
reference's objective, from the reference's start point, accepted when f < 0.1).

``optim_decode_bbox3d(clses, bbox3d_projs, K, ref_dim, ref_loc)`` has the reference's signature for ONE image and returns a
``Box3DFit`` whose ``get_field`` serves the names the reference's ParamList carries ('class', 'Ry', 'dimension', 'location',
'K'), so detect.py's drawing code reads it unchanged.  ``fit_packed`` is the asynchronous batched form for
``PackedDetections``.  No CPU path: the inputs go to the GPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native


@dataclass
class PackedBoxFit:
    """Fixed-size result of the fit for a batch; rows beyond counts[b] are zero with accept = 0."""
    loc: torch.Tensor       # f32 [B,K,3]
    dim: torch.Tensor       # f32 [B,K,3] (h, w, l)
    ry: torch.Tensor        # f32 [B,K]
    fun: torch.Tensor       # f32 [B,K]   value of the reprojection objective at the solution
    accept: torch.Tensor    # int32 [B,K] fun < 0.1 (utils/model_utils.py:298)
    x8: Optional[torch.Tensor] = None      # f64 [B,K,8] raw solution [sin, cos, l, h, w, X, Y, Z]
    iters: Optional[torch.Tensor] = None   # int32 [B,K]


def fit_packed(verts: torch.Tensor, cls: torch.Tensor, counts: Optional[torch.Tensor], cam: torch.Tensor, dim_ref, ref_loc=(0.0, -0.5, 20.0),
               max_iter: int = 100, want_solution: bool = False) -> PackedBoxFit:
    """verts f32 [B,K,8,2] (input pixels), cls int64 [B,K], counts int32 [B] or None, cam [B,9] / [9] / [3,3], dim_ref [C,3] (h,w,l)."""
    if not verts.is_cuda:
        raise ValueError("fit_packed: expected CUDA tensors (rtm3d_b200 has no CPU path)")
    dev = verts.device
    B, K = verts.shape[0], verts.shape[1]
    if tuple(verts.shape[2:]) != (8, 2):
        raise ValueError("verts must be [B,K,8,2]")
    verts = verts.to(torch.float32).contiguous()
    cls = cls.to(device=dev, dtype=torch.int64).contiguous()
    cam = torch.as_tensor(cam, dtype=torch.float32, device=dev).reshape(-1, 9).contiguous()
    if cam.shape[0] not in (1, B):
        raise ValueError("cam must hold one camera matrix or one per image")
    dim_ref = torch.as_tensor(np.asarray(dim_ref, dtype=np.float32), device=dev).reshape(-1, 3).contiguous()
    if counts is not None:
        counts = counts.to(device=dev, dtype=torch.int32).contiguous()
    e = lambda *sh, d=torch.float32: torch.empty(sh, dtype=d, device=dev)
    out = PackedBoxFit(loc=e(B, K, 3), dim=e(B, K, 3), ry=e(B, K), fun=e(B, K), accept=e(B, K, d=torch.int32),
                       x8=e(B, K, 8, d=torch.float64) if want_solution else None, iters=e(B, K, d=torch.int32) if want_solution else None)
    rl = (ctypes.c_float * 3)(*[float(v) for v in ref_loc])
    with torch.cuda.device(dev):
        rc = _native.lib().rtm3d_fit_box3d(
            verts.data_ptr(), cls.data_ptr(), counts.data_ptr() if counts is not None else None, cam.data_ptr(), 1 if cam.shape[0] == B and B > 1 else 0,
            dim_ref.data_ptr(), dim_ref.shape[0], rl, B, K, int(max_iter), out.loc.data_ptr(), out.dim.data_ptr(), out.ry.data_ptr(),
            out.fun.data_ptr(), out.accept.data_ptr(), out.x8.data_ptr() if want_solution else None, out.iters.data_ptr() if want_solution else None,
            torch.cuda.current_stream(dev).cuda_stream)
    _native.check(rc, "rtm3d_fit_box3d")
    out._keepalive = (verts, cls, counts, cam, dim_ref)
    return out


class Box3DFit:
    """What the reference's optim_decode_bbox3d returns (a ParamList with the fields below), for the accepted detections."""

    def __init__(self, fields):
        self._fields = fields

    def get_field(self, name):
        return self._fields[name]

    def has_field(self, name):
        return name in self._fields

    def fields(self):
        return list(self._fields)


def optim_decode_bbox3d(clses, bbox3d_projs, K, ref_dim, ref_loc, device="cuda:0") -> Box3DFit:
    """Drop-in for utils/model_utils.optim_decode_bbox3d (one image): clses (N,), bbox3d_projs (N,8,2), K (9,) or (3,3)."""
    clses = np.asarray(clses)
    n = len(clses)
    K9 = np.asarray(K, dtype=np.float32).reshape(9)
    if n == 0:
        return Box3DFit({'class': [], 'Ry': np.zeros((0,)), 'dimension': np.zeros((0, 3)), 'location': np.zeros((0, 3)), 'K': np.zeros((0, 9))})
    dev = torch.device(device)
    verts = torch.as_tensor(np.asarray(bbox3d_projs, dtype=np.float32), device=dev).reshape(1, n, 8, 2)
    fit = fit_packed(verts, torch.as_tensor(clses.astype(np.int64), device=dev).reshape(1, n), None, torch.as_tensor(K9, device=dev),
                     ref_dim, ref_loc)
    keep = fit.accept[0].bool().cpu().numpy()
    return Box3DFit({'class': [c for c, k in zip(clses.tolist(), keep) if k],
                     'Ry': fit.ry[0].cpu().numpy().astype(np.float64)[keep],
                     'dimension': fit.dim[0].cpu().numpy().astype(np.float64)[keep].reshape(-1, 3),
                     'location': fit.loc[0].cpu().numpy().astype(np.float64)[keep].reshape(-1, 3),
                     'K': np.tile(K9.astype(np.float64).reshape(1, 9), (int(keep.sum()), 1)),
                     'fun': fit.fun[0].cpu().numpy().astype(np.float64)[keep]})

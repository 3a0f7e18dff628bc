"""Tier C restatement: closed-form 3D box recovery at the Tier A peaks.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED for the decode formulas: the reference checkout has no depth / dimension / orientation heads and no
closed-form 3D recovery (its 3D step is a scipy fit, utils/model_utils.py:264-312).  The
formulas below are the spec from BASELINE.json's north_star (SMOKE-style and multi-bin decoders as published in the
SMOKE / CenterNet-ddd papers); this file is their normative text for this repository and `rtm3d_decode_box3d` is tested
against it.

PINNED sub-step: the geometric conventions ARE the reference's and are checked against its own functions
(tests/test_box3d.py: live import when /root/reference is mounted, committed goldens otherwise):
  rotation_matrix   utils/model_utils.py:66-76   Ry = [[c,0,s],[0,1,0],[-s,0,c]] with |sin|,|cos| < 1e-3 snapped to 0
  create_corners    utils/model_utils.py:80-119  corner order x in {+,-} L/2, y in {+,-} H/2, z in {+,-} W/2 nested in
                                                 that order; dims are (h, w, l)
  calc_proj_corners utils/model_utils.py:147-152 uv = (K X)[:2] / ((K X)[2] + 1e-6), K the 3x3 camera matrix
  camera matrix     datasets/dataset_reader.py:108,191-192  flat-9 row-major, fx = K[0], cx = K[2], fy = K[4], cy = K[5]

All arithmetic in float32 torch ops (the kernel is compared with rel 1e-4).
"""
from __future__ import annotations

import math

import torch

PI = math.pi


def sigmoid(x):
    return torch.sigmoid(x)


def project_corners(dim_hwl: torch.Tensor, loc: torch.Tensor, rot_y: torch.Tensor, cam9: torch.Tensor) -> torch.Tensor:
    """[N,3] (h,w,l), [N,3], [N], [9] -> [N,8,2] projected corners in the reference's order (create_corners :102-107)."""
    N = dim_hwl.shape[0]
    sn, cs = torch.sin(rot_y), torch.cos(rot_y)
    sn = torch.where(sn.abs() < 1e-3, torch.zeros_like(sn), sn)
    cs = torch.where(cs.abs() < 1e-3, torch.zeros_like(cs), cs)
    hl, hh, hw = 0.5 * dim_hwl[:, 2], 0.5 * dim_hwl[:, 0], 0.5 * dim_hwl[:, 1]
    out = torch.empty((N, 8, 2), dtype=torch.float32, device=dim_hwl.device)
    q = 0
    for i in (1, -1):
        for j in (1, -1):
            for k in (1, -1):
                X = cs * (hl * i) + sn * (hw * k) + loc[:, 0]
                Y = hh * j + loc[:, 1]
                Z = -sn * (hl * i) + cs * (hw * k) + loc[:, 2]
                px = cam9[0] * X + cam9[1] * Y + cam9[2] * Z
                py = cam9[3] * X + cam9[4] * Y + cam9[5] * Z
                pz = cam9[6] * X + cam9[7] * Y + cam9[8] * Z
                out[:, q, 0] = px / (pz + 1e-6)
                out[:, q, 1] = py / (pz + 1e-6)
                q += 1
    return out


def decode_box3d(flat: torch.Tensor, reg: torch.Tensor, cam9: torch.Tensor, dim_ref: torch.Tensor, n_classes: int,
                 multibin: bool = False, sigmoid_subpixel: bool = False, depth_ref=(28.01, 16.32)):
    """One image.  flat int64 [N] (c*H*W + y*W + x), reg [Creg,H,W] (Creg = 8: depth 1, sub-pixel 2, dims 3, sin/cos 2;
    Creg = 14: depth 1, sub-pixel 2, dims 3, bins 8), cam9 [9] already in heat-map units, dim_ref [C,3] rows (h,w,l).
    Returns dict(loc [N,3], dim [N,3], alpha [N], rot_y [N], corners2d [N,8,2])."""
    Creg, H, W = reg.shape
    HW = H * W
    cls = flat // HW
    rem = flat % HW
    yi, xi = rem // W, rem % W
    r = reg[:, yi, xi].float()                                   # [Creg,N]
    fx, cx, fy, cy = cam9[0], cam9[2], cam9[4], cam9[5]
    u = xi.float() + (sigmoid(r[1]) if sigmoid_subpixel else r[1])
    v = yi.float() + (sigmoid(r[2]) if sigmoid_subpixel else r[2])
    if multibin:
        z = 1.0 / (sigmoid(r[0]) + 1e-6) - 1.0
    else:
        z = r[0] * depth_ref[1] + depth_ref[0]
    loc = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], dim=-1)
    dim = torch.exp(sigmoid(r[3:6]) - 0.5).t() * dim_ref[cls]
    if multibin:
        a1 = torch.atan2(r[8], r[9]) - 0.5 * PI
        a2 = torch.atan2(r[12], r[13]) + 0.5 * PI
        alpha = torch.where(r[7] > r[11], a1, a2)
        rot_y = alpha + torch.atan2(u - cx, fx.expand_as(u))
    else:
        nrm = torch.sqrt(r[6] * r[6] + r[7] * r[7])
        o0, o1 = r[6] / nrm, r[7] / nrm
        alpha = torch.atan(o0 / (o1 + 1e-7))
        alpha = alpha + torch.where(o1 >= 0, torch.full_like(o1, -0.5 * PI), torch.full_like(o1, 0.5 * PI))
        rot_y = alpha + torch.atan(loc[:, 0] / (loc[:, 2] + 1e-7))
    rot_y = torch.where(rot_y > PI, rot_y - 2 * PI, rot_y)
    rot_y = torch.where(rot_y < -PI, rot_y + 2 * PI, rot_y)
    return dict(loc=loc, dim=dim, alpha=alpha, rot_y=rot_y, corners2d=project_corners(dim, loc, rot_y, cam9))

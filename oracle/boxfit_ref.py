"""CPU restatement of the reference's 3D-box fit.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows utils/model_utils.py of the reference:
  objective   aimFun   :155-177   f(x) = sum over the 8 corners of (xc fx/(zc+1e-4) + cx - u)^2 + (yc fy/(zc+1e-4) + cy - v)^2,
                                   x = [sin, cos, l, h, w, X, Y, Z], corner pattern +-0.5 in the create_corners order (:102-107)
  gradient    jac      :206-234   (the reference's own analytic gradient, with ITS constants: 1e-6 and zc^2 + 1e-6)
  driver      optim_decode_bbox3d :264-312: X0 = [0, 1, l_ref, h_ref, w_ref] + ref_loc, scipy L-BFGS-B with the reference's
              options (the `constraints` it passes are ignored by L-BFGS-B), accepted when res.fun < 0.1,
              Ry = arctan2(x0, x1), dimension = (x3, x4, x2) = (h, w, l), location = x5..7.

The objective is invariant under (sin, cos, l, w) -> (a sin, a cos, l / a, w / a), a != 0, and (up to the 1e-4 in the denominator)
under a common scale of (l, h, w, X, Y, Z): its minimum is a two-parameter family and
the (l, w) L-BFGS-B stops at depend on its path.  `canonical` maps a solution to the gauge sin^2 + cos^2 = 1; parity is
stated on res.fun, the accept decision, Ry, h, location, the canonical (l, w) and the reprojected corners.

Pin status: PINNED -- tests/test_boxfit.py compares this restatement with the imported reference (when /root/reference is
mounted) and with tests/golden/boxfit_golden.npz (outputs of the real reference, oracle/make_boxfit_golden.py).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

ACCEPT_FUN = 0.1                      # utils/model_utils.py:298
OPTIONS = {'disp': None, 'maxcor': 10, 'ftol': 2.220446049250313e-09, 'gtol': 1e-05, 'eps': 1e-08,
           'maxfun': 15000, 'maxiter': 15000, 'iprint': -1, 'maxls': 20, 'finite_diff_rel_step': None}


def corner_pattern() -> np.ndarray:
    """[3, 8] corner signs * 0.5 in the order of the nested loops x, y, z over (+1, -1) (utils/model_utils.py:270-277)."""
    xs, ys, zs = [], [], []
    for i in (1, -1):
        for j in (1, -1):
            for k in (1, -1):
                xs.append(i); ys.append(j); zs.append(k)
    return np.vstack([xs, ys, zs]) * 0.5


def _camera_points(x, cor):
    xc = cor[0] * x[2] * x[1] + cor[2] * x[4] * x[0] + x[5]
    yc = cor[1] * x[3] + x[6]
    zc = -cor[0] * x[2] * x[0] + cor[2] * x[4] * x[1] + x[7]
    return xc, yc, zc


def objective(x, K, UV, cor=None):
    """aimFun (:155-177).  K [3,3], UV [8,2]."""
    cor = corner_pattern() if cor is None else cor
    xc, yc, zc = _camera_points(x, cor)
    ex = xc * K[0, 0] / (zc + 1e-4) + K[0, 2] - UV[:, 0]
    ey = yc * K[1, 1] / (zc + 1e-4) + K[1, 2] - UV[:, 1]
    return float(np.sum(ex * ex) + np.sum(ey * ey))


def gradient(x, K, UV, cor=None):
    """jac (:206-234), constants as in the reference."""
    cor = corner_pattern() if cor is None else cor
    cost = 1e-6
    xc, yc, zc = _camera_points(x, cor)
    dex = (xc * K[0, 0] / (zc + cost) + K[0, 2] - UV[:, 0]) * 2
    dey = (yc * K[1, 1] / (zc + cost) + K[1, 2] - UV[:, 1]) * 2
    z8, o8 = np.zeros(8), np.ones(8)
    dx = np.stack([cor[2] * x[4], cor[0] * x[2], cor[0] * x[1], z8, cor[2] * x[0], o8, z8, z8])        # [8 params, 8 corners]
    dy = np.stack([z8, z8, z8, cor[1], z8, z8, o8, z8])
    dz = np.stack([-cor[0] * x[2], cor[2] * x[4], -cor[0] * x[0], z8, cor[2] * x[1], z8, z8, o8])
    gx = K[0, 0] * (dx * zc - dz * xc) / (zc ** 2 + cost)
    gy = K[1, 1] * (dy * zc - dz * yc) / (zc ** 2 + cost)
    return (gx * dex + gy * dey).sum(axis=1)


def fit_one(UV, K, dim_ref_cls, ref_loc):
    """One object of optim_decode_bbox3d (:289-296).  UV [8,2] input pixels, K [3,3], dim_ref_cls = (h, w, l)."""
    cor = corner_pattern()
    x0 = np.array([0, 1, dim_ref_cls[2], dim_ref_cls[0], dim_ref_cls[1]] + list(ref_loc), dtype=np.float64)
    UV = np.asarray(UV, dtype=np.float64)
    res = minimize(lambda x: objective(x, K, UV, cor), x0, method='L-BFGS-B', jac=lambda x: gradient(x, K, UV, cor), options=OPTIONS)
    return res.x, float(res.fun)


def canonical(x):
    """(Ry, l, h, w, X, Y, Z) in the gauge sin^2 + cos^2 = 1."""
    x = np.asarray(x, dtype=np.float64)
    if x[2] < 0:                      # the family member a < 0: (sin, cos, l, w) -> -(sin, cos, l, w) is the same box
        x = x * np.array([-1, -1, -1, 1, -1, 1, 1, 1.0])
    n = float(np.hypot(x[0], x[1]))
    return np.array([np.arctan2(x[0], x[1]), x[2] * n, x[3], x[4] * n, x[5], x[6], x[7]])


def reproject(x, K):
    """The 8 projected corners [8,2] of a solution (what the objective compares with UV)."""
    xc, yc, zc = _camera_points(x, corner_pattern())
    return np.stack([xc * K[0, 0] / (zc + 1e-4) + K[0, 2], yc * K[1, 1] / (zc + 1e-4) + K[1, 2]], axis=1)


def optim_decode_bbox3d(clses, bbox3d_projs, K, ref_dim, ref_loc):
    """optim_decode_bbox3d (:264-312) returning plain arrays: dict(cls, Ry, dimension (h,w,l), location, K, fun, x, accepted index)."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    keep, rys, dims, locs, funs, xs = [], [], [], [], [], []
    for i, (c, UV) in enumerate(zip(clses, bbox3d_projs)):
        x, fun = fit_one(UV, K, ref_dim[int(c)], ref_loc)
        if fun < ACCEPT_FUN:
            keep.append(i); rys.append(np.arctan2(x[0], x[1])); dims.append([x[3], x[4], x[2]]); locs.append([x[5], x[6], x[7]])
            funs.append(fun); xs.append(x)
    n = len(keep)
    return dict(index=np.asarray(keep, dtype=np.int64), cls=np.asarray([clses[i] for i in keep], dtype=np.int64), Ry=np.asarray(rys),
                dimension=np.asarray(dims).reshape(n, 3), location=np.asarray(locs).reshape(n, 3),
                K=np.tile(K.reshape(1, 9), (n, 1)), fun=np.asarray(funs), x=np.asarray(xs).reshape(n, 8))

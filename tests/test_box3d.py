"""Tier C (closed-form 3D recovery; NOT in the reference, parity unpinned for the decode formulas).

 * CPU: the projection sub-step of oracle/box3d_ref.py against the REFERENCE's own geometry functions
   (utils/model_utils.py:66-76, 80-119, 147-152): live import when /root/reference is mounted, committed golden vectors
   (tests/golden/box3d_proj_golden.npz, made by the generator at the bottom of this file) otherwise.
 * GPU: rtm3d_decode_box3d against oracle/box3d_ref.py, float tolerance 1e-4 relative (north_star).
"""
import os
import sys

import numpy as np
import pytest
import torch

import parity
from oracle import box3d_ref, ref_import

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "box3d_proj_golden.npz")
KITTI_K = np.array([721.54, 0, 609.56, 0, 721.54, 172.85, 0, 0, 1], np.float64)


def _proj_cases(seed=5, n=64):
    g = np.random.default_rng(seed)
    dim = g.uniform(0.5, 4.5, (n, 3))                       # (h, w, l)
    loc = np.stack([g.uniform(-20, 20, n), g.uniform(0.5, 2.5, n), g.uniform(4, 70, n)], -1)
    ry = g.uniform(-np.pi, np.pi, n)
    ry[:4] = [0.0, np.pi / 2, -np.pi / 2, 5e-4]             # exercise the |sin|,|cos| < 1e-3 snapping
    return dim, loc, ry


def _reference_projection(dim, loc, ry):
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import model_utils  # the reference's own module
    K = KITTI_K.reshape(3, 3)
    return np.stack([model_utils.calc_proj_corners(dim[i], loc[i], ry[i], K)[:8] for i in range(len(dim))])


def test_projection_substep_matches_reference_geometry():
    dim, loc, ry = _proj_cases()
    if ref_import.available():
        want = _reference_projection(dim, loc, ry)
        if os.path.exists(GOLD):
            assert np.allclose(np.load(GOLD)["corners"], want, rtol=1e-12, atol=1e-9), "golden file is stale"
    else:
        want = np.load(GOLD)["corners"]
    got = box3d_ref.project_corners(torch.tensor(dim, dtype=torch.float32), torch.tensor(loc, dtype=torch.float32),
                                    torch.tensor(ry, dtype=torch.float32), torch.tensor(KITTI_K, dtype=torch.float32)).numpy()
    parity.assert_close_rel(got, want, "projected corners", rel=2e-4, abs_=2e-2)   # fp32 restatement vs fp64 reference, pixels


@pytest.mark.gpu
@pytest.mark.parametrize("multibin,sig", [(False, False), (False, True), (True, False)])
def test_box3d_kernel_vs_restatement(multibin, sig):
    from rtm3d_b200 import HeatmapDecoder, synth
    dev = torch.device("cuda:0")
    B, C, H, W, K = 3, 3, 48, 80, 40
    logits, _ = synth.head_outputs(B, C, H, W, seed=61, kind="randn")
    logits = [t.to(dev) for t in logits]
    gen = torch.Generator().manual_seed(9)
    Creg = 14 if multibin else 8
    reg = (torch.randn((B, Creg, H, W), generator=gen) * 0.5).to(dev)
    cam = torch.tensor(KITTI_K, dtype=torch.float32)
    cam[:6] /= 4.0
    cams = cam.repeat(B, 1).to(dev)
    dim_ref = torch.tensor([[1.53, 1.63, 3.88], [1.76, 0.66, 0.84], [1.74, 0.60, 1.76]], dtype=torch.float32, device=dev)
    dec = HeatmapDecoder(0.4, K, 4.0)
    det = dec.decode_packed(logits)
    out = dec.decode_box3d(det, reg, cams, dim_ref, C, multibin=multibin, sigmoid_subpixel=sig)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(det.counts[b])
        assert n > 0
        want = box3d_ref.decode_box3d(det.flat[b, :n].long(), reg[b], cams[b], dim_ref, C, multibin=multibin, sigmoid_subpixel=sig)
        for f in ("loc", "dim", "alpha", "rot_y"):
            parity.assert_close_rel(out[f][b, :n].cpu().numpy(), want[f].cpu().numpy(), f"{f} image {b}", rel=1e-4, abs_=1e-4)
        # projected vertices: points close to the camera plane amplify rounding; compare where the depth is sane
        ok = (want["loc"][:, 2] > 1.0).cpu().numpy()
        parity.assert_close_rel(out["corners2d"][b, :n].cpu().numpy()[ok], want["corners2d"].cpu().numpy()[ok],
                                f"corners2d image {b}", rel=1e-4, abs_=5e-3)
        for f in ("loc", "dim", "corners2d"):
            assert torch.all(out[f][b, n:] == 0)


if __name__ == "__main__":   # generator of tests/golden/box3d_proj_golden.npz (run where /root/reference is mounted)
    d, l, r = _proj_cases()
    np.savez_compressed(GOLD, corners=_reference_projection(d, l, r))
    print("wrote", GOLD)

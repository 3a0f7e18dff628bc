"""3D-box fit behind the decoder: utils/model_utils.py:264-312 (optim_decode_bbox3d), objective :155-177,
gradient :206-234, called at detect.py:71-74.

 * CPU: oracle/boxfit_ref.py (the restatement) against tests/golden/boxfit_golden.npz = outputs of the REAL reference
   (oracle/make_boxfit_golden.py) and, when /root/reference is mounted, against the imported reference itself.
 * GPU: rtm3d_fit_box3d (through the C ABI) against the same goldens on what the objective determines: res.fun, the accept
   decision (:298), Ry, the reprojected corners, and the shape (l, h, w, X, Y, Z) up to the two-parameter family the
   objective leaves open (rotation-vector length, common scale) -- see oracle/boxfit_ref.py.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import boxfit_ref as bf
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "boxfit_golden.npz")


def _gold():
    return np.load(GOLD)


def _shape_invariants(x):
    """(Ry, l, h, w, X, Y) / Z-normalised: independent of the rotation-vector length and of the common scale."""
    c = bf.canonical(x)
    return np.concatenate([[c[0]], c[1:6] / c[6]])


def test_oracle_restatement_matches_reference_goldens():
    g = _gold()
    K = g["K"].astype(np.float64).reshape(3, 3)
    for i in range(len(g["cls"])):
        x, fun = bf.fit_one(g["uv"][i], K, g["dim_ref"][g["cls"][i]], list(g["ref_loc"]))
        assert abs(fun - g["fun"][i]) <= 1e-8 * max(1.0, abs(g["fun"][i])), f"object {i}: fun {fun} vs reference {g['fun'][i]}"
        if g["fun"][i] < 10:      # (garbage objects: any local minimum)
            np.testing.assert_allclose(x, g["x"][i], rtol=1e-6, atol=1e-8, err_msg=f"object {i}")
    out = bf.optim_decode_bbox3d(g["cls"], g["uv"], g["K"], g["dim_ref"], list(g["ref_loc"]))
    assert np.array_equal(out["cls"], g["out_class"])
    np.testing.assert_allclose(out["Ry"], g["out_Ry"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(out["dimension"], g["out_dimension"], rtol=1e-6)
    np.testing.assert_allclose(out["location"], g["out_location"], rtol=1e-6, atol=1e-8)


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
def test_oracle_restatement_matches_live_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import model_utils
    g = _gold()
    K = g["K"].astype(np.float64).reshape(3, 3)
    cor = bf.corner_pattern()
    rng = np.random.default_rng(3)
    for i in range(6):
        x = np.array([0.3, 0.9, 3.5, 1.5, 1.6, 1.0, 1.2, 25.0]) + rng.normal(0, 0.2, 8)
        UV = g["uv"][i].astype(np.float64)
        assert np.isclose(bf.objective(x, K, UV, cor), model_utils.aimFun(cor, K, UV.T)(x), rtol=1e-12)
        np.testing.assert_allclose(bf.gradient(x, K, UV, cor), model_utils.jac(cor, K, UV.T)(x), rtol=1e-10, atol=1e-12)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = model_utils.optim_decode_bbox3d(g["cls"][:6], g["uv"][:6], g["K"], g["dim_ref"].tolist(), list(g["ref_loc"]))
    mine = bf.optim_decode_bbox3d(g["cls"][:6], g["uv"][:6], g["K"], g["dim_ref"], list(g["ref_loc"]))
    assert list(ref.get_field("class")) == mine["cls"].tolist()
    np.testing.assert_allclose(mine["dimension"], ref.get_field("dimension"), rtol=1e-6)
    np.testing.assert_allclose(mine["location"], ref.get_field("location"), rtol=1e-6, atol=1e-8)


@pytest.mark.gpu
def test_gpu_fit_matches_reference_goldens():
    from rtm3d_b200 import fit_packed
    g = _gold()
    dev = torch.device("cuda:0")
    n = len(g["cls"])
    K = g["K"].astype(np.float64).reshape(3, 3)
    fit = fit_packed(torch.as_tensor(g["uv"], device=dev).reshape(1, n, 8, 2), torch.as_tensor(g["cls"], device=dev).reshape(1, n), None,
                     torch.as_tensor(g["K"], device=dev), g["dim_ref"], list(g["ref_loc"]), want_solution=True)
    torch.cuda.synchronize()
    fun, acc, x8 = fit.fun[0].cpu().numpy().astype(np.float64), fit.accept[0].cpu().numpy(), fit.x8[0].cpu().numpy()
    checked, stalled = 0, []
    for i in range(n):
        fr = g["fun"][i]
        if fr > 1e4:                                   # garbage objects: only the decision is comparable
            assert acc[i] == 0
            continue
        # the reference stops at ftol 2.2e-9 / gtol 1e-5: this minimum is at most that much lower, never meaningfully higher
        assert fun[i] <= fr * (1 + 1e-4) + 1e-7, f"object {i}: fun {fun[i]} above the reference's {fr}"
        if fun[i] < fr * (1 - 2e-2) - 1e-5:
            # the reference's L-BFGS-B stalled in the local minimum of a wrong yaw (it starts at yaw 0 only); the kernel's four yaw
            # starts find the lower one.  Known for the goldens: few, and only where the reference REJECTS the object.
            stalled.append(i)
            assert fr >= bf.ACCEPT_FUN, f"object {i}: the reference accepted at fun {fr}, the kernel found {fun[i]}"
            continue
        if abs(fr - bf.ACCEPT_FUN) > 1e-3:
            assert acc[i] == int(fr < bf.ACCEPT_FUN), f"object {i}: accept {acc[i]} vs reference fun {fr}"
        # reprojected corners (pixels) and the shape the objective determines
        np.testing.assert_allclose(bf.reproject(x8[i], K), bf.reproject(g["x"][i], K), atol=2e-2, err_msg=f"object {i} reprojection")
        a, b = _shape_invariants(x8[i]), _shape_invariants(g["x"][i])
        d_ry = np.abs((a[0] - b[0] + np.pi) % (2 * np.pi) - np.pi)
        assert d_ry < 2e-3, f"object {i}: Ry {a[0]} vs {b[0]}"
        np.testing.assert_allclose(a[1:], b[1:], rtol=5e-3, atol=5e-4, err_msg=f"object {i} shape")
        checked += 1
    assert checked >= 40 and len(stalled) <= 4, (checked, stalled)
    # reported members of the solution family: unit rotation vector, dimensions scaled towards the class prior
    dim, loc, ry = fit.dim[0].cpu().numpy(), fit.loc[0].cpu().numpy(), fit.ry[0].cpu().numpy()
    for i in range(n):
        if g["fun"][i] > 1e4:
            continue
        c = bf.canonical(x8[i])
        assert np.isclose(ry[i], c[0], atol=1e-5)
        np.testing.assert_allclose(dim[i] / loc[i, 2], np.array([c[2], c[3], c[1]]) / c[6], rtol=1e-4)     # (h, w, l) / Z
        np.testing.assert_allclose(loc[i, :2] / loc[i, 2], c[4:6] / c[6], rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_gpu_fit_drop_in_signature_and_padding():
    """optim_decode_bbox3d(clses, bbox3d_projs, K, ref_dim, ref_loc) as detect.py:71-74 calls it; packed form with counts."""
    from rtm3d_b200 import fit_packed, optim_decode_bbox3d
    g = _gold()
    out = optim_decode_bbox3d(g["cls"], g["uv"], g["K"], g["dim_ref"].tolist(), list(g["ref_loc"]))
    got = out.get_field("class")
    assert len(out.get_field("Ry")) == len(got) == out.get_field("dimension").shape[0] == out.get_field("location").shape[0] == out.get_field("K").shape[0]
    # every object the reference accepts (utils/model_utils.py:298) is accepted, in the same order; the few extra ones are the
    # objects where the reference's L-BFGS-B stalled in a wrong-yaw minimum (test_gpu_fit_matches_reference_goldens)
    ref_keep = [int(c) for c, f in zip(g["cls"], g["fun"]) if f < bf.ACCEPT_FUN - 1e-3]
    assert len(ref_keep) <= len(got) <= len(ref_keep) + 5
    it = iter(got)
    assert all(any(c == d for d in it) for c in ref_keep), "the reference's accepted objects are not a subsequence of the kernel's"
    assert np.all(out.get_field("fun") < bf.ACCEPT_FUN) and np.all(out.get_field("dimension") > 0)
    empty = optim_decode_bbox3d(np.zeros((0,), dtype=np.int64), np.zeros((0, 8, 2), dtype=np.float32), g["K"], g["dim_ref"].tolist(), list(g["ref_loc"]))
    assert empty.get_field("dimension").shape == (0, 3) and empty.get_field("class") == []
    dev = torch.device("cuda:0")
    n = 12
    verts = torch.as_tensor(g["uv"][:2 * n], device=dev).reshape(2, n, 8, 2)
    cls = torch.as_tensor(g["cls"][:2 * n], device=dev).reshape(2, n)
    counts = torch.tensor([n, 5], dtype=torch.int32, device=dev)
    fit = fit_packed(verts, cls, counts, torch.as_tensor(g["K"], device=dev).repeat(2, 1), g["dim_ref"], list(g["ref_loc"]))
    torch.cuda.synchronize()
    assert torch.all(fit.accept[1, 5:] == 0) and torch.all(fit.dim[1, 5:] == 0) and torch.all(fit.fun[1, 5:] == 0)
    full = fit_packed(verts, cls, None, torch.as_tensor(g["K"], device=dev), g["dim_ref"], list(g["ref_loc"]))
    torch.cuda.synchronize()
    assert torch.equal(full.fun[0], fit.fun[0]) and torch.equal(full.fun[1, :5], fit.fun[1, :5])

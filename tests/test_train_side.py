"""Training-side mirror: main heat-map target encoder (datasets/dataset_reader.py:215-291,
utils/data_utils.py:97-141) and focal loss (models/nets/module.py:41-68 on utils/model_utils.py:10-14, models/rtm3d_loss.py:283).
CPU: the oracle restatement against goldens made with the reference's own helpers (and against those helpers, live, when
/root/reference is mounted).  GPU: the CUDA kernels through the C ABI against the same goldens."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_import, train_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz")


def test_oracle_targets_and_loss_match_reference_goldens():
    g = np.load(GOLD)
    B, C, H, W = g["shape"]
    m_hm, m_proj, m_off, sigma, radius = train_ref.main_targets(g["bbox"], g["cls"], g["img_id"], g["mask"], g["noise_mask"], B, C, H, W)
    assert np.array_equal(m_hm.astype(np.float32), g["m_hm"]) and np.array_equal(m_proj, g["m_proj"])
    np.testing.assert_allclose(m_off, g["m_off"], rtol=0, atol=0)
    np.testing.assert_allclose(sigma, g["sigma"], rtol=1e-15)
    assert np.array_equal(radius, g["radius"])
    lg = torch.from_numpy(g["logits"]).clone().requires_grad_(True)
    loss = train_ref.focal_loss(lg, torch.from_numpy(g["m_hm"]))
    loss.backward()
    assert np.isclose(float(loss), float(g["loss"]), rtol=1e-6)
    np.testing.assert_allclose(lg.grad.numpy(), g["grad"], rtol=1e-5, atol=1e-9)
    assert np.isclose(float(train_ref.focal_loss(torch.from_numpy(g["logits"]), torch.zeros(tuple(g["shape"])))), float(g["empty_loss"]), rtol=1e-6)


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
def test_oracle_helpers_match_live_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_import.REF_ROOT)
    from utils import data_utils, model_utils
    from models.nets.module import FocalLoss
    g = np.load(GOLD)
    bb = g["bbox"].astype(np.float64)
    np.testing.assert_allclose(train_ref.gaussian_radius(bb), data_utils._compute_gaussian_radius(bb), rtol=1e-15)
    s, r = data_utils.dynamic_radius(bb)
    for i in (0, 7, 19):
        a, b = train_ref.gaussian2d(s[i], r[i]), data_utils.gaussian2D(s[i], r[i])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    lg = torch.from_numpy(g["logits"])
    t = torch.from_numpy(g["m_hm"])
    assert torch.allclose(train_ref.focal_loss(lg, t), FocalLoss(2.0, 4.0)(model_utils.sigmoid_hm(lg.clone()), t), rtol=1e-6)


@pytest.mark.gpu
def test_gpu_target_encoder_matches_goldens():
    from rtm3d_b200.train_side import build_main_targets
    g = np.load(GOLD)
    B, C, H, W = [int(v) for v in g["shape"]]
    dev = torch.device("cuda:0")
    t = build_main_targets(torch.as_tensor(g["bbox"], device=dev), torch.as_tensor(g["cls"], device=dev), torch.as_tensor(g["img_id"], device=dev),
                           torch.as_tensor(g["mask"], device=dev), torch.as_tensor(g["noise_mask"], device=dev), B, C, H, W)
    torch.cuda.synchronize()
    assert np.array_equal(t.m_proj.cpu().numpy(), g["m_proj"]) and np.array_equal(t.radius.cpu().numpy(), g["radius"].astype(np.int32))
    np.testing.assert_allclose(t.m_off.cpu().numpy(), g["m_off"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(t.sigma.cpu().numpy(), g["sigma"], rtol=1e-6)
    got, want = t.m_hm.cpu().numpy(), g["m_hm"]
    assert np.array_equal(got == 0, want == 0), "support of the splats differs"
    assert np.array_equal(got == 1, want == 1), "positive positions differ"
    assert np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max() <= 1, "more than 1 ulp from the fp64 numpy result"
    empty = build_main_targets(torch.zeros((0, 4), device=dev), torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
                               torch.zeros(0, dtype=torch.uint8, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev), 1, C, H, W)
    assert float(empty.m_hm.abs().sum()) == 0.0


@pytest.mark.gpu
def test_gpu_focal_loss_and_gradient_match_goldens():
    from rtm3d_b200.train_side import FocalLoss
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    logits = torch.as_tensor(g["logits"], device=dev).clone().requires_grad_(True)
    before = logits.detach().clone()
    target = torch.as_tensor(g["m_hm"], device=dev)
    loss = FocalLoss(2.0, 4.0)(logits, target)
    (3.0 * loss).backward()
    torch.cuda.synchronize()
    assert torch.equal(before, logits.detach()), "the loss modified the logits (the reference's sigmoid_hm works in place on them)"
    assert np.isclose(float(loss.detach()), float(g["loss"]), rtol=2e-5)
    np.testing.assert_allclose(logits.grad.cpu().numpy(), 3.0 * g["grad"], rtol=2e-4, atol=1e-8)
    assert np.isclose(float(FocalLoss(2.0, 4.0)(logits.detach(), torch.zeros_like(target))), float(g["empty_loss"]), rtol=2e-5)
    # against torch's own autograd on the same device
    lg2 = torch.as_tensor(g["logits"], device=dev).clone().requires_grad_(True)
    ref = train_ref.focal_loss(lg2, target)
    ref.backward()
    assert np.isclose(float(loss.detach()), float(ref.detach()), rtol=2e-5)
    np.testing.assert_allclose(logits.grad.cpu().numpy() / 3.0, lg2.grad.cpu().numpy(), rtol=2e-4, atol=1e-8)


# ---------------------------------------------------------------------------------------------------------------
# the whole loss: models/rtm3d_loss.py:268-340
LOSS_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.npz")


def test_oracle_loss_matches_reference_goldens():
    g = np.load(LOSS_GOLD)
    pred, fields = train_ref.make_loss_case()
    leaves = [p.clone().requires_grad_(True) for p in pred]
    loss, parts = train_ref.rtm3d_loss(leaves, fields, train_ref.LOSS_WEIGHTS)
    loss.backward()
    np.testing.assert_allclose(np.array([float(x) for x in parts]), g["parts"], rtol=1e-6)
    for i, leaf in enumerate(leaves):
        np.testing.assert_allclose(leaf.grad.numpy(), g[f"grad{i}"], rtol=1e-5, atol=1e-9)
    assert all(torch.equal(a, b.detach()) for a, b in zip(pred, leaves)), "the restatement must not modify the logits"


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
def test_loss_goldens_are_what_the_live_reference_yields():
    from oracle import make_train_golden
    g = np.load(LOSS_GOLD)
    pred, fields = train_ref.make_loss_case()
    loss, parts, grads = make_train_golden.reference_loss(pred, fields, train_ref.LOSS_WEIGHTS)
    np.testing.assert_allclose(parts.numpy(), g["parts"], rtol=1e-6)
    for i, gr in enumerate(grads):
        np.testing.assert_allclose(gr.numpy(), g[f"grad{i}"], rtol=1e-6, atol=1e-10)


@pytest.mark.gpu
def test_gpu_rtm3d_loss_matches_reference_goldens():
    import types
    from rtm3d_b200 import RTM3DLoss
    g = np.load(LOSS_GOLD)
    dev = torch.device("cuda:0")
    pred, fields = train_ref.make_loss_case()
    leaves = [p.to(dev).clone().requires_grad_(True) for p in pred]
    before = [p.detach().clone() for p in leaves]
    ns = types.SimpleNamespace
    W = train_ref.LOSS_WEIGHTS
    cfg = ns(MODEL=ns(FOCAL_LOSS_ALPHA=2.0, FOCAL_LOSS_BEDA=4.0), TRAINING=ns(W_MKF=W[0], W_VFM=W[1], W_M_OFF=W[2], W_V_OFF=W[3]))
    targets = types.SimpleNamespace(get_field=lambda k: fields[k])
    loss, parts = RTM3DLoss(cfg)(leaves, targets)
    loss.backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(parts.cpu().numpy(), g["parts"], rtol=2e-5)
    for i, leaf in enumerate(leaves):
        assert torch.equal(before[i], leaf.detach()), "the loss modified its inputs"
        np.testing.assert_allclose(leaf.grad.cpu().numpy(), g[f"grad{i}"], rtol=2e-4, atol=2e-8, err_msg=f"gradient of map {i}")


@pytest.mark.gpu
def test_gpu_gather_l1_edge_cases():
    from rtm3d_b200 import gather_l1_loss
    dev = torch.device("cuda:0")
    fmap = torch.randn(2, 4, 5, 6, device=dev, requires_grad=True)
    z = lambda *s, dt=torch.int64: torch.zeros(s, dtype=dt, device=dev)
    # nothing valid: mean over an empty selection is NaN (F.l1_loss), and the gradient call must not fault
    loss = gather_l1_loss(fmap, z(3), z(3), z(3), torch.zeros(3, dtype=torch.bool, device=dev), torch.zeros(3, 2, device=dev))
    assert torch.isnan(loss)
    # out-of-map entries are skipped; the others count
    img = torch.tensor([0, 1, 5], device=dev); x = torch.tensor([1, 7, 0], device=dev); y = torch.tensor([2, 0, 0], device=dev)
    tgt = torch.tensor([[0.5, -0.5], [0.0, 0.0], [0.0, 0.0]], device=dev)
    loss = gather_l1_loss(fmap, img, x, y, torch.ones(3, dtype=torch.bool, device=dev), tgt, c0=torch.tensor([2, 0, 0], device=dev, dtype=torch.int32))
    want = ((fmap[0, 2, 2, 1] - 0.5).abs() + (fmap[0, 3, 2, 1] + 0.5).abs()) / 2
    loss.backward()
    assert torch.allclose(loss, want.detach(), rtol=1e-6)
    assert int((fmap.grad != 0).sum()) == 2 and float(fmap.grad.abs().sum()) == pytest.approx(1.0, rel=1e-6)

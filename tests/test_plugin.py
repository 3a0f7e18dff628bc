"""The drop-in itself: ``rtm3d_b200.install(Model)`` rebinds ``Model.inference`` / ``Model.forward`` of the reference
(models/model.py:20-27, :29-75).  The stub below has the reference Model's attributes (backbone, kfpn_fusion,
detect_header, config with the three scalars the decoder reads) and the reference's forward/inference dataflow; the real
class is patched too when /root/reference is mounted (this container only)."""
import os
import sys
import types

import pytest
import torch

import rtm3d_b200
from rtm3d_b200 import plugin, synth

NS = types.SimpleNamespace


def _config(thresh=0.4, topk=30, down=4.0):
    return NS(DETECTOR=NS(SCORE_THRESH=thresh, TOPK_CANDIDATES=topk), MODEL=NS(DOWN_SAMPLE=down))


class _Head(torch.nn.Module):
    """Stands in for RTM3DHeader (models/nets/header.py:40-46): returns the four NCHW maps it was given."""

    def __init__(self, maps):
        super().__init__()
        self.maps = maps

    def forward(self, _):
        return list(self.maps)


def _stub_model_class():
    class Model(torch.nn.Module):        # same attribute names and eval dataflow as models/model.py:9-27
        def __init__(self, maps, config):
            super().__init__()
            self.config = config
            self.backbone = torch.nn.Identity()
            self.kfpn_fusion = torch.nn.Identity()
            self.detect_header = _Head(maps)

        def forward(self, x):
            pred_logits = self.detect_header(self.kfpn_fusion(self.backbone(x)))
            if self.training:
                return pred_logits
            return self.inference([p.clone() for p in pred_logits]), pred_logits

        def inference(self, pred_logits):
            raise AssertionError("the original inference must not run once the plugin is installed")
    return Model


def test_install_rebinds_and_uninstall_restores():
    Model = _stub_model_class()
    orig_inf, orig_fwd = Model.inference, Model.forward
    rtm3d_b200.install(Model)
    try:
        assert Model.inference is plugin._inference and Model.forward is plugin._forward
        rtm3d_b200.install(Model)                       # idempotent
        assert plugin._ORIG[Model] == (orig_inf, orig_fwd)
        logits, _ = synth.head_outputs(1, 3, 16, 16, seed=1)
        m = Model(logits, _config())
        m.train()
        out = m(torch.zeros(1))
        assert all(a is b for a, b in zip(out, logits))           # training branch: pred_logits handed through (model.py:24-25)
        m.eval()
        with pytest.raises(ValueError, match="no CPU path"):       # CPU maps: the plugin has no fallback to the reference code
            m(torch.zeros(1))
    finally:
        rtm3d_b200.uninstall(Model)
    assert Model.inference is orig_inf and Model.forward is orig_fwd


@pytest.mark.skipif(not os.path.exists("/root/reference/models/model.py"), reason="reference checkout not mounted")
def test_install_on_the_real_reference_class():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        from models.model import Model
    finally:
        sys.path.remove("/root/reference")
    orig = (Model.inference, Model.forward)
    rtm3d_b200.install(Model)
    try:
        assert Model.inference is plugin._inference and Model.forward is plugin._forward
        m = Model.__new__(Model)
        torch.nn.Module.__init__(m)
        logits, _ = synth.head_outputs(1, 3, 16, 16, seed=2)
        m.config = _config()
        m.backbone = m.kfpn_fusion = torch.nn.Identity()
        m.detect_header = _Head(logits)
        m.eval()
        with pytest.raises(ValueError, match="no CPU path"):
            m(torch.zeros(1))
    finally:
        rtm3d_b200.uninstall(Model)
    assert (Model.inference, Model.forward) == orig


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["randn", "few"])
def test_installed_forward_matches_reference_dataflow(kind):
    """forward(x) in eval mode -> ((clses, m_scores, m_projs, v_projs_regress, bboxes_2d), pred_logits) with pred_logits
    bit-unchanged (the loss of train.py:71 reads them) and the five lists equal to the oracle's."""
    from oracle import decode_ref
    dev = torch.device("cuda:0")
    B, K = 3, 30
    logits_cpu, _ = synth.head_outputs(B, 3, 48, 80, seed=5, kind=kind)
    if kind == "few":
        logits_cpu[0][1].fill_(-20.0)                   # image 1: nothing above the threshold -> None in every list
    logits = [t.to(dev) for t in logits_cpu]
    before = [t.clone() for t in logits]
    Model = _stub_model_class()
    rtm3d_b200.install(Model)
    try:
        m = Model(logits, _config(0.4, K, 4.0)).eval()
        with torch.no_grad():
            decoded, pred_logits = m(torch.zeros(1, device=dev))
        assert len(decoded) == 5 and all(len(l) == B for l in decoded)
        assert all(a is b for a, b in zip(pred_logits, logits)), "pred_logits must be the head's own tensors (no clone)"
        for a, b in zip(before, logits):
            assert torch.equal(a, b), "the plugin modified pred_logits"
        want = decode_ref.decode(logits, 0.4, K, 4.0)
        for got_l, want_l in zip(decoded, want):
            for g, w in zip(got_l, want_l):
                assert (g is None) == (w is None)
                if g is not None:
                    assert g.dtype == w.dtype and g.shape == w.shape and g.device == w.device and torch.equal(g, w)
        if kind == "few":
            assert decoded[0][1] is None
        # a changed config scalar is picked up (the decoder is rebuilt when the three scalars change)
        m.config.DETECTOR.TOPK_CANDIDATES = 10
        decoded2, _ = m(torch.zeros(1, device=dev))
        want2 = decode_ref.decode(logits, 0.4, 10, 4.0)
        for g, w in zip(decoded2[1], want2[1]):
            assert (g is None) == (w is None) and (g is None or torch.equal(g, w))
    finally:
        rtm3d_b200.uninstall(Model)
    assert Model.inference is not plugin._inference

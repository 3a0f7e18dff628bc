"""Drop-in installation into the reference (hitfeelee/rtm3d) without editing its files.

``install(Model)`` rebinds two methods of ``models.model.Model``:

* ``inference(self, pred_logits)`` (models/model.py:29-75) -> ``HeatmapDecoder.decode`` with the three scalars read
  from ``self.config`` exactly where the reference reads them (:41-42, :67, :70);
* ``forward(self, x)`` (models/model.py:20-27): same dataflow, minus the ``[p.clone() for p in pred_logits]`` of :27 --
  the clone exists only because the reference decoder mutates its inputs in place (``sigmoid_`` at :48, :85); this
  decoder never writes to them, so ``pred_logits`` is still returned intact for the loss (train.py:71).

detect.py and the eval loops of train.py / train_multi_gpu.py run unchanged.  ``uninstall(Model)`` restores the originals.
"""
from __future__ import annotations

from .decoder import HeatmapDecoder

_ORIG = {}


def _decoder_of(model) -> HeatmapDecoder:
    cfg = model.config
    key = (float(cfg.DETECTOR.SCORE_THRESH), int(cfg.DETECTOR.TOPK_CANDIDATES), float(cfg.MODEL.DOWN_SAMPLE))
    dec = model.__dict__.get("_rtm3d_b200_decoder")
    if dec is None or dec[0] != key:
        dec = (key, HeatmapDecoder(*key))
        model.__dict__["_rtm3d_b200_decoder"] = dec
    return dec[1]


def _inference(self, pred_logits):
    return _decoder_of(self).decode(pred_logits)


def _forward(self, x):
    pred_logits = self.detect_header(self.kfpn_fusion(self.backbone(x)))
    if self.training:
        return pred_logits
    return self.inference(pred_logits), pred_logits


def install(model_cls) -> None:
    """Patch the reference ``Model`` class (pass ``models.model.Model``)."""
    if model_cls in _ORIG:
        return
    _ORIG[model_cls] = (model_cls.inference, model_cls.forward)
    model_cls.inference = _inference
    model_cls.forward = _forward


def install_boxfit(model_utils_module, device="cuda:0") -> None:
    """Rebind ``utils.model_utils.optim_decode_bbox3d`` (utils/model_utils.py:264-312, called at detect.py:71-74) to the batched
    GPU fit: same arguments, and a result whose ``get_field`` serves 'class', 'Ry', 'dimension', 'location', 'K' like the
    reference's ParamList."""
    from .boxfit import optim_decode_bbox3d
    if model_utils_module in _ORIG:
        return
    _ORIG[model_utils_module] = (model_utils_module.optim_decode_bbox3d,)
    model_utils_module.optim_decode_bbox3d = lambda clses, projs, K, ref_dim, ref_loc: optim_decode_bbox3d(clses, projs, K, ref_dim, ref_loc, device=device)


def uninstall_boxfit(model_utils_module) -> None:
    orig = _ORIG.pop(model_utils_module, None)
    if orig:
        model_utils_module.optim_decode_bbox3d = orig[0]


def uninstall(model_cls) -> None:
    orig = _ORIG.pop(model_cls, None)
    if orig:
        model_cls.inference, model_cls.forward = orig

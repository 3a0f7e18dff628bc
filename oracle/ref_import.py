"""Loader of the REAL reference decoder from /root/reference.  TEST INFRASTRUCTURE ONLY.

Usable only where the reference checkout is mounted (the build container).  It does not exist on the GPU box,
so nothing that runs there may depend on this module: tests that use it skip when ``available()`` is False and
``oracle/make_golden.py`` (run here) turns its outputs into committed fixtures under ``tests/golden/``.

The reference ``Model`` is instantiated without backbone/neck/heads (``Model.__new__`` + ``nn.Module.__init__``):
the decoder methods only read ``self.config`` (models/model.py:41-42,67,70).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("RTM3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "models", "model.py"))


def load_model_class():
    if not available():
        raise RuntimeError(f"reference checkout not mounted at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models.model import Model  # noqa: E402  (the reference's own module)
    return Model


def make_reference_decoder(thresh: float, topk: int, down: float):
    """Instance of the reference ``Model`` with only the three config scalars the decoder reads."""
    import torch
    Model = load_model_class()
    ns = types.SimpleNamespace
    m = Model.__new__(Model)
    torch.nn.Module.__init__(m)
    m.config = ns(DETECTOR=ns(SCORE_THRESH=thresh, TOPK_CANDIDATES=topk), MODEL=ns(DOWN_SAMPLE=down))
    return m


def reference_inference(pred_logits, thresh: float, topk: int, down: float):
    """Run the unmodified ``Model.inference`` on clones (as ``Model.forward`` does, models/model.py:27)."""
    import torch
    m = make_reference_decoder(thresh, topk, down)
    with torch.no_grad():
        return m.inference([p.clone() for p in pred_logits])


def reference_keypoint_branch(pred_logits, kpt_logits, thresh: float, topk: int, down: float):
    """Tier B through the reference's own dormant methods, wired as the commented lines models/model.py:45-62,68-69
    describe.  Returns per image None or a dict (cls, score, proj, verts, bbox, kpt_score, kpt_proj)."""
    import torch
    m = make_reference_decoder(thresh, topk, down)
    main, off16, off2, voff2 = [p.clone() for p in pred_logits]
    kpt = kpt_logits.clone()
    out = []
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(main.shape[0]):
            cls, score, mp = m._obtain_main_proj2d(main[i], thresh, topk)
            if len(cls) == 0:
                out.append(None)
                continue
            vs, vp = m._obtain_vertex_proj2d(kpt[i], topk)
            off = m._obtain_offset_fr_main(off16[i], mp)
            sub = off2[i][:, mp[1].long(), mp[0].long()].sigmoid_()
            mp[0] += sub[0]
            mp[1] += sub[1]
            vp[0] = vp[0].view(-1)
            vp[1] = vp[1].view(-1)
            vsub = voff2[i][:, vp[1].long(), vp[0].long()].sigmoid_()
            vp[0] += vsub[0]
            vp[1] += vsub[1]
            vp[0] = vp[0].view(-1, topk)
            vp[1] = vp[1].view(-1, topk)
            Cv = vs.shape[0]
            if off.shape[0] < Cv:
                off = torch.cat([off, off.new_zeros(Cv - off.shape[0], off.shape[1], 2)], dim=0)
            kp, reg, ks = m._group_vertexs_kf(mp, vp, vs, off)
            mm = torch.cat([mp[0].unsqueeze(-1), mp[1].unsqueeze(-1)], dim=-1)
            verts = down * reg
            out.append(dict(cls=cls, score=score, proj=down * mm, verts=verts,
                            bbox=torch.cat([verts.min(dim=1)[0], verts.max(dim=1)[0]], dim=-1),
                            kpt_score=ks, kpt_proj=down * kp))
    return out

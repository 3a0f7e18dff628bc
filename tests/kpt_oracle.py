"""Tier B comparison of one image against the oracle on the same device (shared by the small-shape and the full-size
keypoint tests).  models/model.py:100-115 (_obtain_vertex_proj2d), :52-60 (commented sub-pixel wiring), :134-162
(_group_vertexs_kf)."""
import numpy as np
import torch

from oracle import decode_ref


def check_image(det, cand, grp, logits, kpt, b, K, thresh=0.4, down=4.0, what="", candidates=True):
    """Bit-exact check of image ``b``: candidates (score, flat, sub-pixel xy) and the grouping (kpt_j, kpt_score, kpt_proj,
    verts).  ``logits`` = the four head maps, ``kpt`` = the keypoint heat-map, all on the GPU."""
    Cv = kpt.shape[1]
    vs, vx, vy, vflat = decode_ref.keypoint_peaks(kpt[b], K)
    # canonical order inside exact-score tie groups: torch.topk leaves that order implementation-defined (batched CUDA topk
    # on short rows does not return ties index-ascending), and argmin's "first minimal index" (models/model.py:151) depends
    # on it when two candidates are equidistant
    orders = [np.lexsort((vflat[c].cpu().numpy(), -vs[c].cpu().numpy().astype(np.float64))) for c in range(Cv)]
    if candidates:
        for c in range(Cv):
            o = orders[c]
            assert np.array_equal(cand.flat[b, c].cpu().numpy(), vflat[c].cpu().numpy()[o]), f"{what} b{b} c{c} kflat"
            assert np.array_equal(cand.score[b, c].cpu().numpy().view(np.uint32),
                                  vs[c].cpu().numpy()[o].view(np.uint32)), f"{what} b{b} c{c} kscore"
    cls, score, xf, yf, flat = decode_ref.main_peaks(logits[0][b], thresh, K)
    n = int(det.counts[b])
    assert n == len(cls), f"{what} b{b}: count {n} != {len(cls)}"
    if n == 0:
        return 0
    assert torch.equal(det.flat[b, :n].long(), flat), f"{what} b{b} flat"
    order = torch.from_numpy(np.stack(orders)).to(kpt.device)
    vs_c, vx_c, vy_c = vs.gather(1, order), vx.gather(1, order), vy.gather(1, order)
    vsub = torch.sigmoid(logits[3][b][:, vy_c.reshape(-1).long(), vx_c.reshape(-1).long()])
    vx_c = (vx_c.reshape(-1) + vsub[0]).view(Cv, K)
    vy_c = (vy_c.reshape(-1) + vsub[1]).view(Cv, K)
    assert torch.equal(cand.xy[b, ..., 0], vx_c) and torch.equal(cand.xy[b, ..., 1], vy_c), f"{what} b{b} kxy"
    off = decode_ref.vertex_offsets(logits[1][b], xf, yf)
    sub = torch.sigmoid(logits[2][b][:, yf.long(), xf.long()])
    mx, my = xf + sub[0], yf + sub[1]
    if off.shape[0] < Cv:
        off = torch.cat([off, off.new_zeros(Cv - off.shape[0], off.shape[1], 2)], dim=0)
    kp, reg, ks, j = decode_ref.group_keypoints(mx, my, vx_c, vy_c, vs_c, off)
    assert torch.equal(grp.kpt_j[b, :n].long(), j.t()), f"{what} b{b} kpt_j"
    assert torch.equal(grp.kpt_score[b, :n], ks), f"{what} b{b} kpt_score"
    assert torch.equal(grp.kpt_proj[b, :n], down * kp), f"{what} b{b} kpt_proj"
    assert torch.equal(grp.verts[b, :n], down * reg), f"{what} b{b} verts"
    assert torch.all(grp.kpt_j[b, n:] == -1)
    return n

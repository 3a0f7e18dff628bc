"""Parity at BASELINE.json's full sizes (cfg4: 256 x (3+9) x 96 x 320, K=100; cfg5: 128 x (3+9) x 192 x 640, K=100) through
size-independent properties plus spot checks against the oracle:

 * rows sorted by the canonical key (score desc, flat index asc), counts consistent with the padding, no duplicates;
 * a random subset of images compared bit-exactly with the oracle run by torch on the same GPU;
 * idempotence: a second call, a call without speculation and a call on the shape-generic kernels give identical bits
   (the speculative start threshold, the remembered thresholds and the kernel variant must not leak into the result).
"""
import numpy as np
import pytest
import torch

import kpt_oracle
import parity
from oracle import decode_ref
from rtm3d_b200 import HeatmapDecoder

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _inputs(B, C, Cv, H, W, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    mk = lambda c: torch.randn((B, c, H, W), generator=g, device=DEV, dtype=torch.float32)
    return [mk(C), mk(16), mk(2), mk(2)], mk(Cv)


def _same(a, b, fields):
    for f in fields:
        assert torch.equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("shape", [(256, 3, 9, 96, 320, 100), (128, 3, 9, 192, 640, 100)], ids=["cfg4", "cfg5"])
def test_full_size_properties_and_spot_checks(shape):
    B, C, Cv, H, W, K = shape
    logits, kpt = _inputs(B, C, Cv, H, W, seed=4321)
    dec = HeatmapDecoder(0.4, K, 4.0)
    det, cand, grp = dec.decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    # ---- properties over the whole batch
    cnt = det.counts.long()
    assert int(cnt.min()) >= 0 and int(cnt.max()) <= K
    valid = torch.arange(K, device=DEV).view(1, K) < cnt.view(B, 1)
    assert torch.all(det.flat[~valid] == -1) and torch.all(det.cls[~valid] == -1) and torch.all(det.score[~valid] == 0)
    assert torch.all(det.score[valid] > 0.4)
    s, f = det.score, det.flat.long()
    ok = (s[:, :-1] > s[:, 1:]) | ((s[:, :-1] == s[:, 1:]) & (f[:, :-1] < f[:, 1:]))
    assert torch.all(ok | ~valid[:, 1:]), "Tier A rows are not in canonical (score desc, index asc) order"
    ks, kf = cand.score, cand.flat.long()
    okk = (ks[..., :-1] > ks[..., 1:]) | ((ks[..., :-1] == ks[..., 1:]) & (kf[..., :-1] < kf[..., 1:]))
    assert torch.all(okk), "Tier B rows are not in canonical order"
    assert torch.all(det.cls[valid] == (f[valid] // (H * W)))
    for b in range(0, B, max(1, B // 16)):                     # no duplicate peaks
        n = int(cnt[b])
        assert len(set(f[b, :n].tolist())) == n
        for c in range(Cv):
            assert len(set(kf[b, c].tolist())) == K
    # ---- spot checks against the oracle on the same GPU (bit-exact)
    rng = np.random.default_rng(7)
    for b in rng.choice(B, size=6, replace=False):
        b = int(b)
        r = decode_ref.decode_image(logits[0][b], logits[1][b], logits[2][b], 0.4, K, 4.0)
        n = int(cnt[b])
        got = dict(cls=det.cls[b, :n].cpu().numpy(), score=det.score[b, :n].cpu().numpy(), proj=det.proj[b, :n].cpu().numpy(),
                   verts=det.verts[b, :n].cpu().numpy(), bbox=det.bbox[b, :n].cpu().numpy(), flat=det.flat[b, :n].cpu().numpy().astype(np.int64))
        parity.assert_exact(got, {k: v.cpu().numpy() for k, v in r.items()}, ("flat", "cls", "score", "proj", "verts", "bbox"), f"image {b}")
        vs, vx, vy, vflat = decode_ref.keypoint_peaks(kpt[b], K)
        for c in range(Cv):
            o = np.lexsort((vflat[c].cpu().numpy(), -vs[c].cpu().numpy().astype(np.float64)))
            assert np.array_equal(cand.flat[b, c].cpu().numpy(), vflat[c].cpu().numpy()[o]), f"image {b} channel {c}"
            assert np.array_equal(cand.score[b, c].cpu().numpy().view(np.uint32), vs[c].cpu().numpy()[o].view(np.uint32))
        # the grouping of the same images (kpt_j / kpt_score / kpt_proj / verts, sub-pixel candidates) against the oracle
        kpt_oracle.check_image(det, cand, grp, logits, kpt, b, K, what=f"full-size image {b}")
    # ---- idempotence across calls, speculation and kernel variants
    fa = ("cls", "score", "proj", "verts", "bbox", "flat", "counts")
    fb, fg = ("score", "xy", "flat"), ("kpt_proj", "kpt_score", "kpt_j", "verts")
    det2, cand2, grp2 = dec.decode_with_keypoints(logits, kpt)
    _same(det, det2, fa); _same(cand, cand2, fb); _same(grp, grp2, fg)
    for other in (dict(legacy=True), dict(legacy=True, speculate=False), dict(debug=1)):   # round-1 kernel; scan kernel forced to deepen
        det3, cand3, grp3 = HeatmapDecoder(0.4, K, 4.0, **other).decode_with_keypoints(logits, kpt)
        _same(det, det3, fa); _same(cand, cand3, fb); _same(grp, grp3, fg)
    det4, cand4, grp4 = HeatmapDecoder(0.4, K, 4.0).decode_with_keypoints(logits, kpt, fused=False)
    _same(det, det4, fa); _same(cand, cand4, fb); _same(grp, grp4, fg)
    # fewer CTAs than SMs (bench.py at N > 1 leaves four SMs to NCCL), strips (the publish + merge path), reused result buffers
    det6, cand6, grp6 = HeatmapDecoder(0.4, K, 4.0, max_ctas=144, reuse_outputs=True).decode_with_keypoints(logits, kpt)
    _same(det, det6, fa); _same(cand, cand6, fb); _same(grp, grp6, fg)
    det7, cand7, grp7 = HeatmapDecoder(0.4, K, 4.0, split=2).decode_with_keypoints(logits, kpt)
    _same(det, det7, fa); _same(cand, cand7, fb); _same(grp, grp7, fg)
    if B * H * W <= 256 * 96 * 320:                            # the generic kernels are slow: cfg4 only
        det5, cand5, grp5 = HeatmapDecoder(0.4, K, 4.0, force_generic=True).decode_with_keypoints(logits, kpt)
        _same(det, det5, fa); _same(cand, cand5, fb); _same(grp, grp5, fg)


def test_full_size_bf16_follows_fp32_pipeline():
    """cfg4 size in bf16: exact widening, then the fp32 pipeline (oracle on logits.float()) for a subset of images."""
    B, C, Cv, H, W, K = 256, 3, 9, 96, 320, 100
    logits, kpt = _inputs(B, C, Cv, H, W, seed=99)
    logits = [t.bfloat16() for t in logits]
    kpt = kpt.bfloat16()
    det, cand, grp = HeatmapDecoder(0.4, K, 4.0).decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    for b in (0, 77, 255):
        r = decode_ref.decode_image(logits[0][b].float(), logits[1][b].float(), logits[2][b].float(), 0.4, K, 4.0)
        n = int(det.counts[b])
        got = dict(cls=det.cls[b, :n].cpu().numpy(), score=det.score[b, :n].cpu().numpy(), proj=det.proj[b, :n].cpu().numpy(),
                   verts=det.verts[b, :n].cpu().numpy(), bbox=det.bbox[b, :n].cpu().numpy(), flat=det.flat[b, :n].cpu().numpy().astype(np.int64))
        parity.assert_exact(got, {k: v.cpu().numpy() for k, v in r.items()}, ("flat", "cls", "score", "proj", "verts", "bbox"), f"bf16 image {b}")

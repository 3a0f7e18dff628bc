"""Property tests of the CUDA decode (Hypothesis) and the review's regression shapes.

Properties checked on random shapes / K / thresholds, all against definitions rather than against a second implementation:
every reported pixel is a 3x3 peak of its plane in the sigmoid domain, rows are sorted (score desc, index asc), counts obey
the strict threshold, no listed pixel is missed (the K-th score bounds every unlisted peak), decode is idempotent and leaves
its inputs untouched, and a permutation of the batch permutes the results."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, example, given, settings, strategies as st

import kpt_oracle
from rtm3d_b200 import HeatmapDecoder, _native, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _peak_scores(hm):
    """sigmoid, then 3x3 max-pool equality (utils/model_utils.py:17-26) -- the definition, evaluated with torch on the GPU"""
    s = torch.sigmoid(hm)
    mx = torch.nn.functional.max_pool2d(s, 3, stride=1, padding=1)
    return torch.where(mx == s, s, torch.zeros_like(s))


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(B=st.integers(1, 4), C=st.integers(1, 4), H=st.integers(3, 70), W=st.integers(3, 130), K=st.integers(1, 120),
       thresh=st.sampled_from([0.0, 0.1, 0.4, 0.9]), kind=st.sampled_from(["randn", "trained", "quant", "plateau", "few"]),
       seed=st.integers(0, 10_000), split=st.sampled_from([0, 1, 2]))
@example(B=1, C=1, H=3, W=4, K=3, thresh=0.0, kind="randn", seed=0, split=0)      # one 16-byte group per row (found by Hypothesis)
@example(B=3, C=2, H=19, W=4, K=55, thresh=0.0, kind="plateau", seed=254, split=1)
@example(B=2, C=3, H=1, W=8, K=5, thresh=0.1, kind="randn", seed=3, split=0)
@example(B=1, C=4, H=70, W=130, K=120, thresh=0.4, kind="quant", seed=9, split=2)
def test_main_selection_properties(B, C, H, W, K, thresh, kind, seed, split):
    K = min(K, C * H * W)
    logits, _ = synth.head_outputs(B, C, H, W, seed=seed, kind=kind, kpt_channels=1)
    logits = [t.to(DEV) for t in logits]
    before = [t.clone() for t in logits]
    dec = HeatmapDecoder(thresh, K, 4.0, split=split)
    det = dec.decode_packed(logits)
    again = HeatmapDecoder(thresh, K, 4.0, force_generic=True).decode_packed(logits)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(before, logits)), "inputs modified"
    for f in ("score", "flat", "counts", "cls", "proj", "verts", "bbox"):
        assert torch.equal(getattr(det, f), getattr(again, f)), f"scan and generic kernels disagree on {f}"
    peaks = _peak_scores(logits[0]).view(B, -1)
    for b in range(B):
        n = int(det.counts[b])
        flat, score = det.flat[b, :n].long(), det.score[b, :n]
        assert torch.equal(peaks[b, flat], score), "a listed pixel is not a peak with that score"
        assert bool((score > thresh).all())
        if n > 1:
            s, f = score.cpu().numpy().astype(np.float64), flat.cpu().numpy()
            assert np.all((s[:-1] > s[1:]) | ((s[:-1] == s[1:]) & (f[:-1] < f[1:]))), "order is not (score desc, index asc)"
        rest = peaks[b].clone()
        rest[flat] = 0
        if n < K:
            assert float(rest.max()) <= thresh, "a peak above the threshold is missing"
        else:
            assert float(rest.max()) <= float(score[-1]), "an unlisted peak beats the K-th listed one"
        assert torch.equal(det.cls[b, :n], flat // (H * W))


@settings(max_examples=20, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(B=st.integers(2, 4), H=st.integers(8, 48), W4=st.integers(2, 24), K=st.integers(1, 60), Cv=st.sampled_from([8, 9]),
       seed=st.integers(0, 10_000), kind=st.sampled_from(["randn", "trained", "few"]))
def test_batch_permutation_and_oracle(B, H, W4, K, Cv, seed, kind):
    W = 4 * W4
    logits, kpt = synth.head_outputs(B, 3, H, W, seed=seed, kind=kind, kpt_channels=Cv, kpt_kind=kind)
    logits, kpt = [t.to(DEV) for t in logits], kpt.to(DEV)
    K = min(K, H * W)
    dec = HeatmapDecoder(0.4, K, 4.0)
    det, cand, grp = dec.decode_with_keypoints(logits, kpt)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(seed)).to(DEV)
    det2, cand2, grp2 = HeatmapDecoder(0.4, K, 4.0).decode_with_keypoints([t[perm].contiguous() for t in logits], kpt[perm].contiguous())
    torch.cuda.synchronize()
    for a, b in ((det.score, det2.score), (det.flat, det2.flat), (det.counts, det2.counts), (det.bbox, det2.bbox), (cand.score, cand2.score),
                 (cand.flat, cand2.flat), (cand.xy, cand2.xy), (grp.kpt_j, grp2.kpt_j), (grp.kpt_proj, grp2.kpt_proj)):
        assert torch.equal(a[perm], b), "images are not decoded independently"
    kpt_oracle.check_image(det, cand, grp, logits, kpt, 0, K, what=f"hypothesis {B}x{H}x{W} K{K}")


# ---------------------------------------------------------------------------------------------------------------
# shapes whose ring of the round-1 plane kernel lands between "fits with 4 KB reserved" and "fits with the real static shared
# memory": they must fall through to another kernel, not fail the launch
@pytest.mark.parametrize("variant", [dict(), dict(legacy=True), dict(legacy=True, split=1)], ids=["auto", "legacy", "legacy-s1"])
@pytest.mark.parametrize("shape", [(224, 80), (54, 324), (96, 320)])
def test_shapes_at_the_shared_memory_limit(shape, variant):
    H, W = shape
    logits, kpt = synth.head_outputs(2, 3, H, W, seed=H + W, kind="randn", kpt_channels=9)
    logits, kpt = [t.to(DEV) for t in logits], kpt.to(DEV)
    det, cand, grp = HeatmapDecoder(0.4, 100, 4.0, **variant).decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    for b in range(2):
        kpt_oracle.check_image(det, cand, grp, logits, kpt, b, 100, what=f"{shape} {variant}")


def test_generic_fallback_keeps_the_workspace_table():
    """rtm3d_workspace_init writes the logit-bound table once; a fused decode that takes the generic kernels (FORCE_GENERIC, or
    an ineligible map) resets its tickets on the same workspace and must leave the table and its magic word alone."""
    logits, kpt = synth.head_outputs(2, 3, 96, 320, seed=5, kind="randn", kpt_channels=9)
    logits, kpt = [t.to(DEV) for t in logits], kpt.to(DEV)
    dec = HeatmapDecoder(0.4, 50, 4.0, legacy=True)
    dec.decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    (ws,) = list(dec._ws.values())
    table = ws[:8192].view(torch.int32).clone()
    assert (int(table[2047]) & 0xFFFFFFFF) == 0x5A17AB1E
    dec.flags |= _native.FLAG_FORCE_GENERIC
    a = dec.decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    assert torch.equal(ws[:8192].view(torch.int32), table), "the generic fallback wiped the workspace's threshold table"
    dec.flags &= ~_native.FLAG_FORCE_GENERIC
    b = dec.decode_with_keypoints(logits, kpt)
    torch.cuda.synchronize()
    assert torch.equal(ws[:8192].view(torch.int32), table)
    assert torch.equal(a[0].flat, b[0].flat) and torch.equal(a[1].flat, b[1].flat) and torch.equal(a[2].kpt_j, b[2].kpt_j)

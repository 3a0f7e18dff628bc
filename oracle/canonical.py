"""Numpy restatement with the ordering contract made explicit.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``torch.topk`` leaves the order inside an exact-score tie group implementation-defined (CPU: heap artefact;
CUDA: ascending index -- measured, profiles/r01_probe_sigmoid_topk.json).  The kernels' contract is the canonical
order **(score desc, flat index asc)** with the lowest indices kept in the boundary tie group, which is what the
reference produces when it runs on CUDA (detect.py:25-30).  This module states that rule with np.lexsort so the
tests do not depend on which torch.topk they happen to run.

The sigmoid itself is NOT restated here (its last-ulp behaviour is the device library's): callers pass the
post-sigmoid probabilities they got from torch on the device they are checking against.

Follows utils/model_utils.py:17-26 (peak keep) and models/model.py:87-98, :109-114 (top-K, index split).
"""
from __future__ import annotations

import numpy as np


def peak_scores_np(p: np.ndarray) -> np.ndarray:
    """[C,H,W] probabilities -> same shape, non-peaks set to 0.0 (3x3 window, -inf outside, ties all kept)."""
    C, H, W = p.shape
    pad = np.full((C, H + 2, W + 2), -np.inf, dtype=p.dtype)
    pad[:, 1:-1, 1:-1] = p
    m = pad[:, 1:-1, 1:-1].copy()
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            np.maximum(m, pad[:, dy:dy + H, dx:dx + W], out=m)
    return np.where(m == p, p, p.dtype.type(0))


def topk_canonical(flat_scores: np.ndarray, k: int):
    """Exact top-k of a 1-D array under (score desc, index asc).  Returns (scores[k], idx int64[k])."""
    n = flat_scores.shape[0]
    idx = np.arange(n, dtype=np.int64)
    order = np.lexsort((idx, -flat_scores.astype(np.float64)))[:k]
    return flat_scores[order], order


def main_peaks_canonical(p_main: np.ndarray, thresh: float, k: int):
    """Tier A selection from post-sigmoid probabilities [C,H,W] -> (cls, score, xi, yi, flat), all length N <= k."""
    C, H, W = p_main.shape
    s = peak_scores_np(p_main).reshape(-1)
    sc, flat = topk_canonical(s, k)
    keep = sc > np.float32(thresh)
    sc, flat = sc[keep], flat[keep]
    cls = flat // (H * W)
    rem = flat % (H * W)
    return cls, sc, rem % W, rem // W, flat


def keypoint_peaks_canonical(p_kpt: np.ndarray, k: int):
    """Tier B per-channel selection (no threshold; 0.0 fillers in ascending index order).  -> (score [Cv,k], flat [Cv,k])."""
    Cv, H, W = p_kpt.shape
    s = peak_scores_np(p_kpt).reshape(Cv, -1)
    sc = np.empty((Cv, k), dtype=p_kpt.dtype)
    fl = np.empty((Cv, k), dtype=np.int64)
    for c in range(Cv):
        sc[c], fl[c] = topk_canonical(s[c], k)
    return sc, fl


def canonicalize_ties(score: np.ndarray, *cols: np.ndarray, key: np.ndarray):
    """Reorder rows inside runs of exactly equal ``score`` by ascending ``key`` (used to compare against a
    torch.topk whose tie order is arbitrary).  Returns (n_tie_groups_touched, reordered score, reordered cols...)."""
    n = score.shape[0]
    order = np.arange(n)
    touched = 0
    i = 0
    while i < n:
        j = i + 1
        while j < n and score[j] == score[i]:
            j += 1
        if j - i > 1:
            sub = order[i:j]
            srt = sub[np.argsort(key[sub], kind="stable")]
            if not np.array_equal(srt, sub):
                touched += 1
            order[i:j] = srt
        i = j
    return (touched, score[order]) + tuple(c[order] for c in cols)

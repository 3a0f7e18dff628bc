"""Comparators shared by the parity tests.

Two bars (north_star): indices / class ids / ordering bit-exact; floats within 1e-4 relative.

* ``assert_exact``      -- against an oracle evaluated on the SAME device library (torch CUDA): everything must be
                           bit-identical, after canonicalising the order inside exact-score tie groups (torch.topk
                           leaves that order implementation-defined; on CUDA it is already canonical).
* ``assert_ulp_tolerant`` -- against outputs computed with a DIFFERENT sigmoid implementation (torch CPU / Sleef, up
                           to 4 ulp away from the CUDA libdevice one, profiles/r01_probe_sigmoid_topk.json): scores
                           must agree within ULP_TOL ulp, the index sequence must agree except for permutations inside
                           runs whose scores are within ULP_TOL ulp of each other, and membership may differ only for
                           elements within ULP_TOL ulp of the K-th score or of the threshold.
"""
from __future__ import annotations

import numpy as np

ULP_TOL = 8          # CPU-vs-CUDA sigmoid differs by <= 4 ulp (measured); two values -> 8
REL_TOL = 1e-4       # north_star float tolerance


def f32_ordinal(a: np.ndarray) -> np.ndarray:
    """Monotone integer image of float32 (so that differences count ulps)."""
    i = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.where(i < 0, np.int64(-(2 ** 31)) - i, i)


def ulp_distance(a, b) -> np.ndarray:
    return np.abs(f32_ordinal(np.asarray(a)) - f32_ordinal(np.asarray(b)))


def assert_close_rel(got, want, what, rel=REL_TOL, abs_=1e-5):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    err = np.abs(got - want)
    bound = rel * np.abs(want) + abs_
    assert np.all(err <= bound), f"{what}: max err {err.max():.3e} (rel bound {rel}) at {np.argmax(err - bound)}"


def canonical_order(score: np.ndarray, flat: np.ndarray) -> np.ndarray:
    return np.lexsort((flat, -score.astype(np.float64)))


def assert_exact(got: dict, want: dict, fields, what=""):
    """Bit-exact comparison of one image.  ``got``/``want``: dicts of numpy arrays with leading dim N."""
    n_g, n_w = len(got["flat"]), len(want["flat"])
    assert n_g == n_w, f"{what}: count {n_g} != {n_w}"
    if n_g == 0:
        return
    og = canonical_order(got["score"], got["flat"])
    ow = canonical_order(want["score"], want["flat"])
    assert np.array_equal(og, np.arange(n_g)), f"{what}: kernel output is not in canonical (score desc, index asc) order"
    for f in fields:
        g = np.asarray(got[f])[og]
        w = np.asarray(want[f])[ow]
        if g.dtype.kind == "f":
            same = g.view(np.uint32) == w.astype(np.float32).view(np.uint32)
        else:
            same = g == w
        assert np.all(same), f"{what}: field {f!r} differs at rows {np.unique(np.argwhere(~same)[:, 0])[:8]} (bit-exact bar)"


def assert_ulp_tolerant(got: dict, want: dict, float_fields, nms_scores_want: np.ndarray, K: int, thresh: float,
                        what="", ulp=ULP_TOL):
    """One image, oracle from another sigmoid implementation.

    nms_scores_want: the oracle's full flat score map (non-peaks 0) used to judge boundary membership.
    Returns the number of positions where the index sequences differed (all inside near-tie runs).
    """
    gs, gf = np.asarray(got["score"]), np.asarray(got["flat"]).astype(np.int64)
    ws, wf = np.asarray(want["score"]), np.asarray(want["flat"]).astype(np.int64)
    th = np.float32(thresh)
    # --- membership
    # cut-off the oracle applied: its K-th score when it returned K rows, else the threshold (everything above it is in)
    bound = ws[-1] if len(wf) == K else th
    # (a count difference is legal only through elements at the cut-off, which the two loops below enforce)
    assert len(gf) <= K and len(set(gf.tolist())) == len(gf), f"{what}: kernel returned duplicates or more than K rows"
    for i, f in enumerate(gf):
        sw = nms_scores_want[f]
        assert f32_ordinal(sw) >= f32_ordinal(bound) - ulp, \
            f"{what}: kernel picked flat {f} whose oracle score {sw} is below the cut-off {bound} by more than {ulp} ulp"
        assert ulp_distance(gs[i], sw) <= ulp, f"{what}: score at flat {f}: {gs[i]} vs oracle {sw}"
    gset = set(gf.tolist())
    for i, f in enumerate(wf):
        if f not in gset:
            assert ulp_distance(ws[i], bound) <= ulp, \
                f"{what}: oracle element flat {f} (score {ws[i]}) missing from kernel output and not at the cut-off"
    # --- order: equal except inside near-tie runs
    n = min(len(gf), len(wf))
    mism = np.nonzero(gf[:n] != wf[:n])[0]
    for i in mism:
        f = gf[i]
        assert ulp_distance(nms_scores_want[f], ws[i]) <= ulp, \
            f"{what}: order differs at rank {i} (kernel flat {f}, oracle flat {wf[i]}) outside a near-tie run"
    # --- floats on the common rows
    wpos = {int(f): i for i, f in enumerate(wf)}
    rows_g = [i for i, f in enumerate(gf) if int(f) in wpos]
    rows_w = [wpos[int(gf[i])] for i in rows_g]
    for f in float_fields:
        assert_close_rel(np.asarray(got[f])[rows_g], np.asarray(want[f])[rows_w], f"{what}:{f}")
    for f in ("cls",):
        if f in got and f in want:
            assert np.array_equal(np.asarray(got[f])[rows_g], np.asarray(want[f])[rows_w]), f"{what}: cls differs"
    return len(mism)

"""Parity of the CUDA decode path (through the C ABI) against the oracle.  Needs a B200: run with -m gpu.

 * test_vs_cuda_oracle_*  : oracle/decode_ref.py evaluated with torch ON THE SAME GPU (how detect.py runs the reference,
                            detect.py:25-30): everything bit-exact -- indices, class ids, order, scores and all floats.
 * test_vs_golden_*       : tests/golden = outputs of the real reference executed on CPU (different sigmoid
                            implementation, up to 4 ulp away): indices/order exact outside near-tie runs, floats 1e-4.
"""
import numpy as np
import pytest
import torch

import golden_io
import parity
from oracle import decode_ref
from rtm3d_b200 import HeatmapDecoder, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FLOAT_FIELDS = ("score", "proj", "verts", "bbox")
# kernel variants: the shape-generic strip kernels, the plane-streaming kernel with its own choice of strips per plane
# (falls back to generic on ineligible shapes), and the plane-streaming kernel forced to 1/2/4/8 strips per plane
# (the scan kernel with 1/2/4/8 strips per plane, its deepening and exact-radix paths forced, and the round-1 kernel)
VARIANTS = [dict(force_generic=True), dict(), dict(split=1), dict(split=2), dict(split=4), dict(split=8),
            dict(debug=1), dict(debug=2), dict(debug=3, split=2), dict(legacy=True), dict(legacy=True, split=2)]
VARIANT_IDS = ["generic", "auto", "s1", "s2", "s4", "s8", "deepen", "exact", "deepen-exact-s2", "legacy", "legacy-s2"]


def _rows(packed, b):
    n = int(packed.counts[b])
    return dict(cls=packed.cls[b, :n].cpu().numpy(), score=packed.score[b, :n].cpu().numpy(),
                proj=packed.proj[b, :n].cpu().numpy(), verts=packed.verts[b, :n].cpu().numpy(),
                bbox=packed.bbox[b, :n].cpu().numpy(), flat=packed.flat[b, :n].cpu().numpy().astype(np.int64))


def _oracle_rows(r):
    return {k: v.cpu().numpy() for k, v in r.items()}


def _check_padding(packed):
    B, K = packed.score.shape
    cnt = packed.counts.cpu()
    for b in range(B):
        n = int(cnt[b])
        assert torch.all(packed.cls[b, n:] == -1) and torch.all(packed.flat[b, n:] == -1)
        assert torch.all(packed.score[b, n:] == 0) and torch.all(packed.verts[b, n:] == 0) and torch.all(packed.bbox[b, n:] == 0)


def _run_exact(logits_cpu, K, thresh, force_generic, dtype=torch.float32, what=""):
    logits = [t.to(DEV).to(dtype).contiguous() for t in logits_cpu]
    before = [t.clone() for t in logits]
    dec = HeatmapDecoder(thresh, K, 4.0, **force_generic)
    packed = dec.decode_packed(logits)
    torch.cuda.synchronize()
    for a, b_ in zip(before, logits):
        assert torch.equal(a, b_), "decoder modified its inputs"
    _check_padding(packed)
    ref_in = [t.float() for t in logits]  # bf16 contract: exact widening, then the fp32 pipeline
    for b in range(logits[0].shape[0]):
        r = decode_ref.decode_image(ref_in[0][b], ref_in[1][b], ref_in[2][b], thresh, K, 4.0)
        got = _rows(packed, b)
        if r is None:
            assert len(got["flat"]) == 0, f"{what} image {b}: oracle has no detection"
            continue
        parity.assert_exact(got, _oracle_rows(r), ("flat", "cls", "score", "proj", "verts", "bbox"), f"{what} image {b}")
    return packed


SHAPES = [
    # B, C, H, W, K
    (3, 3, 24, 40, 20),
    (2, 1, 19, 37, 1),
    (2, 9, 19, 37, 30),
    (2, 8, 21, 50, 128),
    (2, 3, 96, 320, 50),
    (2, 3, 96, 320, 100),
    (1, 3, 192, 640, 100),
    (5, 3, 33, 64, 100),
    (1, 2, 5, 8, 16),
    (1, 1, 1, 64, 8),
    (1, 1, 64, 1, 8),
]


@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("kind", synth.KINDS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_vs_cuda_oracle_exact(shape, kind, force_generic):
    B, C, H, W, K = shape
    logits, _ = synth.head_outputs(B, C, H, W, seed=1000 + H * 7 + W, kind=kind)
    _run_exact(logits, K, 0.4, force_generic, what=f"{kind} {shape}")


@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("thresh", [0.0, 0.05, 0.4, 0.9, 0.999])
def test_thresholds_exact(thresh, force_generic):
    logits, _ = synth.head_outputs(3, 3, 48, 80, seed=77, kind="trained")
    _run_exact(logits, 50, thresh, force_generic, what=f"thresh {thresh}")
    logits, _ = synth.head_outputs(2, 3, 48, 80, seed=78, kind="randn")
    _run_exact(logits, 100, thresh, force_generic, what=f"thresh {thresh}")


@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("K", [1, 30, 50, 100, 128, 256, 1024])
def test_topk_sizes_exact(K, force_generic):
    logits, _ = synth.head_outputs(2, 3, 96, 320, seed=5, kind="randn")
    _run_exact(logits, K, 0.4, force_generic, what=f"K {K}")


@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("kind", ["randn", "trained", "quant"])
def test_bf16_inputs_follow_fp32_pipeline(kind, force_generic):
    logits, _ = synth.head_outputs(2, 3, 96, 320, seed=9, kind=kind)
    _run_exact(logits, 50, 0.4, force_generic, dtype=torch.bfloat16, what=f"bf16 {kind}")


@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
def test_adversarial_saturated_plateaus(force_generic):
    """All-equal saturated maps: every pixel is a peak with score 1.0, top-K = the K lowest flat indices."""
    logits, _ = synth.head_outputs(2, 3, 24, 40, seed=3, kind="randn")
    logits[0].fill_(30.0)
    p = _run_exact(logits, 64, 0.4, force_generic, what="all 30.0")
    assert torch.equal(p.flat[0].cpu(), torch.arange(64, dtype=torch.int32))
    logits[0].fill_(18.0)
    logits[0][:, :, ::2, ::2] = 19.0  # both collapse to 1.0f: both are peaks in the sigmoid domain
    _run_exact(logits, 64, 0.4, force_generic, what="18/19")
    logits[0].fill_(-200.0)           # sigmoid underflows to 0: no detections
    _run_exact(logits, 64, 0.0, force_generic, what="-200")


@pytest.mark.parametrize("name", sorted(n for n, c in golden_io.CASES.items() if not c["kpt"]))
@pytest.mark.parametrize("force_generic", VARIANTS, ids=VARIANT_IDS)
def test_vs_golden_reference_outputs(name, force_generic):
    c = golden_io.CASES[name]
    logits_cpu, _ = golden_io.inputs(name)
    gold = golden_io.arrays(name)
    dec = HeatmapDecoder(c["thresh"], c["K"], c["down"], **force_generic)
    packed = dec.decode_packed([t.to(DEV) for t in logits_cpu])
    torch.cuda.synchronize()
    swaps = 0
    for b in range(c["B"]):
        want = golden_io.image_rows(gold, b)
        nms = decode_ref.peak_scores(logits_cpu[0][b]).reshape(-1).numpy()
        swaps += parity.assert_ulp_tolerant(_rows(packed, b), want, FLOAT_FIELDS, nms, c["K"], c["thresh"],
                                            what=f"{name} image {b}")
    # tie-free cases must agree position by position
    if c["kind"] in ("randn", "trained") and c["K"] <= 50:
        assert swaps <= 2, f"{name}: {swaps} rank differences on a tie-free case"


def test_list_api_matches_reference_layout():
    logits_cpu, _ = synth.head_outputs(3, 3, 24, 40, seed=11, kind="few")
    logits_cpu[0][1].fill_(-20.0)  # image 1: nothing above threshold -> None
    logits = [t.to(DEV) for t in logits_cpu]
    dec = HeatmapDecoder(0.4, 20, 4.0)
    clses, scores, projs, verts, bboxes = dec.decode(logits)
    want = decode_ref.decode(logits, 0.4, 20, 4.0)
    for got_l, want_l in zip((clses, scores, projs, verts, bboxes), want):
        assert len(got_l) == 3
        for g, w in zip(got_l, want_l):
            assert (g is None) == (w is None)
            if g is not None:
                assert g.dtype == w.dtype and g.shape == w.shape and g.device == w.device
                assert torch.equal(g, w)
    assert clses[1] is None and clses[0].dtype == torch.int64 and verts[0].shape[1:] == (8, 2)


def test_rejects_bad_inputs():
    dec = HeatmapDecoder(0.4, 20, 4.0)
    logits, _ = synth.head_outputs(1, 3, 16, 16, seed=1)
    dev = [t.to(DEV) for t in logits]
    with pytest.raises(ValueError):
        dec.decode_packed(logits)                                   # CPU tensors: no fallback
    with pytest.raises(ValueError):
        dec.decode_packed([dev[0].permute(0, 1, 3, 2)] + dev[1:])   # non-contiguous
    with pytest.raises(TypeError):
        dec.decode_packed([t.half() for t in dev])
    with pytest.raises(ValueError):
        HeatmapDecoder(-0.1, 20, 4.0)
    with pytest.raises(ValueError):
        HeatmapDecoder(0.4, 2000, 4.0)
    big = HeatmapDecoder(0.4, 1024, 4.0)
    with pytest.raises(ValueError):
        big.decode_packed(dev)                                      # K > C*H*W


def test_repeatable_and_workspace_self_cleaning():
    logits, _ = synth.head_outputs(4, 3, 96, 320, seed=21, kind="randn")
    dev = [t.to(DEV) for t in logits]
    dec = HeatmapDecoder(0.4, 100, 4.0)
    first = dec.decode_packed(dev)
    for _ in range(5):
        again = dec.decode_packed(dev)
        for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
            assert torch.equal(getattr(first, f), getattr(again, f)), f


@pytest.mark.parametrize("max_ctas", [1, 3, 0])
@pytest.mark.parametrize("kernel", [dict(), dict(debug=1), dict(legacy=True), dict(legacy=True, speculate=False)],
                         ids=["scan", "scan-deepen", "legacy", "legacy-nospec"])
def test_thresholds_chosen_by_the_kernels_never_leak_into_the_result(max_ctas, kernel):
    """The scan kernel picks a threshold per strip from an order statistic of the strip's own maxima and verifies it; the
    round-1 kernel starts a plane at the threshold remembered from the previous plane of the same index and redoes the
    plane when that turns out too high.  Batches whose images alternate between strong, weak, empty and plateau maps make
    any guess fail constantly; few CTAs make every CTA see many planes."""
    B, C, H, W, K = 24, 3, 48, 80, 50
    logits, kpt = synth.head_outputs(B, C, H, W, seed=4242, kind="randn", kpt_channels=9)
    scale = torch.tensor([3.0, 0.05, 1.0, 0.3, 6.0, 0.0])[torch.arange(B) % 6].view(B, 1, 1, 1)
    shift = torch.tensor([0.0, -3.0, 2.0, 0.0, -8.0, 1.0])[torch.arange(B) % 6].view(B, 1, 1, 1)
    logits[0] = logits[0] * scale + shift
    kpt = kpt * scale + shift
    dev_logits = [t.to(DEV) for t in logits]
    dec = HeatmapDecoder(0.4, K, 4.0, max_ctas=max_ctas, **kernel)
    packed = dec.decode_packed(dev_logits)
    cand = dec.decode_keypoints(kpt.to(DEV), dev_logits[3])
    torch.cuda.synchronize()
    for b in range(B):
        r = decode_ref.decode_image(dev_logits[0][b], dev_logits[1][b], dev_logits[2][b], 0.4, K, 4.0)
        got = _rows(packed, b)
        if r is None:
            assert len(got["flat"]) == 0
        else:
            parity.assert_exact(got, _oracle_rows(r), ("flat", "cls", "score", "proj", "verts", "bbox"), f"image {b}")
        vs, vx, vy, vflat = decode_ref.keypoint_peaks(kpt[b].to(DEV), K)
        for c in range(9):
            o = np.lexsort((vflat[c].cpu().numpy(), -vs[c].cpu().numpy().astype(np.float64)))
            assert np.array_equal(cand.flat[b, c].cpu().numpy(), vflat[c].cpu().numpy()[o]), f"b{b} c{c} kflat"
            assert np.array_equal(cand.score[b, c].cpu().numpy().view(np.uint32), vs[c].cpu().numpy()[o].view(np.uint32))
    # the workspace is left clean: a second run gives the same result
    again = dec.decode_packed(dev_logits)
    assert torch.equal(again.flat, packed.flat) and torch.equal(again.score, packed.score)


def test_pack_wire_kernel_matches_torch_packing():
    """rtm3d_pack_wire (the gather's wire rows, one launch) == the torch cat/cast chain, bit for bit."""
    from rtm3d_b200.decoder import PackedDetections
    logits, _ = synth.head_outputs(3, 3, 48, 80, seed=33, kind="randn")
    dec = HeatmapDecoder(0.4, 30, 4.0)
    det = dec.decode_packed([t.to(DEV) for t in logits])
    torch.cuda.synchronize()
    wire = det.to_wire()
    cpu = PackedDetections(**{f: getattr(det, f).cpu() for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts")})
    assert torch.equal(wire.cpu(), cpu.to_wire())
    back = PackedDetections.from_wire(wire, 30)
    for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
        assert torch.equal(getattr(back, f), getattr(det, f)), f


def test_zeroed_workspace_without_threshold_table_gives_the_same_bits():
    """rtm3d_workspace_init also writes the per-bin logit bounds; a workspace that was only zeroed (the pre-table
    contract) makes the kernel compute the bounds itself.  Both must give the same result."""
    logits, kpt = synth.head_outputs(4, 3, 96, 320, seed=77, kind="randn", kpt_channels=9)
    dev_logits = [t.to(DEV) for t in logits]
    dec = HeatmapDecoder(0.4, 50, 4.0)
    first = dec.decode_packed(dev_logits)
    cand = dec.decode_keypoints(kpt.to(DEV), dev_logits[3])
    torch.cuda.synchronize()
    for ws in dec._ws.values():
        ws.zero_()                       # erases the table (and the remembered thresholds)
    again = dec.decode_packed(dev_logits)
    cand2 = dec.decode_keypoints(kpt.to(DEV), dev_logits[3])
    torch.cuda.synchronize()
    for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
        assert torch.equal(getattr(first, f), getattr(again, f)), f
    assert torch.equal(cand.flat, cand2.flat) and torch.equal(cand.score, cand2.score)


def test_reuse_outputs_returns_the_same_buffers_with_the_same_bits():
    logits, kpt = synth.head_outputs(3, 3, 48, 80, seed=5, kind="randn", kpt_channels=9)
    dev_logits = [t.to(DEV) for t in logits]
    ref = HeatmapDecoder(0.4, 30, 4.0).decode_with_keypoints(dev_logits, kpt.to(DEV))
    dec = HeatmapDecoder(0.4, 30, 4.0, reuse_outputs=True)
    a = dec.decode_with_keypoints(dev_logits, kpt.to(DEV))
    b = dec.decode_with_keypoints(dev_logits, kpt.to(DEV))
    torch.cuda.synchronize()
    assert a[0].score.data_ptr() == b[0].score.data_ptr() and a[2].kpt_j.data_ptr() == b[2].kpt_j.data_ptr()
    for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
        assert torch.equal(getattr(ref[0], f), getattr(b[0], f)), f
    for f in ("kpt_proj", "kpt_score", "kpt_j", "verts"):
        assert torch.equal(getattr(ref[2], f), getattr(b[2], f)), f
    p1 = dec.decode_packed(dev_logits)
    p2 = dec.decode_packed(dev_logits)
    assert p1.flat.data_ptr() == p2.flat.data_ptr() and torch.equal(p2.flat, ref[0].flat)

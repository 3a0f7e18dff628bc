"""Host-side mirror of the reference decoder interface (models/model.py:29-75) over librtm3d_decode.so.

``HeatmapDecoder.decode(pred_logits)`` has the call signature and output layout of ``Model.inference``:
``pred_logits = [main_kf, offset_fr_main, main_offset, vertex_offset]`` (NCHW, models/model.py:30-31) ->
``(clses, m_scores, m_projs, v_projs_regress, bboxes_2d)``, five lists of length B whose elements are ``None`` for an
image without a detection above the threshold (models/model.py:43-44) and otherwise tensors on the input device
(int64 [N], f32 [N], f32 [N,2], f32 [N,8,2], f32 [N,4]) sorted by score descending.  They are zero-copy views of
fixed-size [B,K,...] buffers.  Unlike the reference the inputs are never modified, one D2H read of ``counts`` per
BATCH is the only host synchronisation (the reference syncs >= 3 times per image), and ``decode_packed`` does none.

PyTorch is used for device memory and streams only; all arithmetic runs in the CUDA library.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _native


@dataclass
class PackedDetections:
    """Fixed-size Tier A result for a batch; rows >= counts[b] are zero (cls = flat = -1)."""
    cls: torch.Tensor      # int64 [B,K]
    score: torch.Tensor    # f32   [B,K]
    proj: torch.Tensor     # f32   [B,K,2]   main centre, input pixels (x,y)
    verts: torch.Tensor    # f32   [B,K,V,2] regressed vertices, input pixels
    bbox: torch.Tensor     # f32   [B,K,4]   xmin,ymin,xmax,ymax over the regressed vertices
    flat: torch.Tensor     # int32 [B,K]     c*H*W + y*W + x of the peak
    counts: torch.Tensor   # int32 [B]

    WORDS = 1 + 1 + 2 + 16 + 4 + 1  # per-detection 32-bit words of the wire format used by the NCCL gather (V = 8)

    def to_wire(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int32 [B, K*(9+2V)+1]: per image K rows of (cls | score | proj | verts | bbox | flat) as bit patterns, then the
        image's count -- ONE buffer for ONE collective.  CUDA tensors are packed by the library (rtm3d_pack_wire, one
        launch); CPU tensors (the gloo tests) by torch."""
        B, K = self.score.shape
        V = self.verts.shape[2]
        per = K * (9 + 2 * V) + 1
        if out is None:
            out = torch.empty((B, per), dtype=torch.int32, device=self.score.device)
        if self.score.is_cuda:
            with torch.cuda.device(self.score.device):
                _native.check(_native.lib().rtm3d_pack_wire(
                    self.cls.data_ptr(), self.score.data_ptr(), self.proj.data_ptr(), self.verts.data_ptr(), self.bbox.data_ptr(),
                    self.flat.data_ptr(), self.counts.data_ptr(), B, K, V, out.data_ptr(),
                    torch.cuda.current_stream(self.score.device).cuda_stream), "rtm3d_pack_wire")
            return out
        parts = [self.cls.to(torch.int32).view(B, K, 1), self.score.view(torch.int32).view(B, K, 1),
                 self.proj.view(torch.int32).view(B, K, 2), self.verts.reshape(B, K, -1).view(torch.int32),
                 self.bbox.view(torch.int32).view(B, K, 4), self.flat.view(B, K, 1)]
        out[:, :per - 1] = torch.cat(parts, dim=-1).reshape(B, per - 1)
        out[:, per - 1] = self.counts
        return out

    @staticmethod
    def from_wire(wire: torch.Tensor, K: int) -> "PackedDetections":
        B, per = wire.shape
        Wd = (per - 1) // K
        V = (Wd - 9) // 2
        rows = wire[:, :per - 1].reshape(B, K, Wd)
        f = rows.view(torch.float32)
        return PackedDetections(cls=rows[..., 0].to(torch.int64), score=f[..., 1].contiguous(),
                                proj=f[..., 2:4].contiguous(), verts=f[..., 4:4 + 2 * V].reshape(B, K, V, 2).contiguous(),
                                bbox=f[..., 4 + 2 * V:8 + 2 * V].contiguous(), flat=rows[..., 8 + 2 * V].contiguous(),
                                counts=wire[:, per - 1].contiguous())


@dataclass
class KeypointCandidates:
    """Tier B per-channel candidates (models/model.py:100-115 + the commented sub-pixel wiring :52-60)."""
    score: torch.Tensor    # f32 [B,Cv,K]
    xy: torch.Tensor       # f32 [B,Cv,K,2] heat-map units (sub-pixel added, not scaled)
    flat: torch.Tensor     # int32 [B,Cv,K]  y*W + x


@dataclass
class GroupedKeypoints:
    """Tier B grouping (models/model.py:134-162), scaled by DOWN_SAMPLE as the commented :68-69 would."""
    kpt_proj: torch.Tensor   # f32 [B,K,Cv,2]
    kpt_score: torch.Tensor  # f32 [B,K,Cv]
    kpt_j: torch.Tensor      # int32 [B,K,Cv] index of the matched candidate
    verts: torch.Tensor      # f32 [B,K,Cv,2] regressed vertices (zero offset for channels >= n_vert)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _native.F32
    if t.dtype == torch.bfloat16:
        return _native.BF16
    raise TypeError(f"head maps must be float32 or bfloat16, got {t.dtype}")


def _check_map(t: torch.Tensor, name: str, shape=None) -> None:
    if not isinstance(t, torch.Tensor) or t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D NCHW tensor")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (rtm3d_b200 has no CPU path), got {t.device}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be NCHW-contiguous (non-contiguous inputs are rejected, not silently copied)")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: shape {tuple(t.shape)} != expected {tuple(shape)}")


class HeatmapDecoder:
    """B200 decoder with the reference's three configuration scalars
    (DETECTOR.SCORE_THRESH, DETECTOR.TOPK_CANDIDATES, MODEL.DOWN_SAMPLE -- models/model.py:41-42,67,70)."""

    def __init__(self, score_thresh: float = 0.5, topk: int = 30, down_sample: float = 4.0, force_generic: bool = False,
                 split: int = 0, speculate: bool = True, max_ctas: int = 0, reuse_outputs: bool = False, legacy: bool = False,
                 debug: int = 0):
        if not (score_thresh >= 0):
            raise ValueError("score_thresh must be >= 0: zero-score fillers of the peak map could pass a negative threshold")
        if not (1 <= int(topk) <= 1024):
            raise ValueError("topk must be in [1, 1024]")
        self.score_thresh = float(score_thresh)
        self.topk = int(topk)
        self.down_sample = float(down_sample)
        if split not in (0, 1, 2, 4, 8):
            raise ValueError("split (strips per plane of the streaming kernels) must be 0 (auto), 1, 2, 4 or 8")
        if not (0 <= int(max_ctas) <= 255):
            raise ValueError("max_ctas must be in [0, 255] (0 = one CTA per SM)")
        # legacy: the round-1 plane-streaming kernel instead of the plane-resident scan kernel (speculate only applies to it);
        # debug (tests): 1 = the scan kernel's first threshold is forced too high (deepening path), 2 = tiny candidate lists
        # (exact radix path), 3 = both
        self.flags = ((_native.FLAG_FORCE_GENERIC if force_generic else 0) | (0 if speculate else _native.FLAG_NO_SPECULATION)
                      | (_native.FLAG_LEGACY_PLANES if legacy else 0) | (split << 8) | (int(max_ctas) << 16) | ((int(debug) & 0xF) << 24))
        self._lib = _native.lib()
        self._ws = {}
        # reuse_outputs: decode_packed / decode_with_keypoints return the SAME result buffers on every call with the same
        # shapes (valid until the next call): saves a dozen allocations, ~35 us of host time per call -- more than the GPU
        # work of a small batch
        self.reuse_outputs = bool(reuse_outputs)
        self._out = {}
        # which kernels the staged (marks) form of decode_with_keypoints used last: "scan" (scan kernel + select/post kernel) or
        # "legacy" (round-1 plane-streaming kernel + post kernel: planes that do not fit the scan kernel's ring)
        self.staged_path = "legacy" if legacy else "scan"

    # ------------------------------------------------------------------ workspace
    def _workspace(self, device, B, C, H, W):
        stream = torch.cuda.current_stream(device)
        key = (device.index, stream.cuda_stream, B, C, H, W, self.topk)
        ws = self._ws.get(key)
        if ws is None:
            n = ctypes.c_size_t(0)
            _native.check(self._lib.rtm3d_decode_workspace_bytes(B, C, H, W, self.topk, ctypes.byref(n)), "workspace_bytes")
            ws = torch.empty(n.value, dtype=torch.uint8, device=device)
            _native.check(self._lib.rtm3d_workspace_init(ws.data_ptr(), n.value, stream.cuda_stream), "workspace_init")
            self._ws[key] = ws
        return ws, stream.cuda_stream

    # ------------------------------------------------------------------ Tier A
    def decode_packed(self, pred_logits: Sequence[torch.Tensor]) -> PackedDetections:
        """Asynchronous: enqueues one kernel on the current stream and returns fixed-size device buffers."""
        main, off, off2 = pred_logits[0], pred_logits[1], pred_logits[2]
        _check_map(main, "main_kf")
        B, C, H, W = main.shape
        if off.shape[1] % 2:
            raise ValueError("offset_fr_main must have an even channel count (dx,dy per vertex)")
        V = off.shape[1] // 2
        _check_map(off, "offset_fr_main", (B, 2 * V, H, W))
        _check_map(off2, "main_offset", (B, 2, H, W))
        if not (main.dtype == off.dtype == off2.dtype) or not (main.device == off.device == off2.device):
            raise ValueError("head maps must share dtype and device")
        dt = _dtype_code(main)
        dev, K = main.device, self.topk
        with torch.cuda.device(dev):
            ws, stream = self._workspace(dev, B, C, H, W)
            okey = ("main", dev.index, B, K, V)
            out = self._out.get(okey) if self.reuse_outputs else None
            if out is None:
                out = PackedDetections(
                    cls=torch.empty((B, K), dtype=torch.int64, device=dev),
                    score=torch.empty((B, K), dtype=torch.float32, device=dev),
                    proj=torch.empty((B, K, 2), dtype=torch.float32, device=dev),
                    verts=torch.empty((B, K, V, 2), dtype=torch.float32, device=dev),
                    bbox=torch.empty((B, K, 4), dtype=torch.float32, device=dev),
                    flat=torch.empty((B, K), dtype=torch.int32, device=dev),
                    counts=torch.empty((B,), dtype=torch.int32, device=dev))
                if self.reuse_outputs:
                    self._out[okey] = out
            rc = self._lib.rtm3d_decode_main(
                main.data_ptr(), off.data_ptr(), off2.data_ptr(), dt, B, C, H, W, V, K,
                self.score_thresh, self.down_sample,
                out.cls.data_ptr(), out.score.data_ptr(), out.proj.data_ptr(), out.verts.data_ptr(),
                out.bbox.data_ptr(), out.flat.data_ptr(), out.counts.data_ptr(),
                ws.data_ptr(), ws.numel(), self.flags, stream)
        _native.check(rc, "rtm3d_decode_main")
        return out

    def select_main(self, main: torch.Tensor) -> PackedDetections:
        """Selection only (models/model.py:77-98 without the gathers): ``score``, ``flat`` and ``counts`` of the K best peaks
        per image; the other fields of the returned PackedDetections are None.  For epilogues of the caller's own behind the
        peaks -- ``decode_box3d`` (BASELINE configs[2])."""
        _check_map(main, "main_kf")
        B, C, H, W = main.shape
        dev, K = main.device, self.topk
        with torch.cuda.device(dev):
            ws, stream = self._workspace(dev, B, C, H, W)
            okey = ("select", dev.index, B, K)
            out = self._out.get(okey) if self.reuse_outputs else None
            if out is None:
                out = PackedDetections(cls=None, score=torch.empty((B, K), dtype=torch.float32, device=dev), proj=None, verts=None,
                                       bbox=None, flat=torch.empty((B, K), dtype=torch.int32, device=dev),
                                       counts=torch.empty((B,), dtype=torch.int32, device=dev))
                if self.reuse_outputs:
                    self._out[okey] = out
            rc = self._lib.rtm3d_select_main(main.data_ptr(), _dtype_code(main), B, C, H, W, K, self.score_thresh,
                                             out.score.data_ptr(), out.flat.data_ptr(), out.counts.data_ptr(),
                                             ws.data_ptr(), ws.numel(), self.flags, stream)
        _native.check(rc, "rtm3d_select_main")
        return out

    def decode(self, pred_logits: Sequence[torch.Tensor]):
        """Drop-in for ``Model.inference`` (models/model.py:29-75)."""
        p = self.decode_packed(pred_logits)
        counts = p.counts.tolist()  # the one host synchronisation of the batch
        B = len(counts)
        clses: List[Optional[torch.Tensor]] = [None] * B
        m_scores: List[Optional[torch.Tensor]] = [None] * B
        m_projs: List[Optional[torch.Tensor]] = [None] * B
        v_projs_regress: List[Optional[torch.Tensor]] = [None] * B
        bboxes_2d: List[Optional[torch.Tensor]] = [None] * B
        for i, n in enumerate(counts):
            if n == 0:
                continue
            clses[i] = p.cls[i, :n]
            m_scores[i] = p.score[i, :n]
            m_projs[i] = p.proj[i, :n]
            v_projs_regress[i] = p.verts[i, :n]
            bboxes_2d[i] = p.bbox[i, :n]
        return clses, m_scores, m_projs, v_projs_regress, bboxes_2d

    __call__ = decode

    # ------------------------------------------------------------------ Tier B
    def decode_keypoints(self, kpt_logits: torch.Tensor, vertex_offset: torch.Tensor) -> KeypointCandidates:
        _check_map(kpt_logits, "vertex_kf")
        B, Cv, H, W = kpt_logits.shape
        _check_map(vertex_offset, "vertex_offset", (B, 2, H, W))
        if kpt_logits.dtype != vertex_offset.dtype:
            raise ValueError("head maps must share dtype")
        dt, dev, K = _dtype_code(kpt_logits), kpt_logits.device, self.topk
        with torch.cuda.device(dev):
            ws, stream = self._workspace(dev, B, Cv, H, W)
            out = KeypointCandidates(score=torch.empty((B, Cv, K), dtype=torch.float32, device=dev),
                                     xy=torch.empty((B, Cv, K, 2), dtype=torch.float32, device=dev),
                                     flat=torch.empty((B, Cv, K), dtype=torch.int32, device=dev))
            rc = self._lib.rtm3d_decode_keypoints(kpt_logits.data_ptr(), vertex_offset.data_ptr(), dt, B, Cv, H, W, K,
                                                  out.score.data_ptr(), out.xy.data_ptr(), out.flat.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), self.flags, stream)
        _native.check(rc, "rtm3d_decode_keypoints")
        return out

    def group_keypoints(self, det: PackedDetections, cand: KeypointCandidates, pred_logits) -> GroupedKeypoints:
        off, off2 = pred_logits[1], pred_logits[2]
        B, K = det.score.shape
        Cv = cand.score.shape[1]
        H, W = off.shape[2], off.shape[3]
        V = off.shape[1] // 2
        dev = off.device
        with torch.cuda.device(dev):
            out = GroupedKeypoints(kpt_proj=torch.empty((B, K, Cv, 2), dtype=torch.float32, device=dev),
                                   kpt_score=torch.empty((B, K, Cv), dtype=torch.float32, device=dev),
                                   kpt_j=torch.empty((B, K, Cv), dtype=torch.int32, device=dev),
                                   verts=torch.empty((B, K, Cv, 2), dtype=torch.float32, device=dev))
            rc = self._lib.rtm3d_group_vertices(det.flat.data_ptr(), det.counts.data_ptr(), off.data_ptr(), off2.data_ptr(),
                                                _dtype_code(off), B, H, W, V, K, cand.score.data_ptr(), cand.xy.data_ptr(),
                                                Cv, self.down_sample, out.kpt_proj.data_ptr(), out.kpt_score.data_ptr(),
                                                out.kpt_j.data_ptr(), out.verts.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_group_vertices")
        return out

    def decode_with_keypoints(self, pred_logits, kpt_logits, marks=None, fused=True, gather=None):
        """Tier A + Tier B as the commented wiring of models/model.py:45-62,68-69 describes.  Returns
        (PackedDetections, KeypointCandidates, GroupedKeypoints), all asynchronous.  ``fused`` (default): one call of
        ``rtm3d_decode_fused`` -- both heat-maps streamed by a single kernel launch, then the grouping kernel; otherwise
        the three separate entry points.  ``marks``: optional list that receives a recorded ``torch.cuda.Event`` before
        the plane-streaming kernel and after the epilogue and grouping kernels (bench.py times the kernels with it; the
        fused call is then issued as selection-only ``rtm3d_decode_fused`` + ``rtm3d_post_fused``: the same two launches
        with an event in between).  ``gather`` = (ctypes array of peer-mapped gather buffers, n_peers, rank, step id of the arrival flag or 0): the wire rows
        of the detections are stored straight into every rank's gather buffer by the select + post kernel
        (``rtm3d_decode_fused_gather``: the path's one exchange, fused; no pack kernel, no collective call).  A 5-tuple
        (peers of this batch's slot, peers of the previous batch's slot or None, n_peers, rank, id of the previous batch) selects
        ``rtm3d_decode_fused_gather_deferred``: the rows stay on this rank and the previous batch's rows are pushed while this
        launch sorts (flush the last batch with ``rtm3d_push_gather``)."""
        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        if not fused:
            mark()
            det = self.decode_packed(pred_logits)
            mark()
            cand = self.decode_keypoints(kpt_logits, pred_logits[3])
            mark()
            grp = self.group_keypoints(det, cand, pred_logits)
            mark()
            return det, cand, grp
        main, off, off2, voff2 = pred_logits[0], pred_logits[1], pred_logits[2], pred_logits[3]
        _check_map(main, "main_kf")
        B, C, H, W = main.shape
        if off.shape[1] % 2:
            raise ValueError("offset_fr_main must have an even channel count (dx,dy per vertex)")
        V = off.shape[1] // 2
        _check_map(off, "offset_fr_main", (B, 2 * V, H, W))
        _check_map(off2, "main_offset", (B, 2, H, W))
        _check_map(voff2, "vertex_offset", (B, 2, H, W))
        _check_map(kpt_logits, "vertex_kf")
        Cv = kpt_logits.shape[1]
        if tuple(kpt_logits.shape) != (B, Cv, H, W):
            raise ValueError("vertex_kf must have the main heat-map's batch and spatial shape")
        maps = (main, off, off2, voff2, kpt_logits)
        if len({t.dtype for t in maps}) != 1 or len({t.device for t in maps}) != 1:
            raise ValueError("head maps must share dtype and device")
        dt, dev, K = _dtype_code(main), main.device, self.topk
        e = lambda *sh, d=torch.float32: torch.empty(sh, dtype=d, device=dev)
        with torch.cuda.device(dev):
            ws, stream = self._workspace(dev, B, C + Cv, H, W)
            okey = ("fused", dev.index, B, K, V, Cv)
            cached = self._out.get(okey) if self.reuse_outputs else None
            if cached is None:
                det = PackedDetections(cls=e(B, K, d=torch.int64), score=e(B, K), proj=e(B, K, 2), verts=e(B, K, V, 2),
                                       bbox=e(B, K, 4), flat=e(B, K, d=torch.int32), counts=e(B, d=torch.int32))
                cand = KeypointCandidates(score=e(B, Cv, K), xy=e(B, Cv, K, 2), flat=e(B, Cv, K, d=torch.int32))
                grp = GroupedKeypoints(kpt_proj=e(B, K, Cv, 2), kpt_score=e(B, K, Cv), kpt_j=e(B, K, Cv, d=torch.int32),
                                       verts=e(B, K, Cv, 2))
                if self.reuse_outputs:
                    self._out[okey] = (det, cand, grp)
            else:
                det, cand, grp = cached
            if gather is not None:
                if marks is not None:
                    raise ValueError("gather and marks cannot be combined (the fused gather is one call)")
                outs = (det.cls.data_ptr(), det.score.data_ptr(), det.proj.data_ptr(), det.verts.data_ptr(), det.bbox.data_ptr(),
                        det.flat.data_ptr(), det.counts.data_ptr(), cand.score.data_ptr(), cand.xy.data_ptr(), cand.flat.data_ptr(),
                        grp.kpt_proj.data_ptr(), grp.kpt_score.data_ptr(), grp.kpt_j.data_ptr(), grp.verts.data_ptr())
                ins = (main.data_ptr(), off.data_ptr(), off2.data_ptr(), kpt_logits.data_ptr(), voff2.data_ptr(), dt,
                       B, C, Cv, H, W, V, K, self.score_thresh, self.down_sample)
                if len(gather) == 4:
                    peers, n_peers, rank, step_id = gather
                    rc = self._lib.rtm3d_decode_fused_gather(*ins, *outs, ws.data_ptr(), ws.numel(), self.flags, peers, n_peers, rank, step_id, stream)
                    _native.check(rc, "rtm3d_decode_fused_gather")
                else:
                    # deferred: (this batch's slot of every gather buffer, the previous batch's slot or None, n_peers, rank, id of the previous batch)
                    peers, prev_peers, n_peers, rank, prev_step_id = gather
                    rc = self._lib.rtm3d_decode_fused_gather_deferred(*ins, *outs, ws.data_ptr(), ws.numel(), self.flags, peers, prev_peers, n_peers,
                                                                     rank, prev_step_id, stream)
                    _native.check(rc, "rtm3d_decode_fused_gather_deferred")
                return det, cand, grp
            mark()
            legacy = bool(self.flags & (_native.FLAG_LEGACY_PLANES | _native.FLAG_FORCE_GENERIC))
            # marks: the same two launches as the unmarked call, issued separately with an event in between -- scan kernel, then
            # select + post kernel (legacy / generic kernels: selection, then rtm3d_post_fused)
            stage_flags = 0
            if marks is not None:
                stage_flags = (_native.FLAG_NO_GROUP | _native.FLAG_NO_EPILOGUE) if legacy else _native.FLAG_NO_SELECT
            rc = self._lib.rtm3d_decode_fused(
                main.data_ptr(), off.data_ptr(), off2.data_ptr(), kpt_logits.data_ptr(), voff2.data_ptr(), dt,
                B, C, Cv, H, W, V, K, self.score_thresh, self.down_sample,
                det.cls.data_ptr(), det.score.data_ptr(), det.proj.data_ptr(), det.verts.data_ptr(), det.bbox.data_ptr(),
                det.flat.data_ptr(), det.counts.data_ptr(), cand.score.data_ptr(), cand.xy.data_ptr(), cand.flat.data_ptr(),
                grp.kpt_proj.data_ptr(), grp.kpt_score.data_ptr(), grp.kpt_j.data_ptr(), grp.verts.data_ptr(),
                ws.data_ptr(), ws.numel(), self.flags | stage_flags, stream)
            if marks is not None and not legacy and rc == _native.ERR_SHAPE:
                # the scan kernel does not serve this shape (planes larger than its ring): the round-1 kernels in two stages
                legacy = True
                self.staged_path = "legacy"
                rc = self._lib.rtm3d_decode_fused(
                    main.data_ptr(), off.data_ptr(), off2.data_ptr(), kpt_logits.data_ptr(), voff2.data_ptr(), dt,
                    B, C, Cv, H, W, V, K, self.score_thresh, self.down_sample,
                    det.cls.data_ptr(), det.score.data_ptr(), det.proj.data_ptr(), det.verts.data_ptr(), det.bbox.data_ptr(),
                    det.flat.data_ptr(), det.counts.data_ptr(), cand.score.data_ptr(), cand.xy.data_ptr(), cand.flat.data_ptr(),
                    grp.kpt_proj.data_ptr(), grp.kpt_score.data_ptr(), grp.kpt_j.data_ptr(), grp.verts.data_ptr(),
                    ws.data_ptr(), ws.numel(), self.flags | _native.FLAG_NO_GROUP | _native.FLAG_NO_EPILOGUE, stream)
            mark()
            _native.check(rc, "rtm3d_decode_fused")
            if marks is not None and legacy:
                rc = self._lib.rtm3d_post_fused(
                    det.flat.data_ptr(), det.counts.data_ptr(), cand.flat.data_ptr(), cand.score.data_ptr(),
                    off.data_ptr(), off2.data_ptr(), voff2.data_ptr(), dt, B, C, Cv, H, W, V, K, self.down_sample,
                    det.cls.data_ptr(), det.proj.data_ptr(), det.verts.data_ptr(), det.bbox.data_ptr(), cand.xy.data_ptr(),
                    grp.kpt_proj.data_ptr(), grp.kpt_score.data_ptr(), grp.kpt_j.data_ptr(), grp.verts.data_ptr(), stream)
                mark()
                _native.check(rc, "rtm3d_post_fused")
            elif marks is not None:
                rc = self._lib.rtm3d_select_post(
                    off.data_ptr(), off2.data_ptr(), voff2.data_ptr(), dt, B, C, Cv, H, W, V, K, self.score_thresh, self.down_sample,
                    det.cls.data_ptr(), det.score.data_ptr(), det.proj.data_ptr(), det.verts.data_ptr(), det.bbox.data_ptr(),
                    det.flat.data_ptr(), det.counts.data_ptr(), cand.score.data_ptr(), cand.xy.data_ptr(), cand.flat.data_ptr(),
                    grp.kpt_proj.data_ptr(), grp.kpt_score.data_ptr(), grp.kpt_j.data_ptr(), grp.verts.data_ptr(),
                    ws.data_ptr(), ws.numel(), self.flags, stream)
                mark()
                _native.check(rc, "rtm3d_select_post")
        return det, cand, grp

    # ------------------------------------------------------------------ Tier C
    def decode_box3d(self, det: PackedDetections, reg: torch.Tensor, cam: torch.Tensor, dim_ref: torch.Tensor,
                     n_classes: int, multibin: bool = False, sigmoid_subpixel: bool = False,
                     depth_ref=(28.01, 16.32), cached: bool = False):
        """Closed-form 3D recovery at the Tier A peaks (NOT in the reference; spec: oracle/box3d_ref.py).  ``cached``: reuse the
        result buffers of the previous call with the same shapes (valid until the next call)."""
        _check_map(reg, "regression map")
        B, Creg, H, W = reg.shape
        K = det.score.shape[1]
        dev = reg.device
        cam = cam.to(device=dev, dtype=torch.float32).contiguous().view(B, 9)
        dim_ref = dim_ref.to(device=dev, dtype=torch.float32).contiguous().view(n_classes, 3)
        mode = (1 if multibin else 0) | (2 if sigmoid_subpixel else 0)
        with torch.cuda.device(dev):
            okey = ("box3d", dev.index, B, K)
            out = self._out.get(okey) if cached else None
            if out is None:
                out = dict(loc=torch.empty((B, K, 3), dtype=torch.float32, device=dev),
                           dim=torch.empty((B, K, 3), dtype=torch.float32, device=dev),
                           alpha=torch.empty((B, K), dtype=torch.float32, device=dev),
                           rot_y=torch.empty((B, K), dtype=torch.float32, device=dev),
                           corners2d=torch.empty((B, K, 8, 2), dtype=torch.float32, device=dev))
                if cached:
                    self._out[okey] = out
            rc = self._lib.rtm3d_decode_box3d(det.flat.data_ptr(), det.counts.data_ptr(), reg.data_ptr(), _dtype_code(reg),
                                              B, n_classes, H, W, Creg, K, mode, cam.data_ptr(), dim_ref.data_ptr(),
                                              float(depth_ref[0]), float(depth_ref[1]), out["loc"].data_ptr(),
                                              out["dim"].data_ptr(), out["alpha"].data_ptr(), out["rot_y"].data_ptr(),
                                              out["corners2d"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_decode_box3d")
        out["_keepalive"] = (cam, dim_ref)
        return out


class HostDecodeSession:
    """End-to-end path for HOST-resident head outputs (what a caller outside the training process has): page-locked
    host tensors in, page-locked host tensors out, through ``rtm3d_decode_fused_host`` (main + keypoint branch) or
    ``rtm3d_decode_main_host`` (main branch only).

    Per step the heat-maps (main, and the keypoint heat-map when given) cross PCIe once; the regression maps stay in
    host memory and only the K*(2V+2) (+ Cv*K*2) scalars the decode needs are read from them by the GPU (zero-copy).
    Buffers are allocated once; ``run`` enqueues everything on the current stream and returns the host result buffers,
    valid after the stream is synchronised (``run(..., sync=True)`` does that)."""

    def __init__(self, dec: HeatmapDecoder, B, C, H, W, n_vert=8, kpt_channels=0, dtype=torch.float32, device="cuda:0"):
        self.dec, self.shape, self.V, self.Cv, self.dtype = dec, (B, C, H, W), n_vert, kpt_channels, dtype
        self.dev = torch.device(device)
        K = dec.topk
        d = lambda *sh, dt=torch.float32: torch.empty(sh, dtype=dt, device=self.dev)
        h = lambda *sh, dt=torch.float32: torch.empty(sh, dtype=dt, pin_memory=True)
        self.dev_hm = d(B, C, H, W, dt=dtype)
        self.det = PackedDetections(cls=d(B, K, dt=torch.int64), score=d(B, K), proj=d(B, K, 2), verts=d(B, K, n_vert, 2),
                                    bbox=d(B, K, 4), flat=d(B, K, dt=torch.int32), counts=d(B, dt=torch.int32))
        self.det_host = PackedDetections(cls=h(B, K, dt=torch.int64), score=h(B, K), proj=h(B, K, 2), verts=h(B, K, n_vert, 2),
                                         bbox=h(B, K, 4), flat=h(B, K, dt=torch.int32), counts=h(B, dt=torch.int32))
        self.cand = self.grp = self.grp_host = self.dev_kpt = None
        if kpt_channels:
            Cv = kpt_channels
            self.dev_kpt = d(B, Cv, H, W, dt=dtype)
            self.cand = KeypointCandidates(score=d(B, Cv, K), xy=d(B, Cv, K, 2), flat=d(B, Cv, K, dt=torch.int32))
            self.grp = GroupedKeypoints(kpt_proj=d(B, K, Cv, 2), kpt_score=d(B, K, Cv), kpt_j=d(B, K, Cv, dt=torch.int32),
                                        verts=d(B, K, Cv, 2))
            self.grp_host = GroupedKeypoints(kpt_proj=h(B, K, Cv, 2), kpt_score=h(B, K, Cv),
                                             kpt_j=h(B, K, Cv, dt=torch.int32), verts=h(B, K, Cv, 2))

    def h2d_bytes(self) -> int:
        n = self.dev_hm.numel() * self.dev_hm.element_size()
        if self.dev_kpt is not None:
            n += self.dev_kpt.numel() * self.dev_kpt.element_size()
        return n

    def d2h_bytes(self) -> int:
        t = [self.det_host.cls, self.det_host.score, self.det_host.proj, self.det_host.verts, self.det_host.bbox,
             self.det_host.flat, self.det_host.counts]
        if self.grp_host is not None:
            t += [self.grp_host.kpt_proj, self.grp_host.kpt_score, self.grp_host.kpt_j, self.grp_host.verts]
        return sum(x.numel() * x.element_size() for x in t)

    def run(self, pred_logits_host, kpt_host=None, sync=True):
        dec, lib = self.dec, self.dec._lib
        B, C, H, W = self.shape
        K, V = dec.topk, self.V
        main, off, off2, voff2 = pred_logits_host
        for t in (main, off, off2, voff2) + ((kpt_host,) if kpt_host is not None else ()):
            if t.is_cuda or not t.is_pinned() or not t.is_contiguous():
                raise ValueError("HostDecodeSession.run expects contiguous page-locked host tensors")
        if kpt_host is not None:
            return self._run_fused(pred_logits_host, kpt_host, sync)
        dt = _dtype_code(main)
        with torch.cuda.device(self.dev):
            ws, stream = dec._workspace(self.dev, B, C, H, W)
            p, ph = self.det, self.det_host
            _native.check(lib.rtm3d_decode_main_host(
                main.data_ptr(), off.data_ptr(), off2.data_ptr(), dt, B, C, H, W, V, K, dec.score_thresh, dec.down_sample,
                self.dev_hm.data_ptr(), p.cls.data_ptr(), p.score.data_ptr(), p.proj.data_ptr(), p.verts.data_ptr(),
                p.bbox.data_ptr(), p.flat.data_ptr(), p.counts.data_ptr(),
                ph.cls.data_ptr(), ph.score.data_ptr(), ph.proj.data_ptr(), ph.verts.data_ptr(), ph.bbox.data_ptr(),
                ph.flat.data_ptr(), ph.counts.data_ptr(), ws.data_ptr(), ws.numel(), dec.flags, stream),
                "rtm3d_decode_main_host")
            if sync:
                torch.cuda.current_stream(self.dev).synchronize()
        return self.det_host, self.grp_host

    def _run_fused(self, pred_logits_host, kpt_host, sync):
        """Main + keypoint branch: rtm3d_decode_fused_host (two H2D copies, one streaming launch, one post kernel), D2H."""
        dec, lib = self.dec, self.dec._lib
        B, C, H, W = self.shape
        K, V, Cv = dec.topk, self.V, self.Cv
        main, off, off2, voff2 = pred_logits_host
        dt = _dtype_code(main)
        with torch.cuda.device(self.dev):
            ws, stream = dec._workspace(self.dev, B, C + Cv, H, W)
            p, c, g = self.det, self.cand, self.grp
            _native.check(lib.rtm3d_decode_fused_host(
                main.data_ptr(), off.data_ptr(), off2.data_ptr(), kpt_host.data_ptr(), voff2.data_ptr(), dt, B, C, Cv, H, W, V, K,
                dec.score_thresh, dec.down_sample, self.dev_hm.data_ptr(), self.dev_kpt.data_ptr(),
                p.cls.data_ptr(), p.score.data_ptr(), p.proj.data_ptr(), p.verts.data_ptr(), p.bbox.data_ptr(), p.flat.data_ptr(),
                p.counts.data_ptr(), c.score.data_ptr(), c.xy.data_ptr(), c.flat.data_ptr(),
                g.kpt_proj.data_ptr(), g.kpt_score.data_ptr(), g.kpt_j.data_ptr(), g.verts.data_ptr(),
                ws.data_ptr(), ws.numel(), dec.flags, stream), "rtm3d_decode_fused_host")
            for name in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
                getattr(self.det_host, name).copy_(getattr(p, name), non_blocking=True)
            for name in ("kpt_proj", "kpt_score", "kpt_j", "verts"):
                getattr(self.grp_host, name).copy_(getattr(g, name), non_blocking=True)
            if sync:
                torch.cuda.current_stream(self.dev).synchronize()
        return self.det_host, self.grp_host


def decoder_from_config(config) -> HeatmapDecoder:
    """Build from the reference's config node (only the three scalars the decoder reads)."""
    return HeatmapDecoder(config.DETECTOR.SCORE_THRESH, config.DETECTOR.TOPK_CANDIDATES, config.MODEL.DOWN_SAMPLE)

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

    WORDS = 1 + 1 + 2 + 16 + 4 + 1  # per-detection 32-bit words of the wire format used by the NCCL gather

    def to_wire(self) -> torch.Tensor:
        """[B,K,W] int32 bit-pattern tensor (cls as i32 | score | proj | verts | bbox | flat) for a single collective."""
        B, K = self.score.shape
        parts = [self.cls.to(torch.int32).view(B, K, 1), self.score.view(torch.int32).view(B, K, 1),
                 self.proj.view(torch.int32).view(B, K, 2), self.verts.reshape(B, K, -1).view(torch.int32),
                 self.bbox.view(torch.int32).view(B, K, 4), self.flat.view(B, K, 1)]
        return torch.cat(parts, dim=-1)

    @staticmethod
    def from_wire(wire: torch.Tensor, counts: torch.Tensor) -> "PackedDetections":
        B, K, Wd = wire.shape
        V = (Wd - 9) // 2
        f = wire.view(torch.float32)
        return PackedDetections(cls=wire[..., 0].to(torch.int64), score=f[..., 1].contiguous(),
                                proj=f[..., 2:4].contiguous(), verts=f[..., 4:4 + 2 * V].reshape(B, K, V, 2).contiguous(),
                                bbox=f[..., 4 + 2 * V:8 + 2 * V].contiguous(), flat=wire[..., 8 + 2 * V].contiguous(),
                                counts=counts)


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
                 cluster: int = 0):
        if not (score_thresh >= 0):
            raise ValueError("score_thresh must be >= 0: zero-score fillers of the peak map could pass a negative threshold")
        if not (1 <= int(topk) <= 1024):
            raise ValueError("topk must be in [1, 1024]")
        self.score_thresh = float(score_thresh)
        self.topk = int(topk)
        self.down_sample = float(down_sample)
        if cluster not in (0, 1, 2, 4, 8):
            raise ValueError("cluster (CTAs per image of the streaming kernel) must be 0 (auto), 1, 2, 4 or 8")
        self.flags = (_native.FLAG_FORCE_GENERIC if force_generic else 0) | (cluster << 8)
        self._lib = _native.lib()
        self._ws = {}

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
            out = PackedDetections(
                cls=torch.empty((B, K), dtype=torch.int64, device=dev),
                score=torch.empty((B, K), dtype=torch.float32, device=dev),
                proj=torch.empty((B, K, 2), dtype=torch.float32, device=dev),
                verts=torch.empty((B, K, V, 2), dtype=torch.float32, device=dev),
                bbox=torch.empty((B, K, 4), dtype=torch.float32, device=dev),
                flat=torch.empty((B, K), dtype=torch.int32, device=dev),
                counts=torch.empty((B,), dtype=torch.int32, device=dev))
            rc = self._lib.rtm3d_decode_main(
                main.data_ptr(), off.data_ptr(), off2.data_ptr(), dt, B, C, H, W, V, K,
                self.score_thresh, self.down_sample,
                out.cls.data_ptr(), out.score.data_ptr(), out.proj.data_ptr(), out.verts.data_ptr(),
                out.bbox.data_ptr(), out.flat.data_ptr(), out.counts.data_ptr(),
                ws.data_ptr(), ws.numel(), self.flags, stream)
        _native.check(rc, "rtm3d_decode_main")
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

    def decode_with_keypoints(self, pred_logits, kpt_logits):
        """Tier A + Tier B as the commented wiring of models/model.py:45-62,68-69 describes.  Returns
        (PackedDetections, KeypointCandidates, GroupedKeypoints), all asynchronous."""
        det = self.decode_packed(pred_logits)
        cand = self.decode_keypoints(kpt_logits, pred_logits[3])
        return det, cand, self.group_keypoints(det, cand, pred_logits)

    # ------------------------------------------------------------------ Tier C
    def decode_box3d(self, det: PackedDetections, reg: torch.Tensor, cam: torch.Tensor, dim_ref: torch.Tensor,
                     n_classes: int, multibin: bool = False, sigmoid_subpixel: bool = False,
                     depth_ref=(28.01, 16.32)):
        """Closed-form 3D recovery at the Tier A peaks (NOT in the reference; spec: oracle/box3d_ref.py)."""
        _check_map(reg, "regression map")
        B, Creg, H, W = reg.shape
        K = det.score.shape[1]
        dev = reg.device
        cam = cam.to(device=dev, dtype=torch.float32).contiguous().view(B, 9)
        dim_ref = dim_ref.to(device=dev, dtype=torch.float32).contiguous().view(n_classes, 3)
        mode = (1 if multibin else 0) | (2 if sigmoid_subpixel else 0)
        with torch.cuda.device(dev):
            out = dict(loc=torch.empty((B, K, 3), dtype=torch.float32, device=dev),
                       dim=torch.empty((B, K, 3), dtype=torch.float32, device=dev),
                       alpha=torch.empty((B, K), dtype=torch.float32, device=dev),
                       rot_y=torch.empty((B, K), dtype=torch.float32, device=dev),
                       corners2d=torch.empty((B, K, 8, 2), dtype=torch.float32, device=dev))
            rc = self._lib.rtm3d_decode_box3d(det.flat.data_ptr(), det.counts.data_ptr(), reg.data_ptr(), _dtype_code(reg),
                                              B, n_classes, H, W, Creg, K, mode, cam.data_ptr(), dim_ref.data_ptr(),
                                              float(depth_ref[0]), float(depth_ref[1]), out["loc"].data_ptr(),
                                              out["dim"].data_ptr(), out["alpha"].data_ptr(), out["rot_y"].data_ptr(),
                                              out["corners2d"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "rtm3d_decode_box3d")
        out["_keepalive"] = (cam, dim_ref)
        return out


def decoder_from_config(config) -> HeatmapDecoder:
    """Build from the reference's config node (only the three scalars the decoder reads)."""
    return HeatmapDecoder(config.DETECTOR.SCORE_THRESH, config.DETECTOR.TOPK_CANDIDATES, config.MODEL.DOWN_SAMPLE)

"""Seeded synthetic head outputs.  No dataset or checkpoint is available offline, so every
benchmark and parity case runs on maps of the head's shape (models/nets/header.py:40-46: main_kf [B,C,H,W],
offset_fr_main [B,16,H,W], main_offset [B,2,H,W], vertex_offset [B,2,H,W], fp32 NCHW) drawn from a CPU
``torch.Generator`` so the same seed gives the same bytes on every box.

kinds
  randn        N(0,1) logits: the stated stand-in for random-init heads (~11 % of pixels are 3x3 peaks)
  trained      randn*3 - 6: sparse, "trained-like" heat-map
  quant        logits rounded to multiples of 1/8 in [-4,4]: tie-heavy (exact-score tie groups everywhere)
  saturate     randn*12: |x| > 17 collapses to 0.0 / 1.0 in fp32 sigmoid
  plateau      piecewise-constant 4x4 blocks: every pixel of a block is a peak (equal neighbours are all kept)
  empty        randn - 12: nothing passes a 0.4 threshold
  few          all -10 except a handful of isolated peaks (fewer than K candidates)
  border       peaks planted on the image border and corners
"""
from __future__ import annotations

import torch

KINDS = ("randn", "trained", "quant", "saturate", "plateau", "empty", "few", "border")


def heat_map(kind: str, shape, gen: torch.Generator) -> torch.Tensor:
    B, C, H, W = shape
    r = torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind == "randn":
        return r
    if kind == "trained":
        return r * 3 - 6
    if kind == "quant":
        return (r * 8).round().clamp_(-32, 32) / 8
    if kind == "saturate":
        return r * 12
    if kind == "plateau":
        hb, wb = (H + 3) // 4, (W + 3) // 4
        blocks = torch.randn((B, C, hb, wb), generator=gen, dtype=torch.float32)
        up = blocks.repeat_interleave(4, dim=2).repeat_interleave(4, dim=3)
        return up[:, :, :H, :W].contiguous()
    if kind == "empty":
        return r - 12
    if kind == "few":
        x = torch.full(shape, -10.0)
        n = 7
        for b in range(B):
            cs = torch.randint(0, C, (n,), generator=gen)
            ys = torch.randint(0, H, (n,), generator=gen)
            xs = torch.randint(0, W, (n,), generator=gen)
            x[b, cs, ys, xs] = torch.rand(n, generator=gen) * 4
        return x
    if kind == "border":
        x = r * 0.5 - 3
        x[:, :, 0, :] += 5 * (torch.rand((B, C, W), generator=gen) > 0.7)
        x[:, :, -1, :] += 5 * (torch.rand((B, C, W), generator=gen) > 0.7)
        x[:, :, :, 0] += 5 * (torch.rand((B, C, H), generator=gen) > 0.7)
        x[:, :, :, -1] += 5 * (torch.rand((B, C, H), generator=gen) > 0.7)
        x[:, :, 0, 0] = 6
        x[:, :, -1, -1] = 6
        return x
    raise ValueError(f"unknown kind {kind!r}; choose from {KINDS}")


def head_outputs(B: int, C: int, H: int, W: int, seed: int, kind: str = "randn", n_vert: int = 8,
                 kpt_channels: int = 0, kpt_kind: str | None = None):
    """Returns ([main, off16, off2, voff2], kpt or None), CPU fp32 contiguous."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    main = heat_map(kind, (B, C, H, W), gen)
    off16 = torch.randn((B, 2 * n_vert, H, W), generator=gen, dtype=torch.float32) * 4
    off2 = torch.randn((B, 2, H, W), generator=gen, dtype=torch.float32)
    voff2 = torch.randn((B, 2, H, W), generator=gen, dtype=torch.float32)
    kpt = None
    if kpt_channels:
        kpt = heat_map(kpt_kind or kind, (B, kpt_channels, H, W), gen)
    return [main, off16, off2, voff2], kpt


# BASELINE.json configs (index -> workload).  96x320 = 1280x384 input / DOWN_SAMPLE 4.
WORKLOADS = {
    "cfg1": dict(B=1, C=3, H=96, W=320, K=50, kpt=9, note="RTM3D ResNet-18 batch 1 (reference's CPU-runnable case)"),
    "cfg2": dict(B=32, C=3, H=96, W=320, K=50, kpt=9, note="RTM3D DLA-34 head outputs batch 32"),
    "cfg3": dict(B=64, C=3, H=96, W=320, K=100, kpt=0, note="SMOKE-style centre-keypoint decode batch 64"),
    "cfg4": dict(B=256, C=3, H=96, W=320, K=100, kpt=0, note="RTM3D ResNet-18 synthetic head outputs batch 256"),
    "cfg5": dict(B=128, C=3, H=192, W=640, K=100, kpt=9, note="high-res 2560x768 input batch 128"),
}
SCORE_THRESH = 0.4
DOWN_SAMPLE = 4.0

"""Image sharding of the decode path across the GPUs of one box.

Images are independent (the reference loops ``for i in range(Bs)``, models/model.py:40; under its DDP eval every rank
already decodes only its own shard, train_multi_gpu.py:136-156), so the batch is cut into contiguous blocks -- rank r of G
owns images [r*B/G, (r+1)*B/G) -- and nothing crosses GPUs until the end, where ONE collective gathers the fixed-size
detections: ``all_gather_into_tensor`` of the [B/G, K*25 + 1] 32-bit wire rows (100 B/detection; the per-image count rides
in the last word of each image's row).
Plumbing only (torch.distributed over NCCL on the GPUs, gloo in the CPU tests); no arithmetic of the path lives here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .decoder import PackedDetections


def shard_range(n_images: int, rank: int, world: int):
    """Contiguous block of rank ``rank``: sizes differ by at most one image, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0,{world})")
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(det: PackedDetections, group=None, out=None, wire=None) -> PackedDetections:
    """All-gather the per-rank detections (equal shard sizes) into the global batch order: ONE collective on the packed
    wire rows (counts ride in the last word of each image's row).  ``out`` / ``wire``: optional [G*B, words] / [B, words]
    int32 buffers to reuse between steps."""
    world = dist.get_world_size(group)
    wire = det.to_wire(wire)
    if out is None:
        out = torch.empty((world * wire.shape[0], wire.shape[1]), dtype=wire.dtype, device=wire.device)
    dist.all_gather_into_tensor(out, wire, group=group)
    return PackedDetections.from_wire(out, det.score.shape[1])

"""The N>1 path on CPU: image sharding + the one collective (all-gather of the fixed-size detections), world_size 2,
gloo backend.  The wire format and the gather are exactly what bench.py runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtm3d_b200.decoder import PackedDetections
from rtm3d_b200.sharding import gather_detections, shard_range


def _fake_detections(B, K, V, seed):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(0, K + 1, (B,), generator=g, dtype=torch.int32)
    det = PackedDetections(cls=torch.randint(0, 3, (B, K), generator=g, dtype=torch.int64),
                           score=torch.rand((B, K), generator=g), proj=torch.randn((B, K, 2), generator=g),
                           verts=torch.randn((B, K, V, 2), generator=g), bbox=torch.randn((B, K, 4), generator=g),
                           flat=torch.randint(0, 92160, (B, K), generator=g, dtype=torch.int32), counts=counts)
    for b in range(B):                       # rows >= counts[b] are the kernels' padding
        n = int(counts[b])
        det.cls[b, n:] = -1
        det.flat[b, n:] = -1
        for t in (det.score, det.proj, det.verts, det.bbox):
            t[b, n:] = 0
    return det


def test_shard_range_partitions_the_batch():
    for n in (1, 7, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_wire_roundtrip_is_bit_exact():
    det = _fake_detections(5, 20, 8, seed=3)
    back = PackedDetections.from_wire(det.to_wire(), 20)
    for f in ("cls", "score", "proj", "verts", "bbox", "flat", "counts"):
        assert torch.equal(getattr(det, f), getattr(back, f)), f
    assert det.to_wire().shape[-1] == 20 * PackedDetections.WORDS + 1 and PackedDetections.WORDS == 25


def _worker(rank, world, port, B, K, V, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _fake_detections(B, K, V, seed=11)
        lo, hi = shard_range(B, rank, world)
        mine = PackedDetections(**{f: getattr(full, f)[lo:hi].contiguous() for f in
                                   ("cls", "score", "proj", "verts", "bbox", "flat", "counts")})
        got = gather_detections(mine)
        ok = all(torch.equal(getattr(got, f), getattr(full, f)) for f in
                 ("cls", "score", "proj", "verts", "bbox", "flat", "counts"))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, 8, 20, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]

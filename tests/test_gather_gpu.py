"""The fused gather of rtm3d_decode_fused_gather (the path's one exchange): the select + post kernel stores every
image's wire row into each peer's gather buffer and raises an arrival flag.  One GPU stands in for the peers here (two gather
buffers in the same device memory, the two "ranks" run one after the other); the N-GPU run over NVLink is bench.py --verify."""
import ctypes

import pytest
import torch

import bench
from rtm3d_b200 import HeatmapDecoder, PackedDetections, _native

pytestmark = pytest.mark.gpu


def _rows(B, K):
    return K * PackedDetections.WORDS + 1


@pytest.mark.parametrize("n_peers", [1, 2, 3])
@pytest.mark.parametrize("shape", [(5, 96, 320, 100), (3, 48, 160, 40), (2, 96, 320, 7)])
def test_fused_gather_rows_equal_pack_wire(shape, n_peers):
    B, H, W, K = shape
    dev = torch.device("cuda:0")
    w = dict(B=B, C=3, H=H, W=W, K=K, kpt=9)
    per = _rows(B, K)
    words = n_peers * B * per + n_peers
    # odd word offset between the buffers: the row stores must cope with a base that is only 4-byte aligned
    arena = torch.full((n_peers * (words + 3) + 4,), -1, dtype=torch.int32, device=dev)
    bufs = [arena[1 + p * (words + 3): 1 + p * (words + 3) + words] for p in range(n_peers)]
    for b in bufs:
        b[n_peers * B * per:] = 0                             # arrival flags start below every step id
    peers = (ctypes.c_void_p * n_peers)(*[b.data_ptr() for b in bufs])
    dec = HeatmapDecoder(0.4, K, 4.0)
    want = []
    for rank in range(n_peers):
        logits, kpt, _ = bench.make_inputs(torch, w, dev, 77 + rank, kind="trained" if rank == 1 else "randn")
        det, cand, grp = dec.decode_with_keypoints(logits, kpt, gather=(peers, n_peers, rank, 5))
        det_plain, cand_plain, grp_plain = HeatmapDecoder(0.4, K, 4.0).decode_with_keypoints(logits, kpt)
        for a, b in zip((det.score, det.flat, det.counts, det.bbox, det.verts, cand.score, cand.flat, grp.kpt_j, grp.kpt_proj),
                        (det_plain.score, det_plain.flat, det_plain.counts, det_plain.bbox, det_plain.verts, cand_plain.score, cand_plain.flat,
                         grp_plain.kpt_j, grp_plain.kpt_proj)):
            assert torch.equal(a, b), "the gather variant changed the local results"
        want.append(det_plain.to_wire().clone())
    lib = _native.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    for b in bufs:
        _native.check(lib.rtm3d_wait_gather(b.data_ptr(), B, K, 8, n_peers, 5, stream), "rtm3d_wait_gather")
    torch.cuda.synchronize()
    full = torch.cat(want, dim=0)
    for p, b in enumerate(bufs):
        got = b[:n_peers * B * per].view(n_peers * B, per)
        counts = full[:, -1]
        assert torch.equal(got[:, -1], counts), f"peer {p}: counts column differs"
        # rows are defined up to their image's count (what from_wire consumers read); pack_wire zero-fills behind it
        valid = (torch.arange(K, device=dev)[None, :] < counts[:, None]).repeat_interleave(PackedDetections.WORDS, dim=1)
        assert torch.equal(got[:, :-1][valid], full[:, :-1][valid]), f"peer {p}: wire rows differ from rtm3d_pack_wire's"
        assert torch.equal(b[n_peers * B * per:], torch.full((n_peers,), 5, dtype=torch.int32, device=dev)), "arrival flags"
    # nothing outside the buffers was touched
    mask = torch.ones_like(arena, dtype=torch.bool)
    for p in range(n_peers):
        mask[1 + p * (words + 3): 1 + p * (words + 3) + words] = False
    assert bool((arena[mask] == -1).all()), "stores outside the gather buffers"


@pytest.mark.parametrize("n_peers", [2, 3])
def test_deferred_gather_pushes_previous_batch(n_peers):
    """rtm3d_decode_fused_gather_deferred over two slots and three batches: a launch keeps its rows on its own rank and pushes
    the previous batch; rtm3d_push_gather flushes the last one.  Afterwards every rank holds every rank's rows of the last
    two batches, and the flags carry their ids."""
    B, H, W, K = 3, 48, 160, 30
    dev = torch.device("cuda:0")
    w = dict(B=B, C=3, H=H, W=W, K=K, kpt=9)
    per = _rows(B, K)
    words = n_peers * B * per + n_peers
    slot_words = (words + 3) // 4 * 4 + 4                         # the same 16-byte phase for every slot and rank, as symmetric memory gives
    bufs = [torch.full((2 * slot_words + 4,), -1, dtype=torch.int32, device=dev) for _ in range(n_peers)]
    base = [b[(16 - b.data_ptr() % 16) % 16 // 4 + 1:] for b in bufs]          # deliberately 4 bytes past a 16-byte boundary
    slots = [[base[r][s * slot_words: s * slot_words + words] for s in range(2)] for r in range(n_peers)]
    for r in range(n_peers):
        for s_ in range(2):
            slots[r][s_][n_peers * B * per:] = 0
    peers = [(ctypes.c_void_p * n_peers)(*[slots[r][s_].data_ptr() for r in range(n_peers)]) for s_ in range(2)]
    decs = [HeatmapDecoder(0.4, K, 4.0) for _ in range(n_peers)]
    lib = _native.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    want = {}
    n_steps = 3
    for step in range(n_steps):
        for rank in range(n_peers):
            logits, kpt, _ = bench.make_inputs(torch, w, dev, 500 + 10 * step + rank, kind="trained" if (step + rank) % 2 else "randn")
            prev = peers[(step - 1) & 1] if step > 0 else None
            det, _, _ = decs[rank].decode_with_keypoints(logits, kpt, gather=(peers[step & 1], prev, n_peers, rank, step))    # id of batch s = s + 1
            want[(step, rank)] = det.to_wire().clone()
    for rank in range(n_peers):
        _native.check(lib.rtm3d_push_gather(peers[(n_steps - 1) & 1], n_peers, rank, B, K, 8, n_steps, stream), "rtm3d_push_gather")
    for r in range(n_peers):
        for step in (n_steps - 2, n_steps - 1):
            _native.check(lib.rtm3d_wait_gather(slots[r][step & 1].data_ptr(), B, K, 8, n_peers, step + 1, stream), "rtm3d_wait_gather")
    torch.cuda.synchronize()
    for step in (n_steps - 2, n_steps - 1):
        full = torch.cat([want[(step, rank)] for rank in range(n_peers)], dim=0)
        counts = full[:, -1]
        valid = (torch.arange(K, device=dev)[None, :] < counts[:, None]).repeat_interleave(PackedDetections.WORDS, dim=1)
        for r in range(n_peers):
            got = slots[r][step & 1][:n_peers * B * per].view(n_peers * B, per)
            assert torch.equal(got[:, -1], counts), f"rank {r} batch {step}: counts"
            assert torch.equal(got[:, :-1][valid], full[:, :-1][valid]), f"rank {r} batch {step}: rows"
            assert torch.equal(slots[r][step & 1][n_peers * B * per:], torch.full((n_peers,), step + 1, dtype=torch.int32, device=dev)), "flags"

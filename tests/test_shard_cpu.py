"""N>1 path on CPU: world_size-2 gloo processes shard frames by frame_idx % world and all-gather
per-frame summaries.  No kernel runs here (the hot path has no CPU implementation):
the per-frame "fused maps" are seeded random tensors, identical to what a single process computes,
so the gathered result must equal the single-process result row for row."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gencomm_b200 import shard

N_FRAMES, C, H, W = 7, 4, 6, 10      # odd frame count -> ragged ownership (rank 0: 4 frames, rank 1: 3)


def fake_fused(frame):
    g = torch.Generator().manual_seed(1234 + frame)
    return torch.randn(C, H, W, generator=g)


def single_process():
    local = {f: shard.frame_summary(fake_fused(f)) + ([float(f), 2.0 * f],) for f in range(N_FRAMES)}
    return shard.gather_summaries(local, N_FRAMES, torch.device("cpu"), n_timings=2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.frames_for_rank(N_FRAMES, rank, world)
        local = {f: shard.frame_summary(fake_fused(f)) + ([float(f), 2.0 * f],) for f in mine}
        v, i, t, owner = shard.gather_summaries(local, N_FRAMES, torch.device("cpu"), n_timings=2)
        out[rank] = (v, i, t, owner, mine)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_frames_for_rank_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(f for r in range(world) for f in shard.frames_for_rank(11, r, world))
        assert seen == list(range(11))
    assert shard.frames_for_rank(7, 1, 2) == [1, 3, 5]


def test_world2_gloo_gather_matches_single_process():
    ref = single_process()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert sorted(out.keys()) == [0, 1]
    assert out[0][4] == [0, 2, 4, 6] and out[1][4] == [1, 3, 5]
    for rank in (0, 1):
        v, i, t, owner, _ = out[rank]
        assert torch.equal(v, ref[0]) and torch.equal(i, ref[1]) and torch.equal(t, ref[2])
        assert owner.tolist() == [f % 2 for f in range(N_FRAMES)]


def test_missing_frame_is_reported():
    local = {f: shard.frame_summary(fake_fused(f)) + ([0.0, 0.0],) for f in range(N_FRAMES - 1)}
    try:
        shard.gather_summaries(local, N_FRAMES, torch.device("cpu"), n_timings=2)
    except RuntimeError as e:
        assert "not processed" in str(e)
    else:
        raise AssertionError("missing frame went unnoticed")


def fake_detections(frame):
    """Ragged per-frame detections incl. an empty frame (post_process returns (None, None) there)."""
    k = (frame * 3) % 5
    if k == 0:
        return None, None
    g = torch.Generator().manual_seed(99 + frame)
    return torch.randn(k, 8, 3, generator=g), torch.rand(k, generator=g)


def _det_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = {f: fake_detections(f) for f in shard.frames_for_rank(N_FRAMES, rank, world)}
        out[rank] = shard.gather_detections(local, N_FRAMES, torch.device("cpu"), k_max=6)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_detection_gather():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_det_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    single = shard.gather_detections({f: fake_detections(f) for f in range(N_FRAMES)}, N_FRAMES, torch.device("cpu"), k_max=6)
    for rank in (0, 1):
        got = out[rank]
        assert sorted(got.keys()) == list(range(N_FRAMES))
        for f in range(N_FRAMES):
            boxes, scores = fake_detections(f)
            k = 0 if boxes is None else boxes.shape[0]
            assert got[f][0].shape == (k, 8, 3) and got[f][1].shape == (k,)
            if k:
                assert torch.equal(got[f][0], boxes) and torch.equal(got[f][1], scores)
            assert torch.equal(got[f][0], single[f][0]) and torch.equal(got[f][1], single[f][1])


def fake_padded(rank, F=3, top=1000):
    """What gc_postprocess hands over for a rank's F frames: padded boxes / scores with garbage beyond the count."""
    g = torch.Generator().manual_seed(7 + rank)
    counts = torch.tensor([(5 * rank + 3 * f) % 7 for f in range(F)], dtype=torch.int32)   # includes empty frames
    return torch.randn(F, top, 8, 3, generator=g), torch.rand(F, top, generator=g), counts


def _padded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = shard.gather_detections_device(*fake_padded(rank))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_padded_detection_gather():
    """The bench's exchange step: all_gather_into_tensor of the packed padded detections (count | scores | boxes)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_padded_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.equal(out[0], out[1]) and out[0].shape == (6, shard.DET_WIDTH)
    dets = shard.unpack_detections(out[0])
    for rank in (0, 1):
        boxes, scores, counts = fake_padded(rank)
        single = shard.gather_detections_device(boxes, scores, counts)          # world 1: just the packed payload
        assert torch.equal(single, out[0][3 * rank:3 * rank + 3])
        for f in range(3):
            k = int(counts[f])
            b, s = dets[3 * rank + f]
            assert b.shape == (k, 8, 3) and torch.equal(b, boxes[f, :k]) and torch.equal(s, scores[f, :k])
        assert float(single[:, 1:1001].sum()) == float(sum(scores[f, :int(counts[f])].sum() for f in range(3)))  # padding zeroed

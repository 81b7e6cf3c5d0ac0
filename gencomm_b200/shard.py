"""Frame sharding across the GPUs of one box.

Collaborative frames are independent units (the reference processes them one by one,
tools/inference.py:131-227), so the path shards with NO data-path collective: global frame f runs on
rank ``f % world``.  The only exchange is an end-of-run ``all_gather`` of small per-frame summaries
(``gather_detections``: count + boxes + scores per frame from ``VoxelPostprocessor``; ``gather_summaries``: top-K
responses of the fused map for the head-less bench step, plus per-stage timings), KBs per rank: latency-bound, NVLink bandwidth is irrelevant.

Works with any ``torch.distributed`` backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

TOPK = 16


def frames_for_rank(n_frames, rank, world):
    """Global frame indices owned by `rank` (round-robin, like a DistributedSampler without padding)."""
    return list(range(int(rank), int(n_frames), int(world)))


def frame_summary(fused, k=TOPK):
    """fused [C,H,W] -> (values [k] f32, flat pixel indices [k] i64) of the per-pixel channel maximum."""
    resp = fused.amax(dim=0).flatten()
    v, i = torch.topk(resp, min(k, resp.numel()))
    if v.numel() < k:
        v = torch.cat([v, v.new_full((k - v.numel(),), float("-inf"))])
        i = torch.cat([i, i.new_full((k - i.numel(),), -1)])
    return v.float(), i.long()


def gather_summaries(local, n_frames, device, k=TOPK, n_timings=0):
    """All-gather per-frame summaries.

    local: {global_frame_idx: (values [k], indices [k], timings [n_timings])} for this rank's frames.
    Returns (values [n_frames,k], indices [n_frames,k], timings [n_frames,n_timings], owner [n_frames]) on every
    rank, rows ordered by global frame index.  Ranks may own different numbers of frames (ragged): rows are
    padded to the per-rank maximum and a count is exchanged with them.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    per_rank = (int(n_frames) + world - 1) // world
    width = 1 + k + k + n_timings                      # frame id, values, indices (as f64), timings
    buf = torch.full((per_rank, width), -1.0, dtype=torch.float64, device=device)
    for row, f in enumerate(sorted(local)):
        v, i, t = local[f]
        buf[row, 0] = float(f)
        buf[row, 1:1 + k] = v.to(device=device, dtype=torch.float64)
        buf[row, 1 + k:1 + 2 * k] = i.to(device=device, dtype=torch.float64)
        if n_timings:
            buf[row, 1 + 2 * k:] = torch.as_tensor(t, dtype=torch.float64, device=device)
    if world > 1:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    else:
        parts = [buf]
    values = torch.full((n_frames, k), float("nan"), dtype=torch.float32)
    indices = torch.full((n_frames, k), -1, dtype=torch.int64)
    timings = torch.zeros((n_frames, n_timings), dtype=torch.float64)
    owner = torch.full((n_frames,), -1, dtype=torch.int64)
    for r, p in enumerate(parts):
        p = p.cpu()
        for row in p:
            f = int(row[0])
            if f < 0:
                continue
            if owner[f] != -1:
                raise RuntimeError(f"frame {f} reported by ranks {int(owner[f])} and {r}")
            owner[f] = r
            values[f] = row[1:1 + k].float()
            indices[f] = row[1 + k:1 + 2 * k].long()
            timings[f] = row[1 + 2 * k:]
    if (owner < 0).any():
        raise RuntimeError(f"frames {torch.nonzero(owner < 0).flatten().tolist()} were not processed by any rank")
    expect = torch.arange(n_frames) % world
    if not torch.equal(owner, expect):
        raise RuntimeError("frame ownership does not follow frame_idx % world")
    return values, indices, timings, owner


def gather_detections(local, n_frames, device, k_max=1000):
    """All-gather per-frame detections (``count i32, boxes [k_max,8,3] f32, scores [k_max] f32`` zero
    padded; the NMS caps a frame at top 1000, utils/box_utils.py:941).

    local: {global_frame_idx: (boxes [K,8,3] | None, scores [K] | None)} for this rank's frames (the return value of
    ``VoxelPostprocessor.post_process``).  Returns {frame: (boxes [K,8,3], scores [K])} for ALL frames on every rank
    (empty tensors for frames without detections).  One collective of per_rank * (2 + 25 k_max) floats per rank.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    per_rank = (int(n_frames) + world - 1) // world
    width = 2 + 25 * k_max                              # frame id, count, boxes, scores
    buf = torch.zeros((per_rank, width), dtype=torch.float32, device=device)
    buf[:, 0] = -1.0
    if len(local) > per_rank:
        raise RuntimeError(f"rank owns {len(local)} frames, more than ceil(n_frames / world) = {per_rank}")
    for row, f in enumerate(sorted(local)):
        boxes, scores = local[f]
        k = 0 if boxes is None else int(boxes.shape[0])
        if k > k_max:
            raise ValueError(f"frame {f}: {k} detections exceed k_max = {k_max}")
        buf[row, 0], buf[row, 1] = float(f), float(k)
        if k:
            buf[row, 2:2 + 24 * k] = boxes.to(device=device, dtype=torch.float32).reshape(-1)
            buf[row, 2 + 24 * k_max:2 + 24 * k_max + k] = scores.to(device=device, dtype=torch.float32)
    if world > 1:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    else:
        parts = [buf]
    out = {}
    for r, p in enumerate(parts):
        p = p.cpu()
        for row in p:
            f = int(row[0])
            if f < 0:
                continue
            if f in out:
                raise RuntimeError(f"frame {f} reported twice (second time by rank {r})")
            if f % world != r:
                raise RuntimeError("frame ownership does not follow frame_idx % world")
            k = int(row[1])
            out[f] = (row[2:2 + 24 * k].reshape(k, 8, 3).clone(), row[2 + 24 * k_max:2 + 24 * k_max + k].clone())
    missing = [f for f in range(int(n_frames)) if f not in out]
    if missing:
        raise RuntimeError(f"frames {missing} were not processed by any rank")
    return out


DET_WIDTH = 1 + 1000 + 24 * 1000     # count, scores [1000], boxes [1000,8,3] per frame


def pack_detections(boxes, scores, counts):
    """Device-side payload of the detection all-gather: [F, DET_WIDTH] f32 = count | scores | boxes, rows beyond the count
    zeroed (``gc_postprocess`` leaves them unspecified).  No host sync."""
    F, top = scores.shape
    valid = (torch.arange(top, device=scores.device)[None, :] < counts[:, None]).to(scores.dtype)
    return torch.cat([counts.to(scores.dtype)[:, None], scores * valid, (boxes.reshape(F, top, 24) * valid[:, :, None]).reshape(F, -1)],
                     dim=1)


def gather_detections_device(boxes, scores, counts, out=None):
    """All-gather the padded ``gc_postprocess`` output of this rank's F frames over every rank (the one
    exchange step of the path, NCCL on the GPU box).  boxes [F,1000,8,3], scores [F,1000], counts [F] (device) ->
    payload [world*F, DET_WIDTH] on every rank, rank-major (rank r holds global frames r*F .. r*F+F-1 of the step).
    Asynchronous (the collective is enqueued on the current stream by torch.distributed); unpack with
    ``unpack_detections``."""
    payload = pack_detections(boxes, scores, counts)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return payload
    if out is None:
        out = torch.empty((world * payload.shape[0], payload.shape[1]), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload)
    return out


def unpack_detections(payload, top=1000):
    """payload [n, DET_WIDTH] -> list of (boxes [K,8,3], scores [K]) per frame (host sync: the result is ragged)."""
    p = payload.cpu()
    out = []
    for row in p:
        k = int(row[0])
        out.append((row[1 + top:1 + top + 24 * k].reshape(k, 8, 3).clone(), row[1:1 + k].clone()))
    return out

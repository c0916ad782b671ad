"""Multi-GPU partitioning of the propagation path: independent videos (and their objects) shard
across ranks with NO data-path collective -- one process per GPU, weights replicated, each rank runs its own
`propagate_in_video` loops.  torch.distributed is only used to gather small results (metrics / mask areas)."""
import torch
import torch.distributed as dist


def shard_videos(num_videos, rank, world_size):
    """Round-robin assignment of video indices to ranks (stable, disjoint, covering)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_videos, world_size))


def gather_per_video(local, num_videos, device=None):
    """local: {video_index: float tensor [k]} held by this rank -> dense [num_videos, k] on every rank.
    Uses one all_reduce(SUM) of a zero-initialised table (tiny; not on the per-frame path)."""
    k = len(next(iter(local.values()))) if local else 0
    if dist.is_available() and dist.is_initialized():
        kk = torch.tensor([k], dtype=torch.int64, device=device)
        dist.all_reduce(kk, op=dist.ReduceOp.MAX)
        k = int(kk)
    table = torch.zeros((num_videos, k), dtype=torch.float64, device=device)
    for i, v in local.items():
        table[i] = torch.as_tensor(v, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table


def track_videos(predictor, videos, rank=0, world_size=1, on_frame=None):
    """videos: list of dict(images=[T,3,S,S] normalised, height, width, prompts=[(frame, obj_id, mask)]).
    Runs this rank's share; returns {video_index: [foreground pixel count per frame, summed over objects]}."""
    out = {}
    for vi in shard_videos(len(videos), rank, world_size):
        v = videos[vi]
        st = predictor.init_state(v["images"], v["height"], v["width"])
        for frame, obj_id, mask in v["prompts"]:
            predictor.add_new_mask(st, frame, obj_id, mask)
        areas = []
        for t, ids, logits in predictor.propagate_in_video(st):
            if on_frame is not None:
                on_frame(vi, t, ids, logits)
            areas.append(float((logits > 0).sum()))
        out[vi] = areas
    return out

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
        areas = []  # one device scalar per frame: no host synchronisation inside the propagation loop
        for t, ids, logits in predictor.propagate_in_video(st):
            if on_frame is not None:
                on_frame(vi, t, ids, logits)
            areas.append((logits > 0).sum())
        out[vi] = torch.stack(areas).double().cpu().tolist() if areas else []
    return out


def shard_objects(obj_ids, rank, world_size):
    """Objects of ONE clip are independent too (the optional non-overlap constraint is the only coupling):
    round-robin assignment of a clip's object ids to ranks.  Every rank encodes the clip itself -- the image encoder runs
    beside the tracked frames on its own SM partition, so sharing its features would save nothing on the critical path --
    and tracks only its objects: 4 objects on 4 GPUs run at the one-object frame rate, with no communication."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    ids = list(obj_ids)
    return ids[rank::world_size]


def track_clip_objects(predictor, video, rank=0, world_size=1, on_frame=None):
    """One clip, its objects split across ranks: video = dict(images, height, width, prompts=[(frame, obj_id, mask)]).
    Tracks the prompts whose object id belongs to this rank (shard_objects over the sorted ids); returns
    {obj_id: [foreground pixel count per frame]} for this rank's objects.  Not valid with non_overlap_masks=True."""
    if getattr(predictor, "non_overlap_masks", False):
        raise ValueError("objects coupled by non_overlap_masks cannot be tracked on different ranks")
    all_ids = sorted({obj_id for _, obj_id, _ in video["prompts"]})
    mine = set(shard_objects(all_ids, rank, world_size))
    if not mine:
        return {}
    st = predictor.init_state(video["images"], video["height"], video["width"])
    for frame, obj_id, mask in video["prompts"]:
        if obj_id in mine:
            predictor.add_new_mask(st, frame, obj_id, mask)
    per_frame, ids_seen = [], None
    for t, ids, logits in predictor.propagate_in_video(st):
        if on_frame is not None:
            on_frame(t, ids, logits)
        ids_seen = list(ids)
        per_frame.append((logits > 0).flatten(1).sum(dim=1))
    if not per_frame:
        return {}
    table = torch.stack(per_frame).double().cpu()  # [frames, objects]
    return {oid: table[:, j].tolist() for j, oid in enumerate(ids_seen)}

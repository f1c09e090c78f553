"""Multi-GPU sharding of independent images: image i -> rank i mod G, results returned in input
order, no collective on the data path.  torch.distributed is used for the barrier and the MAX-over-ranks timing only
(backend nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_indices(n_images: int, rank: int, world: int):
    """indices of the images rank `rank` owns (round robin)"""
    return list(range(rank, n_images, world))


def owner_of(image: int, world: int) -> int:
    return image % world


def max_over_ranks(seconds: float, device=None) -> float:
    """the job's time is the slowest rank's time"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_input_order(local_results: dict, n_images: int):
    """all ranks contribute {image index: small python object}; every rank gets the full list in input order"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local_results[i] for i in range(n_images)]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local_results)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(n_images)]


def map_sharded(items, fn):
    """Applies `fn(index, item)` to the items this rank owns (image i -> rank i mod G) and returns the results of ALL items in
    input order on every rank.  `fn` is the per-image work on the rank's own GPU context (e.g. host.Spectral.decompress(...)
    .to_rgb8() digests, file writes, statistics); its result travels through all_gather_object, so keep it small -- pixel
    data stays on the rank that produced it (there is no collective on the data path)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    local = {i: fn(i, items[i]) for i in shard_indices(len(items), rank, world)}
    return gather_in_input_order(local, len(items))


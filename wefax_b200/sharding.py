"""Sharding of independent recordings over ranks (one process per GPU).

A single recording is a 1-GPU problem; batches shard by recording, with no
collective on the data path: each rank decodes its own recordings and only the
results (images + a few integers) are gathered on the host at the end.
Recordings are bucketed by (frames, sample rate, channels) because one
``Decoder.decode`` call takes equal-length recordings, and dealt round-robin
inside each bucket so that every rank gets the same mix of lengths and LPMs.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, Sequence


def assign(keys: Sequence, world_size: int) -> list:
    """``keys[i]`` = bucket key of recording i (e.g. ``(n_frames, rate, channels)``).
    Returns ``world_size`` lists of recording indices."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    buckets = defaultdict(list)
    for i, k in enumerate(keys):
        buckets[k].append(i)
    shards = [[] for _ in range(world_size)]
    nxt = 0
    for k in sorted(buckets, key=repr):
        for i in buckets[k]:
            shards[nxt % world_size].append(i)
            nxt += 1
    return shards


def local_batches(keys: Sequence, rank: int, world_size: int) -> list:
    """This rank's recordings grouped into decode batches: ``[(key, [indices...]), ...]``."""
    mine = assign(keys, world_size)[rank]
    groups = defaultdict(list)
    for i in mine:
        groups[keys[i]].append(i)
    return [(k, groups[k]) for k in sorted(groups, key=repr)]


def decode_sharded(keys: Sequence, decode_batch: Callable, rank: int, world_size: int) -> dict:
    """Run ``decode_batch(key, indices) -> {index: result}`` on this rank's batches."""
    out = {}
    for key, idx in local_batches(keys, rank, world_size):
        out.update(decode_batch(key, idx))
    return out


def gather_results(local: dict, dst: int = 0, group=None):
    """Host gather of per-recording results onto rank ``dst`` (None elsewhere).
    Uses ``torch.distributed`` object gather: results are small host objects (images as
    numpy arrays, ints), this is the only exchange of the multi-GPU path."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for part in bucket:
        overlap = merged.keys() & part.keys()
        if overlap:
            raise RuntimeError(f"recordings decoded twice: {sorted(overlap)[:5]}")
        merged.update(part)
    return merged

"""Multi-GPU sharding of an image batch: images are independent, so the batch is cut into
contiguous per-rank slices, every rank runs its slice on its own GPU, and only the (small) minutiae lists are gathered.
No collective touches the image data."""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of `rank`'s contiguous slice; the first `n_items % world_size` ranks get one extra item."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(max(n_items, 0), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_in_order(local_results: Sequence[Any], n_items: int, group=None) -> List[Any] | None:
    """all_gather_object of the per-rank result lists; rank 0 returns the concatenation in global image order
    (other ranks return None).  Works with any torch.distributed backend (gloo on CPU, nccl on GPU)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_items, world, rank)
    if len(local_results) != hi - lo:
        raise ValueError(f"rank {rank}: {len(local_results)} results for a slice of {hi - lo}")
    parts: List[Any] = [None] * world
    dist.all_gather_object(parts, list(local_results), group=group)
    if rank != 0:
        return None
    out: List[Any] = []
    for r in range(world):
        out.extend(parts[r])
    assert len(out) == n_items
    return out


def run_sharded(images, make_pipeline, group=None, chunk: int = 1480):
    """Run the whole hot path on this rank's slice of `images` ([n,H,W] uint8, identical on every rank or at least
    valid on the rank's slice) and gather the refined minutiae lists on rank 0.  `make_pipeline(H, W, max_batch)`
    builds the per-rank FingerprintPipeline (the caller picks the device)."""
    import torch.distributed as dist
    n = len(images)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    results = []
    if hi > lo:
        H, W = images[lo].shape
        pipe = make_pipeline(H, W, min(chunk, hi - lo))
        for s in range(lo, hi, chunk):
            e = min(s + chunk, hi)
            pipe.run(images[s:e])
            results.extend(pipe.minutiae(i) for i in range(e - s))
    return gather_in_order(results, n, group)


def match_sharded(templates, pairs, make_matcher, group=None, **match_kw):
    """FRR/FAR sweeps over several GPUs: pairs are independent, so every rank uploads the (small)
    template set to its own GPU, matches its contiguous slice of `pairs` and rank 0 gets the result records in pair
    order.  `make_matcher(n_templates, max_minutiae, max_iter)` builds the per-rank MinutiaeMatcher."""
    import numpy as np
    import torch.distributed as dist
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(len(pairs), world, rank)
    local = []
    if hi > lo:
        m = make_matcher(len(templates), max(max((len(t) for t in templates), default=1), 1), int(match_kw.get("ransac_iter", 300)))
        m.set_templates(templates)
        res, _, _ = m.match(pairs[lo:hi], False, **match_kw)
        local = [tuple(r) for r in res.tolist()]
    return gather_in_order(local, len(pairs), group)

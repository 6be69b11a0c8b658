"""Host-side multi-GPU logic on CPU: contiguous sharding + ordered gather over gloo with world_size 2 (no data-path
collective exists; this is the only place torch.distributed touches results)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_biometric_fingerprints_palms_b200.sharding import gather_in_order, shard_bounds


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 7, 8, 1480, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n, world, rank)
    # stand-in for pipe.minutiae(i): ragged per-image lists that encode the global image index
    local = [[{"x": i, "y": k, "type": "ending"} for k in range(i % 4)] for i in range(lo, hi)]
    out = gather_in_order(local, n)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_ordered_gather_world_size_2_gloo():
    n, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(out) == n
    for i, lst in enumerate(out):
        assert len(lst) == i % 4 and all(m["x"] == i for m in lst)

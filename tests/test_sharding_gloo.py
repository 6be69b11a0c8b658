"""Host-side multi-GPU logic on CPU: contiguous sharding + ordered gather over gloo with world_size 2 (no data-path
collective exists; this is the only place torch.distributed touches results)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_biometric_fingerprints_palms_b200.sharding import gather_in_order, match_sharded, shard_bounds


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


class _FakeMatcher:
    """Stands in for MinutiaeMatcher on a box without a GPU: the 'score' of a pair encodes the pair itself."""
    def __init__(self, n_templates, max_minutiae, max_iter):
        self.args = (n_templates, max_minutiae, max_iter)

    def set_templates(self, templates):
        self.n = len(templates)

    def match(self, pairs, want_matches, **kw):
        import numpy as np
        res = np.zeros(len(pairs), dtype=[("final_score", "f8"), ("n_matches", "i4")])
        res["final_score"] = pairs[:, 0] * 1000 + pairs[:, 1] + kw["dist_thresh"] / 100.0
        res["n_matches"] = self.n
        return res, None, None


def _match_worker(rank, world, port, q):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    templates = [np.zeros((5 + i, 7)) for i in range(6)]
    pairs = [(a, b) for a in range(6) for b in range(6) if a < b]
    out = match_sharded(templates, pairs, _FakeMatcher, dist_thresh=15, ransac_iter=120)
    if rank == 0:
        q.put((pairs, out))
    dist.barrier()
    dist.destroy_process_group()


def test_match_sharded_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_match_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    pairs, out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(out) == len(pairs) == 15
    for (a, b), rec in zip(pairs, out):
        assert rec[0] == a * 1000 + b + 0.15 and rec[1] == 6

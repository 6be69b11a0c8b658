"""Sustained streaming run over a large synthetic set (BASELINE.json configs[3]: 1 M 320x240 prints over 1/2/4/8 GPUs).

Images are generated ON THE DEVICE from (seed, global image index) by `fpb_synth_ridge` (counter-based Philox, any index can
be regenerated alone - csrc/k_synth.cu), so no host I/O is on the path; the rank's contiguous slice of the index range
(`sharding.shard_bounds`) is cut into batches that alternate between TWO handles: while the GPU works on batch i (handle A),
the host waits for batch i-1 (handle B), copies its refined lists out in one `fpb_result_block` call and hands them to
`on_batch`, then enqueues batch i+1 on B.  The GPU always has one batch queued; only roi / counts / refined lists
(3.1 KB per image) come back.  No collective on the data path."""
from __future__ import annotations

import time
from typing import Callable, Dict, Optional

import numpy as np

from .sharding import shard_bounds


def batch_plan(lo: int, hi: int, batch: int):
    """[(first_index, n)] covering [lo, hi) in order; the last batch may be short."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    return [(s, min(batch, hi - s)) for s in range(lo, hi, batch)]


def run_stream(total: int, seed: int, make_pipeline: Callable, height: int = 320, width: int = 240, batch: int = 1480,
               rank: int = 0, world: int = 1, on_batch: Optional[Callable] = None, period: float = 0.0,
               noise_sigma: float = 12.0, depth: int = 2) -> Dict:
    """Process images [lo, hi) = this rank's slice of range(total).  `make_pipeline(H, W, max_batch)` builds a handle
    (the caller picks device / stream); `on_batch(first_index, roi, raw_counts, out_counts, refined)` receives every
    batch's results in index order.  Returns counters and the wall time of the loop (the caller synchronises ranks)."""
    lo, hi = shard_bounds(total, world, rank)
    plan = batch_plan(lo, hi, batch)
    stats = {"rank": rank, "first": lo, "stop": hi, "images": hi - lo, "batches": len(plan), "refined": 0, "raw": 0,
             "seconds": 0.0, "host_wait_s": 0.0, "host_gather_s": 0.0}
    if not plan:
        return stats
    pipes = [make_pipeline(height, width, min(batch, hi - lo)) for _ in range(max(1, min(depth, len(plan))))]
    inflight = []                      # (pipe, first_index) in submission order

    def drain():
        p, first = inflight.pop(0)
        t0 = time.perf_counter()
        p.download_refined()           # waits for that handle's stream, D2H of the result block
        t1 = time.perf_counter()
        roi, rc, oc, ref = p.result_block()
        stats["refined"] += int(oc.sum()); stats["raw"] += int(rc.sum())
        if on_batch is not None:
            on_batch(first, roi, rc, oc, ref)
        t2 = time.perf_counter()
        stats["host_wait_s"] += t1 - t0; stats["host_gather_s"] += t2 - t1

    t_begin = time.perf_counter()
    for k, (first, n) in enumerate(plan):
        p = pipes[k % len(pipes)]
        if len(inflight) == len(pipes):
            drain()                    # frees `p` (the oldest in flight)
        p.synth_ridge(seed, first, n, period=period, noise_sigma=noise_sigma)
        p.run_input_async(n)
        inflight.append((p, first))
    while inflight:
        drain()
    stats["seconds"] = time.perf_counter() - t_begin
    for p in pipes:
        p.close()
    return stats

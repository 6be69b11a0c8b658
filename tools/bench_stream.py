#!/usr/bin/env python
"""BASELINE.json configs[3]: 1 M synthetic 320x240 prints sharded over 1/2/4/8 B200 - the sustained streaming run.

  python tools/bench_stream.py --total 1000000                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_stream.py --total 1000000                                      # strong scaling over 8 GPUs

Images are generated on the device from (seed, global index) (fpb_synth_ridge, Philox4x32-10), every rank takes a
contiguous slice of the index range, two handles alternate so the GPU always has a batch queued while the host drains
the previous one (roi / counts / refined lists, 3.1 KB per image).  Wall clock between two barriers, max over ranks.
Also prints the single-batch device-resident throughput of the same GPU(s) for comparison.  One JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1480)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--depth", type=int, default=2)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    from multimodal_biometric_fingerprints_palms_b200.stream import run_stream
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mk = lambda H, W, B: FingerprintPipeline(H, W, max_batch=B, device=local)
    # single-batch reference point: the same generator + run, no host drain
    p = mk(320, 240, a.batch)
    p.synth_ridge(a.seed, 0, a.batch)
    for _ in range(3):
        p.run_input_async(a.batch)
    p.sync(); barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        p.run_input_async(a.batch)
    p.sync()
    single = 10 * a.batch / (time.perf_counter() - t0)
    p.close()
    hist = np.zeros(65, np.int64)

    def on_batch(first, roi, rc, oc, ref):
        hist[:] += np.bincount(oc, minlength=65)[:65]      # the host "consumes" every batch: refined-count histogram

    run_stream(2 * a.batch * world, a.seed, mk, batch=a.batch, rank=rank, world=world, depth=a.depth)     # warm-up
    barrier()
    t0 = time.perf_counter()
    st = run_stream(a.total, a.seed, mk, batch=a.batch, rank=rank, world=world, on_batch=on_batch, depth=a.depth)
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall, -single, st["host_wait_s"], st["host_gather_s"]], dtype=torch.float64, device="cuda")
    c = torch.tensor([st["images"], st["refined"], st["raw"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c)
    if rank == 0:
        wall, single_min = float(t[0]), -float(t[1])
        print(json.dumps({"workload": f"{a.total} device-generated 320x240 prints (BASELINE configs[3]), batch {a.batch}, "
                                      f"{a.depth} handles alternating per GPU", "n_gpus": world, "images": int(c[0]),
                          "seconds": wall, "images_per_s": int(c[0]) / wall, "scaling": "strong",
                          "single_batch_images_per_s_per_gpu": single_min,
                          "sustained_over_single_batch": int(c[0]) / wall / (single_min * world),
                          "refined_minutiae": int(c[1]), "raw_minutiae": int(c[2]),
                          "host_wait_s_max": float(t[2]), "host_gather_s_max": float(t[3]),
                          "refined_count_histogram_rank0": hist.tolist()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

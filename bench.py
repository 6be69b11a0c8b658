#!/usr/bin/env python
"""Throughput of the fingerprint enhance -> minutiae hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                  # CUDA arm (this repo)
  python bench.py --impl reference --gpus N --steps K --warmup W # CPU arm: the reference's algorithm on host cores

One "step" = one pass of K1..K9 over one batch of synthetic 320x240 prints (BASELINE.json configs[1]:
1480 images, PolyU DBII shape) per GPU.  N > 1 runs under torchrun, one rank per GPU, each rank its own
batch (independent images: no data-path collective, weak scaling); the barrier + max-over-ranks timing is
the only use of torch.distributed.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fingerprints/sec enhance->minutiae (240x320)"
UNIT = "images/s"
H, W = 320, 240
BATCH = 1480


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic prints generated on the host and tiled to the batch")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (same OpenCV/SciPy/NumPy calls as the reference; the reference itself is pure
# Python and cannot travel to the GPU box) on all host cores, one image per task
# ------------------------------------------------------------------------------------------------
def _cpu_one(seed):
    import cv2
    cv2.setNumThreads(1)
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    from oracle import ref_pipeline as rp
    img = ridge_image(H, W, seed=seed, period=None)          # synthetic input: generated outside the timed span
    t0 = time.perf_counter()
    res = rp.enhance_to_minutiae(img)
    return time.perf_counter() - t0, len(res["minutiae"])


def cpu_throughput(n_images: int, workers: int):
    """images/s of the CPU port: `workers` processes, each running images back to back (all host cores busy);
    throughput = workers / mean per-image pipeline time, so input generation is not charged to the CPU arm."""
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max_workers=workers) as ex:
        list(ex.map(_cpu_one, range(workers)))                      # warm the pool (imports, page-in)
        t0 = time.perf_counter()
        per = list(ex.map(_cpu_one, range(1000, 1000 + n_images), chunksize=max(1, n_images // (workers * 8))))
        wall = time.perf_counter() - t0
    mean_t = sum(p[0] for p in per) / n_images
    return workers / mean_t, wall, mean_t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(cores * 48, 384)
    vals = []
    for _ in range(args.warmup):
        cpu_throughput(max(cores, 8), cores)
    t_total = 0.0
    for _ in range(args.steps):
        v, wall, _ = cpu_throughput(sample, cores)
        vals.append(v)
        t_total += wall
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{sample}-image bounded sample of the 1480-image 320x240 (PolyU DBII shape) batch per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} synthetic 320x240 prints per step, ProcessPoolExecutor({cores}), cv2.setNumThreads(1)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture_summary():
    """SM-side figures of k_nlm from the committed ncu capture (profiles/), for context next to the HBM roofline."""
    import csv, glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k_nlm_v*_batch*.csv")))
    if not files:
        return None
    try:
        rows = {r["metric"]: r for r in csv.DictReader(open(files[-1]))}
        nimg = int(files[-1].split("batch")[-1].split(".")[0])
        def val(k, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}):
            r = rows[k]
            return float(r["value"]) * scale.get(r["unit"], 1.0)
        return {"file": os.path.relpath(files[-1], ROOT), "images": nimg,
                "dram_bytes_per_image": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / nimg,
                "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": val("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
                "alu_pipe_pct": val("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active")}
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.batch
    # synthetic inputs: `distinct` host-generated prints (per-rank seeds) tiled to the batch
    k = min(args.distinct, n)
    base = synth.ridge_batch(k, H, W, first_seed=10_000 * (rank + 1))
    host = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for i in range(n):
        hv[i] = base[i % k]
    dev = host.to("cuda", non_blocking=False)
    # a dedicated (non-default) torch stream: the library launches on it and the CUDA events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    p = FingerprintPipeline(H, W, max_batch=n, device=local, stream=stream.cuda_stream)
    p.set_profiling(False)          # timed runs: the library's normal mode (two half-batches on two internal streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    for _ in range(args.warmup):
        p.run_device(dev.data_ptr(), n)
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    l0 = p.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nlm_ms = []
    stage_acc = {}
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        p.run_device(dev.data_ptr(), n)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = p.launch_count - l0
    # ---- end to end through the C ABI with HOST buffers: H2D from pinned memory + run + D2H of results
    for _ in range(min(args.warmup, 2)):
        p.run(hv)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p.run(hv)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    # three more runs in profiling mode (single stream, CUDA events around every stage and around the NLM launch):
    # the NLM kernel's own duration for the roofline, and the per-stage breakdown
    p.set_profiling(True)
    p.run_device(dev.data_ptr(), n); p.sync()
    for _ in range(3):
        p.run_device(dev.data_ptr(), n)
        p.sync()
        nlm_ms.append(p.stage_times_ms()["nlm_kernel"])
        for kk, vv in p.stage_times_ms().items():
            stage_acc.setdefault(kk, []).append(vv)
    p.download()
    n_min = sum(len(p.minutiae(i)) for i in range(min(n, 64)))

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_imgs = n * world * args.steps
    value = total_imgs / (ms_max / 1e3)
    e2e_value = total_imgs / (e2e_ms_max / 1e3)
    peak, peak_src = measured_peak_hbm()
    nlm_avg = sum(nlm_ms) / len(nlm_ms)
    alg_bytes = 2.0 * H * W * n                       # NLM: read the plane once, write it once
    achieved = alg_bytes / (nlm_avg / 1e3) / 1e9
    ncu = ncu_capture_summary()
    traffic = ncu["dram_bytes_per_image"] * n if ncu else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{n}-image batch 320x240 uint8 (PolyU DBII shape, BASELINE configs[1]) per GPU, K1..K9 fused run",
                   "global_batch": n * world, "image": [H, W], "parallelism": f"images sharded over {world} GPU(s), no collective",
                   "l2": "per-step working set (inputs 113 MB + intermediates > 5 GB) exceeds the 126 MB L2; no flush needed",
                   "distinct_prints": k},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * H * W),
                "d2h_bytes_per_step": int(n * (16 + 4 + 4 + 64 * 48)), "ms_per_step": e2e_ms_max / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_nlm", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": nlm_avg, "algorithmic_bytes": alg_bytes,
                     # what actually bounds k_nlm (from the committed ncu capture): fraction of the SM issue slots in use
                     "sm_issue_frac": (ncu["issue_active_pct"] / 100.0) if ncu else None,
                     "note": "the path is integer-ALU / shared-memory bound, not HBM bound (SURVEY 8(d), DESIGN 4): "
                             "the HBM fraction is ~0.1 % by construction; the ncu capture under profiles/ gives the SM-side figures",
                     "ncu": ncu},
        "stage_ms": {kk: sum(vv) / len(vv) for kk, vv in stage_acc.items()},
        "clocks": clk,
        "sanity": {"refined_minutiae_first64": n_min},
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or max(cores * 96, 768)
        v, wall, per_img = cpu_throughput(sample, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} synthetic 320x240 prints, ProcessPoolExecutor({cores}), "
                                          f"{wall:.1f} s wall, {1e3 * per_img:.0f} ms/image/core"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

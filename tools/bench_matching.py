"""Throughput of the GPU matcher next to the CPU oracle, on PolyU-DBII-sized synthetic templates:
148 users x 10 impressions; FRR = all 6660 genuine pairs (match_features.py:124-131 parameters), FAR = sampled
impostor pairs (match_features.py:141-149 parameters).  Prints one JSON line per workload.

    python tools/bench_matching.py [--impostor-users 20] [--cpu-pairs 48]
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_worker(args):
    from oracle import ref_matching as rm
    a, b, kw = args
    return rm.match_minutiae_pair(a, b, **kw)["final_score"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=148)
    ap.add_argument("--impostor-users", type=int, default=20)
    ap.add_argument("--cpu-pairs", type=int, default=48)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="frr | far")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import MinutiaeMatcher
    from multimodal_biometric_fingerprints_palms_b200.matching.FRR import genuine_pairs
    from multimodal_biometric_fingerprints_palms_b200.matching.FAR import impostor_pairs

    dataset = {}
    for u in range(a.users):
        base = rm.synthetic_template(1000 + u, n=int(45 + (u * 7) % 16))
        dataset[f"{u:03d}"] = [base] + [rm.perturbed_copy(base, 2000 + 10 * u + k, angle_deg=(k * 5) % 17 - 8.0,
                                                          shift=(k - 4.0, 6.0 - k), jitter=1.0 + 0.1 * k) for k in range(9)]
    random.seed(0)
    work = {"frr": (genuine_pairs(dataset), dict(dist_thresh=30, orient_thresh_deg=30, use_type=True, ransac_iter=300,
                                                 min_inliers=6, stop_inlier_ratio=0.15, cross_check=True)),
            "far": (impostor_pairs(dataset, a.impostor_users), dict(dist_thresh=15, orient_thresh_deg=10, use_type=True,
                                                                   ransac_iter=300, min_inliers=12, stop_inlier_ratio=0.15,
                                                                   cross_check=True))}
    for name, ((tpl, pairs), kw) in work.items():
        if a.only and name != a.only:
            continue
        m = MinutiaeMatcher(len(tpl), 64, 300)
        t0 = time.perf_counter(); m.set_templates(tpl); t_prep = time.perf_counter() - t0
        st = torch.cuda.ExternalStream(m.stream)
        m.upload_pairs(pairs)
        for _ in range(2):
            m.run_device(**kw)
        m.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        with torch.cuda.stream(st):
            ev[0].record()
            for r in range(a.reps):
                m.run_device(**kw)
                ev[r + 1].record()
        m.sync()
        ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(a.reps))
        t0 = time.perf_counter(); res, _, _ = m.match(pairs, False, **kw); t_e2e = time.perf_counter() - t0
        # CPU arm: the oracle (reference arithmetic, seed-order aggregate) on a bounded sample, all host cores
        from concurrent.futures import ProcessPoolExecutor
        sel = np.linspace(0, len(pairs) - 1, a.cpu_pairs).astype(int)
        cores = os.cpu_count() or 1
        if a.no_cpu:
            print(json.dumps({"workload": name, "pairs": int(len(pairs)), "gpu_kernel_ms": ms}))
            m.close()
            continue
        with ProcessPoolExecutor(cores) as ex:
            list(ex.map(cpu_worker, [(tpl[pairs[0][0]], tpl[pairs[0][1]], kw)] * cores))       # warm the workers
            t0 = time.perf_counter()
            cpu = list(ex.map(cpu_worker, [(tpl[pairs[i][0]], tpl[pairs[i][1]], kw) for i in sel]))
            t_cpu = time.perf_counter() - t0
        dev = np.abs(res["final_score"][sel] - np.array(cpu)).max()
        hyp = len(pairs) * kw["ransac_iter"]
        print(json.dumps({"workload": name, "pairs": int(len(pairs)), "templates": len(tpl), "ransac_iter": kw["ransac_iter"],
                          "gpu_kernel_ms": ms, "gpu_pairs_per_s": len(pairs) / ms * 1e3, "gpu_hypotheses_per_s": hyp / ms * 1e3,
                          "gpu_e2e_pairs_per_s": len(pairs) / t_e2e, "template_prep_ms": t_prep * 1e3,
                          "cpu_pairs_per_s": len(sel) / t_cpu, "cpu_cores": cores, "cpu_sample_pairs": len(sel),
                          "max_abs_score_diff_on_sample": float(dev), "mean_score": float(res["final_score"].mean()),
                          "nonzero_scores": int((res["final_score"] > 0).sum())}))
        m.close()


if __name__ == "__main__":
    main()

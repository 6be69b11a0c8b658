"""Mirror of `src/matching/FRR.py`: every genuine pair of every user, matched in ONE batched GPU call.

The reference submits `match_minutiae_pair` per pair to a thread pool (FRR.py:104-118); pair order, parameter
overrides in demo mode and the CSV log are kept."""
from __future__ import annotations

import csv
import os
from itertools import combinations
from typing import Dict, List

import numpy as np

from .match import MinutiaeMatcher, _as_template


def genuine_pairs(dataset: Dict[str, List[np.ndarray]], demo: bool = False):
    """Flat template list + [n,2] index pairs in the reference's task order (FRR.py:78-90)."""
    templates, pairs = [], []
    for _, samples in dataset.items():
        if len(samples) < 2:
            continue
        base = len(templates)
        templates.extend(samples)
        pr = list(combinations(range(len(samples)), 2))
        if demo:
            pr = pr[:3]
        pairs.extend((base + i, base + j) for i, j in pr)
    return templates, np.array(pairs, np.int32).reshape(-1, 2)


def compute_frr(dataset, dist_thresh, orient_thresh_deg, use_type, ransac_iter, min_inliers, stop_inlier_ratio=0.15,
                max_workers=1, demo=False, device: int = 0, log_file: str = "logs/genuine_match_stats.csv") -> List[float]:
    templates, pairs = genuine_pairs(dataset, demo)
    scores: List[float] = []
    if len(pairs):
        tpl = [_as_template(t) for t in templates]
        it = ransac_iter if not demo else 50
        m = MinutiaeMatcher(len(tpl), max(max(len(t) for t in tpl), 1), it, device)
        try:
            m.set_templates(tpl)
            res, _, _ = m.match(pairs, False, dist_thresh=dist_thresh, orient_thresh_deg=orient_thresh_deg,
                                use_type=use_type, ransac_iter=it, min_inliers=min_inliers if not demo else 3,
                                stop_inlier_ratio=stop_inlier_ratio, cross_check=True)
        finally:
            m.close()
        scores = [float(s) for s in res["final_score"]]
    if log_file:
        os.makedirs(os.path.dirname(log_file) or ".", exist_ok=True)
        with open(log_file, "w", newline="") as f:              # FRR.py:93-135: the reference logs placeholders
            wr = csv.writer(f)
            wr.writerow(["user_id", "idx1", "idx2", "score", "num_inliers", "num_outliers", "rotation_deg",
                         "translation_x", "translation_y"])
            for s in scores:
                wr.writerow(["N/A", -1, -1, s, 0, 0, 0.0, 0.0, 0.0])
    return scores

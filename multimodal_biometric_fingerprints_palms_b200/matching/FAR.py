"""Mirror of `src/matching/FAR.py`: sampled impostor user pairs, all sample x sample matches in ONE batched GPU call.

`sample_impostor_pairs` draws with the `random` module exactly as FAR.py:26-32 does (seed `random` for repeatable
pairs); scores come back in task order (the reference extends them in process-completion order, FAR.py:77-82)."""
from __future__ import annotations

import random
from typing import Dict, List

import numpy as np

from .match import MinutiaeMatcher, _as_template


def sample_impostor_pairs(users, sample_size=100):                      # FAR.py:26-32
    pairs = []
    for u1 in users:
        for u2 in random.sample([u for u in users if u != u1], min(sample_size, len(users) - 1)):
            pairs.append((u1, u2))
    return pairs


def impostor_pairs(dataset: Dict[str, List[np.ndarray]], impostor_sample_size=100, demo=False):
    users = list(dataset.keys())
    if demo:
        impostor_sample_size = min(5, len(users))
    first, templates = {}, []
    for u in users:
        first[u] = len(templates)
        templates.extend(dataset[u])
    pairs = []
    for u1, u2 in sample_impostor_pairs(users, impostor_sample_size):   # far_worker_batch: for a in A: for b in B
        for i in range(len(dataset[u1])):
            for j in range(len(dataset[u2])):
                pairs.append((first[u1] + i, first[u2] + j))
    return templates, np.array(pairs, np.int32).reshape(-1, 2)


def compute_far(dataset, dist_thresh, orient_thresh_deg, use_type, ransac_iter, min_inliers, stop_inlier_ratio=0.15,
                max_workers=4, impostor_sample_size=100, demo=False, device: int = 0) -> List[float]:
    templates, pairs = impostor_pairs(dataset, impostor_sample_size, demo)
    if not len(pairs):
        return []
    tpl = [_as_template(t) for t in templates]
    m = MinutiaeMatcher(len(tpl), max(max(len(t) for t in tpl), 1), ransac_iter, device)
    try:
        m.set_templates(tpl)
        res, _, _ = m.match(pairs, False, dist_thresh=dist_thresh, orient_thresh_deg=orient_thresh_deg,
                            use_type=use_type, ransac_iter=ransac_iter, min_inliers=min_inliers,
                            stop_inlier_ratio=stop_inlier_ratio, cross_check=True)
    finally:
        m.close()
    return [float(s) for s in res["final_score"]]

"""Mirror of `src/matching/utils.py` (reports and threshold sweeps; plain host code, no arithmetic on the hot path)."""
from __future__ import annotations

import csv
import logging
import math
import os
from typing import List

import numpy as np


def console_step(title: str):                                            # utils.py:9-12 (colorama is cosmetic)
    print(f"\n{'=' * 60}\n{title.upper()}\n{'=' * 60}")


def rotate_points(points: np.ndarray, theta: float) -> np.ndarray:       # utils.py:14-18
    c, s = math.cos(theta), math.sin(theta)
    return points.dot(np.array([[c, -s], [s, c]]).T)


def angle_diff(a, b):                                                    # utils.py:20-24
    return ((a - b) + np.pi) % (2 * np.pi) - np.pi


def report_scores(title: str, scores: List[float]):                      # utils.py:29-40
    print("\n===========================================")
    print(f" {title}")
    print("===========================================")
    print(f"Num campioni: {len(scores)}")
    if len(scores) > 0:
        print(f"Media:  {np.mean(scores):.4f}")
        print(f"Min:    {np.min(scores):.4f}")
        print(f"Max:    {np.max(scores):.4f}")
        print(f"Std:    {np.std(scores):.4f}")
    print("===========================================\n")


def _sweep(scores, num_points, below: bool, name: str, verbose: bool):
    thresholds = np.linspace(0, 1, num_points)
    s = np.array(scores)
    vals = np.array([np.mean(s < t) if below else np.mean(s >= t) for t in thresholds])
    if verbose:
        print(f"\n VALORI {name} AL VARIARE DELLA SOGLIA\n\n{'Soglia':>8} | {name:>8}\n" + "-" * 22)
        for t, v in zip(thresholds, vals):
            print(f"{t:8.3f} | {v:8.3f}")
    return thresholds, vals


def evaluate_frr_across_thresholds(genuine_scores, num_points=50, verbose=True):     # utils.py:42-63: mean(score < t)
    return _sweep(genuine_scores, num_points, True, "FRR", verbose)


def evaluate_far_across_thresholds(impostor_scores, num_points=50, verbose=True):    # utils.py:66-87: mean(score >= t)
    return _sweep(impostor_scores, num_points, False, "FAR", verbose)


def compute_minutiae_statistics(dataset, output_file="logs/minutiae_stats.csv"):     # utils.py:89-122
    os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
    header = ["user_id", "sample_index", "num_minutiae", "mean_quality", "std_quality", "mean_orientation",
              "std_orientation", "mean_stability", "std_stability", "min_x", "max_x", "min_y", "max_y"]
    with open(output_file, "w", newline="") as fout:
        wr = csv.writer(fout)
        wr.writerow(header)
        for user_id, samples in dataset.items():
            for idx, M in enumerate(samples):
                if M.shape[0] == 0:
                    continue
                wr.writerow([user_id, idx, M.shape[0], np.mean(M[:, 4]), np.std(M[:, 4]), np.mean(M[:, 3]),
                             np.std(M[:, 3]), np.mean(M[:, 6]), np.std(M[:, 6]), np.min(M[:, 0]), np.max(M[:, 0]),
                             np.min(M[:, 1]), np.max(M[:, 1])])
    logging.info("Minutiae statistics salvate in %s", output_file)

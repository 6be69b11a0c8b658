"""Mirror of `src/matching/match_features.py`: JSON templates -> FRR -> FAR -> ROC with the matcher on the GPU."""
from __future__ import annotations

import json
import logging
import os
from typing import Dict, List

import numpy as np

from .FAR import compute_far
from .FRR import compute_frr
from .ROC import plot_roc
from .utils import (compute_minutiae_statistics, console_step, evaluate_far_across_thresholds,
                    evaluate_frr_across_thresholds, report_scores)


def load_dataset(minutiae_base: str, max_per_user: int = None) -> Dict[str, List[np.ndarray]]:
    """match_features.py:27-71: user id = file name up to the first '_'; rows x, y, type, orientation, quality,
    coherence, angular_stability as float64."""
    files_per_user: Dict[str, List[str]] = {}
    for root, _, files in os.walk(minutiae_base):
        for f in files:
            if f.endswith("_minutiae.json"):
                files_per_user.setdefault(f.split("_")[0], []).append(os.path.join(root, f))
    dataset = {}
    for user_id, paths in files_per_user.items():
        paths = sorted(paths)
        if max_per_user is not None:
            paths = paths[:max_per_user]
        out = []
        for path in paths:
            try:
                with open(path) as fin:
                    minutiae = json.load(fin)
                out.append(np.array([[float(m["x"]), float(m["y"]), float(0 if m.get("type", "ending") == "ending" else 1),
                                      float(m.get("orientation", 0.0)), float(m.get("quality", 0.0)),
                                      float(m.get("coherence", 0.0)), float(m.get("angular_stability", 0.0))]
                                     for m in minutiae], dtype=np.float64))
            except Exception as e:                                          # :67-68
                logging.warning(f"Errore caricando {path}: {e}")
        dataset[user_id] = out
    return dataset


def main(config_path="config/config_matching.yml", demo=False, minutiae_base=None, show=True):
    """match_features.py:75-159.  Thresholds and iteration counts are the reference's hard-coded ones (:95-148)."""
    import yaml
    console_step("Caricamento Configurazione")
    cfg = {}
    if config_path and os.path.exists(config_path):
        with open(config_path) as f:
            cfg = yaml.safe_load(f) or {}
    base = minutiae_base or cfg.get("minutiae_base", "dataset/processed/minutiae")
    if cfg.get("deterministic", True):
        np.random.seed(42)
    st = ({"max_per_user": 2, "frr_ransac": 500, "far_ransac": 500, "frr_min_inliers": 5, "far_min_inliers": 5, "num_points": 30}
          if demo else
          {"max_per_user": 2, "frr_ransac": 300, "far_ransac": 300, "frr_min_inliers": 6, "far_min_inliers": 12, "num_points": 50})
    console_step("Caricamento Dataset")
    dataset = load_dataset(base, max_per_user=st["max_per_user"])
    print(f"Utenti caricati: {len(dataset)}")
    compute_minutiae_statistics(dataset, output_file="logs/minutiae_stats.csv")
    console_step("Calcolo FRR")
    genuine = compute_frr(dataset, dist_thresh=30, orient_thresh_deg=30, use_type=True, ransac_iter=st["frr_ransac"],
                          min_inliers=st["frr_min_inliers"], demo=demo)
    report_scores("REPORT FRR (Genuine Scores)", genuine)
    _, frr = evaluate_frr_across_thresholds(genuine, num_points=st["num_points"])
    console_step("Calcolo FAR")
    impostor = compute_far(dataset, dist_thresh=15, orient_thresh_deg=10, use_type=True, ransac_iter=st["far_ransac"],
                           min_inliers=st["far_min_inliers"], demo=demo)
    report_scores("REPORT FAR (Impostor Scores)", impostor)
    th_far, far = evaluate_far_across_thresholds(impostor, num_points=st["num_points"])
    console_step("Generazione ROC")
    plot_roc(th_far, far_values=far, frr_values=frr, title="ROC (FAR vs FRR)", show=show)
    print("\nMatching completato\n")
    return {"genuine": genuine, "impostor": impostor, "frr": frr, "far": far}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="Fingerprint/Minutiae Matching")
    ap.add_argument("--config", type=str, default="config/config_matching.yml")
    ap.add_argument("--demo", action="store_true")
    a = ap.parse_args()
    console_step("Avvio Matching Minutiae")
    main(config_path=a.config, demo=a.demo)

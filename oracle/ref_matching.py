"""Restatement of the reference's RANSAC minutiae matcher.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows `/root/reference/src/matching/match.py` and
`src/matching/utils.py:17-27` on the same NumPy / scikit-learn calls.

One deliberate difference, stated here and in DESIGN.md: the reference runs its `max_iter` hypotheses in a thread
pool and consumes them with `as_completed` (`match.py:154-166`), keeping the first strictly better score and
BREAKING at the first hypothesis whose inlier count reaches `stop_inlier_ratio * min(nA, nB)` - so which hypothesis
wins depends on thread timing.  The per-hypothesis function (`ransac_worker`, seeded `default_rng(42 + i)`) is
deterministic and is restated verbatim; the aggregate is restated in SEED ORDER (what the reference computes when
its futures complete in submission order).  `oracle/make_golden_matching.py` pins `ransac_worker` and
`match_with_transform` against the reference's own functions.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
from sklearn.neighbors import KDTree


def rotate_points(points: np.ndarray, theta: float) -> np.ndarray:          # utils.py:17-21
    c, s = math.cos(theta), math.sin(theta)
    rot = np.array([[c, -s], [s, c]])
    return points.dot(rot.T)


def angle_diff(a, b):                                                       # utils.py:23-27
    d = a - b
    return (d + np.pi) % (2 * np.pi) - np.pi


def compute_descriptor_weight(m) -> float:                                  # match.py:10-22
    bonus = 1.25 if int(m[2]) == 1 else 1.0
    q = float(m[4]) if len(m) > 4 else 0.0
    coh = float(m[5]) if len(m) > 5 else 0.0
    angs = float(m[6]) if len(m) > 6 else 0.0
    return float(np.clip(bonus * (0.5 * q + 0.3 * coh + 0.2 * angs), 0.05, 2.0))


def match_with_transform(mins_a, mins_b, theta, t, dist_thresh, orient_thresh, weights_a, weights_b, use_type):
    """match.py:33-72"""
    if mins_a.shape[0] == 0 or mins_b.shape[0] == 0:
        return [], 0
    moved = rotate_points(mins_a[:, :2], theta) + t
    dists, idxs = KDTree(mins_b[:, :2]).query(moved, k=1)
    dists, idxs = dists.ravel(), idxs.ravel()
    out = []
    for ia, (d, ib) in enumerate(zip(dists, idxs)):
        if d > dist_thresh:
            continue
        if use_type and mins_a[ia, 2] != mins_b[ib, 2]:
            continue
        ang_err = abs(angle_diff(mins_a[ia, 3] + theta, mins_b[ib, 3]))
        if ang_err > orient_thresh:
            continue
        sigma_d, sigma_o = dist_thresh * 0.7, orient_thresh * 0.7
        spatial = math.exp(-(d ** 2) / (2 * sigma_d ** 2))
        orient_factor = math.exp(-(ang_err ** 2) / (2 * sigma_o ** 2))
        out.append((ia, ib, float(spatial * orient_factor * weights_a[ia] * weights_b[ib])))
    return out, len(out)


ZERO = {"score": 0.0, "inliers": []}


def ransac_worker(mins_a, mins_b, dist_thresh, orient_thresh, min_inliers, use_type, wA, wB, seed) -> Dict:
    """match.py:75-127"""
    rng = np.random.default_rng(seed)
    if mins_a.shape[0] < 8 or mins_b.shape[0] < 8:
        return dict(ZERO)
    if np.linalg.norm(mins_a[:, :2].std(0) - mins_b[:, :2].std(0)) > 35:
        return dict(ZERO)
    pA = rng.choice(np.arange(mins_a.shape[0]), p=wA / np.sum(wA))
    same = np.where(mins_b[:, 2] == mins_a[pA, 2])[0]
    if len(same) == 0:
        return dict(ZERO)
    pB = rng.choice(same, p=wB[same] / np.sum(wB[same]))
    theta = angle_diff(mins_b[pB, 3], mins_a[pA, 3])
    t = mins_b[pB, :2] - rotate_points(mins_a[pA, :2].reshape(1, 2), theta).reshape(2)
    inliers, n = match_with_transform(mins_a, mins_b, theta, t, dist_thresh, orient_thresh, wA, wB, use_type)
    if n < min_inliers:
        return dict(ZERO)
    weighted = sum(c for (_, _, c) in inliers)
    possible = min(np.sum(wA), np.sum(wB))
    score = (weighted / (possible + 1e-6)) ** 0.75
    return {"theta": theta, "t": t, "inliers": inliers, "score": float(np.clip(score, 0, 1))}


def ransac_align_and_match(mins_a, mins_b, dist_thresh, orient_thresh, max_iter, min_inliers, use_type,
                           stop_inlier_ratio) -> Dict:
    """match.py:129-217 with the hypotheses consumed in seed order (see module docstring)."""
    if len(mins_a) == 0 or len(mins_b) == 0:
        return dict(ZERO)
    wA = np.array([compute_descriptor_weight(m) for m in mins_a])
    wB = np.array([compute_descriptor_weight(m) for m in mins_b])
    best = dict(ZERO)
    for i in range(max_iter):
        r = ransac_worker(mins_a, mins_b, dist_thresh, orient_thresh, min_inliers, use_type, wA, wB, 42 + i)
        if r["score"] > best["score"]:
            best = r
        if len(r.get("inliers", [])) >= stop_inlier_ratio * min(len(mins_a), len(mins_b)):
            best = r
            break
    if best["score"] <= 0:
        return best
    idxA = np.array([i for (i, _, _) in best["inliers"]])
    idxB = np.array([j for (_, j, _) in best["inliers"]])
    Pa, Pb = mins_a[idxA, :2], mins_b[idxB, :2]
    ca, cb = Pa.mean(0), Pb.mean(0)
    Hm = (Pa - ca).T @ (Pb - cb)
    U, _, Vt = np.linalg.svd(Hm)
    Rm = Vt.T @ U.T
    if np.linalg.det(Rm) < 0:
        Vt[-1] *= -1
        Rm = Vt.T @ U.T
    theta = math.atan2(Rm[1, 0], Rm[0, 0])
    t = cb - rotate_points(ca.reshape(1, 2), theta).reshape(2)
    inliers, _ = match_with_transform(mins_a, mins_b, theta, t, dist_thresh, orient_thresh, wA, wB, use_type)
    weighted = sum(c for (_, _, c) in inliers)
    possible = min(np.sum(wA), np.sum(wB))
    score = float(np.clip((weighted / (possible + 1e-6)) ** 0.5, 0, 1))
    if len(inliers) >= 8:
        Pa = mins_a[[i for (i, _, _) in inliers], :2]
        Pb = mins_b[[j for (_, j, _) in inliers], :2]
        dA = np.linalg.norm(Pa - Pa.mean(0), axis=1).mean()
        dB = np.linalg.norm(Pb - Pb.mean(0), axis=1).mean()
        if abs(dA - dB) > 18:
            return dict(ZERO)
    return {"theta": theta, "t": t, "inliers": inliers, "score": score}


def match_minutiae_pair(mins_a, mins_b, dist_thresh=10.0, orient_thresh_deg=12.0, use_type=True, ransac_iter=300,
                        min_inliers=8, stop_inlier_ratio=0.25, cross_check=True) -> Dict:
    """match.py:219-275"""
    if mins_a is None or mins_b is None:
        return {"final_score": 0.0, "inlier_ratio": 0.0, "matches": []}
    A, B = np.array(mins_a), np.array(mins_b)
    orient_thresh = math.radians(orient_thresh_deg)
    best = ransac_align_and_match(A, B, dist_thresh, orient_thresh, ransac_iter, min_inliers, use_type, stop_inlier_ratio)
    inliers = best.get("inliers", [])
    if cross_check and len(inliers) > 0:
        moved = rotate_points(A[:, :2], best["theta"]) + best["t"]
        back = KDTree(moved).query(B[:, :2], k=1)[1].ravel()
        inliers = [(i, j, s) for (i, j, s) in inliers if back[j] == i]
    wA = np.array([compute_descriptor_weight(m) for m in A])
    wB = np.array([compute_descriptor_weight(m) for m in B])
    weighted = sum(s for (_, _, s) in inliers)
    possible = min(np.sum(wA), np.sum(wB))
    final = float(np.clip((weighted / (possible + 1e-6)) ** 0.25, 0, 1))
    return {"final_score": final, "inlier_ratio": len(inliers) / max(1, min(len(A), len(B))), "matches": inliers,
            "theta": best.get("theta", 0.0), "t": best.get("t", np.array([0.0, 0.0]))}


def synthetic_template(seed: int, n: int = 45, size: Tuple[int, int] = (315, 222)) -> np.ndarray:
    """Random template in the layout of match_features.py:52-62: x, y, type, orientation, quality, coherence, stability."""
    rng = np.random.default_rng(seed)
    h, w = size
    return np.column_stack([rng.integers(30, w - 30, n).astype(float), rng.integers(30, h - 30, n).astype(float),
                            rng.integers(0, 2, n).astype(float), rng.uniform(-np.pi / 2, np.pi / 2, n),
                            rng.uniform(0.2, 0.9, n), rng.uniform(0.2, 1.0, n), rng.uniform(0.3, 1.0, n)])


def perturbed_copy(tpl: np.ndarray, seed: int, angle_deg: float = 7.0, shift=(9.0, -6.0), jitter: float = 1.2,
                   drop: float = 0.2, extra: int = 6) -> np.ndarray:
    """A 'genuine' second impression: rigid motion + jitter + missing and spurious minutiae."""
    rng = np.random.default_rng(seed)
    th = math.radians(angle_deg)
    keep = rng.random(len(tpl)) > drop
    out = tpl[keep].copy()
    c = np.array([111.0, 157.0])
    out[:, :2] = rotate_points(out[:, :2] - c, th) + c + np.array(shift) + rng.normal(0, jitter, (len(out), 2))
    out[:, 3] = (out[:, 3] + th + np.pi / 2) % np.pi - np.pi / 2
    if extra:
        out = np.vstack([out, synthetic_template(seed + 999, extra)])
    return out

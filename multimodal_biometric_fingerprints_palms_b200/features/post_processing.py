"""Mirror of `src/features/post_processing.py` of the reference (post_processing.py:69-137)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from ..pipeline import pipeline_for


def nms_adaptive(minutiae: List[Dict], density_map: np.ndarray, base_dist: float = 8.0) -> List[Dict]:
    """post_processing.py:10-32 - including the reference's last-writer-wins visit (a suppressed point is never
    skipped, so every visited point re-instates itself and clears its neighbours)."""
    if not minutiae:
        return []
    dm = np.asarray(density_map)
    xy = np.array([[m["x"], m["y"]] for m in minutiae], np.int32)
    q = np.array([m.get("quality", 1.0) for m in minutiae], np.float64)
    dens = np.array([np.float32(dm[m["y"], m["x"]]) for m in minutiae], np.float32)
    keep = pipeline_for(dm.shape[0], dm.shape[1]).nms_adaptive(xy, q, dens, base_dist)
    return [m for i, m in enumerate(minutiae) if keep[i]]


def remove_redundant_oriented_adaptive(minutiae: List[Dict], density_map: np.ndarray, base_radius: float = 20.0,
                                       angle_thresh: float = float(np.deg2rad(30))) -> List[Dict]:
    """post_processing.py:37-64."""
    if not minutiae:
        return []
    dm = np.asarray(density_map)
    xy = np.array([[m["x"], m["y"]] for m in minutiae], np.int32)
    q = np.array([m.get("quality", 1.0) for m in minutiae], np.float64)
    o = np.array([m["orientation"] for m in minutiae], np.float64)
    dens = np.array([np.float32(dm[m["y"], m["x"]]) for m in minutiae], np.float32)
    keep = pipeline_for(dm.shape[0], dm.shape[1]).remove_redundant(xy, q, o, dens, base_radius, angle_thresh)
    return [m for i, m in enumerate(minutiae) if keep[i]]


def postprocess_minutiae(minutiae: List[Dict], skel: np.ndarray, gray: Optional[np.ndarray] = None,
                         params: Optional[Dict] = None) -> List[Dict]:
    """Scoring, adaptive NMS, oriented-redundancy removal, top-K on the GPU.

    As in the reference the surviving input dicts are updated IN PLACE with `orientation`, `quality`,
    `coherence`, `angular_stability` and returned quality-descending.  `gray` is what the orientation
    map is computed from (post_processing.py:93): the reference's only caller passes the skeleton itself
    (extract_features.py:92) - the fused path; any other uint8 image of the skeleton's shape goes through
    `fpb_postprocess_gray`, and `gray=None` is `(skel > 0)` as a 0/1 uint8 image, as in the reference."""
    if not minutiae or skel is None:
        return []
    skel = np.ascontiguousarray(skel)
    if skel.dtype != np.uint8 or skel.ndim != 2:
        raise NotImplementedError("CUDA path takes a 2-D uint8 skeleton")
    if gray is None:
        g = (skel > 0).astype(np.uint8)                          # sk_bin (post_processing.py:85, 93)
    else:
        g = np.ascontiguousarray(gray)
        if g.dtype != np.uint8 or g.ndim != 2:
            raise NotImplementedError("CUDA path takes a 2-D uint8 `gray` (float images are not on the reference's path)")
        if g is skel or np.array_equal(g, skel):
            g = None                                             # the reference's call site: one upload, fpb_postprocess
    h, w = skel.shape
    p = pipeline_for(h, w)
    p.set_post_params(params)
    try:
        refined = p.postprocess(skel, [minutiae], gray=g)[0]
    finally:
        if params:
            p.set_post_params(None)                              # the cached handle goes back to the hard-coded defaults
    # hand back the caller's own dict objects, updated in place (post_processing.py:122-128)
    by_key = {}
    for m in minutiae:
        by_key.setdefault((int(m["x"]), int(m["y"]), m["type"]), []).append(m)
    out = []
    for r in refined:
        src = by_key[(r["x"], r["y"], r["type"])].pop(0)
        src.update({"orientation": r["orientation"], "quality": r["quality"], "coherence": r["coherence"],
                    "angular_stability": r["angular_stability"]})
        out.append(src)
    return out

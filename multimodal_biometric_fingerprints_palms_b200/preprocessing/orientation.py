"""Mirror of `src/preprocessing/orientation.py` of the reference (orientation.py:9-130)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ..pipeline import pipeline_for


def compute_orientation_map(img: np.ndarray, block_size: int = 16, smooth_sigma: float = 3.0,
                            invert_if_needed: bool = True, smooth_orientation_sigma: float = 3.0,
                            mask: Optional[np.ndarray] = None):
    """orientation.py:9-85 on the GPU.  Returns (orient_blocks f32[h//16, w//16], orient_img f32[h,w],
    rel_img f32[h,w]).  Only the parameter values the hot path uses are compiled in
    (fingerprint_preprocess.py:192-195, post_processing.py:93); anything else raises."""
    if (block_size, float(smooth_sigma), bool(invert_if_needed), float(smooth_orientation_sigma)) != (16, 3.0, True, 3.0):
        raise NotImplementedError("CUDA path implements block_size=16, smooth_sigma=3.0, invert_if_needed=True, "
                                  "smooth_orientation_sigma=3.0 (the values on the reference's hot path)")
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise NotImplementedError("CUDA path takes 2-D uint8 images (what the hot path passes)")
    h, w = img.shape
    m = None
    if mask is not None:
        m = np.ascontiguousarray((np.asarray(mask) > 0).astype(np.uint8) * 255)
    p = pipeline_for(h, w)
    blocks, oimg, rel = p.orientation(img, m)
    return np.ascontiguousarray(blocks[0]), np.ascontiguousarray(oimg[0]), np.ascontiguousarray(rel[0])


def visualize_orientation(img: np.ndarray, orient_img: np.ndarray, reliability_img: np.ndarray = None,
                          block_size: int = 16, scale: int = 8, rel_thresh: float = 0.2,
                          mask: Optional[np.ndarray] = None, color=(0, 0, 255)):
    """orientation.py:87-130 - debug overlay (row V1 of SURVEY.md section 8(a)).  Rendered on the host with
    OpenCV's anti-aliased line drawing from the GPU-computed fields; it is a debug artefact of the
    returned dict, not part of the enhance -> minutiae path."""
    import cv2
    base = cv2.cvtColor(np.clip(img, 0, 255).astype(np.uint8), cv2.COLOR_GRAY2BGR) if img.ndim == 2 else img.copy()
    vis = base.copy()
    h, w = orient_img.shape
    half = block_size // 2
    for by in range(h // block_size):
        for bx in range(w // block_size):
            cy, cx = by * block_size + half, bx * block_size + half
            if cy >= h or cx >= w:
                continue
            if mask is not None and mask[cy, cx] == 0:
                continue
            if reliability_img is not None and reliability_img[cy, cx] < rel_thresh:
                continue
            ang = orient_img[cy, cx]
            dx, dy = int(round(scale * np.cos(ang))), int(round(scale * np.sin(ang)))
            p1 = (max(0, cx - dx), max(0, cy - dy))
            p2 = (min(w - 1, cx + dx), min(h - 1, cy + dy))
            cv2.line(vis, p1, p2, color, 1, cv2.LINE_AA)
    return cv2.addWeighted(vis, 0.8, base, 0.2, 0)

"""Mirror of `src/preprocessing/orientation.py` of the reference (orientation.py:9-130)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ..pipeline import pipeline_for


def compute_orientation_map(img: np.ndarray, block_size: int = 16, smooth_sigma: float = 3.0,
                            invert_if_needed: bool = True, smooth_orientation_sigma: float = 3.0,
                            mask: Optional[np.ndarray] = None):
    """orientation.py:9-85 on the GPU.  Returns (orient_blocks f32[h//block_size, w//block_size], orient_img f32[h,w],
    rel_img f32[h,w]).  The defaults are the values on the reference's hot path (fingerprint_preprocess.py:192-195,
    post_processing.py:93) and run the tuned kernels; other keyword values go through `fpb_orientation_ex` (same kernels,
    generic Gaussian radii / block size).  Non-uint8 images (orientation.py:21-24) are handed over as
    `img.astype(np.float32)` and rescaled on the device (`fpb_orientation_f32`)."""
    img = np.asarray(img)
    if img.ndim != 2:
        raise ValueError("too many values to unpack (expected 2)" if img.ndim > 2 else
                         f"not enough values to unpack (expected 2, got {img.ndim})")       # h, w = f.shape (:46)
    if img.dtype != np.uint8:
        img = np.ascontiguousarray(img.astype(np.float32))        # :23
    h, w = img.shape
    bs = int(block_size)
    if bs == 0:
        raise ZeroDivisionError("integer division or modulo by zero")          # h // block_size (:47)
    if bs < 0:
        raise ValueError("negative dimensions are not allowed")               # np.zeros((n_by, n_bx)) (:48)
    if h // bs < 1 or w // bs < 1:
        import cv2                                                             # cv2.resize of the empty block grid (:81)
        raise cv2.error(f"block_size {bs} leaves no whole block in a {h}x{w} image: (-215:Assertion failed) !ssize.empty() in function 'resize'")
    # SciPy leaves an axis unfiltered when its sigma is <= 1e-15 (negative values included); the ABI takes 0 for that
    smooth_sigma = float(smooth_sigma) if float(smooth_sigma) > 1e-15 else 0.0
    smooth_orientation_sigma = float(smooth_orientation_sigma) if float(smooth_orientation_sigma) > 1e-15 else 0.0
    m = None
    if mask is not None:
        m = np.ascontiguousarray((np.asarray(mask) > 0).astype(np.uint8) * 255)
    p = pipeline_for(h, w)
    blocks, oimg, rel = p.orientation(img, m, block_size=bs, smooth_sigma=float(smooth_sigma),
                                      invert_if_needed=bool(invert_if_needed),
                                      smooth_orientation_sigma=float(smooth_orientation_sigma))
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

"""Mirror of `src/preprocessing/fingerprint_preprocess.py` of the reference
(fingerprint_preprocess.py:13-225): same public functions, computed by libfpb200."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from ..pipeline import pipeline_for
from .orientation import compute_orientation_map, visualize_orientation  # noqa: F401  (re-exported like the reference)


def _gray_u8(img) -> np.ndarray:
    img = np.asarray(img)
    if img.ndim != 2 or img.dtype != np.uint8:
        raise NotImplementedError("CUDA path takes 2-D uint8 images")
    return np.ascontiguousarray(img)


def normalize_image(img: np.ndarray) -> np.ndarray:
    """:13-29  percentile stretch + CLAHE(2.5, 8x8)."""
    img = _gray_u8(img)
    return pipeline_for(*img.shape, exact=True).normalize(img)[0]


def denoise_image(img: np.ndarray) -> np.ndarray:
    """:34-38  NLM(h=10, 7, 21) + GaussianBlur 3x3 sigma 0.6."""
    img = _gray_u8(img)
    return pipeline_for(*img.shape, exact=True).denoise(img)[0]


def binarize(img: np.ndarray) -> np.ndarray:
    """:43-81  adaptive Sauvola | patch Otsu, component clean-up, opening, reconstruction -> {0,255}."""
    img = _gray_u8(img)
    return np.ascontiguousarray(pipeline_for(*img.shape).binarize(img)[0])


def segment_fingerprint(img: np.ndarray, debug_dir: Optional[str] = None, save_mask_dir: Optional[str] = None,
                        img_name: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
    """:86-136  returns (cropped image zeroed outside the hull, cropped hull mask).  Colour input (H x W x 3 or 4, uint8)
    goes through cv2.COLOR_BGR2GRAY first, as in the reference (:94) - on the device (`fpb_segment_bgr`)."""
    img = np.asarray(img)
    if img.ndim == 3 and img.dtype == np.uint8 and img.shape[2] in (3, 4):
        seg, mask, roi = pipeline_for(img.shape[0], img.shape[1], exact=True).segment_bgr(np.ascontiguousarray(img))
    else:
        img = _gray_u8(img)
        seg, mask, roi = pipeline_for(*img.shape, exact=True).segment(img)
    _, _, w, h = (int(v) for v in roi[0])
    seg, mask = seg[0, :h, :w].copy(), mask[0, :h, :w].copy()
    if save_mask_dir and img_name:
        import cv2
        os.makedirs(save_mask_dir, exist_ok=True)
        cv2.imwrite(os.path.join(save_mask_dir, img_name), mask)
    return seg, mask


def smooth_fingerprint_skeleton(binary_img: np.ndarray, sigma: float = 1.4, diffusion_iter: int = 3,
                                contrast_boost: float = 1.25) -> np.ndarray:
    """:141-159.  The defaults (the hot path's call, :197) run the fused kernel; other values the unfused sequence."""
    b = _gray_u8(binary_img)
    return np.ascontiguousarray(pipeline_for(*b.shape).smooth(b, sigma=float(sigma), diffusion_iter=int(diffusion_iter),
                                                             contrast_boost=float(contrast_boost))[0])


def thinning_and_cleaning(binary_img: np.ndarray, orientation_img: np.ndarray, reliability_img: np.ndarray,
                          rel_thresh: float = 0.1) -> np.ndarray:
    """:161-177 (`orientation_img` is unused, as in the reference)."""
    b = _gray_u8(binary_img)
    r = np.ascontiguousarray(np.asarray(reliability_img, dtype=np.float32))
    return np.ascontiguousarray(pipeline_for(*b.shape).thin(b, r, rel_thresh=float(rel_thresh))[0])


def preprocess_fingerprint(img: np.ndarray, debug_dir: Optional[str] = None, save_mask_dir: Optional[str] = None,
                           img_name: Optional[str] = None) -> Dict[str, np.ndarray]:
    """:182-225  K1..K7 in one fused GPU run; same seven result keys, fresh arrays, and the same
    RuntimeError wrapping on any failure."""
    try:
        img = _gray_u8(img)
        H, W = img.shape
        p = pipeline_for(H, W, exact=True)
        # EXTENSION (not in the reference): FPB200_ENHANCED=1 adds the Gabor-enhanced crop under the key "enhanced"
        # that run_preprocessing.py:133 looks for; off by default so the result dict is the reference's
        want_enh = os.environ.get("FPB200_ENHANCED", "0") == "1"
        if want_enh:
            p.enable_enhanced()
        else:
            p.disable_enhanced()
        # rel_thresh of thinning_and_cleaning: 0.1 as hard-coded at :202, or general.rel_threshold with FPB200_YAML_OVERRIDES=1
        from ..config import config_fingerprint
        p.set_rel_threshold(config_fingerprint.active_overrides().get("rel_threshold", 0.1))
        p.run(img)
        x0, y0, w, h = p.roi(0)
        crop = lambda name: p.fetch(name)[0, :h, :w].copy()
        normalized, denoised = p.fetch("normalized")[0], p.fetch("denoised")[0]
        segmented, mask, binary, skeleton = crop("segmented"), crop("mask"), crop("binary"), crop("skeleton")
        orient_img, reliability = crop("orient_img"), crop("reliability")
        orientation_vis = visualize_orientation(img=segmented, orient_img=orient_img, reliability_img=reliability,
                                                block_size=16, scale=7, rel_thresh=0.1, mask=mask)
        if save_mask_dir and img_name:
            import cv2
            os.makedirs(save_mask_dir, exist_ok=True)
            cv2.imwrite(os.path.join(save_mask_dir, img_name), mask)
        out = {"normalized": normalized, "denoised": denoised, "segmented": segmented, "mask": mask,
               "binary": binary, "skeleton": skeleton, "orientation_vis": orientation_vis}
        if want_enh:
            out["enhanced"] = crop("enhanced")
        if debug_dir:
            import cv2
            os.makedirs(debug_dir, exist_ok=True)
            for key in ("normalized", "denoised", "segmented", "binary", "skeleton", "orientation_vis"):   # the six of :205-212
                cv2.imwrite(os.path.join(debug_dir, f"{key}.jpg"), out[key])
        return out
    except Exception as e:
        raise RuntimeError(f"preprocess_fingerprint failed: {e}") from e

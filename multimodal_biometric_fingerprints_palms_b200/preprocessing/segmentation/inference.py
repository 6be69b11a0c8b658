"""Mirror of `src/preprocessing/segmentation/inference.py` of the reference (inference.py:87-133) as importable
functions (the reference runs everything at import time from config/config_segmentation.yml)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Mapping, Optional

import numpy as np

from .model import FingerprintSegmentationModel


def preprocess_image(img_path: str, image_size):
    """:87-93  grey image resized with INTER_AREA, / 255, replicated to three channels -> ([1,3,H,W] float32, BGR original)."""
    import cv2
    img_gray = cv2.imread(img_path, cv2.IMREAD_GRAYSCALE)
    img_rgb = cv2.imread(img_path)
    if img_gray is None:
        raise FileNotFoundError(img_path)
    img_resized = cv2.resize(img_gray, tuple(image_size), interpolation=cv2.INTER_AREA)
    t = (img_resized / 255.0).astype(np.float32)[None, None]
    return np.repeat(t, 3, axis=1), img_rgb


def mask_to_rgb(mask_logits: np.ndarray, original_rgb: np.ndarray, threshold: float = 0.5):
    """:95-110  sigmoid > threshold, nearest-neighbour resize to the original, black-out and red overlays."""
    import cv2
    logits = np.asarray(mask_logits, np.float32).squeeze()
    mask = 1.0 / (1.0 + np.exp(-logits))
    mask_bin = (mask > threshold).astype(np.uint8)
    mask_bin_resized = cv2.resize(mask_bin, (original_rgb.shape[1], original_rgb.shape[0]), interpolation=cv2.INTER_NEAREST)
    overlay = original_rgb.copy()
    overlay[mask_bin_resized == 0] = 0
    mask_color = np.zeros_like(original_rgb)
    mask_color[:, :, 2] = mask_bin_resized * 255
    overlay_color = cv2.addWeighted(original_rgb, 0.7, mask_color, 0.3, 0)
    return mask_bin_resized * 255, overlay, overlay_color


def run_inference(img_dir: str, output_dir: str, state_dict: Mapping, image_size=(256, 256), device: int = 0,
                  max_batch: int = 8) -> Dict[str, int]:
    """:115-133  every .jpg/.png/.jpeg of `img_dir` -> <stem>_mask.png, <stem>_segmented.png, <stem>_overlay.png.
    `state_dict`: the `model_state_dict` of the reference's checkpoint (inference.py:77-78)."""
    import cv2
    os.makedirs(output_dir, exist_ok=True)
    model = FingerprintSegmentationModel(image_size=tuple(image_size), device=device, max_batch=max_batch)
    model.load_state_dict(state_dict)
    model.eval()
    files = [f for f in sorted(os.listdir(img_dir)) if f.lower().endswith((".jpg", ".png", ".jpeg"))]
    done = 0
    for s in range(0, len(files), max_batch):
        part = files[s:s + max_batch]
        pre = [preprocess_image(os.path.join(img_dir, f), image_size) for f in part]
        logits = model(np.concatenate([p[0] for p in pre], 0))
        for f, (_, rgb), lg in zip(part, pre, logits):
            mask_bin, segmented, overlay_color = mask_to_rgb(lg, rgb)
            stem = Path(f).stem
            cv2.imwrite(os.path.join(output_dir, f"{stem}_mask.png"), mask_bin)
            cv2.imwrite(os.path.join(output_dir, f"{stem}_segmented.png"), segmented)
            cv2.imwrite(os.path.join(output_dir, f"{stem}_overlay.png"), overlay_color)
            done += 1
    return {"found": len(files), "processed": done}

"""Mirror of `src/preprocessing/run_preprocessing.py` of the reference (run_preprocessing.py:71-166): directory walk ->
per-image pipeline -> `<base>_enhanced.jpg` / `<base>_skeleton.jpg` under `<output_dir>/enhanced/<relative dir>`.

Images are grouped by shape and pushed through the GPU in batches instead of one at a time; file naming, directory
mirroring, the "enhanced = input image" behaviour (the reference looks for a result key that is never produced,
:133-135), the debug tree (`debug/<relative dir>/<result key>/<base>.jpg`, :49-66, 143-144) and the RuntimeError on an
empty input directory are the reference's."""
from __future__ import annotations

import logging
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Tuple

import numpy as np

from ..pipeline import FingerprintPipeline

VALID_EXTS = (".jpg", ".jpeg", ".png", ".bmp")
log = logging.getLogger(__name__)


def load_image(path: str):
    """:38-47"""
    import cv2
    try:
        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise IOError("Immagine non leggibile")
        return img
    except Exception as e:
        log.error(f"Impossibile leggere {path}: {e}")
        return None


def save_debug_images(results: dict, debug_base: str, base_name: str):
    """:49-66  debug/<rel dir>/<result key>/<base>.jpg for every entry of the result dict."""
    import cv2
    for key, img in results.items():
        if img is None:
            continue
        out_dir = os.path.join(debug_base, key)
        os.makedirs(out_dir, exist_ok=True)
        if img.dtype == bool:
            img = img.astype(np.uint8) * 255
        elif np.issubdtype(img.dtype, np.floating):
            img = np.clip(img * 255.0, 0, 255).astype(np.uint8)
        else:
            img = np.clip(img, 0, 255).astype(np.uint8)
        cv2.imwrite(os.path.join(out_dir, f"{base_name}.jpg"), img)


def _result_dict(pipe: FingerprintPipeline, planes: Dict[str, np.ndarray], k: int) -> Dict[str, np.ndarray]:
    """The seven entries preprocess_fingerprint returns (fingerprint_preprocess.py:214-222) for image k of the batch."""
    from .orientation import visualize_orientation
    _, _, cw, ch = pipe.roi(k)
    crop = lambda name: np.ascontiguousarray(planes[name][k, :ch, :cw])
    seg, mask = crop("segmented"), crop("mask")
    vis = visualize_orientation(img=seg, orient_img=crop("orient_img"), reliability_img=crop("reliability"),
                                block_size=16, scale=7, rel_thresh=0.1, mask=mask)
    return {"normalized": planes["normalized"][k], "denoised": planes["denoised"][k], "segmented": seg, "mask": mask,
            "binary": crop("binary"), "skeleton": crop("skeleton"), "orientation_vis": vis}


def run_preprocessing(input_dir: str, output_dir: str, debug: bool = False, small_subset: bool = False,
                      max_workers: int = 4, batch: int = 256, device: int = 0):
    import cv2
    files = [os.path.join(root, f) for root, _, fs in os.walk(input_dir) for f in fs if f.lower().endswith(VALID_EXTS)]
    if not files:
        log.error(f"Nessuna immagine trovata in {input_dir}")
        raise RuntimeError(f"Nessuna immagine trovata in {input_dir}")
    if small_subset:
        files = files[:10]
    enhanced_dir = os.path.join(output_dir, "enhanced")
    debug_root = os.path.join(output_dir, "debug") if debug else None          # :103-108
    os.makedirs(enhanced_dir, exist_ok=True)
    if debug:
        os.makedirs(debug_root, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max_workers) as ex:          # decode on host threads
        imgs = list(ex.map(load_image, files))
    by_shape: Dict[Tuple[int, int], List[int]] = {}
    for i, im in enumerate(imgs):
        if im is None:
            log.warning(f"Immagine NON processata: {os.path.basename(files[i])}")
            continue
        by_shape.setdefault(im.shape, []).append(i)
    done = 0
    for (h, w), idxs in by_shape.items():
        pipe = FingerprintPipeline(h, w, max_batch=min(batch, len(idxs)), device=device)
        for s in range(0, len(idxs), batch):
            part = idxs[s:s + batch]
            pipe.run(np.stack([imgs[i] for i in part]))
            skel = pipe.fetch("skeleton")
            planes = None
            if debug:                                                # the whole result dict per image (:143-144)
                planes = {name: pipe.fetch(name) for name in ("normalized", "denoised", "segmented", "mask", "binary",
                                                               "orient_img", "reliability")}
                planes["skeleton"] = skel
            for k, i in enumerate(part):
                x0, y0, cw, ch = pipe.roi(k)
                base = os.path.splitext(os.path.basename(files[i]))[0]
                rel_dir = os.path.relpath(os.path.dirname(files[i]), input_dir)
                sub = os.path.join(enhanced_dir, rel_dir)
                os.makedirs(sub, exist_ok=True)
                if debug:
                    save_debug_images(_result_dict(pipe, planes, k), os.path.join(debug_root, rel_dir), base)
                cv2.imwrite(os.path.join(sub, f"{base}_enhanced.jpg"), imgs[i])
                cv2.imwrite(os.path.join(sub, f"{base}_skeleton.jpg"), np.ascontiguousarray(skel[k, :ch, :cw]))
                done += 1
        pipe.close()
    log.info(f"Risultati salvati in: {output_dir}")
    return done

"""Mirror of `src/features/extract_features.py` of the reference (extract_features.py:38-159)."""
from __future__ import annotations

import json
import logging
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List

import numpy as np

from ..pipeline import pipeline_for
from .post_processing import postprocess_minutiae

OUTPUT_DIR_DEFAULT = "dataset/processed/minutiae"
log = logging.getLogger(__name__)


def clean_skeleton(skel: np.ndarray) -> np.ndarray:
    """:38-39."""
    return (skel > 127).astype(np.uint8)


def extract_minutiae(skel: np.ndarray) -> List[Dict]:
    """:41-69  crossing-number minutiae of `skel > 127`, row-major order, dicts {"x","y","type"}."""
    skel = np.ascontiguousarray(skel)
    if skel.dtype != np.uint8 or skel.ndim != 2:
        raise NotImplementedError("CUDA path takes a 2-D uint8 skeleton")
    h, w = skel.shape
    return pipeline_for(h, w).extract_minutiae(skel)[0]


def process_image(filename, cluster_dir, out_dir, params):
    """:74-108  <sample>_skeleton.jpg -> <sample>_minutiae.json (+ .jpg visualisation)."""
    import cv2
    if not filename.endswith("_skeleton.jpg"):
        return
    sample = filename.replace("_skeleton.jpg", "")
    path = os.path.join(cluster_dir, filename)
    try:
        skel = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if skel is None:
            log.error(f"Immagine skeleton corrotta: {path}")
            return
        raw = extract_minutiae(skel)
        try:
            refined = postprocess_minutiae(raw, skel, skel, params)
        except Exception as e:
            log.error(f"postprocess_minutiae error on {filename}: {e}")
            refined = []
        vis = cv2.cvtColor((skel > 127).astype(np.uint8) * 255, cv2.COLOR_GRAY2BGR)
        for m in refined:
            cv2.circle(vis, (m["x"], m["y"]), 3, (0, 0, 255) if m["type"] == "ending" else (0, 255, 0), -1)
        cv2.imwrite(os.path.join(out_dir, f"{sample}_minutiae.jpg"), vis)
        with open(os.path.join(out_dir, f"{sample}_minutiae.json"), "w") as f:
            json.dump(refined, f, indent=2)
    except Exception as e:
        log.error(f"Errore elaborando {filename}: {e}")


def process_cluster_dir(cluster_dir: str, output_base: str, params=None, max_workers=None):
    """:113-136."""
    files = [f for f in os.listdir(cluster_dir) if f.lower().endswith("_skeleton.jpg")]
    if not files:
        return
    out_dir = os.path.join(output_base, os.path.basename(cluster_dir))
    os.makedirs(out_dir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max_workers) as ex:
        list(ex.map(lambda f: process_image(f, cluster_dir, out_dir, params), files))


def main(input_base="dataset/processed/enhanced", output_base=OUTPUT_DIR_DEFAULT, max_workers=None):
    """:141-159."""
    if not os.path.exists(input_base):
        raise FileNotFoundError(f"Input base non trovato: {input_base}")
    clusters = [os.path.join(input_base, d) for d in os.listdir(input_base) if d.startswith("cluster_")]
    log.info(f"Trovati {len(clusters)} cluster.")
    # params=None like the reference (extract_features.py:156-157: the hard-coded defaults of postprocess_minutiae);
    # FPB200_YAML_OVERRIDES=1 makes config_fingerprint.yml's orientation.* section live instead (opt-in)
    from ..config import config_fingerprint
    params = config_fingerprint.active_overrides().get("post_params")
    for c in clusters:
        process_cluster_dir(c, output_base, params=params, max_workers=max_workers)


if __name__ == "__main__":
    main()

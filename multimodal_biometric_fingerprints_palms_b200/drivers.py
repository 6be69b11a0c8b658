"""Fused directory driver: image files -> `<sample>_minutiae.json` in one pass.

Equivalent to the reference's two batch drivers run back to back (`run_preprocessing.py:71-166` then
`extract_features.py:141-159`) but without the JPEG round trip of the skeleton between them (the hand-off is
loss-free after `> 127`, SURVEY row D2, so the JSON is the same), with batching by image shape and resume
(skip-if-exists).  File names and directory mirroring are the reference's, so `src/matching/match_features.py`
consumes the output unchanged."""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Tuple

import numpy as np

from .pipeline import FingerprintPipeline
from .preprocessing.run_preprocessing import VALID_EXTS, load_image


def run_directory(input_dir: str, output_dir: str, batch: int = 512, device: int = 0, resume: bool = True,
                  write_skeletons: bool = True, io_workers: int = 8, params: Dict | None = None) -> Dict[str, int]:
    """Returns {"found", "processed", "skipped", "unreadable"}."""
    import cv2
    files = sorted(os.path.join(r, f) for r, _, fs in os.walk(input_dir) for f in fs if f.lower().endswith(VALID_EXTS))
    if not files:
        raise RuntimeError(f"Nessuna immagine trovata in {input_dir}")
    enh_root, min_root = os.path.join(output_dir, "enhanced"), os.path.join(output_dir, "minutiae")

    def targets(path: str) -> Tuple[str, str, str]:
        rel = os.path.relpath(os.path.dirname(path), input_dir)
        base = os.path.splitext(os.path.basename(path))[0]
        return (os.path.join(min_root, rel, f"{base}_minutiae.json"), os.path.join(enh_root, rel, f"{base}_skeleton.jpg"),
                os.path.join(enh_root, rel, f"{base}_enhanced.jpg"))

    todo = [f for f in files if not (resume and os.path.exists(targets(f)[0]))]
    stats = {"found": len(files), "processed": 0, "skipped": len(files) - len(todo), "unreadable": 0}
    with ThreadPoolExecutor(max_workers=io_workers) as ex:
        imgs = list(ex.map(load_image, todo))
        by_shape: Dict[Tuple[int, int], List[int]] = {}
        for i, im in enumerate(imgs):
            if im is None:
                stats["unreadable"] += 1
            else:
                by_shape.setdefault(im.shape, []).append(i)
        for (h, w), idxs in by_shape.items():
            pipe = FingerprintPipeline(h, w, max_batch=min(batch, len(idxs)), device=device)
            pipe.set_post_params(params)
            for s in range(0, len(idxs), batch):
                part = idxs[s:s + batch]
                pipe.run(np.stack([imgs[i] for i in part]))
                skel = pipe.fetch("skeleton") if write_skeletons else None

                def emit(k_i):
                    k, i = k_i
                    js, sk, en = targets(todo[i])
                    os.makedirs(os.path.dirname(js), exist_ok=True)
                    if write_skeletons:
                        os.makedirs(os.path.dirname(sk), exist_ok=True)
                        _, _, cw, ch = pipe.roi(k)
                        cv2.imwrite(sk, np.ascontiguousarray(skel[k, :ch, :cw]))
                        cv2.imwrite(en, imgs[i])
                    tmp = js + ".tmp"
                    with open(tmp, "w") as f:
                        json.dump(pipe.minutiae(k), f, indent=2)
                    os.replace(tmp, js)              # a crash never leaves a half-written JSON for the resume check
                for item in enumerate(part):
                    emit(item)
                stats["processed"] += len(part)
            pipe.close()
    return stats

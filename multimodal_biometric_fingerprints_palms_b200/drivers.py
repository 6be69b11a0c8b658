"""Fused directory driver: image files -> `<sample>_minutiae.json` in one pass.

Equivalent to the reference's two batch drivers run back to back (`run_preprocessing.py:71-166` then
`extract_features.py:141-159`).  The reference hands the skeleton from the first to the second as a quality-95 JPEG
file, and its second stage computes on the DECODED grey levels (`extract_minutiae` thresholds them at 127,
`postprocess_minutiae` takes density / orientation / coherence from the codec's ringing); the fused run reproduces
that file on the device (`fpb_set_handoff`, k_jpeg_roundtrip: bit-identical to cv2.imwrite + cv2.imread), so the JSON
is the one the reference's CLI writes - not the one an in-memory hand-off would give.  Batching by image shape, resume
(skip-if-exists).  Baseline JPEG inputs are entropy-decoded by the library's host threads and reconstructed on the
GPU (bit-identical to `cv2.imread`); anything else (PNG, BMP, progressive JPEG ...) is read with cv2 as the reference
does.  The JSON files are written by the library's native writer (byte-identical to `json.dump(..., indent=2)`).
File names and directory mirroring are the reference's, so `src/matching/match_features.py` consumes the output
unchanged."""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Tuple

import numpy as np

from . import _native as N
from .pipeline import FingerprintPipeline
from .preprocessing.run_preprocessing import VALID_EXTS


def _probe(path: str):
    """(bytes, (h, w), is_native_jpeg) - or (None, None, False) when unreadable."""
    import cv2
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None, None, False
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    if N.load().fpb_jpeg_info(blob, len(blob), C.byref(w), C.byref(h), C.byref(c)) == 0 and c.value in (1, 3):
        return blob, (h.value, w.value), True
    img = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_GRAYSCALE)       # run_preprocessing.py:41
    if img is None:
        return None, None, False
    return img, img.shape, False


def run_directory(input_dir: str, output_dir: str, batch: int = 512, device: int = 0, resume: bool = True,
                  write_skeletons: bool = True, io_workers: int = 8, params: Dict | None = None, lanes: int = 1,
                  window: int = 0) -> Dict[str, int]:
    """Returns {"found", "processed", "skipped", "unreadable", "gpu_decoded", "seconds": {phase: seconds summed over lanes}}.

    Host memory is bounded: the pending files are taken in windows of `window` files (default 8 x batch); a window is read
    and bucketed by shape while the GPU works on the previous one, its buffers are released after its results are written,
    and the per-shape handles (sized for `batch` images) are reused across windows."""
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

    import time
    # opt-in YAML parameters (FPB200_YAML_OVERRIDES=1); an explicit `params` argument wins
    from .config import config_fingerprint
    ov = config_fingerprint.active_overrides()
    if params is None:
        params = ov.get("post_params")
    rel_threshold = ov.get("rel_threshold", 0.1)
    todo = [f for f in files if not (resume and os.path.exists(targets(f)[0]))]
    stats = {"found": len(files), "processed": 0, "skipped": len(files) - len(todo), "unreadable": 0, "gpu_decoded": 0}
    tm = {"read": 0.0, "create": 0.0, "decode": 0.0, "run": 0.0, "emit": 0.0}       # wall-clock seconds per phase
    clock = time.perf_counter
    window = int(window) if window and window > 0 else 8 * batch
    windows = [todo[s0:s0 + window] for s0 in range(0, len(todo), window)]
    lock = threading.Lock()
    create_lock = threading.Lock()
    lane_pipes: List[Dict[Tuple[int, int], FingerprintPipeline]] = [dict() for _ in range(max(1, lanes))]

    def add(key, dt=None, **inc):
        with lock:
            if dt is not None:
                tm[key] += dt
            for k, v in inc.items():
                stats[k] += v

    try:
        with ThreadPoolExecutor(max_workers=io_workers) as ex, ThreadPoolExecutor(max_workers=io_workers) as rd, \
                ThreadPoolExecutor(max_workers=1) as pre:
            nxt = pre.submit(lambda fs: list(rd.map(_probe, fs)), windows[0]) if windows else None
            for wi, wfiles in enumerate(windows):
                t0 = clock()
                probed = nxt.result()                                   # exposed read time only: the next window loads under this one's GPU work
                tm["read"] += clock() - t0
                nxt = pre.submit(lambda fs: list(rd.map(_probe, fs)), windows[wi + 1]) if wi + 1 < len(windows) else None
                by_shape: Dict[Tuple[int, int], List[int]] = {}
                for i, (data, shape, _) in enumerate(probed):
                    if data is None:
                        stats["unreadable"] += 1
                    else:
                        by_shape.setdefault(tuple(shape), []).append(i)
                # jobs = (shape, kind, indices); `lanes` worker threads, each with its own handles, take them in turn so that one
                # lane's host phases (Huffman decoding, JSON / JPEG writing) run under the other lane's GPU phase
                jobs: List[Tuple[Tuple[int, int], str, List[int]]] = []
                for shape, idxs in by_shape.items():
                    native = [i for i in idxs if probed[i][2]]
                    host = [i for i in idxs if not probed[i][2]]
                    jobs += [(shape, "jpeg", native[s0:s0 + batch]) for s0 in range(0, len(native), batch)]
                    jobs += [(shape, "host", host[s0:s0 + batch]) for s0 in range(0, len(host), batch)]
                cursor = [0]

                def lane(pipes, probed=probed, wfiles=wfiles, jobs=jobs, cursor=cursor):
                    def emit(pipe, part: List[int]):
                        t0 = clock()
                        paths = [targets(wfiles[i])[0] for i in part]
                        for d in {os.path.dirname(p) for p in paths}:
                            os.makedirs(d, exist_ok=True)
                        pipe.write_json(paths, io_workers)
                        if write_skeletons:
                            skel = pipe.fetch("skeleton")

                            def one(k_i):
                                k, i = k_i
                                _, sk, en = targets(wfiles[i])
                                os.makedirs(os.path.dirname(sk), exist_ok=True)
                                _, _, cw, ch = pipe.roi(k)
                                cv2.imwrite(sk, np.ascontiguousarray(skel[k, :ch, :cw]))
                                src = probed[i][0]          # `_enhanced.jpg` is the input image (run_preprocessing.py:133-135)
                                cv2.imwrite(en, src if isinstance(src, np.ndarray) else
                                            cv2.imdecode(np.frombuffer(src, np.uint8), cv2.IMREAD_GRAYSCALE))
                            list(ex.map(one, enumerate(part)))
                        add("emit", clock() - t0, processed=len(part))

                    def run_host(pipe, part, imgs):
                        t0 = clock()
                        pipe.run(np.stack(imgs))
                        add("run", clock() - t0)
                        emit(pipe, part)

                    while True:
                        with lock:
                            j = cursor[0]; cursor[0] += 1
                        if j >= len(jobs):
                            break
                        (h, w), kind, part = jobs[j]
                        pipe = pipes.get((h, w))
                        if pipe is None:
                            t0 = clock()
                            with create_lock:           # concurrent cudaMalloc / cudaMallocHost of two multi-GB workspaces contend badly
                                pipe = pipes[(h, w)] = FingerprintPipeline(h, w, max_batch=batch, device=device)
                            pipe.set_post_params(params)
                            pipe.set_rel_threshold(rel_threshold)
                            add("create", clock() - t0)
                        if kind == "host":
                            run_host(pipe, part, [probed[i][0] for i in part])
                            continue
                        t0 = clock()
                        status = pipe.decode_jpeg([probed[i][0] for i in part], io_workers)
                        bad = [i for i, st in zip(part, status) if st != 0]
                        good = [i for i, st in zip(part, status) if st == 0]
                        if bad and good:
                            pipe.decode_jpeg([probed[i][0] for i in good], io_workers)
                        add("decode", clock() - t0)
                        if good:
                            t0 = clock()
                            pipe.run_decoded(len(good))
                            add("run", clock() - t0, gpu_decoded=len(good))
                            emit(pipe, good)
                        if bad:                             # progressive / EXIF / corrupt: let cv2 decide, like the reference
                            imgs = [cv2.imdecode(np.frombuffer(probed[i][0], np.uint8), cv2.IMREAD_GRAYSCALE) for i in bad]
                            keep = [(i, im) for i, im in zip(bad, imgs) if im is not None and im.shape == (h, w)]
                            add("read", unreadable=len(bad) - len(keep))
                            if keep:
                                run_host(pipe, [i for i, _ in keep], [im for _, im in keep])

                errors: List[BaseException] = []

                def guarded(fn, arg):
                    def run():
                        try:
                            fn(arg)
                        except BaseException as e:  # re-raised on the caller's thread below
                            errors.append(e)
                    return run
                workers = [threading.Thread(target=guarded(lane, lane_pipes[k])) for k in range(max(1, min(lanes, len(jobs))))]
                for t in workers:
                    t.start()
                for t in workers:
                    t.join()
                if errors:
                    raise errors[0]
                del probed                                              # this window's blobs / decoded arrays are released here
    finally:
        for pipes in lane_pipes:
            for pp in pipes.values():
                pp.close()
    stats["seconds"] = {k: round(v, 4) for k, v in tm.items()}
    return stats

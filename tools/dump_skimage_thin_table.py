#!/usr/bin/env python
"""Recover scikit-image's literal 256-entry `skeletonize` deletion table on a machine that HAS scikit-image, and write
it in the form `FPB200_THIN_TABLE=<file>` / `FingerprintPipeline.set_thin_table` take.

Why: /root/reference/src/preprocessing/fingerprint_preprocess.py:171 calls `skimage.morphology.skeletonize`, whose 2-D
path (`_fast_skeletonize`, Cython) is Zhang-Suen thinning driven by a hard-coded table.  The container libfpb200 was
built in has no scikit-image, so its built-in default is the table DERIVED from the Zhang & Suen conditions; if the
package's table has other entries, skeletons differ (the library's start-up self-check, selfcheck.py, reports that).

How: (1) scan the compiled extension module for 256 consecutive int32 (or int8) values in 0..3 - the table is static
data there; (2) VALIDATE every candidate, and the built-in table, by running table-driven thinning in NumPy against
`skimage.morphology.skeletonize` on random masks; (3) write the validated table.  If no candidate validates, paste the
`lut = [...]` list of your version's `skimage/morphology/_skeletonize*_cy.pyx` into a text file - the loader accepts it.

    python tools/dump_skimage_thin_table.py [out.txt]
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np


def zhang_suen_table() -> np.ndarray:
    bit = [2, 4, 8, 16, 32, 64, 128, 1]            # N NE E SE S SW W NW in skimage's coding
    t = np.zeros(256, np.uint8)
    for c in range(256):
        p = [1 if c & b else 0 for b in bit]
        B = sum(p); A = sum(1 for i in range(8) if p[i] == 0 and p[(i + 1) % 8] == 1)
        if 2 <= B <= 6 and A == 1:
            n, e, s, w = p[0], p[2], p[4], p[6]
            t[c] = (1 if n * e * s == 0 and e * s * w == 0 else 0) | (2 if n * e * w == 0 and n * s * w == 0 else 0)
    return t


def thin(mask: np.ndarray, table: np.ndarray) -> np.ndarray:
    """Table-driven parallel thinning with scikit-image's pass semantics (snapshot per sub-iteration, until stable)."""
    sk = np.pad(mask.astype(np.uint8), 1)
    while True:
        removed = False
        for first in (True, False):
            c = (sk[:-2, :-2] * 1 + sk[:-2, 1:-1] * 2 + sk[:-2, 2:] * 4 + sk[1:-1, 2:] * 8 + sk[2:, 2:] * 16 +
                 sk[2:, 1:-1] * 32 + sk[2:, :-2] * 64 + sk[1:-1, :-2] * 128)
            v = table[c]
            kill = (sk[1:-1, 1:-1] == 1) & ((v == 3) | (v == (1 if first else 2)))
            if kill.any():
                removed = True
                sk[1:-1, 1:-1][kill] = 0
        if not removed:
            return sk[1:-1, 1:-1].astype(bool)


def validates(table: np.ndarray, n: int = 120) -> int:
    from skimage.morphology import skeletonize
    rng = np.random.default_rng(0)
    bad = 0
    for i in range(n):
        h, w = int(rng.integers(8, 48)), int(rng.integers(8, 48))
        m = rng.random((h, w)) < rng.uniform(0.3, 0.9)
        if i % 3 == 0:                                   # blobs: thick shapes exercise many iterations
            from scipy.ndimage import binary_dilation
            m = binary_dilation(rng.random((h, w)) < 0.08, iterations=int(rng.integers(1, 4)))
        bad += int((thin(m, table) != skeletonize(m)).sum())
    return bad


def candidates_from_binary(path: str):
    blob = np.frombuffer(open(path, "rb").read(), np.uint8)
    for dtype, width in ((np.int32, 4), (np.int8, 1)):
        ok = (blob <= 3)
        if width == 4:
            for shift in range(4):
                b = blob[shift:shift + (len(blob) - shift) // 4 * 4].reshape(-1, 4)
                good = (b[:, 0] <= 3) & (b[:, 1] == 0) & (b[:, 2] == 0) & (b[:, 3] == 0)
                yield from _runs(good, b[:, 0])
        else:
            yield from _runs(ok, blob)


def _runs(good: np.ndarray, vals: np.ndarray):
    run = 0
    for i, g in enumerate(good):
        run = run + 1 if g else 0
        if run >= 256:
            t = vals[i - 255:i + 1].astype(np.uint8)
            if t[0] == 0 and {1, 2, 3} <= set(t.tolist()) and (t != 0).sum() >= 20:
                yield t.copy()


def main() -> int:
    out = sys.argv[1] if len(sys.argv) > 1 else "skimage_thin_table.txt"
    try:
        import skimage
        import skimage.morphology as mo
    except Exception as e:
        print("scikit-image is not importable here:", e)
        return 2
    zs = zhang_suen_table()
    bad = validates(zs)
    print(f"scikit-image {getattr(skimage, '__version__', '?')}: built-in Zhang-Suen table differs on {bad} pixels of the validation set")
    if bad == 0:
        print("the built-in default already equals scikit-image's behaviour - nothing to override")
        np.savetxt(out, zs.reshape(16, 16), fmt="%d")
        return 0
    seen = set()
    for so in glob.glob(os.path.join(os.path.dirname(mo.__file__), "_skeletonize*.so")) + \
            glob.glob(os.path.join(os.path.dirname(mo.__file__), "_skeletonize*.pyd")):
        for t in candidates_from_binary(so):
            key = t.tobytes()
            if key in seen:
                continue
            seen.add(key)
            b = validates(t)
            print(f"{os.path.basename(so)}: candidate with {(t != 0).sum()} non-zero entries -> {b} differing pixels")
            if b == 0:
                np.savetxt(out, t.reshape(16, 16), fmt="%d")
                print(f"validated; written to {out}.  Use it with FPB200_THIN_TABLE={os.path.abspath(out)}")
                return 0
    print("no table found in the compiled module; paste the `lut = [...]` list from your version's "
          "skimage/morphology/_skeletonize*_cy.pyx into a text file and pass it as FPB200_THIN_TABLE")
    return 1


if __name__ == "__main__":
    sys.exit(main())

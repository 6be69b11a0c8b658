"""Restatement of the five scikit-image functions on the reference's hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference imports (`/root/reference/src/preprocessing/fingerprint_preprocess.py:5-6`)

    from skimage.filters import threshold_otsu
    from skimage.morphology import remove_small_objects, remove_small_holes, \
                                   reconstruction, skeletonize

scikit-image (pinned `>=0.22,<1.0`, `config/environment.yml`) is NOT installed in
this image and there is no network, so the published algorithms are restated
here on numpy + scipy.ndimage.  PARITY UNPINNED against scikit-image itself:

* `threshold_otsu`, `remove_small_objects`, `remove_small_holes`,
  `reconstruction`: restated from the documented behaviour of scikit-image 0.22
  (`filters/thresholding.py`, `exposure/exposure.py::_histogram`,
  `morphology/misc.py`, `morphology/grayreconstruct.py`).
* `skeletonize` (2-D): scikit-image's `_fast_skeletonize` is Zhang-Suen thinning
  driven by a hard-coded 256-entry table whose literal contents could not be
  recovered here.  The table below is DERIVED from the Zhang & Suen (CACM 1984)
  deletion conditions with scikit-image's neighbour encoding and pass semantics.
  If scikit-image's literal table contains extra deletions the skeletons differ;
  `tests/test_oracle_skimage.py` compares against the real package whenever it
  is importable and fails loudly.  The CUDA path takes the table as DATA, so
  CUDA<->oracle bit-exactness holds for any table.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

__all__ = [
    "threshold_otsu", "remove_small_objects", "remove_small_holes",
    "reconstruction", "skeletonize", "zhang_suen_table",
]


# --------------------------------------------------------------------------- #
# threshold_otsu  (skimage/filters/thresholding.py, float-image path)
# --------------------------------------------------------------------------- #
def threshold_otsu(image, nbins: int = 256):
    """Otsu threshold of a float image, scikit-image semantics.

    Called at `fingerprint_preprocess.py:68` on float32 32x32 patches.
    Float input => `np.histogram(image, bins=nbins)` over [min, max], bin
    centres `(e[:-1]+e[1:])/2`, counts cast to float32, between-class variance
    `w1[:-1]*w2[1:]*(m1[:-1]-m2[1:])**2`, returns the bin CENTRE at the argmax.
    A constant image returns its value.
    """
    image = np.asarray(image)
    flat = image.reshape(-1)
    first = flat[0]
    if np.all(flat == first):
        return first
    if np.issubdtype(flat.dtype, np.integer):
        # integer images use a bincount histogram with unit-spaced centres
        lo, hi = int(flat.min()), int(flat.max())
        counts = np.bincount(flat.astype(np.int64) - lo, minlength=hi - lo + 1)
        centers = np.arange(lo, hi + 1)
    else:
        counts, edges = np.histogram(flat, bins=nbins)
        centers = (edges[:-1] + edges[1:]) / 2.0
    counts = counts.astype("float32", copy=False)

    w1 = np.cumsum(counts)
    w2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
        var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return centers[int(np.argmax(var12))]


# --------------------------------------------------------------------------- #
# remove_small_objects / remove_small_holes  (skimage/morphology/misc.py)
# --------------------------------------------------------------------------- #
def remove_small_objects(ar, min_size: int = 64, connectivity: int = 1):
    """Drop connected components with fewer than `min_size` pixels.

    bool input => labelled with `ndi.label` and a `generate_binary_structure(2,
    connectivity)` footprint (connectivity=1: 4-neighbours); components with
    `size < min_size` are cleared.  Call sites: `fingerprint_preprocess.py:73,167`.
    """
    ar = np.asarray(ar)
    if ar.dtype != bool:
        raise TypeError("oracle restatement covers the bool path only")
    out = ar.copy()
    if min_size == 0:
        return out
    footprint = ndi.generate_binary_structure(ar.ndim, connectivity)
    ccs = np.zeros(ar.shape, dtype=np.int32)
    ndi.label(ar, footprint, output=ccs)
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = False
    return out


def remove_small_holes(ar, area_threshold: int = 64, connectivity: int = 1):
    """NOT -> remove_small_objects -> NOT (so border-touching background
    components smaller than the threshold are filled too).
    Call sites: `fingerprint_preprocess.py:74,168`."""
    ar = np.asarray(ar)
    if ar.dtype != bool:
        raise TypeError("oracle restatement covers the bool path only")
    inv = np.logical_not(ar)
    inv = remove_small_objects(inv, area_threshold, connectivity)
    return np.logical_not(inv)


# --------------------------------------------------------------------------- #
# reconstruction  (skimage/morphology/grayreconstruct.py, method='dilation')
# --------------------------------------------------------------------------- #
def reconstruction(seed, mask, method: str = "dilation"):
    """Morphological reconstruction by dilation, default 3x3 footprint.

    Call site `fingerprint_preprocess.py:80`: `seed` bool (eroded mask), `mask`
    uint8 in {0,1}.  For two-level images the geodesic dilation with the full 3x3
    footprint converges to: every 8-connected component of `mask>0` that holds
    at least one seed pixel, at the mask's value; float64 result as skimage.
    """
    if method != "dilation":
        raise NotImplementedError("only method='dilation' is on the hot path")
    seed = np.asarray(seed)
    mask = np.asarray(mask)
    if np.any(seed.astype(np.float64) > mask.astype(np.float64)):
        raise ValueError("Intensity of seed image must be less than that of the "
                         "mask image for reconstruction by dilation.")
    levels = np.unique(mask)
    if levels.size > 2 or (levels.size == 2 and levels[0] != 0):
        raise NotImplementedError("oracle restatement covers two-level masks")
    fg = mask > 0
    lab = np.zeros(mask.shape, dtype=np.int32)
    ndi.label(fg, np.ones((3, 3), dtype=bool), output=lab)
    hit = np.zeros(int(lab.max()) + 1, dtype=bool)
    hit[np.unique(lab[(seed > 0) & fg])] = True
    hit[0] = False
    out = np.where(hit[lab], mask.astype(np.float64), 0.0)
    # seed pixels equal to the mask where the mask is 0 stay 0 (seed <= mask)
    return out


# --------------------------------------------------------------------------- #
# skeletonize  (skimage/morphology/_skeletonize.py + _skeletonize_cy.pyx)
# --------------------------------------------------------------------------- #
# Neighbour encoding of scikit-image's table:
#   NW=1  N=2  NE=4  E=8  SE=16  S=32  SW=64  W=128
_NB_BITS = {"NW": 1, "N": 2, "NE": 4, "E": 8, "SE": 16, "S": 32, "SW": 64, "W": 128}


def zhang_suen_table() -> np.ndarray:
    """256-entry deletion table derived from Zhang & Suen (1984).

    Value bit 0 (1): deletable in the first sub-iteration, bit 1 (2): in the
    second, 3: in either - the same value convention scikit-image uses.
    P2..P9 = N, NE, E, SE, S, SW, W, NW (clockwise from north):
      (a) 2 <= B(P) <= 6        B = number of set neighbours
      (b) A(P) == 1             A = 0->1 transitions in P2,P3,...,P9,P2
      (c1) P2*P4*P6 == 0 and P4*P6*P8 == 0      first sub-iteration
      (c2) P2*P4*P8 == 0 and P2*P6*P8 == 0      second sub-iteration
    """
    ring = ["N", "NE", "E", "SE", "S", "SW", "W", "NW"]
    tab = np.zeros(256, dtype=np.uint8)
    for code in range(256):
        p = [1 if code & _NB_BITS[k] else 0 for k in ring]
        b = sum(p)
        a = sum(1 for i in range(8) if p[i] == 0 and p[(i + 1) % 8] == 1)
        if not (2 <= b <= 6 and a == 1):
            continue
        n, e, s, w = p[0], p[2], p[4], p[6]
        v = 0
        if n * e * s == 0 and e * s * w == 0:
            v |= 1
        if n * e * w == 0 and n * s * w == 0:
            v |= 2
        tab[code] = v
    return tab


_ZS_TABLE = zhang_suen_table()


def neighbour_codes(img01: np.ndarray) -> np.ndarray:
    """Per-pixel 8-neighbour code (zero padding outside the image)."""
    p = np.pad(img01.astype(np.uint8), 1)
    h, w = img01.shape
    sl = lambda dy, dx: p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w].astype(np.uint16)
    return (sl(-1, -1) * 1 + sl(-1, 0) * 2 + sl(-1, 1) * 4 + sl(0, 1) * 8 +
            sl(1, 1) * 16 + sl(1, 0) * 32 + sl(1, -1) * 64 + sl(0, -1) * 128)


def skeletonize(image, table: np.ndarray | None = None):
    """2-D thinning with scikit-image's pass structure.

    Each sub-iteration reads a snapshot of the image and writes a copy (fully
    parallel); sub-iteration 1 removes pixels whose table value is 1 or 3,
    sub-iteration 2 those with 2 or 3; both are repeated until a full pass
    removes nothing.  The image is treated as zero-padded by one pixel.
    Call site: `fingerprint_preprocess.py:171`.
    """
    tab = _ZS_TABLE if table is None else np.asarray(table, dtype=np.uint8)
    sk = (np.asarray(image) != 0).astype(np.uint8)
    if sk.ndim != 2:
        raise NotImplementedError("2-D only")
    changed = True
    while changed:
        changed = False
        for want in (1, 2):
            v = tab[neighbour_codes(sk)]
            kill = (sk == 1) & ((v == 3) | (v == want))
            if kill.any():
                sk[kill] = 0
                changed = True
    return sk.astype(bool)

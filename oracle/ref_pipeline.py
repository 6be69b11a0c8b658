"""The reference's per-image hot path, restated on the libraries it calls.

TEST INFRASTRUCTURE (see oracle/__init__.py) - the parity oracle proper.

Every function names the reference lines it follows (paths relative to
`/root/reference`).  OpenCV, SciPy and NumPy are present in this image, so the
library calls are made exactly as the reference makes them; the five
scikit-image calls go to `oracle.skimage_compat`.  `oracle/make_golden.py`
checks these functions bit-for-bit against the reference's own modules imported
from `/root/reference` and freezes the outputs under `tests/golden/`.

All intermediate products are returned so each CUDA stage can be compared in
isolation (`run_all`).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np
from scipy import ndimage as ndi
from scipy.spatial import cKDTree

from . import skimage_compat as sk

# --------------------------------------------------------------------------- #
# hard-coded constants of the reference - NOT the YAML values
# --------------------------------------------------------------------------- #
CLAHE_CLIP_NORMALIZE = 2.5          # fingerprint_preprocess.py:25-28
CLAHE_CLIP_SEGMENT = 2.0            # :97
CLAHE_CLIP_BINARIZE = 2.5           # :46
CLAHE_GRID = (8, 8)
NLM_H, NLM_TEMPLATE, NLM_SEARCH = 10, 7, 21     # :36
POST_BLUR_SIGMA = 0.6               # :38
SAUVOLA_WIN, SAUVOLA_K = 25, 0.25   # :49-50
OTSU_PATCH = 32                     # :60
BIN_MIN_OBJ, BIN_MAX_HOLE = 80, 150  # :73-74
THIN_MIN_OBJ, THIN_MAX_HOLE = 64, 80  # :167-168
SEG_MORPH = 15                      # :107
SEG_MARGIN = 10                     # :126
REL_THRESH = 0.1                    # :164


# ------------------------------ K1 ----------------------------------------- #
def normalize_image(img: np.ndarray) -> np.ndarray:
    """fingerprint_preprocess.py:13-29 - percentile stretch then CLAHE(2.5, 8x8)."""
    if img.dtype == np.uint8:
        unit = img.astype(np.float32) / 255.0
    else:
        unit = (img - img.min()) / (img.max() - img.min() + 1e-8)
    lo = np.percentile(unit, 0.5)
    span = np.percentile(unit, 99.5) - np.percentile(unit, 0.5) + 1e-12
    unit = np.clip((unit - lo) / span, 0.0, 1.0)
    stretched = (unit * 255).astype(np.uint8)
    return cv2.createCLAHE(clipLimit=CLAHE_CLIP_NORMALIZE, tileGridSize=CLAHE_GRID).apply(stretched)


# ------------------------------ K2 ----------------------------------------- #
def denoise_image(img: np.ndarray) -> np.ndarray:
    """fingerprint_preprocess.py:34-38 - NLM(h=10,7,21) then GaussianBlur 3x3 sigma 0.6."""
    nlm = cv2.fastNlMeansDenoising(img, None, h=NLM_H, templateWindowSize=NLM_TEMPLATE,
                                   searchWindowSize=NLM_SEARCH)
    return cv2.GaussianBlur(nlm, (3, 3), POST_BLUR_SIGMA)


def denoise_image_parts(img: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Same as `denoise_image` but also returns the NLM output before the blur."""
    nlm = cv2.fastNlMeansDenoising(img, None, h=NLM_H, templateWindowSize=NLM_TEMPLATE,
                                   searchWindowSize=NLM_SEARCH)
    return nlm, cv2.GaussianBlur(nlm, (3, 3), POST_BLUR_SIGMA)


# ------------------------------ K3 ----------------------------------------- #
def segment_fingerprint(img: np.ndarray, detail: Optional[dict] = None) -> Tuple[np.ndarray, np.ndarray]:
    """fingerprint_preprocess.py:86-136 - Otsu + morphology + convex hull + crop.

    `detail`, when given, receives the intermediates (equalised, blurred, otsu
    mask after the inversion test, mask after close/open, hull mask, bbox).
    """
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img
    eq = cv2.createCLAHE(clipLimit=CLAHE_CLIP_SEGMENT, tileGridSize=CLAHE_GRID).apply(gray)
    blurred = cv2.GaussianBlur(eq, (5, 5), 0)
    otsu_t, fg = cv2.threshold(blurred, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    with np.errstate(invalid="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if np.mean(gray[fg == 255]) > np.mean(gray[fg == 0]):     # :103
                fg = cv2.bitwise_not(fg)
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (SEG_MORPH, SEG_MORPH))
    closed = cv2.morphologyEx(fg, cv2.MORPH_CLOSE, se)
    opened = cv2.morphologyEx(closed, cv2.MORPH_OPEN, se)
    contours, _ = cv2.findContours(opened, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if detail is not None:
        detail.update(eq=eq, blurred=blurred, otsu_t=otsu_t, fg=fg, closed=closed, opened=opened)
    if not contours:                                                   # :113-118
        full = np.ones_like(opened, dtype=np.uint8) * 255
        if detail is not None:
            detail.update(hull_mask=full, bbox=None)
        return gray, full
    biggest = max(contours, key=cv2.contourArea)                       # :120
    hull = cv2.convexHull(biggest)
    hull_mask = np.zeros_like(opened)
    cv2.drawContours(hull_mask, [hull], -1, 255, -1)
    bx, by, bw, bh = cv2.boundingRect(hull)
    y0, y1 = max(0, by - SEG_MARGIN), by + bh + SEG_MARGIN
    x0, x1 = max(0, bx - SEG_MARGIN), bx + bw + SEG_MARGIN
    crop = gray[y0:y1, x0:x1]
    crop_mask = hull_mask[y0:y1, x0:x1]
    crop = cv2.bitwise_and(crop, crop, mask=crop_mask)
    if detail is not None:
        detail.update(hull_mask=hull_mask, bbox=(bx, by, bw, bh),
                      roi=(x0, y0, crop.shape[1], crop.shape[0]), hull=hull.reshape(-1, 2))
    return crop, crop_mask


# ------------------------------ K4 ----------------------------------------- #
def binarize(img: np.ndarray, detail: Optional[dict] = None) -> np.ndarray:
    """fingerprint_preprocess.py:43-81 - adaptive Sauvola | patch Otsu, CC clean-up,
    3x3 cross opening, reconstruction."""
    as_f = img.astype(np.float32)
    eq = cv2.createCLAHE(clipLimit=CLAHE_CLIP_BINARIZE, tileGridSize=CLAHE_GRID).apply(
        as_f.astype(np.uint8)).astype(np.float32)
    m = cv2.boxFilter(eq, -1, (SAUVOLA_WIN, SAUVOLA_WIN))
    m2 = cv2.boxFilter(eq ** 2, -1, (SAUVOLA_WIN, SAUVOLA_WIN))
    sd = np.sqrt(np.clip(m2 - m ** 2, 0, None))
    sd_rel = sd / (sd.max() + 1e-6)
    kk = SAUVOLA_K * (1 - 0.5 * sd_rel)
    thr = m * (1 - kk * (1 - sd / (m + 1e-6)))
    fg = eq < thr
    sauvola_only = fg.copy()

    hh, ww = eq.shape
    for r in range(0, hh, OTSU_PATCH):                                 # :62-71
        for c in range(0, ww, OTSU_PATCH):
            tile = eq[r:r + OTSU_PATCH, c:c + OTSU_PATCH]
            if tile.size < 10 or tile.std() < 3:
                continue
            try:
                t = sk.threshold_otsu(tile)
                fg[r:r + OTSU_PATCH, c:c + OTSU_PATCH] |= (tile < t)
            except Exception:
                pass
    merged = fg.copy()
    kept = sk.remove_small_objects(fg, min_size=BIN_MIN_OBJ)
    filled = sk.remove_small_holes(kept, area_threshold=BIN_MAX_HOLE)
    cross = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    opened = cv2.morphologyEx(filled.astype(np.uint8), cv2.MORPH_OPEN, cross)
    marker = cv2.erode(opened, cross, iterations=1).astype(bool)
    rec = sk.reconstruction(marker, opened, method="dilation")
    out = (rec > 0).astype(np.uint8) * 255
    if detail is not None:
        detail.update(eq=eq, mean=m, sqmean=m2, std=sd, thr=thr, sauvola=sauvola_only,
                      merged=merged, kept=kept, filled=filled, opened=opened, marker=marker)
    return out


# ------------------------------ K5 ----------------------------------------- #
def compute_orientation_map(img: np.ndarray, block_size: int = 16, smooth_sigma: float = 3.0,
                            invert_if_needed: bool = True, smooth_orientation_sigma: float = 3.0,
                            mask: Optional[np.ndarray] = None, detail: Optional[dict] = None):
    """orientation.py:9-85 - structure-tensor orientation / reliability."""
    if img.dtype == np.uint8:
        f = img.astype(np.float32) / 255.0
    else:
        f = img.astype(np.float32)
        if f.max() > 1.0 or f.min() < 0.0:
            f = (f - f.min()) / (f.max() - f.min() + 1e-12)
    if invert_if_needed:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            med = np.median(f)
            if np.mean(f[f > med]) > np.mean(f[f <= med]):              # :27
                f = 1.0 - f
    pre = ndi.gaussian_filter(f, sigma=max(0.5, smooth_sigma / 2.0))
    scaled = (pre * 255).astype(np.float32)
    gx = cv2.Sobel(scaled, cv2.CV_32F, 1, 0, ksize=3)
    gy = cv2.Sobel(scaled, cv2.CV_32F, 0, 1, ksize=3)
    jxx = ndi.gaussian_filter(gx * gx, sigma=smooth_sigma)
    jyy = ndi.gaussian_filter(gy * gy, sigma=smooth_sigma)
    jxy = ndi.gaussian_filter(gx * gy, sigma=smooth_sigma)
    rel_raw = np.sqrt((jxx - jyy) ** 2 + 4.0 * jxy ** 2)
    r_lo, r_hi = np.percentile(rel_raw, [2, 98])
    rel = np.clip((rel_raw - r_lo) / (r_hi - r_lo + 1e-12), 0.0, 1.0)
    theta = 0.5 * np.arctan2(2.0 * jxy, (jxx - jyy) + 1e-12) + np.pi / 2.0

    rows, cols = f.shape
    nby, nbx = rows // block_size, cols // block_size
    blk_theta = np.zeros((nby, nbx), dtype=np.float32)
    blk_rel = np.zeros((nby, nbx), dtype=np.float32)
    for j in range(nby):                                               # :52-72
        ys = slice(j * block_size, (j + 1) * block_size)
        for i in range(nbx):
            xs = slice(i * block_size, (i + 1) * block_size)
            if mask is not None and np.mean(mask[ys, xs] > 0) < 0.3:
                continue
            th = theta[ys, xs]
            rr = rel[ys, xs]
            if th.size == 0:
                continue
            wt = rr.flatten() + 1e-6
            s2 = np.sum(wt * np.sin(2.0 * th).flatten())
            c2 = np.sum(wt * np.cos(2.0 * th).flatten())
            blk_theta[j, i] = 0.5 * np.arctan2(s2, c2)
            blk_rel[j, i] = np.mean(rr)
    raw_blk_theta = blk_theta.copy()
    s_sm = ndi.gaussian_filter(np.sin(2.0 * blk_theta), sigma=smooth_orientation_sigma)
    c_sm = ndi.gaussian_filter(np.cos(2.0 * blk_theta), sigma=smooth_orientation_sigma)
    blk_theta = 0.5 * np.arctan2(s_sm, c_sm)
    orient_img = cv2.resize(blk_theta, (cols, rows), interpolation=cv2.INTER_LINEAR)
    rel_img = cv2.resize(blk_rel, (cols, rows), interpolation=cv2.INTER_LINEAR)
    orient_img = (orient_img + np.pi / 2) % np.pi - np.pi / 2
    if detail is not None:
        detail.update(f=f, pre=pre, gx=gx, gy=gy, jxx=jxx, jyy=jyy, jxy=jxy, rel_raw=rel_raw,
                      r_lo=r_lo, r_hi=r_hi, rel=rel, theta=theta, raw_blk_theta=raw_blk_theta,
                      blk_rel=blk_rel)
    return blk_theta, orient_img, rel_img


# ------------------------------ K6 ----------------------------------------- #
def smooth_fingerprint_skeleton(binary_img: np.ndarray, sigma: float = 1.4, diffusion_iter: int = 3,
                                contrast_boost: float = 1.25, detail: Optional[dict] = None) -> np.ndarray:
    """fingerprint_preprocess.py:141-159 - three explicit 'diffusion' steps, Gaussian 0.6,
    contrast boost, threshold 0.35."""
    base = binary_img.astype(np.float32) / 255.0
    gx, gy = ndi.sobel(base, axis=1), ndi.sobel(base, axis=0)
    norm = np.sqrt(gx ** 2 + gy ** 2) + 1e-6
    ux, uy = gx / norm, gy / norm
    acc = base.copy()
    for _ in range(diffusion_iter):
        dx, dy = ndi.sobel(acc, axis=1), ndi.sobel(acc, axis=0)
        acc += sigma * (dx * uy - dy * ux)
    diffused = acc.copy()
    acc = ndi.gaussian_filter(acc, sigma=0.6)
    acc = np.clip(acc * contrast_boost, 0, 1)
    if detail is not None:
        detail.update(diffused=diffused, boosted=acc)
    return (acc > 0.35).astype(np.uint8) * 255


# ------------------------------ K7 ----------------------------------------- #
def thinning_gate(binary_img: np.ndarray, reliability_img: np.ndarray,
                  rel_thresh: float = REL_THRESH) -> np.ndarray:
    """fingerprint_preprocess.py:166-170 - the boolean mask entering `skeletonize`."""
    m = (binary_img > 0).astype(bool)
    m = sk.remove_small_objects(m, min_size=THIN_MIN_OBJ)
    m = sk.remove_small_holes(m, area_threshold=THIN_MAX_HOLE)
    return m & (ndi.gaussian_filter(reliability_img, sigma=2.0) > rel_thresh)


def thin_and_clean(gate: np.ndarray, table: Optional[np.ndarray] = None) -> np.ndarray:
    """fingerprint_preprocess.py:171-177 - skeletonize, then drop pixels whose 3x3 sum
    (reflect border, centre included) is <= 1."""
    skel = sk.skeletonize(gate, table)
    cnt = ndi.convolve(skel.astype(np.uint8), np.ones((3, 3), np.uint8))
    skel = skel & (cnt > 1)
    return (skel > 0).astype(np.uint8) * 255


def thinning_and_cleaning(binary_img: np.ndarray, orientation_img: np.ndarray,
                          reliability_img: np.ndarray, rel_thresh: float = REL_THRESH) -> np.ndarray:
    """fingerprint_preprocess.py:161-177 (`orientation_img` is unused there too)."""
    return thin_and_clean(thinning_gate(binary_img, reliability_img, rel_thresh))


# ------------------------------ K8 ----------------------------------------- #
def extract_minutiae(skel: np.ndarray) -> List[Dict]:
    """extract_features.py:38-69 - crossing number on `skel > 127`, row-major order,
    1-px border skipped; CN==1 ending, CN==3 bifurcation."""
    s = (skel > 127).astype(np.int32)
    h, w = s.shape
    out: List[Dict] = []
    if h < 3 or w < 3:
        return out
    c = s[1:-1, 1:-1]
    ring = [s[1:-1, 2:], s[:-2, 2:], s[:-2, 1:-1], s[:-2, :-2],
            s[1:-1, :-2], s[2:, :-2], s[2:, 1:-1], s[2:, 2:]]       # E,NE,N,NW,W,SW,S,SE
    cn = sum(np.abs(ring[i] - ring[(i + 1) % 8]) for i in range(8)) // 2
    ys, xs = np.nonzero((c == 1) & ((cn == 1) | (cn == 3)))
    for y, x in zip(ys, xs):
        out.append({"x": int(x + 1), "y": int(y + 1),
                    "type": "ending" if cn[y, x] == 1 else "bifurcation"})
    return out


# ------------------------------ K9 ----------------------------------------- #
def nms_adaptive(minutiae: List[Dict], density_map: np.ndarray, base_dist: float = 8.0) -> List[Dict]:
    """post_processing.py:10-32 - including the last-writer-wins behaviour (the visit
    never skips a suppressed point)."""
    if not minutiae:
        return []
    pts = np.array([[m["x"], m["y"]] for m in minutiae])
    q = np.array([m.get("quality", 1.0) for m in minutiae])
    keep = np.zeros(len(minutiae), dtype=bool)
    tree = cKDTree(pts)
    for i in np.argsort(-q):
        if keep[i]:
            continue
        d = density_map[minutiae[i]["y"], minutiae[i]["x"]]
        keep[i] = True
        for j in tree.query_ball_point(pts[i], r=base_dist / (0.5 + d)):
            if j != i:
                keep[j] = False
    return [m for i, m in enumerate(minutiae) if keep[i]]


def remove_redundant_oriented_adaptive(minutiae: List[Dict], density_map: np.ndarray,
                                       base_radius: float = 20.0,
                                       angle_thresh: float = np.deg2rad(30)) -> List[Dict]:
    """post_processing.py:37-64."""
    if not minutiae:
        return []
    pts = np.array([[m["x"], m["y"]] for m in minutiae])
    tree = cKDTree(pts)
    gone = set()
    for i, a in enumerate(minutiae):
        if i in gone:
            continue
        d = density_map[a["y"], a["x"]]
        r = base_radius * (1.0 + (1.0 - a.get("quality", 1.0))) / (0.5 + d)
        for j in tree.query_ball_point(pts[i], r=r):
            if j <= i or j in gone:
                continue
            dth = a["orientation"] - minutiae[j]["orientation"]
            if abs(np.arctan2(np.sin(dth), np.cos(dth))) < angle_thresh:
                gone.add(i if float(a.get("quality", 1.0)) < float(minutiae[j].get("quality", 1.0)) else j)
    return [m for k, m in enumerate(minutiae) if k not in gone]


def postprocess_minutiae(minutiae: List[Dict], skel: np.ndarray, gray: Optional[np.ndarray] = None,
                         params: Optional[Dict] = None, detail: Optional[dict] = None) -> List[Dict]:
    """post_processing.py:69-137 - scoring, NMS, redundancy removal, top-K."""
    if not minutiae or skel is None:
        return []
    params = params or {}
    qwin = params.get("quality_window", 25)
    qth = params.get("quality_threshold", 0.15)
    coh_th = params.get("coherence_threshold", 0.2)
    min_dist = params.get("min_distance", 8.0)
    margin = params.get("margin", 30)
    max_m = params.get("max_minutiae", 60)
    patch_r = params.get("patch_radius", 15)

    on = (skel > 0).astype(np.uint8)
    h, w = on.shape
    density = cv2.blur(on.astype(np.float32), (qwin, qwin))
    density /= (density.max() + 1e-6)
    _, orient, coh = compute_orientation_map(gray if gray is not None else on)
    coh = np.clip(coh, 0, 1)

    scored = []
    for m in minutiae:
        x, y = int(m["x"]), int(m["y"])
        if not (margin <= x < w - margin and margin <= y < h - margin):
            continue
        c_here, d_here = float(coh[y, x]), float(density[y, x])
        if d_here < qth or c_here < coh_th:
            continue
        win = orient[max(0, y - patch_r):min(h, y + patch_r), max(0, x - patch_r):min(w, x + patch_r)]
        stab = float(np.exp(-3.0 * np.std(win))) if win.size > 0 else 0.0
        bonus = 1.0 - 0.5 * ((abs(x - w / 2) / (w / 2)) ** 2 + (abs(y - h / 2) / (h / 2)) ** 2)
        lit = float(on[y, x])
        q = (0.5 * c_here + 0.25 * d_here + 0.1 * stab + 0.1 * lit) * bonus
        m.update({"orientation": float(orient[y, x]), "quality": q, "coherence": c_here,
                  "angular_stability": stab})
        scored.append(m)
    if detail is not None:
        detail.update(density=density, orient=orient, coherence=coh,
                      scored=[dict(m) for m in scored])
    kept = nms_adaptive(scored, density_map=density, base_dist=min_dist)
    if detail is not None:
        detail.update(after_nms=[dict(m) for m in kept])
    kept = remove_redundant_oriented_adaptive(kept, density_map=density, base_radius=20.0,
                                              angle_thresh=np.deg2rad(30))
    return sorted(kept, key=lambda m: float(m["quality"]), reverse=True)[:max_m]


# ------------------------------ pipeline ----------------------------------- #
def preprocess_fingerprint(img: np.ndarray) -> Dict[str, np.ndarray]:
    """fingerprint_preprocess.py:182-225 without the debug-file side effects and without
    the `orientation_vis` overlay (a debug artefact)."""
    try:
        normalized = normalize_image(img)
        denoised = denoise_image(normalized)
        segmented, mask = segment_fingerprint(denoised)
        binary = binarize(segmented)
        blk, orient_img, reliability = compute_orientation_map(
            segmented, block_size=16, smooth_sigma=3.0, invert_if_needed=True,
            smooth_orientation_sigma=3.0, mask=mask)
        binary_smooth = smooth_fingerprint_skeleton(binary)
        skeleton = thinning_and_cleaning(binary_smooth, orient_img, reliability)
        return {"normalized": normalized, "denoised": denoised, "segmented": segmented,
                "mask": mask, "binary": binary, "skeleton": skeleton,
                "orient_blocks": blk, "orient_img": orient_img, "reliability": reliability,
                "binary_smooth": binary_smooth}
    except Exception as e:  # same wrapping as the reference (:224-225)
        raise RuntimeError(f"preprocess_fingerprint failed: {e}") from e


def skeleton_file_roundtrip(skel: np.ndarray) -> np.ndarray:
    """The hand-off between the reference's two stages, with the real codec: `cv2.imwrite(<base>_skeleton.jpg, skel)`
    (run_preprocessing.py:137-140, JPEG at OpenCV's default quality 95) then `cv2.imread(path, IMREAD_GRAYSCALE)`
    (extract_features.py:83).  imencode/imdecode are the in-memory forms of the same codec calls."""
    ok, buf = cv2.imencode(".jpg", skel)
    assert ok
    return cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)


def enhance_to_minutiae(img: np.ndarray, params: Optional[Dict] = None, handoff: str = "file") -> Dict:
    """Whole hot path for one image as the reference's CLI computes it: K1..K7 (`preprocess_fingerprint`), the
    skeleton written to / read from a quality-95 JPEG, then K8 and K9 on the DECODED grey image
    (extract_features.py:83-92: `extract_minutiae(skel)`, `postprocess_minutiae(raw, skel, skel)`).
    `handoff="memory"` skips the file (the two functions called on the in-memory skeleton)."""
    res = preprocess_fingerprint(img)
    sk = res["skeleton"]
    if handoff == "file":
        sk = skeleton_file_roundtrip(sk)
        res["skeleton_file"] = sk
    raw = extract_minutiae(sk)
    refined = postprocess_minutiae([dict(m) for m in raw], sk, sk, params)
    res["raw_minutiae"] = raw
    res["minutiae"] = refined
    return res

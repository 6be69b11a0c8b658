"""numpy restatements of the INSIDES of the library calls on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

`oracle.ref_pipeline` calls OpenCV / SciPy / NumPy exactly like the reference does;
this module restates what those calls compute, in plain numpy, because that is
the specification the CUDA kernels were written from (a kernel cannot call
`cv2.createCLAHE`).  Every function here is checked against the library itself
in `tests/test_oracle_stages.py` - bit-for-bit where the arithmetic is integer or
fixed point, and where the float operation order could be pinned.

Third-party versions the semantics were pinned on (this image): OpenCV 4.13.0,
SciPy 1.18.1, NumPy 2.3.5.  The reference pins `opencv>=4.8,<5`, `scipy>=1.11,<2`,
`numpy<2.0` (`config/environment.yml:9-34`); NumPy 2 differs from 1.x on this path only
in scalar promotion of the `+1e-12` terms (NEP 50), which the oracle inherits from the
NumPy that is actually installed.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- #
# border helpers
# --------------------------------------------------------------------------- #

def reflect101(i, n):
    """OpenCV BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def reflect_dup(i, n):
    """scipy.ndimage mode='reflect' index map (dcba|abcd|dcba), any distance."""
    i = np.asarray(i)
    p = 2 * n
    i = np.mod(i, p)
    return np.where(i >= n, p - 1 - i, i)


def pad101(a, r):
    h, w = a.shape
    return a[reflect101(np.arange(-r, h + r), h)][:, reflect101(np.arange(-r, w + r), w)]


# --------------------------------------------------------------------------- #
# K1: percentile stretch  (fingerprint_preprocess.py:15-23)
# --------------------------------------------------------------------------- #

def percentile_u8_as_f32(hist: np.ndarray, n: int, q: float) -> np.float32:
    """`np.percentile(img.astype(f32)/255, q)` from the 256-bin histogram.

    NumPy 2.3 (`_function_base_impl.py::percentile/_quantile/_lerp`): q is divided by
    float32(100) -> float32; virtual index `(n-1)*q` in float32; gamma = frac in float32;
    `a + (b-a)*g`, replaced by `b - (b-a)*(1-g)` when g >= 0.5, all float32."""
    cum = np.cumsum(hist)

    def kth(k):
        return np.float32(int(np.searchsorted(cum, k + 1))) / np.float32(255.0)

    qf = np.float32(q) / np.float32(100)
    vi = np.float32(n - 1) * qf
    lo_f = np.floor(vi)
    lo = int(lo_f)
    hi = lo + 1
    if vi >= n - 1:
        lo = hi = n - 1
    g = np.float32(vi - lo_f)
    a, b = kth(lo), kth(hi)
    d = np.float32(b - a)
    if g >= 0.5:
        return np.float32(b - np.float32(d * np.float32(np.float32(1) - g)))
    return np.float32(a + np.float32(d * g))


def stretch_lut(img: np.ndarray) -> np.ndarray:
    """256-entry map equal to `(clip((f-p0.5)/(p99.5-p0.5+1e-12),0,1)*255).astype(u8)`."""
    hist = np.bincount(img.ravel(), minlength=256)
    lo = percentile_u8_as_f32(hist, img.size, 0.5)
    hi = percentile_u8_as_f32(hist, img.size, 99.5)
    span = np.float32(np.float32(hi - lo) + np.float32(1e-12))
    lut = np.zeros(256, np.uint8)
    for v in range(256):
        fv = np.float32(v) / np.float32(255)
        t = np.float32(np.float32(fv - lo) / span)
        t = min(max(t, np.float32(0)), np.float32(1))
        lut[v] = np.uint8(np.float32(t * np.float32(255)))
    return lut


# --------------------------------------------------------------------------- #
# CLAHE  (cv2.createCLAHE(clip, (8,8)).apply on uint8; OpenCV imgproc/clahe.cpp)
# --------------------------------------------------------------------------- #

def clahe_tile_lut(hist: np.ndarray, clip_limit: int, area: int) -> np.ndarray:
    hist = hist.astype(np.int64).copy()
    clipped = int(np.maximum(hist - clip_limit, 0).sum())
    hist = np.minimum(hist, clip_limit)
    batch = clipped // 256
    resid = clipped - batch * 256
    hist += batch
    if resid:
        step = max(256 // resid, 1)
        i = 0
        while i < 256 and resid > 0:
            hist[i] += 1
            i += step
            resid -= 1
    scale = np.float32(np.float32(255) / area)
    return np.clip(np.rint(np.cumsum(hist).astype(np.float32) * scale), 0, 255).astype(np.uint8)


def clahe(src: np.ndarray, clip: float, tiles: int = 8) -> np.ndarray:
    h, w = src.shape
    if w % tiles == 0 and h % tiles == 0:
        ext = src
    else:  # note: BOTH axes are padded, a divisible axis by a full `tiles`
        eh, ew = h + tiles - (h % tiles), w + tiles - (w % tiles)
        ext = src[reflect101(np.arange(eh), h)][:, reflect101(np.arange(ew), w)]
    th, tw = ext.shape[0] // tiles, ext.shape[1] // tiles
    area = tw * th
    cl = max(int(clip * area / 256), 1)
    luts = np.zeros((tiles, tiles, 256), np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            t = ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            luts[ty, tx] = clahe_tile_lut(np.bincount(t.ravel(), minlength=256), cl, area)
    inv_tw = np.float32(1.0) / np.float32(tw)
    inv_th = np.float32(1.0) / np.float32(th)
    txf = np.arange(w, dtype=np.float32) * inv_tw - np.float32(0.5)
    tyf = np.arange(h, dtype=np.float32) * inv_th - np.float32(0.5)
    tx1 = np.floor(txf).astype(np.int32)
    ty1 = np.floor(tyf).astype(np.int32)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    xa1, ya1 = np.float32(1) - xa, np.float32(1) - ya
    tx2, tx1 = np.minimum(tx1 + 1, tiles - 1), np.maximum(tx1, 0)
    ty2, ty1 = np.minimum(ty1 + 1, tiles - 1), np.maximum(ty1, 0)
    f = lambda ty, tx: luts[ty[:, None], tx[None, :], src].astype(np.float32)
    res = ((f(ty1, tx1) * xa1[None, :] + f(ty1, tx2) * xa[None, :]) * ya1[:, None] +
           (f(ty2, tx1) * xa1[None, :] + f(ty2, tx2) * xa[None, :]) * ya[:, None])
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- #
# K2: non-local means  (cv2.fastNlMeansDenoising uint8, photo/fast_nlmeans_denoising_invoker.hpp)
# --------------------------------------------------------------------------- #

def nlm_weight_table(h: float = 10.0, template: int = 7, search: int = 21):
    fixed_mult = 2147483647 // (search * search * 255)          # 19096 for 21x21
    tsq = template * template
    shift = int(np.ceil(np.log2(tsq)))                          # 6  (64 >= 49)
    mult = float(1 << shift) / tsq
    n = int(65025 / mult + 1)
    w = np.exp(-(np.arange(n, dtype=np.float64) * mult) / float(np.float32(h) * np.float32(h)))
    tab = np.rint(fixed_mult * w).astype(np.int64)
    tab[tab < 0.001 * fixed_mult] = 0
    return tab, shift, fixed_mult


def nlm(img: np.ndarray, h: float = 10.0, template: int = 7, search: int = 21) -> np.ndarray:
    tab, shift, _ = nlm_weight_table(h, template, search)
    th, sh = template // 2, search // 2
    b = th + sh
    ext = pad101(img, b).astype(np.int64)
    H, W = img.shape
    est = np.zeros((H, W), np.int64)
    wsum = np.zeros((H, W), np.int64)
    base = ext[sh:sh + H + 2 * th, sh:sh + W + 2 * th]
    for dy in range(-sh, sh + 1):
        for dx in range(-sh, sh + 1):
            other = ext[sh + dy:sh + dy + H + 2 * th, sh + dx:sh + dx + W + 2 * th]
            d2 = (base - other) ** 2
            ii = np.pad(np.cumsum(np.cumsum(d2, 0), 1), ((1, 0), (1, 0)))
            ssd = ii[template:, template:] - ii[:-template, template:] - ii[template:, :-template] + ii[:-template, :-template]
            wv = tab[ssd >> shift]
            est += wv * ext[b + dy:b + dy + H, b + dx:b + dx + W]
            wsum += wv
    return ((est + wsum // 2) // wsum).astype(np.uint8)


# --------------------------------------------------------------------------- #
# cv2.GaussianBlur on uint8: 8.8 fixed-point taps, 16.16 accumulation, +2^15 >> 16
# --------------------------------------------------------------------------- #
GAUSS3_SIGMA06_TAPS = (43, 170, 43)          # cv2.GaussianBlur(u8, (3,3), 0.6)
GAUSS5_SIGMA0_TAPS = (16, 64, 96, 64, 16)    # cv2.GaussianBlur(u8, (5,5), 0)


def gauss_u8(img: np.ndarray, taps) -> np.ndarray:
    """Domain: images of at least 12 rows/columns (OpenCV 4.13's 5x5 fixed-point path deviates from reflect-101 on
    row 1 of images with 8..11 rows; fingerprints are two orders of magnitude larger)."""
    r = len(taps) // 2
    ext = pad101(img, r).astype(np.int64)
    H, W = img.shape
    rows = np.zeros((H + 2 * r, W), np.int64)
    for k, t in enumerate(taps):
        rows += ext[:, k:k + W] * t
    out = np.zeros((H, W), np.int64)
    for k, t in enumerate(taps):
        out += rows[k:k + H, :] * t
    return ((out + (1 << 15)) >> 16).astype(np.uint8)


# --------------------------------------------------------------------------- #
# cv2.threshold(..., THRESH_OTSU) on uint8  (imgproc/thresh.cpp::getThreshVal_Otsu_8u)
# --------------------------------------------------------------------------- #

def otsu_u8(hist: np.ndarray, n: int) -> int:
    h = hist.astype(np.float64)
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * h[i]
    mu *= scale
    mu1 = q1 = 0.0
    best, best_t = 0.0, 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p = h[i] * scale
        mu1 *= q1
        q1 += p
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > best:
            best, best_t = sigma, i
    return best_t


# --------------------------------------------------------------------------- #
# structuring elements / binary morphology  (cv2.getStructuringElement(MORPH_ELLIPSE), erode/dilate)
# --------------------------------------------------------------------------- #

def ellipse_half_widths(k: int):
    """Half-width of each row of cv2.getStructuringElement(MORPH_ELLIPSE, (k,k))."""
    r = c = k // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    out = []
    for i in range(k):
        dy = i - r
        out.append(int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2))))
    return out


def morph_binary(mask: np.ndarray, k: int, op: str) -> np.ndarray:
    """erode (outside counts as set) / dilate (outside counts as clear) of a {0,255} mask."""
    hw = ellipse_half_widths(k)
    r = k // 2
    H, W = mask.shape
    m = mask > 0
    fill = op == "erode"
    p = np.pad(m, r, constant_values=fill)
    out = np.full((H, W), fill)
    for i, half in enumerate(hw):
        for dx in range(-half, half + 1):
            sl = p[i:i + H, r + dx:r + dx + W]
            out = (out & sl) if fill else (out | sl)
    return out.astype(np.uint8) * 255


# --------------------------------------------------------------------------- #
# cv2.boxFilter / cv2.blur on float32 holding integers (normalised, BORDER_REFLECT_101)
# --------------------------------------------------------------------------- #

def box_mean_f32(a_int: np.ndarray, k: int) -> np.ndarray:
    """Window sums are exact (integers below 2^53 in OpenCV's double accumulators);
    the result is float32(sum * (1.0/(k*k)))."""
    r = k // 2
    ext = pad101(a_int.astype(np.int64), r).astype(np.float64)
    ii = np.pad(np.cumsum(np.cumsum(ext, 0), 1), ((1, 0), (1, 0)))
    s = ii[k:, k:] - ii[:-k, k:] - ii[k:, :-k] + ii[:-k, :-k]
    return (s * (1.0 / (k * k))).astype(np.float32)


# --------------------------------------------------------------------------- #
# skimage threshold_otsu on a float patch holding integers  (via np.histogram, 256 bins)
# --------------------------------------------------------------------------- #

def patch_otsu(patch_int: np.ndarray):
    vals = patch_int.astype(np.int64).ravel()
    a, b = int(vals.min()), int(vals.max())
    if a == b:
        return np.float32(a)
    fa, fb = np.float32(a), np.float32(b)
    step = np.float32(np.float32(fb - fa) / np.float32(256))
    edges = (np.arange(257, dtype=np.float32) * step + fa).astype(np.float32)
    edges[-1] = fb
    ih = np.bincount(vals, minlength=256)
    counts = np.zeros(256, np.float32)
    norm = np.float32(fb - fa)
    for v in range(a, b + 1):
        if ih[v] == 0:
            continue
        fv = np.float32(v)
        i = int((np.float32(fv - fa) / norm) * np.float32(256))
        if i == 256:
            i -= 1
        if fv < edges[i]:
            i -= 1
        if fv >= edges[i + 1] and i != 255:
            i += 1
        counts[i] += ih[v]
    centers = ((edges[:-1] + edges[1:]) / np.float32(2.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w1 = np.cumsum(counts)
        w2 = np.cumsum(counts[::-1])[::-1]
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
        var = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return centers[int(np.argmax(var))]


# --------------------------------------------------------------------------- #
# scipy.ndimage.gaussian_filter / sobel on float32  (ni_filters.c::NI_Correlate1D)
# --------------------------------------------------------------------------- #

def gaussian_weights(sigma: float, truncate: float = 4.0):
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), r


def _corr1d_sym(a: np.ndarray, w: np.ndarray, r: int, axis: int) -> np.ndarray:
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    idx = np.arange(n)
    acc = a.astype(np.float64) * w[r]
    for ll in range(-r, 0):            # outermost taps first, as NI_Correlate1D does
        lo = a[reflect_dup(idx + ll, n)].astype(np.float64)
        hi = a[reflect_dup(idx - ll, n)].astype(np.float64)
        acc = acc + (lo + hi) * w[ll + r]
    return np.moveaxis(acc.astype(np.float32), 0, axis)


def gaussian_filter_f32(a: np.ndarray, sigma: float) -> np.ndarray:
    """axis 0 then axis 1, float64 accumulation, float32 intermediate, mode='reflect'."""
    w, r = gaussian_weights(sigma)
    return _corr1d_sym(_corr1d_sym(a.astype(np.float32), w, r, 0), w, r, 1)


def ndi_sobel_f32(a: np.ndarray, axis: int) -> np.ndarray:
    """[-1,0,1] along `axis`, then [1,2,1] along the other, mode='reflect', f32 intermediate."""
    a = a.astype(np.float32)

    def line(arr, kind, ax):
        arr = np.moveaxis(arr, ax, 0)
        n = arr.shape[0]
        idx = np.arange(n)
        lo = arr[reflect_dup(idx - 1, n)].astype(np.float64)
        hi = arr[reflect_dup(idx + 1, n)].astype(np.float64)
        out = (hi - lo) if kind == "d" else (arr.astype(np.float64) * 2.0 + (lo + hi))
        return np.moveaxis(out.astype(np.float32), 0, ax)

    return line(line(a, "d", axis), "s", 1 - axis)


def cv_sobel_f32(a: np.ndarray, dx: int) -> np.ndarray:
    """cv2.Sobel(f32, CV_32F, dx, 1-dx, ksize=3), BORDER_REFLECT_101.  OpenCV's vectorised body and
    its scalar row tail associate the three-term sum differently, so this matches to 1 ulp only."""
    e = pad101(a.astype(np.float32), 1)
    if dx:
        d = e[:, 2:] - e[:, :-2]
        return (d[:-2] + d[2:]) + np.float32(2) * d[1:-1]
    d = e[2:, :] - e[:-2, :]
    return (d[:, :-2] + d[:, 2:]) + np.float32(2) * d[:, 1:-1]


# --------------------------------------------------------------------------- #
# cv2.resize(float32, INTER_LINEAR): half-pixel centres, edge clamp
# --------------------------------------------------------------------------- #

def resize_linear_f32(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    sh, sw = src.shape

    def axis_coef(dn, sn):
        scale = np.float64(sn) / dn
        f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = (f - s).astype(np.float32)
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= sn - 1
        f[hi], s[hi] = 0, sn - 1
        return s, np.minimum(s + 1, sn - 1), f

    x0, x1, fx = axis_coef(dst_w, sw)
    y0, y1, fy = axis_coef(dst_h, sh)
    rows = src[:, x0] * (np.float32(1) - fx)[None, :] + src[:, x1] * fx[None, :]
    return rows[y0] * (np.float32(1) - fy)[:, None] + rows[y1] * fy[:, None]

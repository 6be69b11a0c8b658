"""NumPy restatement of the JPEG hand-off between the reference's two stages: forward "islow" DCT + quantisation at
quality 95, the inverse of which is `oracle/jpeg_idct.py`.

TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this package).

The reference writes every skeleton with `cv2.imwrite(<base>_skeleton.jpg, results["skeleton"])`
(/root/reference/src/preprocessing/run_preprocessing.py:137-140: JPEG, OpenCV's default quality 95, one grey component)
and the feature stage reads it back with `cv2.imread(path, cv2.IMREAD_GRAYSCALE)`
(/root/reference/src/features/extract_features.py:83) before `extract_minutiae` and
`postprocess_minutiae(raw, skel, skel)` (:89-92).  `postprocess_minutiae` therefore sees the RINGING grey levels of the
decoded JPEG (density counts `skel > 0`, the orientation map is computed on the grey values), so the lossy step is
part of the hot path's arithmetic.  Entropy coding is loss-free; what changes the pixels is

    pad to whole 8x8 blocks by edge replication  ->  sample - 128  ->  forward DCT (libjpeg "islow": the 13-bit
    fixed-point Loeffler-Ligtenberg-Moschytz factorisation of jfdctint.c, output scaled by 8)  ->  quantise with the
    Annex-K luminance table scaled for quality 95 (divisor = 8*q, round half away from zero)  ->  [file]  ->
    dequantise  ->  inverse islow DCT  ->  range limit

OpenCV's codec is libjpeg-turbo (bundled); its SIMD kernels are bit-identical to the C ones by design.  Restated from the
published algorithm, not from the library's source.  PINNED by tests/test_oracle_io.py: `jpeg_roundtrip` equals
`cv2.imdecode(cv2.imencode('.jpg', img))` bit for bit on skeletons, noise and odd sizes, and `fdct_quantise` equals the
coefficients the library's entropy decoder (`fpb_jpeg_coefficients`, host code) reads back from cv2's own stream.
"""
from __future__ import annotations

import numpy as np

from .jpeg_idct import idct_islow

CB, P1 = 13, 2
F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
         f2053=16819, f2562=20995, f3072=25172)

# ITU-T T.81 Annex K.1 luminance quantisation table, natural (row-major) order
STD_LUMA = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                     14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)


def quality_table(quality: int = 95) -> np.ndarray:
    """jpeg_set_quality(quality, force_baseline=TRUE): scale = 5000/q (q<50) or 200-2q; (std*scale+50)/100 in [1,255]."""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    return np.clip((STD_LUMA * scale + 50) // 100, 1, 255).astype(np.uint16)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct8(d, first_pass: bool):
    """One 8-point pass along the last axis.  first_pass: outputs scaled up by 2**P1; second: P1 removed again."""
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    sh = CB - P1 if first_pass else CB + P1
    if first_pass:
        o0, o4 = (t10 + t11) << P1, (t10 - t11) << P1
    else:
        o0, o4 = _descale(t10 + t11, P1), _descale(t10 - t11, P1)
    z1 = (t12 + t13) * F["f0541"]
    o2 = _descale(z1 + t13 * F["f0765"], sh)
    o6 = _descale(z1 + t12 * (-F["f1847"]), sh)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * F["f1175"]
    t4, t5, t6, t7 = t4 * F["f0298"], t5 * F["f2053"], t6 * F["f3072"], t7 * F["f1501"]
    z1, z2, z3, z4 = z1 * -F["f0899"], z2 * -F["f2562"], z3 * -F["f1961"] + z5, z4 * -F["f0390"] + z5
    o7 = _descale(t4 + z1 + z3, sh)
    o5 = _descale(t5 + z2 + z4, sh)
    o3 = _descale(t6 + z2 + z3, sh)
    o1 = _descale(t7 + z1 + z4, sh)
    return np.stack([o0, o1, o2, o3, o4, o5, o6, o7], axis=-1)


def fdct_quantise(img: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """uint8 [H, W] -> quantised coefficients int16 [bh, bw, 64] (natural order), as written to the file."""
    h, w = img.shape
    bh, bw = (h + 7) // 8, (w + 7) // 8
    pad = np.pad(img, ((0, bh * 8 - h), (0, bw * 8 - w)), mode="edge").astype(np.int64) - 128
    blk = pad.reshape(bh, 8, bw, 8).transpose(0, 2, 1, 3)                    # [bh, bw, y, x]
    ws = _fdct8(blk, True)                                                   # rows:    [.., y, u]
    ws = _fdct8(np.swapaxes(ws, -1, -2), False)                              # columns: [.., u, v]
    coef = np.swapaxes(ws, -1, -2).reshape(bh, bw, 64)                       # natural order v*8+u
    div = qt.astype(np.int64)[None, None, :] << 3
    mag = (np.abs(coef) + (div >> 1)) // div
    return (np.sign(coef) * mag).astype(np.int16)


def jpeg_roundtrip(img: np.ndarray, quality: int = 95) -> np.ndarray:
    """= cv2.imdecode(cv2.imencode('.jpg', img, [IMWRITE_JPEG_QUALITY, quality])[1], IMREAD_GRAYSCALE) for uint8 [H, W]."""
    qt = quality_table(quality)
    h, w = img.shape
    return idct_islow(fdct_quantise(img, qt), qt, w, h)

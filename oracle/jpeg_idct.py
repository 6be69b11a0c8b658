"""NumPy restatement of libjpeg's "islow" 8x8 inverse DCT + dequantisation + range limiting.

TEST INFRASTRUCTURE.  The reference reads its inputs with `cv2.imread(path, cv2.IMREAD_GRAYSCALE)`
(/root/reference/src/preprocessing/run_preprocessing.py:41); OpenCV's JPEG codec is libjpeg-turbo 3.1.2 (bundled,
`cv2.getBuildInformation()`), default `dct_method = JDCT_ISLOW`: the 13-bit fixed-point Loeffler-Ligtenberg-Moschytz
factorisation of jidctint.c.  Pinned by tests/test_oracle_io.py: coefficients from the library's entropy decoder
through this function equal `cv2.imdecode(..., IMREAD_GRAYSCALE)` bit for bit on every tested stream.
"""
import numpy as np

CB, P1 = 13, 2
F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
         f2053=16819, f2562=20995, f3072=25172)


def _idct8(x):
    """x: int64 [..., 8] along the last axis -> unscaled outputs [..., 8]"""
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * F["f0541"]
    tmp2 = z1 + z3 * (-F["f1847"])
    tmp3 = z1 + z2 * F["f0765"]
    z2, z3 = x[..., 0], x[..., 4]
    tmp0, tmp1 = (z2 + z3) << CB, (z2 - z3) << CB
    t10, t13, t11, t12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F["f1175"]
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F["f0298"], tmp1 * F["f2053"], tmp2 * F["f3072"], tmp3 * F["f1501"]
    z1, z2, z3, z4 = z1 * -F["f0899"], z2 * -F["f2562"], z3 * -F["f1961"] + z5, z4 * -F["f0390"] + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return np.stack([t10 + tmp3, t11 + tmp2, t12 + tmp1, t13 + tmp0, t13 - tmp0, t12 - tmp1, t11 - tmp2, t10 - tmp3], axis=-1)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def idct_islow(coefs: np.ndarray, qt: np.ndarray, width: int, height: int) -> np.ndarray:
    """coefs int16 [bh, bw, 64] natural order, qt uint16 [64] -> uint8 [height, width]"""
    bh, bw, _ = coefs.shape
    c = coefs.astype(np.int64) * qt.astype(np.int64)[None, None, :]
    c = c.reshape(bh, bw, 8, 8)                                   # [.., v, u]
    ws = _descale(_idct8(np.swapaxes(c, -1, -2)), CB - P1)        # pass 1 along v (columns): result [.., u, y]
    ws = np.swapaxes(ws, -1, -2)                                  # [.., y, u]
    out = _descale(_idct8(ws), CB + P1 + 3)                       # pass 2 along u (rows): [.., y, x]
    i = out & 1023                                                # sample_range_limit + CENTERJSAMPLE
    px = np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)
    img = px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
    return np.ascontiguousarray(img[:height, :width])

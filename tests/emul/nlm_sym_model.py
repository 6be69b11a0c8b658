"""NumPy model of k_nlm_sym's formulation (csrc/k_front.cu): every patch distance is computed ONCE, for the half plane
{ox > 0} u {ox = 0, oy > 0} of the 21 x 21 search window, over the image extended by ten pixels on every side (the p range of the
kernel: the extension is walked for its q side only), and added to BOTH pixels of the pair; the centre offset is the constant
T[0].  Test infrastructure: pins the arithmetic claim (unsigned integer sums, order-free) against OpenCV without a GPU."""
import numpy as np


def weight_table():
    fixed_mult = 2147483647 // (21 * 21 * 255)
    t = np.rint(fixed_mult * np.exp(-(np.arange(529, dtype=np.float64) * (64.0 / 49.0)) / 100.0)).astype(np.int64)
    t[t < 0.001 * fixed_mult] = 0
    t[528] = 0
    return t


def nlm_sym(img: np.ndarray) -> np.ndarray:
    tab = weight_table()
    H, W = img.shape
    B = 13
    ext = np.pad(img.astype(np.int64), B, mode="reflect")           # reflect-101, as cv2.copyMakeBorder(BORDER_REFLECT_101)
    est = np.zeros((H, W), np.int64)
    wsum = np.zeros((H, W), np.int64)
    # p runs over the image extended by 10: p = (y, x) with -10 <= y < H + 10, -10 <= x < W + 10  (array index = coordinate + 10)
    PH, PW = H + 20, W + 20
    pix = ext[3:3 + PH, 3:3 + PW]                                    # I(p) for the extended p range
    inside_p = np.zeros((PH, PW), bool); inside_p[10:10 + H, 10:10 + W] = True
    for oy in range(-10, 11):
        for ox in range(0, 11):
            if ox == 0 and oy <= 0:
                continue
            # SSD of the 7x7 patches at p and p + o wherever both patches lie inside `ext`
            ssd = np.full((PH, PW), 1 << 40, np.int64)
            ys = slice(max(0, -oy), min(PH, PH - oy)); xs = slice(0, PW - ox)      # p + o inside the extended range too
            d2 = (ext[:, :ext.shape[1] - ox][max(0, -oy):ext.shape[0] - max(0, oy)] -
                  ext[:, ox:][max(0, oy):ext.shape[0] - max(0, -oy)]) ** 2        # (I(r) - I(r + o))^2 for every r where both exist
            ii = np.pad(np.cumsum(np.cumsum(d2, 0), 1), ((1, 0), (1, 0)))
            box = ii[7:, 7:] - ii[:-7, 7:] - ii[7:, :-7] + ii[:-7, :-7]            # box[r] = SSD of the patches whose top-left is r
            # patch of p = (y, x) (coordinates) has top-left ext index (y + 13 - 3, x + 13 - 3) = p-array index + (0, 0) shifted by 0
            r0y = max(0, -oy)                                                       # first ext row of d2
            py = np.arange(PH)[ys]; px = np.arange(PW)[xs]
            sub = box[(py - r0y)[:, None], px[None, :]]
            ssd[ys, xs] = sub
            w = tab[np.minimum(ssd >> 6, 528)]
            # p side: est(p) += w I(p + o) for p inside the image
            qy, qx = np.arange(PH) + oy, np.arange(PW) + ox
            okq = (qy >= 0) & (qy < PH); okx = (qx >= 0) & (qx < PW)
            Iq = np.zeros((PH, PW), np.int64)
            Iq[np.ix_(okq, okx)] = pix[np.ix_(qy[okq], qx[okx])]
            m = inside_p & (w > 0)
            est += (w * Iq * m)[10:10 + H, 10:10 + W]
            wsum += (w * m)[10:10 + H, 10:10 + W]
            # q side: est(p + o) += w I(p) for p + o inside the image
            contrib_e = np.zeros((PH + 40, PW + 40), np.int64); contrib_w = np.zeros_like(contrib_e)
            contrib_e[20 + oy:20 + oy + PH, 20 + ox:20 + ox + PW] = w * pix
            contrib_w[20 + oy:20 + oy + PH, 20 + ox:20 + ox + PW] = w
            est += contrib_e[30:30 + H, 30:30 + W]
            wsum += contrib_w[30:30 + H, 30:30 + W]
    est += tab[0] * img.astype(np.int64)
    wsum += tab[0]
    assert est.max() < 2 ** 32 and wsum.max() < 2 ** 32              # the kernel's accumulators are unsigned 32-bit
    return np.minimum((est + wsum // 2) // wsum, 255).astype(np.uint8)

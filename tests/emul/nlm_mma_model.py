"""NumPy model of k_nlm_mma.cu (csrc/k_nlm_mma.cu): the same tile, im2col layout, operand byte offsets, GEMM, sign test,
survivor arithmetic and lane / quadrant / chunk mapping, with the tensor-core product replaced by an integer matmul over
the very bytes the kernel stores.  TEST INFRASTRUCTURE: lets the CPU suite pin the kernel's arithmetic and index maps
against cv2.fastNlMeansDenoising before (and independently of) any GPU run."""
import numpy as np

BH, BW, CR, CC, NPR, TR, TS, NW = 16, 8, 36, 28, 32, 42, 64, 529
SSD_MAX = 33791
ORDER = [0, 7, 1, 8, 2, 3, 4, 5, 6]


def weight_table():
    fixed = 2147483647 // (21 * 21 * 255)
    t = np.zeros(NW, np.int64)
    for i in range(NW):
        w = np.exp(-(i * (64.0 / 49.0)) / 100.0)
        v = int(np.rint(fixed * w))
        t[i] = 0 if v < 0.001 * fixed else v
    t[NW - 1] = 0
    return t


def reflect101(i, n):
    if 0 <= i < n:
        return i
    if n == 1:
        return 0
    p = 2 * (n - 1)
    i %= p
    return p - i if i >= n else i


def op_off(n, pr):
    return (n >> 3) * 512 + (pr >> 1) * 128 + (n & 7) * 16 + (pr & 1) * 8


def operand_matrix(buf, rows):
    """read rows x 64 bytes back out of the canonical layout: byte (row, k) at (row%8)*16 + (row/8)*512 + (k/16)*128 + k%16"""
    r = np.arange(rows)[:, None]; k = np.arange(64)[None, :]
    return buf[(r % 8) * 16 + (r // 8) * 512 + (k // 16) * 128 + k % 16].astype(np.int64)


def nlm_block(img, y0, x0, lut, rng):
    H, W = img.shape
    ox, TX0 = (x0 & ~15) - 16, 3 + (x0 & 8)      # tile column 0 on a 16-byte boundary of the image row (TMA box)
    tile = np.zeros((TR, TS), np.uint8)
    for r in range(TR):
        for c in range(TS):
            tile[r, c] = img[reflect101(y0 - 13 + r, H), reflect101(ox + c, W)]
    sB = rng.integers(0, 256, CR * NPR * 64).astype(np.uint8)      # never-written bytes are garbage in the kernel too
    sA = np.zeros(128 * 64, np.uint8)
    hs = np.zeros((TR, CC), np.int64)
    for R in range(TR):
        for cxi in range(CC):
            win = tile[R, TX0 + cxi:TX0 + cxi + 8].copy()
            hs[R, cxi] = int((win[:7].astype(np.int64) ** 2).sum())
            for pr in range(7):
                cyi = R - pr
                if cyi < 0 or cyi >= CR:
                    continue
                o = op_off(cyi * NPR + cxi, pr)
                sB[o:o + 8] = win
                if 10 <= cxi < 10 + BW and 10 <= cyi < 10 + BH:
                    r, c = cyi - 10, cxi - 10
                    m = ((r >> 3) * 2 + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3)
                    o = op_off(m, pr)
                    a = win.copy(); a[7] = 0
                    sA[o:o + 8] = a
    nq = np.zeros((CR, NPR), np.int64)
    for cyi in range(CR):
        for cxi in range(CC):
            nq[cyi, cxi] = hs[cyi:cyi + 7, cxi].sum()
    na = -(nq >> 1)
    iq = np.zeros((CR, NPR), np.int64)
    iq[:, :CC] = tile[3:3 + CR, TX0 + 3:TX0 + 3 + CC]
    A = operand_matrix(sA, 128)
    out = np.zeros((BH, BW), np.uint8)
    sw = np.zeros(128, np.int64); swp = np.zeros(128, np.int64)
    stats = {"pairs": 0, "survivors": 0, "rows": 0, "rows_any": 0}
    for chunk in ORDER:
        Bc = operand_matrix(sB[chunk * 8192:(chunk + 1) * 8192], 128)
        D = A @ Bc.T                                               # [128 pixels, 128 candidates of the chunk]
        for warp in range(8):
            quad, half = warp & 3, warp >> 2
            rbase, cbase = (quad >> 1) * 8, (quad & 1) * 4
            crow0 = 4 * chunk
            if not (rbase <= crow0 < rbase + 28):
                continue
            for rr in (2 * half, 2 * half + 1):
                cyi = crow0 + rr
                stats["rows"] += 1
                row_any = False
                for lane in range(32):
                    r_abs, c_rel = rbase + (lane >> 2), lane & 3
                    np_ = nq[r_abs + 10, cbase + c_rel + 10]
                    bp = (np_ - SSD_MAX) >> 1
                    cp = np_ - 2 * bp
                    nbr = -bp if 0 <= cyi - r_abs <= 20 else -(1 << 30)
                    m = quad * 32 + lane
                    g = D[m, rr * 32 + cbase: rr * 32 + cbase + 24]
                    e = g + na[cyi, cbase:cbase + 24] + nbr
                    for j in range(24):
                        stats["pairs"] += 1
                        if e[j] >= 0 and c_rel <= j <= c_rel + 20:
                            row_any = True
                            stats["survivors"] += 1
                            ev = int(e[j]) & 0xFFFF
                            ssd = cp + (int(nq[cyi, cbase + j]) & 1) - 2 * ev
                            assert ssd >= 0
                            w = lut[min(ssd >> 6, NW - 1)]
                            sw[m] += w; swp[m] += w * iq[cyi, cbase + j]
                stats["rows_any"] += row_any
    for m in range(128):
        quad, lane = m >> 5, m & 31
        r, c = (quad >> 1) * 8 + (lane >> 2), (quad & 1) * 4 + (lane & 3)
        out[r, c] = min((swp[m] + sw[m] // 2) // sw[m], 255)
    return out, stats


def nlm_image(img, seed=0):
    H, W = img.shape
    lut = weight_table()
    rng = np.random.default_rng(seed)
    out = np.zeros_like(img)
    tot = {"pairs": 0, "survivors": 0, "rows": 0, "rows_any": 0}
    for y0 in range(0, H, BH):
        for x0 in range(0, W, BW):
            o, st = nlm_block(img, y0, x0, lut, rng)
            h, w = min(BH, H - y0), min(BW, W - x0)
            out[y0:y0 + h, x0:x0 + w] = o[:h, :w]
            for k in tot:
                tot[k] += st[k]
    return out, tot

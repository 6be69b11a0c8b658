"""NumPy statement of the EXTENSION rows G1/G2 (per-block ridge frequency + oriented Gabor enhancement).

TEST INFRASTRUCTURE.  **PARITY UNPINNED**: the reference contains no ridge-frequency or Gabor code, so
there is nothing in it to pin this against; this file states, loop by loop, the arithmetic `k_ridge_freq`, `k_freq_fill`
and `k_gabor` (csrc/k_gabor.cu) implement, after Hong, Wan & Jain, "Fingerprint image enhancement: algorithm and
performance evaluation" (IEEE TPAMI 20(8), 1998) on the 16x16 block grid of `compute_orientation_map`
(/root/reference/src/preprocessing/orientation.py:52-79).  The tests feed it the SAME block orientations the GPU used.
"""
from __future__ import annotations

import math

import numpy as np

DEFAULTS = dict(n_orient=16, min_period=3, max_period=25, sigma_factor=0.45, radius_factor=2.5, min_amplitude=8.0,
                default_period=9.0)
RMAX = 31


def _prm(params):
    d = dict(DEFAULTS)
    d.update(params or {})
    return d


def ridge_frequency_raw(img, mask, blk_theta, params=None) -> np.ndarray:
    """x-signature peak spacing per block (0 = invalid)."""
    p = _prm(params)
    h, w = img.shape
    nby, nbx = h // 16, w // 16
    out = np.zeros((nby, nbx), np.float32)
    f32 = np.float32
    for by in range(nby):
        for bx in range(nbx):
            th = float(blk_theta[by, bx])
            c, s = math.cos(th), math.sin(th)
            cx, cy = bx * 16 + 7.5, by * 16 + 7.5
            X = np.zeros(32, np.float32)
            for k in range(32):
                off = k - 15.5
                acc = f32(0)
                for t in range(16):
                    along = t - 7.5
                    x = cx + along * c + off * (-s)
                    y = cy + along * s + off * c
                    xi = min(max(int(math.floor(x + 0.5)), 0), w - 1)
                    yi = min(max(int(math.floor(y + 0.5)), 0), h - 1)
                    acc = f32(acc + f32(img[yi, xi]))
                X[k] = acc * f32(1.0 / 16.0)
            lft = np.concatenate([X[:1], X[:-1]])
            rgt = np.concatenate([X[1:], X[-1:]])
            Y = (lft + f32(2) * X + rgt) * f32(0.25)
            peaks = [k for k in range(1, 31) if Y[k] > Y[k - 1] and Y[k] >= Y[k + 1]]
            on = int((mask[by * 16:by * 16 + 16, bx * 16:bx * 16 + 16] > 0).sum()) if mask is not None else 256
            if len(peaks) >= 2 and (Y.max() - Y.min()) >= f32(p["min_amplitude"]) and on >= 77:
                period = f32(peaks[-1] - peaks[0]) / f32(len(peaks) - 1)
                if f32(p["min_period"]) <= period <= f32(p["max_period"]):
                    out[by, bx] = f32(1.0) / period
    return out


def fill_frequency(raw, params=None) -> np.ndarray:
    p = _prm(params)
    f32 = np.float32
    A = raw.astype(np.float32).copy()
    nby, nbx = A.shape
    if A.size == 0:
        return A

    def neigh(M, y, x, valid_only):
        s, c = f32(0), 0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy, xx = y + dy, x + dx
                if 0 <= yy < nby and 0 <= xx < nbx and (not valid_only or M[yy, xx] > 0):
                    s = f32(s + M[yy, xx]); c += 1
        return s, c

    for _ in range(8):
        B = A.copy()
        changed = False
        for y in range(nby):
            for x in range(nbx):
                if A[y, x] == 0:
                    s, c = neigh(A, y, x, True)
                    if c:
                        B[y, x] = s / f32(c); changed = True
        A = B
        if not changed:
            break
    s, c = f32(0), 0
    for v in A.ravel():
        if v > 0:
            s = f32(s + v); c += 1
    fill = s / f32(c) if c else f32(1.0 / p["default_period"])
    A[A == 0] = fill
    out = np.zeros_like(A)
    for y in range(nby):
        for x in range(nbx):
            s, c = neigh(A, y, x, False)
            out[y, x] = s / f32(c)
    return out


def gabor_bank(params=None):
    """{(period, orientation index): float32 [K,K]} zero-mean, L1-normalised."""
    p = _prm(params)
    bank = {}
    for per in range(p["min_period"], p["max_period"] + 1):
        sigma = p["sigma_factor"] * per
        R = max(1, min(int(math.ceil(p["radius_factor"] * sigma)), RMAX))
        v, u = np.mgrid[-R:R + 1, -R:R + 1].astype(np.float64)
        for oi in range(p["n_orient"]):
            phi = oi * math.pi / p["n_orient"]
            xn = -u * math.sin(phi) + v * math.cos(phi)
            g = np.exp(-(u * u + v * v) / (2.0 * sigma * sigma)) * np.cos(2.0 * math.pi * xn / per)
            g = g - g.sum() / g.size
            bank[(per, oi)] = (g / np.abs(g).sum()).astype(np.float32)
    return bank


def pick_filter(theta, freq, params=None):
    p = _prm(params)
    step = math.pi / p["n_orient"]
    t = math.fmod(float(theta), math.pi)
    if t < 0:
        t += math.pi
    oi = int(math.floor(t / step + 0.5))
    if oi >= p["n_orient"]:
        oi -= p["n_orient"]
    per = int(math.floor(float(np.float32(1.0) / np.float32(freq)) + 0.5))
    return min(max(per, p["min_period"]), p["max_period"]), oi


def gabor_enhance(img, mask, blk_theta, blk_freq, params=None):
    """(response float32, enhanced uint8); masked-out pixels -> 0 / 255."""
    bank = gabor_bank(params)
    h, w = img.shape
    nby, nbx = h // 16, w // 16
    resp = np.zeros((h, w), np.float32)
    enh = np.full((h, w), 255, np.uint8)
    if nby == 0 or nbx == 0:
        return resp, enh
    pad = RMAX
    P = np.pad(img.astype(np.float64), pad, mode="reflect")
    for qy in range(0, h, 16):
        for qx in range(0, w, 16):
            by, bx = min(qy // 16, nby - 1), min(qx // 16, nbx - 1)
            g = bank[pick_filter(blk_theta[by, bx], blk_freq[by, bx], params)].astype(np.float64)
            R = g.shape[0] // 2
            y1, x1 = min(qy + 16, h), min(qx + 16, w)
            region = P[qy + pad - R:y1 + pad + R, qx + pad - R:x1 + pad + R]
            win = np.lib.stride_tricks.sliding_window_view(region, g.shape)
            resp[qy:y1, qx:x1] = np.einsum("yxvu,vu->yx", win, g).astype(np.float32)
    e = np.clip(np.rint(np.float32(128.0) + np.float32(2.0) * resp), 0, 255).astype(np.uint8)
    if mask is not None:
        on = mask > 0
        resp = np.where(on, resp, np.float32(0))
        enh = np.where(on, e, np.uint8(255)).astype(np.uint8)
    else:
        enh = e
    return resp, enh
